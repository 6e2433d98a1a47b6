"""ctypes binding of libiper_b200.so — the C-ABI boundary (include/iper_b200.h).

There is deliberately no fallback: if the shared library is missing or a symbol is absent this module raises at
import time, and every op raises RuntimeError with iper_last_error() on a non-zero status.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libiper_b200.so")

c_void_p, c_int, c_float, c_ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong
c_size_t = ctypes.c_size_t

IPER_CONV_S1, IPER_CONV_S2, IPER_CONVT_4S2, IPER_CONV_ROW5 = 0, 1, 2, 3
IPER_EPI_PLANES, IPER_EPI_F32, IPER_EPI_SPADE, IPER_EPI_HEADS = 0, 1, 2, 3


class ConvGemmDesc(ctypes.Structure):
    """Mirror of iper_conv_gemm_desc (include/iper_b200.h) — field order and types must match exactly."""
    _fields_ = [
        ("a", c_void_p), ("a_planes", c_int), ("a_plane_stride", c_ll),
        ("N", c_int), ("H", c_int), ("W", c_int),
        ("a_pitch", c_int), ("a_coff", c_int), ("Cin", c_int),
        ("mode", c_int), ("ksize", c_int),
        ("w", c_void_p), ("w_planes", c_int), ("w_plane_stride", c_ll),
        ("rows", c_int), ("block_n", c_int),
        ("epi", c_int), ("bias", c_void_p), ("relu", c_int),
        ("out", c_void_p), ("out_planes", c_int), ("out_plane_stride", c_ll), ("out_pitch", c_int), ("out_coff", c_int),
        ("x", c_void_p), ("x_planes", c_int), ("x_plane_stride", c_ll), ("x_pitch", c_int), ("x_coff", c_int),
        ("mean_rstd", c_void_p), ("spade_C", c_int),
        ("bg", c_void_p), ("bg_batch_stride", c_ll), ("img", c_void_p), ("mask", c_void_p), ("pred", c_void_p),
        ("max_ctas", c_int),
        ("w8", c_void_p), ("wl8", c_void_p), ("cross_scale", c_float), ("tiles_m", c_int),
        ("stats_ws", c_void_p), ("cta_pair", c_int),
        ("w_scale_inv", c_void_p),
    ]


class AdamSeg(ctypes.Structure):
    """Mirror of iper_adam_seg (include/iper_b200.h)."""
    _fields_ = [("offset", c_ll), ("numel", c_ll), ("co", c_int), ("ci", c_int), ("taps", c_int), ("co_pad", c_int),
                ("ci_pad", c_int), ("reserved", c_int), ("fwd_offset", c_ll), ("dgrad_offset", c_ll)]


# name -> argtypes; every function returns int status except iper_last_error
SIGNATURES = {
    "iper_abi_version": [],
    "iper_rasterize_faces": [c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p, c_size_t,
                             c_void_p],
    "iper_raster_frames": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_float,
                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                           c_void_p, c_void_p, c_size_t, c_void_p],
    "iper_flow_from_fim_wim": [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "iper_encode_fim": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "iper_vis_f2pts": [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p],
    "iper_flow_resize": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "iper_conv_gemm": [ctypes.POINTER(ConvGemmDesc), c_void_p],
    "iper_conv_direct": [ctypes.POINTER(ConvGemmDesc), c_void_p, c_int, c_void_p],
    "iper_conv_stem": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_ll, c_int,
                       c_int, c_void_p, c_void_p],
    "iper_conv_stem_tc": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_ll, c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int,
                          c_int, c_void_p, c_void_p],
    "iper_stem_im2col": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_ll, c_int, c_int, c_void_p],
    "iper_instnorm_finalize": [c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p],
    "iper_instnorm_stats": [c_void_p, c_int, c_ll, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                            c_void_p],
    "iper_instnorm_apply": [c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                            c_ll, c_int, c_int, c_void_p, c_int, c_ll, c_int, c_int, c_void_p],
    "iper_tanh_nhwc_to_nchw": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "iper_lbs_shape": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "iper_lbs_frames": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                        c_void_p, c_void_p, c_void_p, c_void_p],
    "iper_warp_attention": [c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                            c_int, c_void_p, c_int, c_ll, c_int, c_int, c_void_p],
    "iper_warp_nhwc": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "iper_nchw_to_planes": [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_ll, c_int, c_int, c_void_p],
    "iper_planes_to_nchw": [c_void_p, c_int, c_ll, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "iper_nhwc_f32_to_nchw": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "iper_pred_to_u8": [c_void_p, c_int, c_int, c_void_p, c_void_p],
    "iper_morph": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "iper_conv_bf16": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                       c_void_p],
    "iper_conv_transposed_bf16": [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p],
    "iper_conv_wgrad_bf16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_ll, c_ll, c_int,
                             c_int, c_void_p],
    "iper_bias_grad_bf16": [c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p],
    "iper_adam_pack": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_float, c_float,
                       c_float, c_void_p, c_int, c_void_p, c_void_p, c_void_p],
    "iper_pad_nhwc_bf16": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_ll, c_ll, c_ll, c_ll, c_int, c_void_p, c_void_p],
    "iper_thin_wgrad_bf16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_ll, c_ll,
                             c_void_p],
    "iper_warp_bf16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "iper_warp_bwd_bf16": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
    "iper_att_combine_bf16": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_int, c_void_p, c_void_p, c_void_p],
    "iper_att_combine_bwd_bf16": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p],
    "iper_norm_stats_bf16": [c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p],
    "iper_norm_apply_bf16": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_float, c_int, c_float, c_void_p, c_void_p],
    "iper_norm_bwd_bf16": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_float, c_int, c_float, c_void_p, c_void_p,
                           c_void_p, c_void_p, c_void_p],
    "iper_gen_create": [c_void_p, c_int, c_int, c_int, c_void_p],
    "iper_gen_load_weight": [c_void_p, ctypes.c_char_p, c_void_p, c_void_p, c_int],
    "iper_gen_pack": [c_void_p, c_void_p, c_size_t, c_void_p],
    "iper_gen_forward_src": [c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p],
    "iper_gen_forward_tsf": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_size_t, c_void_p],
    "iper_canny_edges": [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_void_p,
                         c_void_p, c_void_p, c_void_p, c_void_p],
    "iper_morph_image": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "iper_uv_warp": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p],
    "iper_uv_merge": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p],
}


# entry points whose return type is not the int status: name -> (restype, argtypes)
OTHER_SIGNATURES = {
    "iper_last_error": (ctypes.c_char_p, []),
    "iper_conv_halo_plan": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_int]),
    "iper_raster_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "iper_raster_set_contraction": (c_int, [c_int]),
    "iper_raster_get_contraction": (c_int, []),
    "iper_vis_f2pts_workspace_bytes": (c_size_t, [c_int, c_int]),
    "iper_gen_destroy": (None, [c_void_p]),
    "iper_gen_packed_bytes": (c_size_t, [c_void_p]),
    "iper_gen_src_cache_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "iper_gen_src_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "iper_gen_tsf_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "ipercore_b200: %s is missing — build it with `python -m ipercore_b200.build` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for these kernels." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = ABI drift; fail loudly
        fn.argtypes = argtypes
        fn.restype = c_int
    for name, (restype, argtypes) in OTHER_SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    return lib


lib = _load()


_LAUNCHES = 0


def launch_count():
    """Number of libiper_b200 kernel launches issued so far by this process (a lower bound: every C-ABI op is at least one launch)."""
    return _LAUNCHES


def check(status, what=""):
    global _LAUNCHES
    _LAUNCHES += 1
    if status != 0:
        msg = lib.iper_last_error()
        raise RuntimeError("iper_b200 %s failed (status %d): %s" % (what, status, msg.decode() if msg else "?"))
