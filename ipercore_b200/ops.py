"""Tensor-level wrappers over the C ABI.  torch is used only as allocator and stream provider: every call passes
raw device pointers + the current CUDA stream to libiper_b200.so (see include/iper_b200.h for reference citations).
"""
import math

import torch

from . import _lib
from ._lib import (ConvGemmDesc, IPER_CONV_ROW5, IPER_CONV_S1, IPER_CONV_S2, IPER_CONVT_4S2, IPER_EPI_F32, IPER_EPI_HEADS,
                   IPER_EPI_PLANES, IPER_EPI_SPADE, check, lib)

# nmr.py:225 eye z, converted to float32 inside nr.look_at
EYE_Z = float(torch.tensor(-(1.0 / math.tan(math.radians(30.0)) + 1.0), dtype=torch.float32))
NEAR, FAR = 0.1, 100.0   # neural_renderer defaults used by rasterize_face_index_map_and_weight_map


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _req(t, dtype, name):
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise ValueError("%s must be a contiguous CUDA %s tensor (got %s, cuda=%s, contiguous=%s)" %
                         (name, dtype, t.dtype, t.is_cuda, t.is_contiguous()))
    return t


FMT_H, FMT_HL, FMT_H8 = 1, 2, 3      # planes formats (include/iper_b200.h, common.cuh)


class Planes:
    """NHWC activation tensor in one of the planes formats; `P` is the FORMAT passed to the C ABI:
      1: data (1,N,H,W,pitch) fp16            value = hi
      2: data (2,N,H,W,pitch) fp16            value = hi + lo
      3: data (2,...) fp16 storage; plane 1 holds two e4m3 planes (a8 = e4m3(8x), l8 = e4m3(2^14 (x-hi)))
    A `Planes` may be a channel window [coff, coff+C) of a wider buffer (used for the decoder's concatenations)."""

    def __init__(self, data, C=None, coff=0, fmt=None):
        assert data.dim() == 5 and data.dtype == torch.float16 and data.is_contiguous()
        self.data = data
        _, self.N, self.H, self.W, self.pitch = data.shape
        self.P = data.shape[0] if fmt is None else fmt
        assert data.shape[0] == (1 if self.P == 1 else 2)
        self.C = self.pitch if C is None else C
        self.coff = coff

    @staticmethod
    def empty(fmt, N, H, W, C, device, pitch=None):
        return Planes(torch.empty((1 if fmt == 1 else 2, N, H, W, pitch or C), dtype=torch.float16, device=device),
                      C=C, fmt=fmt)

    def window(self, coff, C):
        return Planes(self.data, C=C, coff=self.coff + coff, fmt=self.P)

    @property
    def plane_stride(self):
        return self.N * self.H * self.W * self.pitch

    def ptr(self):
        return self.data.data_ptr()

    def to_nchw(self):
        out = torch.empty((self.N, self.C, self.H, self.W), dtype=torch.float32, device=self.data.device)
        check(lib.iper_planes_to_nchw(self.ptr(), self.P, self.plane_stride, self.N, self.C, self.H * self.W,
                                      self.pitch, self.coff, out.data_ptr(), _stream()), "planes_to_nchw")
        return out

    @staticmethod
    def from_nchw(x, P, pitch=None, out=None):
        x = _req(x.float().contiguous(), torch.float32, "x")
        N, C, H, W = x.shape
        if out is None:
            out = Planes.empty(P, N, H, W, C, x.device, pitch=pitch)
        check(lib.iper_nchw_to_planes(x.data_ptr(), N, C, H * W, out.ptr(), out.P, out.plane_stride, out.pitch,
                                      out.coff, _stream()), "nchw_to_planes")
        return out


# ------------------------------------------------------------------------------------------------------------------
# raster / flow
# ------------------------------------------------------------------------------------------------------------------
def _raster_workspace(B, nf, from_verts, device):
    """Caller-owned rasteriser scratch (tile bins + projected corners), sized by the library."""
    return torch.empty((max(int(lib.iper_raster_workspace_bytes(B, nf, from_verts)), 1),), dtype=torch.uint8, device=device)


def rasterize_faces(faces, image_size, near=NEAR, far=FAR):
    """neural_renderer.rasterize_face_index_map_and_weight_map(faces, image_size, False) (nmr.py:337,356)."""
    faces = _req(faces, torch.float32, "faces")
    B, nf = faces.shape[:2]
    fim = torch.empty((B, image_size, image_size), dtype=torch.int32, device=faces.device)
    wim = torch.empty((B, image_size, image_size, 3), dtype=torch.float32, device=faces.device)
    ws = _raster_workspace(B, nf, 0, faces.device)
    check(lib.iper_rasterize_faces(faces.data_ptr(), B, nf, image_size, near, far, fim.data_ptr(), wim.data_ptr(),
                                   _ptr(ws), ws.numel(), _stream()), "rasterize_faces")
    return fim, wim


def raster_frames(verts, cams, faces, image_size, want_fim=True, want_f2pts=True, fused=None):
    """render_fim_wim (nmr.py:319-342) for a batch; `fused` = dict(map_fn, f_uvs2img, uv_img, src_f2pts) additionally
    produces tsf_inputs (B,6,S,S) and Tst (B,ns,S,S,2) in the same launch (flowcomposition.py:206-248, 514-582)."""
    verts = _req(verts, torch.float32, "verts"); cams = _req(cams, torch.float32, "cams")
    faces = _req(faces, torch.int32, "faces")
    B, nv = verts.shape[:2]; nf = faces.shape[0]; S = image_size; dev = verts.device
    fim = torch.empty((B, S, S), dtype=torch.int32, device=dev) if want_fim else None
    wim = torch.empty((B, S, S, 3), dtype=torch.float32, device=dev) if want_fim else None
    f2pts = torch.empty((B, nf, 3, 2), dtype=torch.float32, device=dev) if want_f2pts else None
    tsf = Tst = None
    map_fn = f_uv = uv_img = src_f2pts = None; ns = 0
    if fused is not None:
        map_fn = _req(fused["map_fn"], torch.float32, "map_fn"); f_uv = _req(fused["f_uvs2img"], torch.float32, "f_uvs2img")
        uv_img = _req(fused["uv_img"], torch.float32, "uv_img"); src_f2pts = _req(fused["src_f2pts"], torch.float32, "src_f2pts")
        assert map_fn.shape == (nf + 1, 3) and f_uv.shape == (nf, 3, 2) and uv_img.shape[-3:] == (3, S, S)
        ns = src_f2pts.shape[0]
        tsf = torch.empty((B, 6, S, S), dtype=torch.float32, device=dev)
        Tst = torch.empty((B, ns, S, S, 2), dtype=torch.float32, device=dev)
    ws = _raster_workspace(B, nf, 1, dev)
    check(lib.iper_raster_frames(verts.data_ptr(), cams.data_ptr(), faces.data_ptr(), B, nv, nf, S, EYE_Z, NEAR, FAR,
                                 _ptr(fim), _ptr(wim), _ptr(f2pts), _ptr(map_fn), _ptr(f_uv), _ptr(uv_img),
                                 _ptr(src_f2pts), ns, _ptr(tsf), _ptr(Tst), _ptr(ws), ws.numel(), _stream()),
          "raster_frames")
    return dict(fim=fim, wim=wim, f2pts=f2pts, tsf_inputs=tsf, Tst=Tst)


def cal_bc_transform(src_f2pts, fims, wims):
    """SMPLRenderer.cal_bc_transform (nmr.py:713-757): item i combines src_f2pts[i] with fims[i]/wims[i]."""
    src_f2pts = _req(src_f2pts.contiguous(), torch.float32, "src_f2pts")
    fims = _req(fims.contiguous(), torch.int32, "fims"); wims = _req(wims.contiguous(), torch.float32, "wims")
    b, S = fims.shape[0], fims.shape[1]; nf = src_f2pts.shape[1]
    T = torch.empty((b, S, S, 2), dtype=torch.float32, device=fims.device)
    check(lib.iper_flow_from_fim_wim(src_f2pts.data_ptr(), 1, fims.data_ptr(), wims.data_ptr(), b, 1, nf, S,
                                     T.data_ptr(), _stream()), "flow_from_fim_wim")
    return T


def encode_fim(fim, map_fn, transpose=True):
    fim = _req(fim.contiguous(), torch.int32, "fim"); map_fn = _req(map_fn.contiguous(), torch.float32, "map_fn")
    nb, S = fim.shape[0], fim.shape[1]; ch = map_fn.shape[1]; nf = map_fn.shape[0] - 1
    shape = (nb, ch, S, S) if transpose else (nb, S, S, ch)
    out = torch.empty(shape, dtype=torch.float32, device=fim.device)
    check(lib.iper_encode_fim(fim.data_ptr(), map_fn.data_ptr(), nb, nf, ch, S, int(transpose), out.data_ptr(),
                              _stream()), "encode_fim")
    return out


def vis_f2pts(f2pts, fims, face_k_nearest):
    """SMPLRenderer.get_vis_f2pts (nmr.py:639-681) on (bs,nf,3,2|3) / (nf,3,2|3) corners and (bs,S,S) / (S,S) face maps."""
    single = f2pts.dim() == 3
    if single:
        f2pts, fims = f2pts[None], fims[None]
    f2pts = _req(f2pts.float().contiguous(), torch.float32, "f2pts")
    fims = _req(fims.int().contiguous(), torch.int32, "fims")
    nb = _req(face_k_nearest.long().contiguous(), torch.int64, "face_k_nearest")
    B, nf = f2pts.shape[:2]
    S = fims.shape[-1]
    if fims.shape[0] != B or tuple(nb.shape[:1]) != (nf,):
        raise ValueError("vis_f2pts: f2pts %s, fims %s, face_k_nearest %s do not match" % (tuple(f2pts.shape), tuple(fims.shape), tuple(nb.shape)))
    out = torch.empty_like(f2pts)
    ws = torch.empty((max(int(lib.iper_vis_f2pts_workspace_bytes(B, nf)), 1),), dtype=torch.uint8, device=f2pts.device)
    check(lib.iper_vis_f2pts(f2pts.data_ptr(), f2pts[0, 0].numel(), fims.data_ptr(), nb.data_ptr(), nb.shape[1], B, nf, S,
                             out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "vis_f2pts")
    return out[0] if single else out


def flow_resize(T, h, w):
    """LWB.resize_trans (attlwb_spade_resunet.py:175-182) on (..., S, S, 2) -> (..., h, w, 2)."""
    T = _req(T.contiguous(), torch.float32, "T")
    lead = T.shape[:-3]; S = T.shape[-2]
    n = int(torch.tensor(lead).prod()) if len(lead) else 1
    out = torch.empty((*lead, h, w, 2), dtype=torch.float32, device=T.device)
    check(lib.iper_flow_resize(T.data_ptr(), n, S, h, w, out.data_ptr(), _stream()), "flow_resize")
    return out


# ------------------------------------------------------------------------------------------------------------------
# generator building blocks
# ------------------------------------------------------------------------------------------------------------------
def _fill_desc(a, mode, ksize, rows, block_n, epi, bias=None, relu=False, out=None, x=None, mean_rstd=None,
               spade_C=0, heads=None, max_ctas=0, tiles_m=0, stats_ws=None, cta_pair=0):
    d = ConvGemmDesc()
    d.a = a.ptr(); d.a_planes = a.P; d.a_plane_stride = a.plane_stride
    d.N, d.H, d.W = a.N, a.H, a.W
    d.a_pitch, d.a_coff, d.Cin = a.pitch, a.coff, a.C
    d.mode, d.ksize = mode, ksize
    d.rows, d.block_n = rows, block_n
    d.epi = epi; d.bias = _ptr(bias); d.relu = int(relu)
    if out is not None:
        if isinstance(out, Planes):
            d.out = out.ptr(); d.out_planes = out.P; d.out_plane_stride = out.plane_stride
            d.out_pitch = out.pitch; d.out_coff = out.coff
        else:  # fp32 NHWC tensor (N,H,W,pitch)
            d.out = out.data_ptr(); d.out_planes = 1; d.out_plane_stride = 0; d.out_pitch = out.shape[-1]; d.out_coff = 0
    if x is not None:
        d.x = x.ptr(); d.x_planes = x.P; d.x_plane_stride = x.plane_stride; d.x_pitch = x.pitch; d.x_coff = x.coff
    d.mean_rstd = _ptr(mean_rstd); d.spade_C = spade_C
    if heads is not None:
        bg = heads.get("bg")
        d.bg = _ptr(bg); d.bg_batch_stride = 0 if bg is None or bg.shape[0] == 1 else bg[0].numel()
        d.img = _ptr(heads.get("img")); d.mask = _ptr(heads.get("mask")); d.pred = _ptr(heads.get("pred"))
    d.max_ctas = max_ctas
    d.tiles_m = tiles_m
    d.stats_ws = _ptr(stats_ws)
    d.cta_pair = int(cta_pair)
    return d


def conv_gemm(a, wpack, mode, ksize, rows, block_n, epi, **kw):
    """tcgen05 implicit-GEMM convolution (conv_tc.cu). `wpack` = PackedW from the pack_* helpers below."""
    d = _fill_desc(a, mode, ksize, rows, block_n, epi, **kw)
    d.w = wpack.w.data_ptr(); d.w_planes = wpack.fmt; d.w_plane_stride = wpack.w[0].numel()
    if wpack.scale != 1.0:
        if wpack.scale_inv.device != wpack.w.device:
            wpack.scale_inv = wpack.scale_inv.to(wpack.w.device)
        d.w_scale_inv = wpack.scale_inv.data_ptr()
    if wpack.fmt == FMT_H8:
        d.w8 = wpack.w8.data_ptr(); d.wl8 = wpack.wl8.data_ptr(); d.cross_scale = wpack.cross_scale
    check(lib.iper_conv_gemm(d, _stream()), "conv_gemm")


def conv_direct(a, w_f32, mode, ksize, Cout, epi, **kw):
    """CUDA-core cross-check convolution (ops.cu) on the reference's own fp32 weight layout."""
    d = _fill_desc(a, mode, ksize, Cout, 64, epi, **kw)
    w_f32 = _req(w_f32, torch.float32, "w_f32")
    check(lib.iper_conv_direct(d, w_f32.data_ptr(), Cout, _stream()), "conv_direct")


def conv_stem(x_nchw, w_f32, bias, out, stats_ws=None):
    x_nchw = _req(x_nchw, torch.float32, "x"); w_f32 = _req(w_f32, torch.float32, "w")
    N, Cin, H, W = x_nchw.shape
    check(lib.iper_conv_stem(x_nchw.data_ptr(), N, Cin, H, W, w_f32.data_ptr(), _ptr(bias), w_f32.shape[0], out.ptr(),
                             out.P, out.plane_stride, out.pitch, out.coff, _ptr(stats_ws), _stream()), "conv_stem")


def conv_stem_tc(x_nchw, wpack, bias, out, stats_ws=None):
    """First encoder layer on the tensor cores, A operand built in shared memory (iper_conv_stem_tc); wpack = pack_stem_weight."""
    x_nchw = _req(x_nchw, torch.float32, "x")
    N, Cin, H, W = x_nchw.shape
    if wpack.scale != 1.0 and wpack.scale_inv.device != wpack.w.device:
        wpack.scale_inv = wpack.scale_inv.to(wpack.w.device)
    check(lib.iper_conv_stem_tc(x_nchw.data_ptr(), N, Cin, H, W, wpack.w.data_ptr(), wpack.fmt, wpack.w[0].numel(),
                                wpack.scale_inv.data_ptr() if wpack.scale != 1.0 else 0, _ptr(bias), out.ptr(), out.P,
                                out.plane_stride, out.pitch, out.coff, _ptr(stats_ws), _stream()), "conv_stem_tc")
    return out


def stem_im2col(x_nchw, out):
    """(N,Cin<=7,H,W) fp32 -> Planes (N,H/2,W/2,64): 3x3/s2/p1 patches, channel k = (ky*3+kx)*Cin + ci (iper_stem_im2col)."""
    x_nchw = _req(x_nchw, torch.float32, "x")
    N, Cin, H, W = x_nchw.shape
    check(lib.iper_stem_im2col(x_nchw.data_ptr(), N, Cin, H, W, out.ptr(), out.P, out.plane_stride, out.pitch, out.coff,
                               _stream()), "stem_im2col")
    return out


def pack_stem_weight(w, P):
    """Conv2d(Cin<=7 -> Cout, 3x3) weight -> (P, Cout, 64), K = (ky*3+kx)*Cin + ci zero-padded to 64 (matches stem_im2col)."""
    Cout, Cin = w.shape[:2]
    m = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).float()
    return PackedW(torch.cat([m, m.new_zeros(Cout, 64 - 9 * Cin)], 1), P)


def stats_workspace(N, C, device):
    return torch.empty((N, C, 2), dtype=torch.float64, device=device)


def instnorm_finalize(ws, HW, eps=1e-5):
    """fp64 (sum, sumsq) workspace filled by a producer's fused statistics -> mean_rstd (N,C,2) fp32."""
    N, C = ws.shape[:2]
    out = torch.empty((N, C, 2), dtype=torch.float32, device=ws.device)
    check(lib.iper_instnorm_finalize(ws.data_ptr(), N, C, HW, eps, out.data_ptr(), _stream()), "instnorm_finalize")
    return out


def instnorm_stats(x, eps=1e-5, out=None):
    if out is None:
        out = torch.empty((x.N, x.C, 2), dtype=torch.float32, device=x.data.device)
    ws = torch.empty((x.N, x.C, 2), dtype=torch.float64, device=x.data.device)
    check(lib.iper_instnorm_stats(x.ptr(), x.P, x.plane_stride, x.N, x.H * x.W, x.C, x.pitch, x.coff, eps,
                                  ws.data_ptr(), out.data_ptr(), _stream()), "instnorm_stats")
    return out


def instnorm_apply(x, stats, out, relu=False, res=None):
    """out = [res +] [relu](IN(x)) on Planes (bg_inpaintor.py ResidualBlock / conv-IN-ReLU blocks)."""
    r = res
    check(lib.iper_instnorm_apply(x.ptr(), x.P, x.plane_stride, x.pitch, x.coff, stats.data_ptr(), x.N, x.H * x.W, x.C,
                                  int(relu), 0 if r is None else r.ptr(), 0 if r is None else r.P,
                                  0 if r is None else r.plane_stride, 0 if r is None else r.pitch,
                                  0 if r is None else r.coff, out.ptr(), out.P, out.plane_stride, out.pitch, out.coff,
                                  _stream()), "instnorm_apply")
    return out


def tanh_nhwc_to_nchw(x, C):
    N, H, W, pitch = x.shape
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=x.device)
    check(lib.iper_tanh_nhwc_to_nchw(x.data_ptr(), N, H * W, C, pitch, out.data_ptr(), _stream()), "tanh_nhwc_to_nchw")
    return out


def warp_attention(xt, kv, bias_v, T, out):
    """xt Planes (B,h,w,C) target features; kv (ns,h,w,2C+64) f32 source maps [(Wq^T Wk)x | Wv x | (Wk^T bq).x | pad];
    T (B,ns,h,w,2) f32 -> out Planes (B,h,w,C) (include/iper_b200.h: iper_warp_attention)."""
    B, h, w, C = xt.N, xt.H, xt.W, xt.C
    ns = kv.shape[0]
    assert kv.shape[-1] == 2 * C + 64 and kv.dtype == torch.float32 and kv.is_contiguous()
    if tuple(T.shape) != (B, ns, h, w, 2) or tuple(kv.shape[1:3]) != (h, w):
        raise ValueError("warp_attention: flow %s does not match (B=%d, ns=%d, h=%d, w=%d, 2) — the ns source maps are "
                         "shared by all B frames" % (tuple(T.shape), B, ns, h, w))
    check(lib.iper_warp_attention(xt.ptr(), xt.P, xt.plane_stride, xt.pitch, xt.coff, kv.data_ptr(), bias_v.data_ptr(),
                                  T.data_ptr(), B, ns, h, w, C, out.ptr(), out.P, out.plane_stride, out.pitch, out.coff,
                                  _stream()), "warp_attention")


def attention_source_weight(wq, bq, wk, wv):
    """Conv2d-shaped (2C+64, C, 1, 1) weight of the source-side projection [Wq^T Wk | Wv | Wk^T bq | 0...] (see
    iper_warp_attention); wq/wk/wv (C,C,1,1), bq (C,)."""
    C = wq.shape[0]
    Wq, Wk = wq.reshape(C, C).double(), wk.reshape(C, C).double()
    rows = torch.cat([Wq.t() @ Wk, wv.reshape(C, C).double(), (Wk.t() @ bq.double())[None],
                      torch.zeros(63, C, dtype=torch.float64, device=wq.device)], 0)
    return rows.float().reshape(2 * C + 64, C, 1, 1)


def warp_nhwc(src, T):
    ns, h, w, C = src.shape; B = T.shape[0]
    out = torch.empty((B, ns, h, w, C), dtype=torch.float32, device=src.device)
    check(lib.iper_warp_nhwc(src.data_ptr(), T.data_ptr(), B, ns, h, w, C, out.data_ptr(), _stream()), "warp_nhwc")
    return out


def nhwc_f32_to_nchw(x):
    N, H, W, C = x.shape
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=x.device)
    check(lib.iper_nhwc_f32_to_nchw(x.data_ptr(), N, C, H * W, C, 0, out.data_ptr(), _stream()), "nhwc_f32_to_nchw")
    return out


def pred_to_u8(pred, out=None):
    B, _, S, _ = pred.shape
    if out is None:
        out = torch.empty((B, S, S, 3), dtype=torch.uint8, device=pred.device)
    check(lib.iper_pred_to_u8(pred.data_ptr(), B, S, out.data_ptr(), _stream()), "pred_to_u8")
    return out


MORPH_ERODE, MORPH_DILATE, MORPH_SOFT_DILATE = 0, 1, 2


def morph(mask, ks, mode):
    """morph_ops.morph / soft_dilate (morph_ops.py:7-61) on a (N,1,H,W) fp32 mask; mode MORPH_ERODE / _DILATE / _SOFT_DILATE."""
    mask = _req(mask.contiguous(), torch.float32, "mask")
    if mask.dim() != 4 or mask.shape[1] != 1:
        raise ValueError("morph expects a (N,1,H,W) mask, got %s" % (tuple(mask.shape),))
    N, _, H, W = mask.shape
    out = torch.empty_like(mask)
    check(lib.iper_morph(mask.data_ptr(), N, H, W, int(ks), int(mode), out.data_ptr(), _stream()), "morph")
    return out


# ------------------------------------------------------------------------------------------------------------------
# weight packing (one-time, at load): reference layouts -> K-major fp16 planes for the tcgen05 kernel
# ------------------------------------------------------------------------------------------------------------------
ACT_S8, ACT_SL8 = 8.0, 16384.0       # activation e4m3 scales (common.cuh)


def split_planes(w, P):
    """fp32 -> (P, ...) fp16: plane0 = fp16(w), plane1 = fp16(w - plane0).  (P=3 keeps only the fp16 plane here.)"""
    hi = w.to(torch.float16)
    if P != 2:
        return hi[None].contiguous()
    lo = (w - hi.float()).to(torch.float16)
    return torch.stack([hi, lo], 0).contiguous()


def _to_e4m3(t):
    return t.clamp(-448.0, 448.0).to(torch.float8_e4m3fn)


def e4m3_roundtrip(t, scale):
    """value the kernels see for an e4m3 plane: e4m3(t*scale)/scale (test helper)."""
    return _to_e4m3(t.float() * scale).float() / scale


class PackedW:
    """K-major packed weight matrix (rows_total, K) in planes format `fmt` (1, 2 or 3)."""

    def __init__(self, m, fmt):
        m = m.float().contiguous()
        self.fmt, self.rows_total, self.K = fmt, m.shape[0], m.shape[1]
        # power-of-two pre-scale (formats 1/2): max|w| * s in [128, 256) keeps hi AND lo planes of small weights out of fp16's
        # subnormal range; the conv epilogue multiplies the accumulator by 1/s (exact).  See iper_conv_gemm_desc.w_scale_inv
        self.scale = 1.0
        mx = float(m.abs().max()) if m.numel() else 0.0
        if fmt in (FMT_H, FMT_HL) and mx > 0 and math.isfinite(mx):
            self.scale = 2.0 ** math.floor(math.log2(256.0 / mx))
        self.scale_inv = torch.tensor([1.0 / self.scale], dtype=torch.float32, device=m.device)
        m = m * self.scale
        self.w = split_planes(m, fmt)
        self.w8 = self.wl8 = None
        self.cross_scale = 0.0
        self.sW = 1.0
        if fmt == FMT_H8:
            hi = self.w[0].float()
            mx = float(hi.abs().max())
            self.sW = 2.0 ** math.floor(math.log2(240.0 / mx)) if mx > 0 else 1.0
            self.w8 = _to_e4m3(hi * self.sW).view(torch.uint8).contiguous()
            self.wl8 = _to_e4m3((m - hi) * (self.sW * 2048.0)).view(torch.uint8).contiguous()
            self.cross_scale = 1.0 / (ACT_SL8 * self.sW)

    def to(self, device):
        self.w = self.w.to(device)
        self.scale_inv = self.scale_inv.to(device)
        if self.w8 is not None:
            self.w8, self.wl8 = self.w8.to(device), self.wl8.to(device)
        return self

    def effective(self):
        """(w_main, w_for_lo_term, wlo_term) fp32 matrices the three MMA groups multiply with (test helper)."""
        hi = self.w[0].float().cpu() / self.scale            # de-scaled: the values the layer effectively multiplies with
        if self.fmt == 1:
            return hi, None, None
        if self.fmt == 2:
            return hi, hi, self.w[1].float().cpu() / self.scale
        return hi, self.w8.cpu().view(torch.float8_e4m3fn).float() / self.sW, \
            self.wl8.cpu().view(torch.float8_e4m3fn).float() / (self.sW * 2048.0)


def pack_conv_weight(w, P, pad_rows_to=None):
    """Conv2d weight (Cout,Cin,k,k) -> (P, rows, k*k*Cin) with K ordered (tap=ky*k+kx, cin)."""
    Cout, Cin, k, _ = w.shape
    m = w.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin).float()
    if pad_rows_to is not None and pad_rows_to > Cout:
        m = torch.cat([m, m.new_zeros(pad_rows_to - Cout, m.shape[1])], 0)
    return PackedW(m, P)


def pack_heads_weight(w_img, w_mask, P):
    """tsf_img_reg (3,64,5,5) + tsf_att_reg (1,64,5,5) -> (P, 32, 5*64) for IPER_CONV_ROW5:
    row n = dx*4 + o (o = r,g,b,mask; rows 20..31 zero), K = (dy, cin)."""
    w = torch.cat([w_img, w_mask], 0).float()                  # (4, C, 5, 5) [o, c, dy, dx]
    C = w.shape[1]
    m = w.permute(3, 0, 2, 1).reshape(5 * 4, 5 * C)            # (dx, o) x (dy, c)
    m = torch.cat([m, m.new_zeros(32 - 20, 5 * C)], 0)
    return PackedW(m, P)


def pack_convT_weight(w, P):
    """ConvTranspose2d(4,2,1) weight (Cin,Cout,4,4) -> (P, 4*Cout, 4*Cin): phase p=py*2+px, tap (ta,tb), K=(ta*2+tb, cin).
    ky = {py=0: (1,3), py=1: (0,2)}[ta], same for kx — matches the producer's tap decode in conv_tc.cu."""
    Cin, Cout = w.shape[:2]
    kidx = {0: (1, 3), 1: (0, 2)}
    phases = []
    for py in range(2):
        for px in range(2):
            taps = []
            for ta in range(2):
                for tb in range(2):
                    taps.append(w[:, :, kidx[py][ta], kidx[px][tb]].t())     # (Cout, Cin)
            phases.append(torch.cat(taps, 1))                               # (Cout, 4*Cin)
    return PackedW(torch.cat(phases, 0), P)                                  # (4*Cout, 4*Cin)


def pack_spade_weight(wg, bg_, wb, bb, P, block_n):
    """mlp_gamma / mlp_beta (C,128,3,3) -> rows interleaved per tile: [gamma cb | beta cb] with cb = block_n/2."""
    C = wg.shape[0]; cb = block_n // 2
    mg = wg.permute(0, 2, 3, 1).reshape(C, -1).float(); mb = wb.permute(0, 2, 3, 1).reshape(C, -1).float()
    rows, bias = [], []
    for t in range(C // cb):
        rows += [mg[t * cb:(t + 1) * cb], mb[t * cb:(t + 1) * cb]]
        bias += [bg_[t * cb:(t + 1) * cb], bb[t * cb:(t + 1) * cb]]
    return PackedW(torch.cat(rows, 0), P), torch.cat(bias, 0).float().contiguous()
