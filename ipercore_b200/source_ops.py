"""One-time-per-source stage of ``Imitator.source_setup`` on the B200 kernels (csrc/source.cu).

Mirrors the source-side methods of ``FlowComposition`` (iPERCore/models/flowcomposition.py): ``make_morph_image`` (:335-386,
with ``cal_top_k_ids`` :264-293 and ``morph_image`` :295-333 folded into one kernel), ``make_uv_img`` (:87-137) and the
``CannyFilter`` they use (iPERCore/tools/utils/morphology/canny_ops.py) — same argument meaning, same outputs.
``ipercore_b200.patch.install()`` swaps them into the reference class; nothing here touches the per-frame path.
"""
import ctypes

import numpy as np
import torch

from . import ops
from ._lib import check, lib
from .ops import _req, _stream


def canny_constants(k_gaussian=3, mu=0, sigma=1):
    """Filter taps of CannyFilter.__init__ (canny_ops.py:9-127): normalised 3x3 gaussian on a [-1,1] grid, the x/(x^2+y^2)
    sobel, eight 3x3 directional kernels (+1 centre, -1 at the neighbour 45*k degrees counter-clockwise from +x; the reference
    builds them by rotating a 5x5 kernel with cv2 — the result is this constant set) and the 1.25 hysteresis taps."""
    g1 = np.linspace(-1, 1, k_gaussian)
    x, y = np.meshgrid(g1, g1)
    d = (x ** 2 + y ** 2) ** 0.5
    g = np.exp(-(d - mu) ** 2 / (2 * sigma ** 2)) / (2 * np.pi * sigma ** 2)
    g = g / np.sum(g)
    r = np.linspace(-1, 1, 3)
    sx, sy = np.meshgrid(r, r)
    den = sx ** 2 + sy ** 2
    den[:, 1] = 1
    sobel = sx / den
    nb = [(0, 1), (-1, 1), (-1, 0), (-1, -1), (0, -1), (1, -1), (1, 0), (1, 1)]        # (dy, dx) of the -1 tap
    thin = np.zeros((8, 3, 3))
    for k, (dy, dx) in enumerate(nb):
        thin[k, 1, 1] = 1
        thin[k, 1 + dy, 1 + dx] = -1
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)       # the reference stores them in fp32 conv weights
    return dict(gaussian=f32(g), sobel_x=f32(sobel), directional=f32(thin), hysteresis=np.float32(1.25))


_CANNY = None


def canny_edges(sil, low=0.1, high=0.9):
    """thin_edges of CannyFilter()(sil, low, high, True) for a (N,1,H,W) map -> (N,1,H,W) in {0,1}."""
    global _CANNY
    if _CANNY is None:
        _CANNY = canny_constants()
    sil = _req(sil.float().contiguous(), torch.float32, "sil")
    if sil.dim() != 4 or sil.shape[1] != 1:
        raise ValueError("canny_edges expects a (N,1,H,W) map, got %s" % (tuple(sil.shape),))
    N, _, H, W = sil.shape
    dev = sil.device
    mag = torch.empty((N, H, W), dtype=torch.float32, device=dev)
    tri = torch.empty_like(mag)
    ori = torch.empty((N, H, W), dtype=torch.int8, device=dev)
    edges = torch.empty((N, 1, H, W), dtype=torch.float32, device=dev)
    c = _CANNY
    hp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    check(lib.iper_canny_edges(sil.data_ptr(), N, H, W, hp(c["gaussian"]), hp(c["sobel_x"]), hp(c["directional"]),
                               float(c["hysteresis"]), float(low), float(high), mag.data_ptr(), ori.data_ptr(), tri.data_ptr(),
                               edges.data_ptr(), _stream()), "canny_edges")
    return edges


def morph_image(src_img, confidant_sil, outpad_sil, edges):
    """flowcomposition.py:264-333 for a batch: src_img (N,3,H,W), sils / edges (N,1,H,W) -> morphed image (N,3,H,W)."""
    src_img = _req(src_img.float().contiguous(), torch.float32, "src_img")
    N, _, H, W = src_img.shape
    conf = _req(confidant_sil.float().contiguous(), torch.float32, "confidant_sil")
    outp = _req(outpad_sil.float().contiguous(), torch.float32, "outpad_sil")
    edges = _req(edges.float().contiguous(), torch.float32, "edges")
    for t in (conf, outp, edges):
        if t.numel() != N * H * W:
            raise ValueError("morph_image: mask of %s does not match images %s" % (tuple(t.shape), tuple(src_img.shape)))
    cnt = torch.empty((N,), dtype=torch.int32, device=src_img.device)
    lst = torch.empty((N, H * W), dtype=torch.int32, device=src_img.device)
    out = torch.empty_like(src_img)
    check(lib.iper_morph_image(src_img.data_ptr(), conf.data_ptr(), outp.data_ptr(), edges.data_ptr(), N, H, W, cnt.data_ptr(),
                               lst.data_ptr(), out.data_ptr(), _stream()), "morph_image")
    return out


def make_morph_image(src_img, confidant_sil, outpad_sil, erode_ks=3, dilate_ks=11):
    """FlowComposition.make_morph_image (flowcomposition.py:335-386) on explicit silhouettes."""
    if erode_ks > 0:
        confidant_sil = ops.morph(confidant_sil.float().contiguous(), erode_ks, ops.MORPH_ERODE)
    if dilate_ks > 0:
        outpad_sil = ops.morph(outpad_sil.float().contiguous(), dilate_ks, ops.MORPH_DILATE)
    return morph_image(src_img, confidant_sil, outpad_sil, canny_edges(confidant_sil, 0.1, 0.9))


def make_uv_img(src_img, obj_f2pts, only_vis_obj_f2pts, uv_fim, uv_wim, dilate_ks=13):
    """FlowComposition.make_uv_img (flowcomposition.py:87-137): src_img (bs,ns,3,h,w), corner sets (bs*ns,nf,3,2),
    uv_fim (.,h,w) / uv_wim (.,h,w,3) (the constant UV-layout maps; only the first item is read) -> (bs,3,h,w)."""
    bs, ns, _, h, w = src_img.shape
    N = bs * ns
    src = _req(src_img.reshape(N, 3, h, w).float().contiguous(), torch.float32, "src_img")
    a = _req(obj_f2pts.float().contiguous(), torch.float32, "obj_f2pts")
    b = _req(only_vis_obj_f2pts.float().contiguous(), torch.float32, "only_vis_obj_f2pts")
    fim = _req(uv_fim.reshape(-1, h, w)[0].int().contiguous(), torch.int32, "uv_fim")
    wim = _req(uv_wim.reshape(-1, h, w, 3)[0].float().contiguous(), torch.float32, "uv_wim")
    nf = a.shape[1]
    if a.shape[0] != N or tuple(b.shape) != tuple(a.shape) or tuple(a.shape[2:]) != (3, 2):
        raise ValueError("make_uv_img: corner sets %s / %s do not match %d source images" % (tuple(a.shape), tuple(b.shape), N))
    dev = src.device
    src_warp = torch.empty((N, 3, h, w), dtype=torch.float32, device=dev)
    vis = torch.empty((N, 1, h, w), dtype=torch.float32, device=dev)
    check(lib.iper_uv_warp(src.data_ptr(), a.data_ptr(), b.data_ptr(), fim.data_ptr(), wim.data_ptr(), N, nf, h, w,
                           src_warp.data_ptr(), vis.data_ptr(), _stream()), "uv_warp")
    vis = ops.morph(vis, dilate_ks, ops.MORPH_DILATE)
    uv = torch.empty((bs, 3, h, w), dtype=torch.float32, device=dev)
    check(lib.iper_uv_merge(src_warp.data_ptr(), vis.data_ptr(), bs, ns, h, w, uv.data_ptr(), _stream()), "uv_merge")
    return uv
