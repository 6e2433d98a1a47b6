"""CPU restatement of the one-time-per-source stage (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Follows /root/reference/iPERCore/tools/utils/morphology/canny_ops.py:129-192 (CannyFilter.forward, C = 1, hysteresis on),
iPERCore/models/flowcomposition.py:264-386 (cal_top_k_ids / morph_image / make_morph_image) and :87-137 (make_uv_img).
Pinned against outputs of the reference's own functions captured during a real Imitator.source_setup run
(tests/golden/source_S96.npz, make_golden.py `source`).
"""
import numpy as np


def _corr3(x, k):
    """3x3 cross-correlation with zero padding, row-major tap order, float32 multiply-add chain."""
    H, W = x.shape[-2:]
    p = np.zeros(x.shape[:-2] + (H + 2, W + 2), np.float32)
    p[..., 1:-1, 1:-1] = x
    out = np.zeros_like(x, dtype=np.float32)
    for dy in range(3):
        for dx in range(3):
            out = (out + np.float32(k[dy, dx]) * p[..., dy:dy + H, dx:dx + W]).astype(np.float32)
    return out


def canny_edges(img, consts, low=0.1, high=0.9):
    """img (N,H,W) float32 -> thin_edges (N,H,W) in {0,1} (canny_ops.py:129-192)."""
    img = np.asarray(img, np.float32)
    blurred = _corr3(img, consts["gaussian"])
    gx = _corr3(blurred, consts["sobel_x"]); gy = _corr3(blurred, consts["sobel_x"].T)
    mag = np.sqrt((gx * gx + gy * gy).astype(np.float32)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        ori = (np.arctan((gy / gx).astype(np.float32)).astype(np.float32) * np.float32(360.0 / np.pi) + np.float32(180.0)).astype(np.float32)
        ori = (np.round(ori / np.float32(45.0)) * np.float32(45.0)).astype(np.float32)
        pidx = (ori / np.float32(45.0)) % 8
    directional = np.stack([_corr3(mag, consts["directional"][k]) for k in range(8)], 0)
    thin = mag.copy()
    for pos in range(4):
        oriented = (pidx == pos) | (pidx == pos + 4)
        is_max = np.minimum(directional[pos], directional[pos + 4]) > 0.0
        thin[(~is_max) & oriented] = 0.0
    lo = thin > low; hi = thin > high
    tri = (lo * np.float32(0.5) + hi * np.float32(0.5)).astype(np.float32)
    weak = tri == 0.5
    hyst = _corr3(tri, np.full((3, 3), consts["hysteresis"], np.float32)) > 1
    return (hi | (weak & hyst)).astype(np.float32)


def morph_image(src_img, confidant_sil, outpad_sil, edges):
    """(3,H,W), (H,W), (H,W), (H,W) -> (3,H,W): flowcomposition.py:264-333 with top_k = 3; ties -> lowest row-major index."""
    H, W = edges.shape
    out = (src_img * confidant_sil[None]).astype(np.float32)
    b = np.argwhere(edges != 0); u = np.argwhere((outpad_sil * (1 - confidant_sil)) != 0)
    if len(b) < 3 or len(u) == 0:
        return out
    d = ((u[:, None, :] - b[None, :, :]) ** 2).sum(-1).astype(np.int64)
    key = d * (H * W) + (b[:, 0] * W + b[:, 1])[None, :]
    ids = np.argsort(key, axis=1)[:, :3]
    val = np.take_along_axis(d, ids, 1).astype(np.float32)
    w = val / val.sum(1, keepdims=True)
    nn = b[ids]                                                 # (n1, 3, 2)
    rgb = src_img[:, nn[..., 0], nn[..., 1]]                    # (3, n1, 3)
    out[:, u[:, 0], u[:, 1]] = (rgb * w[None]).sum(-1)
    return out


def _grid_sample(img, gx, gy):
    """bilinear, zeros padding, align_corners=False; img (C,H,W), gx/gy (H,W) -> (C,H,W)."""
    C, H, W = img.shape
    ix = ((gx + 1) * W - 1) * np.float32(0.5); iy = ((gy + 1) * H - 1) * np.float32(0.5)
    x0 = np.floor(ix); y0 = np.floor(iy)
    out = np.zeros((C,) + gx.shape, np.float32)
    for xx, yy, w in ((x0, y0, (x0 + 1 - ix) * (y0 + 1 - iy)), (x0 + 1, y0, (ix - x0) * (y0 + 1 - iy)),
                      (x0, y0 + 1, (x0 + 1 - ix) * (iy - y0)), (x0 + 1, y0 + 1, (ix - x0) * (iy - y0))):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        xi = np.clip(xx, 0, W - 1).astype(np.int64); yi = np.clip(yy, 0, H - 1).astype(np.int64)
        out += img[:, yi, xi] * (w * ok).astype(np.float32)[None]
    return out


def make_uv_img(src_img, obj_f2pts, only_vis_obj_f2pts, uv_fim, uv_wim, dilate_ks=13):
    """src_img (ns,3,h,w), corner sets (ns,nf,3,2), uv_fim (h,w), uv_wim (h,w,3) -> (3,h,w) (flowcomposition.py:87-137, bs=1)."""
    from . import flow_ref, morph_ref
    ns, _, h, w = src_img.shape
    fim = np.repeat(uv_fim[None], ns, 0); wim = np.repeat(uv_wim[None], ns, 0)
    T = flow_ref.cal_bc_transform(obj_f2pts, fim, wim); Tv = flow_ref.cal_bc_transform(only_vis_obj_f2pts, fim, wim)
    ones = np.ones((1, h, w), np.float32)
    src_warp = np.stack([_grid_sample(src_img[i], T[i, ..., 0], T[i, ..., 1]) for i in range(ns)])
    vis = np.stack([_grid_sample(ones, Tv[i, ..., 0], Tv[i, ..., 1]) for i in range(ns)])       # (ns,1,h,w)
    vis = morph_ref.morph(vis, dilate_ks, 1)
    vis_sum = vis[1:].sum(0)
    temp = (src_warp[1:] * vis[1:]).sum(0) / (vis_sum + np.float32(1e-5))
    front_invisible = (1 - vis[0]) * (vis_sum >= 1)
    return (src_warp[0] * (1 - front_invisible) + temp * front_invisible).astype(np.float32)
