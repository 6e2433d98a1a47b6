"""Seeded synthetic workload for the motion-imitation hot path (TEST INFRASTRUCTURE / bench input generator data).

Follows SURVEY.md §8d: a 6 890-vertex SMPL template (from tests/golden/smpl_template.npz, extracted from the
reference's assets/configs/pose3d/mapper_uv.txt by tests/golden/make_golden.py) is posed rigidly per frame;
sources are the template rotated +-20 deg about y.  Everything is a pure function of (seed, sizes) using numpy's
PCG64 stream, so the GPU box regenerates bit-identical inputs without /root/reference.

This module only *generates inputs*; it is imported by bench.py for workload generation and by tests/.
"""
import os

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_template():
    z = np.load(os.path.join(_GOLDEN, "smpl_template.npz"))
    return {k: z[k] for k in z.files}


def _rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float64)


def _rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], np.float64)


def base_verts(tpl):
    """Template in the HMR/SPIN convention the reference feeds the renderer (image-aligned, +y down)."""
    v = tpl["verts"].astype(np.float64)
    return v @ _rot_x(np.pi).T


def pose_sweep(tpl, n_frames, total=300, start=0):
    """V_t = R_y(theta_t) R_x(phi_t) V + 0.01 sin(2 pi t/300 + x); cams s,tx,ty."""
    V = base_verts(tpl)
    cams = np.zeros((n_frames, 3), np.float32)
    verts = np.zeros((n_frames, V.shape[0], 3), np.float32)
    for i in range(n_frames):
        t = start + i
        th = 2 * np.pi * t / total
        ph = np.radians(10.0) * np.sin(2 * np.pi * t / 75.0)
        R = _rot_y(th) @ _rot_x(ph)
        Vt = V @ R.T + 0.01 * np.sin(2 * np.pi * t / total + V[:, 0:1])
        verts[i] = Vt.astype(np.float32)
        cams[i] = [0.9 + 0.1 * np.sin(2 * np.pi * t / 150.0), 0.1 * np.sin(2 * np.pi * t / 97.0),
                   -0.28 + 0.1 * np.cos(2 * np.pi * t / 61.0)]
    return cams, verts


def source_views(tpl, ns=2):
    """ns source views: template rotated by +-20 deg (then +-40, ...) about y; cam fixed."""
    V = base_verts(tpl)
    cams = np.tile(np.array([[0.95, 0.0, -0.28]], np.float32), (ns, 1))
    verts = np.zeros((ns, V.shape[0], 3), np.float32)
    for i in range(ns):
        a = np.radians(20.0 * (i // 2 + 1)) * (1 if i % 2 == 0 else -1)
        verts[i] = (V @ _rot_y(a).T).astype(np.float32)
    return cams, verts


def uniform(shape, seed, lo=-1.0, hi=1.0):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.uniform(lo, hi, size=shape).astype(np.float32)


def smooth_image(shape, seed):
    """Band-limited image-like tensor in [-1,1]: low-res noise upsampled bilinearly + a little fine noise."""
    *lead, H, W = shape
    rng = np.random.Generator(np.random.PCG64(seed))
    g = max(H // 16, 2)
    low = rng.uniform(-1, 1, size=(*lead, g, g))
    ys = np.linspace(0, g - 1, H); xs = np.linspace(0, g - 1, W)
    y0 = np.floor(ys).astype(int).clip(0, g - 2); x0 = np.floor(xs).astype(int).clip(0, g - 2)
    fy = (ys - y0)[:, None]; fx = (xs - x0)[None, :]
    a = low[..., y0][..., :, x0] if False else low[..., y0, :][..., :, x0]
    b = low[..., y0, :][..., :, x0 + 1]
    c = low[..., y0 + 1, :][..., :, x0]
    d = low[..., y0 + 1, :][..., :, x0 + 1]
    img = (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy
    img = 0.9 * img + 0.1 * rng.uniform(-1, 1, size=img.shape)
    return img.astype(np.float32)


def morph_masks(seed=5):
    """Binary (3,1,96,80) test masks for the mask-morphology row: a blob, scattered speckles, a mask touching the border."""
    rng = np.random.Generator(np.random.PCG64(seed))
    m = np.zeros((3, 1, 96, 80), dtype=np.float32)
    yy, xx = np.mgrid[0:96, 0:80]
    m[0, 0] = ((yy - 50) ** 2 / 900.0 + (xx - 38) ** 2 / 400.0 < 1.0)
    m[1, 0] = rng.random((96, 80)) < 0.03
    m[2, 0, :40, :] = 1.0
    m[2, 0, 40:, 60:] = rng.random((56, 20)) < 0.5
    return m
