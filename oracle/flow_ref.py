"""CPU restatement of the per-frame geometry path (projection -> faces -> fim/wim -> f2pts, cond, flow, syn image).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  numpy for the integer/gather arithmetic, torch (CPU)
only for ``grid_sample`` / ``interpolate``, the same library calls the reference makes.
Pinned against the reference's own code by tests/golden/make_golden.py (flow_*.npz fixtures).
"""
import numpy as np

from . import raster

# nmr.py:225  self.eye = [0, 0, -(1. / np.tan(np.radians(self.viewing_angle)) + 1)], viewing_angle=30,
# converted to float32 inside nr.look_at.
EYE_Z = np.float32(-(1.0 / np.tan(np.radians(30.0)) + 1.0))


def project(cam, verts):
    """nmr.py:34-52 orthographic_proj_withz_idrot, then the y flip (nmr.py:331) and nr.look_at (nmr.py:333).

    With eye=(0,0,EYE_Z), at=(0,0,0), up=(0,1,0) the look_at rotation is exactly the identity, so look_at is
    the translation z -> z - EYE_Z.  cam (N,3) [s,tx,ty]; verts (N,V,3) -> (N,V,3) float32.
    """
    cam = np.asarray(cam, np.float32)
    verts = np.asarray(verts, np.float32)
    s = cam[:, 0].reshape(-1, 1, 1)
    t = cam[:, 1:3].reshape(-1, 1, 2)
    xy = s * (verts[:, :, :2] + t)          # scale * (X + trans), float32
    xy = xy.copy()
    xy[:, :, 1] *= np.float32(-1.0)          # proj_verts[:, :, 1] *= -1
    z = verts[:, :, 2:3] - EYE_Z             # vertices - eye
    return np.concatenate([xy, z], axis=2).astype(np.float32)


def vertices_to_faces(verts, faces):
    """nr.vertices_to_faces: (N,V,3),(F,3) -> (N,F,3,3)."""
    return np.ascontiguousarray(verts[:, np.asarray(faces, np.int64)])


def render_fim_wim(cam, verts, faces, image_size):
    """SMPLRenderer.render_fim_wim (nmr.py:319-342): returns f2pts (N,F,3,2), fim (N,S,S) i32, wim (N,S,S,3)."""
    fv = vertices_to_faces(project(cam, verts), faces)
    fim, wim = raster.rasterize_fim_wim(fv, image_size)
    f2pts = fv[:, :, :, 0:2].copy()
    f2pts[:, :, :, 1] *= np.float32(-1.0)    # nmr.py:340 flips y back
    return f2pts, fim, wim


def encode_fim(fim, map_fn):
    """SMPLRenderer.encode_fim (nmr.py:390-401): cond = map_fn[fim] (index -1 -> last row), NHWC -> NCHW."""
    enc = np.asarray(map_fn, np.float32)[np.asarray(fim, np.int64)]
    return np.ascontiguousarray(enc.transpose(0, 3, 1, 2))


def cal_bc_transform(src_f2pts, dst_fims, dst_wims):
    """SMPLRenderer.cal_bc_transform (nmr.py:713-757): T[p] = sum_k wim[p,k] * src_f2pts[fim[p],k,:], bg = -2.

    The reference multiplies then ``.sum(dim=1)`` over the 3 corners; a 3-term float32 sum in index order.
    """
    src_f2pts = np.asarray(src_f2pts, np.float32)
    bs, S = dst_fims.shape[0], dst_fims.shape[1]
    T = np.full((bs, S * S, 2), -2.0, np.float32)
    for i in range(bs):
        fim = np.asarray(dst_fims[i], np.int64).reshape(-1)
        w = np.asarray(dst_wims[i], np.float32).reshape(-1, 3)
        m = fim != -1
        prod = src_f2pts[i][fim[m]] * w[m][:, :, None]          # (n,3,2)
        T[i, m] = (prod[:, 0] + prod[:, 1]) + prod[:, 2]
    return T.reshape(bs, S, S, 2)


def grid_sample(img, grid):
    """F.grid_sample(img, grid) with torch defaults (bilinear, zeros, align_corners=False)."""
    import torch
    import torch.nn.functional as F
    out = F.grid_sample(torch.from_numpy(np.ascontiguousarray(img)), torch.from_numpy(np.ascontiguousarray(grid)),
                        mode="bilinear", padding_mode="zeros", align_corners=False)
    return out.numpy()


def make_tsf_inputs(uv_img, f_uvs2img, fim, wim, cond):
    """FlowComposition.make_tsf_inputs (flowcomposition.py:206-248) for bs=1 source, nt frames.

    uv_img (1,3,S,S); f_uvs2img (F,3,2); fim/wim/cond for nt frames -> (nt,6,S,S)."""
    nt = fim.shape[0]
    f2uvs = np.repeat(np.asarray(f_uvs2img, np.float32)[None], nt, axis=0)
    Tuv2t = cal_bc_transform(f2uvs, fim, wim)
    syn = grid_sample(np.repeat(uv_img, nt, axis=0), Tuv2t)
    return np.concatenate([syn, cond], axis=1), Tuv2t


def make_trans_flow(src_f2pts, fim, wim):
    """FlowComposition.make_trans_flow (flowcomposition.py:514-582), temporal=False, one target frame:
    the frame's fim/wim are repeated ns times and combined with each source's f2pts -> Tst (ns,S,S,2)."""
    ns = src_f2pts.shape[0]
    return cal_bc_transform(src_f2pts, np.repeat(fim[None], ns, 0), np.repeat(wim[None], ns, 0))


def frame_inputs(cam, verts, faces, map_fn, f_uvs2img, uv_img, src_f2pts, image_size):
    """Rows a1-a8 of SURVEY.md §8 for a batch of target frames (each frame independent).

    Returns dict(fim, wim, f2pts, cond, tsf_inputs (N,6,S,S), Tst (N,ns,S,S,2))."""
    f2pts, fim, wim = render_fim_wim(cam, verts, faces, image_size)
    cond = encode_fim(fim, map_fn)
    tsf_inputs, Tuv2t = make_tsf_inputs(uv_img, f_uvs2img, fim, wim, cond)
    Tst = np.stack([make_trans_flow(src_f2pts, fim[i], wim[i]) for i in range(fim.shape[0])], 0)
    return dict(fim=fim, wim=wim, f2pts=f2pts, cond=cond, tsf_inputs=tsf_inputs, Tuv2t=Tuv2t, Tst=Tst)
