"""Seam B1 — drop-in for the subset of the third-party `neural_renderer` package that iPERCore's hot path calls
(iPERCore/tools/human_digitalizer/renders/nmr.py:8 `import neural_renderer as nr`; call sites nmr.py:333-337,356):

    nr.look_at(vertices, eye)                                  pure tensor glue (torch)
    nr.vertices_to_faces(vertices, faces)                      pure tensor glue (torch)
    nr.rasterize_face_index_map_and_weight_map(faces, S, False)  -> libiper_b200 tiled rasteriser (raster.cu)

`ipercore_b200.patch.install()` registers this module as `sys.modules["neural_renderer"]`.  The visualisation-only
entry points of the upstream package (rasterize, lighting, rasterize_silhouettes, rasterize_depth — nmr.py:271-387,
not on the hot path) raise NotImplementedError.
"""
import torch
import torch.nn.functional as F

from . import ops

DEFAULT_NEAR, DEFAULT_FAR = 0.1, 100.0


def look_at(vertices, eye, at=(0, 0, 0), up=(0, 1, 0)):
    """Rotate/translate vertices into the camera frame looking from `eye` at `at` (upstream nr.look_at)."""
    dev = vertices.device
    as_t = lambda v: (v if torch.is_tensor(v) else torch.tensor(v, dtype=torch.float32)).to(dev).float()
    eye, at, up = as_t(eye), as_t(at), as_t(up)
    bs = vertices.shape[0]
    eye, at, up = [v[None].repeat(bs, 1) if v.dim() == 1 else v for v in (eye, at, up)]
    z_axis = F.normalize(at - eye, eps=1e-5)
    x_axis = F.normalize(torch.cross(up, z_axis, dim=1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=1), eps=1e-5)
    r = torch.stack((x_axis, y_axis, z_axis), dim=1)
    return torch.matmul(vertices - eye[:, None, :], r.transpose(1, 2))


def vertices_to_faces(vertices, faces):
    """(bs,nv,3), (bs,nf,3) int -> (bs,nf,3,3)."""
    bs, nv = vertices.shape[:2]
    faces = faces.long() + (torch.arange(bs, device=vertices.device) * nv)[:, None, None]
    return vertices.reshape(bs * nv, 3)[faces]


def rasterize_face_index_map_and_weight_map(faces, image_size=256, anti_aliasing=False, near=DEFAULT_NEAR,
                                            far=DEFAULT_FAR, **_ignored):
    """faces (N,F,3,3) f32 CUDA -> (fim (N,S,S) int32, wim (N,S,S,3) float32); current stream, no host sync."""
    if anti_aliasing:
        raise NotImplementedError("anti_aliasing=True is never used for the index/weight maps (nmr.py:337,356)")
    if not faces.is_cuda:
        raise RuntimeError("ipercore_b200 rasteriser needs CUDA tensors (there is no CPU path)")
    return ops.rasterize_faces(faces.float().contiguous(), int(image_size), float(near), float(far))


def _not_on_hot_path(name):
    def fn(*a, **k):
        raise NotImplementedError("neural_renderer.%s is visualisation/preprocessing only and is not provided by "
                                  "ipercore_b200 (SURVEY.md §2a)" % name)
    fn.__name__ = name
    return fn


rasterize = _not_on_hot_path("rasterize")
lighting = _not_on_hot_path("lighting")
rasterize_silhouettes = _not_on_hot_path("rasterize_silhouettes")
rasterize_depth = _not_on_hot_path("rasterize_depth")
rasterize_face_index_map = lambda faces, image_size=256, anti_aliasing=False, near=DEFAULT_NEAR, far=DEFAULT_FAR, **k: \
    rasterize_face_index_map_and_weight_map(faces, image_size, anti_aliasing, near, far)[0]
