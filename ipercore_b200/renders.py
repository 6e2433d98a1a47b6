"""Seam B2 — SMPLRenderer-compatible module on the fused B200 raster/flow kernels.

Mirrors the hot-path surface of iPERCore/tools/human_digitalizer/renders/nmr.py (class BaseSMPLRenderer :127-757,
SMPLRenderer :766-988): same constructor arguments, same buffer names (smpl_faces, obj_faces, map_fn, front_map_fn,
f_img2uvs, face_k_nearest, f_uvs2img, coords, img2uv_sampler), and the methods FlowComposition calls
(iPERCore/models/flowcomposition.py:60-76,139-248,514-582): render_fim_wim, render_uv_fim_wim, encode_fim,
encode_front_fim, cal_bc_transform, get_vis_f2pts, get_f_uvs2img.  `frame_inputs` is the fused batched entry point
the engine uses (one launch per batch of target frames).
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import mesh_tables as mt
from . import ops


class SMPLRenderer(nn.Module):
    def __init__(self, face_path="assets/checkpoints/pose3d/smpl_faces.npy",
                 fim_enc_path="assets/configs/pose3d/mapper_fim_enc.txt",
                 uv_map_path="assets/configs/pose3d/mapper_uv.txt",
                 part_path="assets/configs/pose3d/smpl_part_info.json",
                 front_path="assets/configs/pose3d/front_body.json",
                 head_path="assets/configs/pose3d/head.json",
                 facial_path="assets/configs/pose3d/front_facial.json",
                 map_name="uv_seg", tex_size=3, image_size=256, anti_aliasing=True, fill_back=False,
                 background_color=(0, 0, 0), viewing_angle=30, near=0.1, far=25.0, has_front=False, top_k=5,
                 tables=None):
        super().__init__()
        if fill_back:
            raise NotImplementedError("fill_back=True is not used by FlowComposition (flowcomposition.py:67)")
        if map_name != "uv_seg":
            raise NotImplementedError("only map_name='uv_seg' (deploy.toml:45) is on the B200 path")
        self.image_size, self.map_name, self.tex_size = image_size, map_name, tex_size
        self.fill_back, self.anti_aliasing, self.background_color = fill_back, anti_aliasing, background_color
        self.near, self.far, self.viewing_angle = near, far, viewing_angle
        self.eye = [0, 0, -(1. / np.tan(np.radians(viewing_angle)) + 1)]
        if viewing_angle != 30:
            raise NotImplementedError("the fused projection is specialised to viewing_angle=30 (nmr.py:137,225)")
        if tables is None:
            tables = self.build_tables(face_path, fim_enc_path, uv_map_path, part_path, head_path, tex_size,
                                       has_front, top_k)
        self.nf = self.base_nf = int(tables["faces"].shape[0])
        reg = lambda name, arr, dt: self.register_buffer(name, torch.as_tensor(np.asarray(arr)).to(dt))
        reg("smpl_faces", tables["faces"], torch.int32)
        reg("obj_faces", tables["obj_faces"], torch.int32)
        reg("map_fn", tables["map_fn"], torch.float32)
        if tables.get("front_map_fn") is not None:
            reg("front_map_fn", tables["front_map_fn"], torch.float32)
        else:
            self.front_map_fn = None
        reg("f_uvs2img", tables["f_uvs2img"], torch.float32)
        if tables.get("f_img2uvs") is not None:
            reg("f_img2uvs", tables["f_img2uvs"], torch.float32)
        if tables.get("face_k_nearest") is not None:
            reg("face_k_nearest", tables["face_k_nearest"], torch.int64)
        if tables.get("img2uv_sampler") is not None:
            reg("img2uv_sampler", tables["img2uv_sampler"], torch.float32)
        step = 1 if tex_size == 1 else 1 / (tex_size - 1)
        ab = torch.arange(0, 1 + step, step, dtype=torch.float32)
        xv, yv = torch.meshgrid([ab, ab], indexing="ij")
        self.register_buffer("coords", torch.stack([xv.flatten(), yv.flatten()], dim=0))

    @staticmethod
    def build_tables(face_path, fim_enc_path, uv_map_path, part_path, head_path, tex_size, has_front, top_k):
        enc, uv = mt.load_obj(fim_enc_path), mt.load_obj(uv_map_path)
        faces = np.load(face_path).astype(np.int32) if os.path.exists(face_path) else uv["faces"]
        nf = faces.shape[0]
        f_img2uvs = mt.face_uv_corners(enc, z=1)
        t = dict(faces=faces, obj_faces=enc["faces"], map_fn=mt.uv_seg_mapping(enc), f_img2uvs=f_img2uvs,
                 f_uvs2img=mt.face_uv_corners(uv, z=1)[:, :, 0:2], img2uv_sampler=mt.uv_sampler(uv, tex_size))
        if has_front and os.path.exists(head_path):
            t["front_map_fn"] = mt.face_flag_mapping(nf, head_path)      # nmr.py:186-188 builds it with map_name "head"
        if os.path.exists(part_path):
            t["face_k_nearest"] = mt.part_k_nearest_faces(f_img2uvs, mt.part_face_ids(nf, part_path), top_k)
        return t

    # ---- hot-path methods (names, arguments and return layouts as in nmr.py) -------------------------------------
    def set_img_size(self, image_size):
        self.image_size = image_size

    def render_fim_wim(self, cam, vertices, smpl_faces=True):
        """nmr.py:319-342 -> (f2pts (N,F,3,2), fim (N,S,S) int32, wim (N,S,S,3)); any N."""
        faces = self.smpl_faces if smpl_faces else self.obj_faces
        out = ops.raster_frames(vertices.float().contiguous(), cam.float().contiguous(), faces, self.image_size)
        return out["f2pts"], out["fim"], out["wim"]

    def render_uv_fim_wim(self, bs):
        """nmr.py:344-358: rasterise the UV layout itself (constant; computed once, repeated bs times)."""
        f = self.f_img2uvs.clone()
        f[:, :, 1] *= -1
        fim, wim = ops.rasterize_faces(f[None].contiguous(), self.image_size)
        return fim.repeat(bs, 1, 1), wim.repeat(bs, 1, 1, 1)

    def encode_fim(self, cam=None, vertices=None, fim=None, transpose=True, map_fn=None):
        assert (cam is not None and vertices is not None) or fim is not None
        if fim is None:
            _, fim, _ = self.render_fim_wim(cam, vertices)
        return ops.encode_fim(fim.int(), self.map_fn if map_fn is None else map_fn, transpose), fim

    def encode_front_fim(self, fim, transpose=True):
        return ops.encode_fim(fim.int(), self.front_map_fn, transpose)

    def cal_bc_transform(self, src_f2pts, dst_fims, dst_wims):
        """nmr.py:713-757 -> (bs,S,S,2); background = -2 exactly."""
        return ops.cal_bc_transform(src_f2pts.float(), dst_fims.int(), dst_wims.float())

    def get_f_uvs2img(self, bs):
        return self.f_uvs2img.repeat(bs, 1, 1, 1)

    def get_vis_f2pts(self, f2pts, fims):
        """nmr.py:639-681: visible faces (minus the smallest unique value of the map, as `fim.unique()[1:]` does) and their
        UV-nearest neighbours keep their coordinates, the rest are set to -2 — iper_vis_f2pts, no host sync."""
        return ops.vis_f2pts(f2pts, fims, self.face_k_nearest)

    # ---- fused engine entry point -------------------------------------------------------------------------------
    def frame_inputs(self, cam, vertices, uv_img, src_f2pts, want_fim=False):
        """rows a1-a8 for a batch of independent target frames in ONE launch: tsf_inputs (B,6,S,S), Tst (B,ns,S,S,2)."""
        uv = uv_img.reshape(-1, 3, self.image_size, self.image_size)[0].float().contiguous()
        fused = dict(map_fn=self.map_fn, f_uvs2img=self.f_uvs2img, uv_img=uv, src_f2pts=src_f2pts.float().contiguous())
        return ops.raster_frames(vertices.float().contiguous(), cam.float().contiguous(), self.smpl_faces,
                                 self.image_size, want_fim=want_fim, want_f2pts=False, fused=fused)
