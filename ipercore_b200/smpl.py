"""Device SMPL / SMPLH body model — the step that feeds the rasteriser every frame.

Mirrors the hot-path surface of ``iPERCore/tools/human_digitalizer/bodynets`` (``SMPLH.forward`` batch_smplh.py:137-180,
``BaseSMPL.get_details`` base_smpl.py:107-142, ``link`` :28-50) on the LBS kernels of lbs.cu: same argument meaning
(theta = [cam(3) | pose(72 or 156) | shape(10)], offsets, links_ids) and the same dict keys for the fields the per-frame
path reads (cam, pose, shape, verts, j3d, j2d, theta).  Shape-dependent quantities (v_shaped, rest joints) are cached per
(betas, offsets): in run_imitator every target frame uses the SOURCE shape (imitator.py:248-256).

Model parameters come from the user's SMPL pkl (through the reference's own loader) or from arrays; the hand-pose PCA
path (``use_pca``: 78-dim poses, batch_smplh.py:160-168) is expanded with the model's PCA components before the kernels.
"""
import numpy as np
import torch
import torch.nn as nn

from ._lib import check, lib
from .ops import _ptr, _stream


class SMPLHDevice(nn.Module):
    def __init__(self, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights, hands_mean=None,
                 left_hand_components=None, right_hand_components=None, use_pca=False):
        super().__init__()
        f = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32).contiguous()
        self.nv, self.nj, self.nb = int(v_template.shape[0]), int(J_regressor.shape[0]), int(shapedirs.shape[-1])
        self.register_buffer("v_template", f(v_template))
        self.register_buffer("shapedirs", f(shapedirs).reshape(self.nv, 3, self.nb).contiguous())
        self.register_buffer("posedirs", f(posedirs))                                   # ((nj-1)*9, nv*3)
        self.register_buffer("J_regressor", f(J_regressor))
        self.register_buffer("parents", torch.as_tensor(np.asarray(parents), dtype=torch.int32))
        self.register_buffer("weights_t", f(lbs_weights).t().contiguous())              # (nj, nv)
        if hands_mean is not None:
            self.register_buffer("hands_mean", f(hands_mean))
        else:
            self.hands_mean = None
        # hand-pose PCA (batch_smplh.py:25-26, 105-131, 160-168): 78-dim poses = 66 body + 6 + 6 PCA coefficients
        self.use_pca = bool(use_pca)
        if left_hand_components is not None and right_hand_components is not None:
            self.register_buffer("left_hand_components", f(left_hand_components))
            self.register_buffer("right_hand_components", f(right_hand_components))
        else:
            self.left_hand_components = self.right_hand_components = None
        if self.use_pca and self.left_hand_components is None:
            raise ValueError("use_pca needs the left/right hand PCA components")
        assert self.posedirs.shape == ((self.nj - 1) * 9, self.nv * 3)
        self._shape_key, self._shape_cache = None, None
        self._pinned = False

    @classmethod
    def from_reference(cls, smplh):
        """Build from an instantiated reference ``SMPLH`` / ``SMPL`` module (its registered buffers)."""
        n = lambda t: None if t is None else (t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t))
        return cls(n(smplh.v_template), n(smplh.shapedirs), n(smplh.posedirs), n(smplh.J_regressor), n(smplh.parents),
                   n(smplh.lbs_weights), n(getattr(smplh, "hands_mean", None)), n(getattr(smplh, "left_hand_components", None)),
                   n(getattr(smplh, "right_hand_components", None)), use_pca=bool(getattr(smplh, "use_pca", False)))

    # ---- one-time per shape ----------------------------------------------------------------------------------------
    def pin_shape(self, betas, offsets=0):
        """Compute the shape-dependent quantities ONCE (one host read of the 10 betas) and keep them for every following
        forward()/get_details() call until unpin_shape(): run_imitator poses every target frame with the SOURCE shape
        (imitator.py:248-256), so the per-batch path then has no host synchronisation at all."""
        self._pinned = False
        self._shape(betas.reshape(-1, self.nb)[0], offsets)
        self._pinned = True

    def unpin_shape(self):
        self._pinned = False

    def _shape(self, betas, offsets):
        if self._pinned:
            return self._shape_cache
        dev = self.v_template.device
        off = None
        if torch.is_tensor(offsets) and offsets.numel() > 1:
            off = offsets.to(dev).float().reshape(self.nv, 3).contiguous()
        key = (betas.detach().cpu().numpy().tobytes(), None if off is None else (off.data_ptr(), off._version))
        if self._shape_key != key:
            v_shaped = torch.empty((self.nv, 3), dtype=torch.float32, device=dev)
            J = torch.empty((self.nj, 3), dtype=torch.float32, device=dev)
            b = betas.to(dev).float().reshape(-1).contiguous()
            check(lib.iper_lbs_shape(self.v_template.data_ptr(), _ptr(off), self.shapedirs.data_ptr(), b.data_ptr(),
                                     self.J_regressor.data_ptr(), self.nv, self.nb, self.nj, v_shaped.data_ptr(),
                                     J.data_ptr(), _stream()), "lbs_shape")
            self._shape_key, self._shape_cache = key, (v_shaped, J, off)
        return self._shape_cache

    @staticmethod
    def link_source(nv, links_ids, device):
        """src_of[v] for the 2-D links form (base_smpl.py:43-44: verts[:, ids[:,0]] = verts[:, ids[:,1]])."""
        if links_ids is None:
            return None
        ids = torch.as_tensor(np.asarray(links_ids)).long()
        if ids.dim() != 2:
            raise NotImplementedError("per-sample links_ids (N,nv,3) are not supported on the device path")
        src = torch.arange(nv, dtype=torch.int32)
        src[ids[:, 0]] = ids[:, 1].int()
        return src.to(device)

    # ---- per batch of frames -----------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, beta, theta, offsets=0, links_ids=None, get_skin=True):
        """beta (1|N,10) — one shape for the batch; theta (N,72|156) axis-angle -> (verts (N,nv,3), joints (N,nj,3), full_pose)."""
        dev = self.v_template.device
        theta = theta.to(dev).float()
        N = theta.shape[0]
        if theta.shape[1] == 72:
            if self.hands_mean is None:
                raise ValueError("72-dim pose needs hands_mean (SMPLH)")
            theta = torch.cat([theta[:, :66], self.hands_mean[None].expand(N, -1)], dim=1)
        if self.use_pca:        # batch_smplh.py:160-168: the last 12 values are PCA coefficients of the two hands
            lh = theta[:, -12:-6] @ self.left_hand_components
            rh = theta[:, -6:] @ self.right_hand_components
            theta = torch.cat([theta[:, :-12], lh, rh], dim=1)
        if theta.shape[1] != self.nj * 3:
            raise ValueError("pose has %d dims, model has %d joints" % (theta.shape[1], self.nj))
        if not self._pinned and beta.dim() == 2 and beta.shape[0] > 1 and not bool((beta == beta[0:1]).all()):
            raise NotImplementedError("the device path shares one shape across the batch (run_imitator uses the source shape)")
        v_shaped, J, _ = self._shape(beta.reshape(-1, self.nb)[0], offsets)
        theta = theta.contiguous()
        pf = torch.empty((N, (self.nj - 1) * 9), dtype=torch.float32, device=dev)
        A = torch.empty((N, self.nj, 12), dtype=torch.float32, device=dev)
        joints = torch.empty((N, self.nj, 3), dtype=torch.float32, device=dev)
        verts = torch.empty((N, self.nv, 3), dtype=torch.float32, device=dev)
        src_of = self.link_source(self.nv, links_ids, dev)
        check(lib.iper_lbs_frames(theta.data_ptr(), N, self.nj, v_shaped.data_ptr(), J.data_ptr(), self.parents.data_ptr(),
                                  self.posedirs.data_ptr(), self.weights_t.data_ptr(), _ptr(src_of), self.nv,
                                  pf.data_ptr(), A.data_ptr(), joints.data_ptr(), verts.data_ptr(), _stream()), "lbs_frames")
        return verts, joints, theta

    def get_details(self, theta, offsets=0, links_ids=None):
        """base_smpl.py:107-142: theta (N, 3 + pose + 10) -> dict(theta, cam, pose, shape, verts, j3d, j2d)."""
        cam, pose, shape = theta[:, 0:3], theta[:, 3:-10].contiguous(), theta[:, -10:].contiguous()
        verts, j3d, _ = self.forward(shape, pose, offsets=offsets, links_ids=links_ids)
        cam_d = cam.to(verts.device).float()
        j2d = cam_d[:, None, 0:1] * (j3d[:, :, :2] + cam_d[:, None, 1:3])      # batch_orth_proj_idrot
        return {"theta": theta, "cam": cam_d, "pose": pose, "shape": shape, "verts": verts, "j2d": j2d, "j3d": j3d}
