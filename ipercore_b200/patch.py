"""install() — swap the B200 kernels into an importable iPERCore so its services run unchanged (seam B4).

    import ipercore_b200.patch as p; p.install()          # before `import iPERCore.models` / services
    from iPERCore.services.run_imitator import run_imitator

What is replaced, at the reference's own seams:
  B1  ``sys.modules["neural_renderer"]``  -> ipercore_b200.neural_renderer  (nmr.py:8 imports it as ``nr``)
  B2  ``BaseSMPLRenderer/SMPLRenderer.render_fim_wim``, ``cal_bc_transform``, ``encode_fim``, ``get_vis_f2pts`` (nmr.py:319-342,
      390-401, 639-681, 713-757, and the bs==3 loop :892-918) -> CUDA kernels; every other method and all buffers stay the
      reference's
  B3  ``NetworksFactory.get_by_name("AttLWB-SPADE", cfg=..., temporal=...)`` (networks/__init__.py:14-16)
      -> ipercore_b200.generator.AttentionLWBGenerator (loads the same checkpoints); with ``temporal=true`` the upstream bs=1
      loop (TemporalFIFO recurrence) drives it frame by frame
  B2' ``iPERCore.tools.utils.morphology.morph / soft_dilate`` (morph_ops.py:7-61; source_setup masks,
      flowcomposition.py:121,176-180,258) -> iper_morph (separable box sum + threshold); custom ``kernel=`` calls, CPU
      tensors and even / > 63 sizes keep going to the reference implementation
  B2" ``FlowComposition.make_morph_image`` / ``make_uv_img`` (flowcomposition.py:87-137, 264-386; CannyFilter of
      canny_ops.py) -> ipercore_b200.source_ops (Canny, top-3 nearest-boundary colour fill, UV merge kernels)
  B4  ``Imitator.inference`` (models/imitator.py:327-382) -> batched FrameEngine when ``temporal`` is false and the
      generator is ours; same arguments, same ``pred_{:0>8}.png`` outputs / returned list.
Nothing else of iPERCore is touched: options, preprocessing, personalisation, source_setup, video fusion stay upstream.
"""
import os
import sys

import numpy as np
import torch

_INSTALLED = False
_DEVICE_LBS = True


def install(precision="fp16x2", batch=16, patch_inference=True, device_lbs=True):
    """device_lbs: run SMPLH linear blend skinning of the target frames on the LBS kernels (ipercore_b200.smpl) built
    from the reference body model's own buffers, instead of the reference's per-frame torch implementation."""
    global _INSTALLED, _DEVICE_LBS
    _DEVICE_LBS = device_lbs
    if _INSTALLED:
        return
    from . import neural_renderer as nr
    from . import ops
    sys.modules["neural_renderer"] = nr                                   # B1
    from iPERCore.tools.human_digitalizer.renders import nmr               # noqa: E402  (imports `nr` -> ours)

    def render_fim_wim(self, cam, vertices, smpl_faces=True):             # B2 (any batch size, incl. 3)
        faces = self.smpl_faces if smpl_faces else self.obj_faces
        out = ops.raster_frames(vertices.float().contiguous(), cam.float().contiguous(), faces.int().contiguous(),
                                self.image_size)
        return out["f2pts"], out["fim"], out["wim"]

    def cal_bc_transform(self, src_f2pts, dst_fims, dst_wims):
        return ops.cal_bc_transform(src_f2pts.float(), dst_fims.int(), dst_wims.float())

    def encode_fim(self, cam=None, vertices=None, fim=None, transpose=True, map_fn=None):
        assert (cam is not None and vertices is not None) or fim is not None
        return ops.encode_fim(fim.int(), (self.map_fn if map_fn is None else map_fn).float(), transpose), fim

    def get_vis_f2pts(self, f2pts, fims):                                  # nmr.py:639-681 (two torch.unique + host syncs)
        return ops.vis_f2pts(f2pts, fims, self.face_k_nearest)

    for cls in (nmr.BaseSMPLRenderer, nmr.SMPLRenderer):
        cls.render_fim_wim = render_fim_wim
        cls.cal_bc_transform = cal_bc_transform
        cls.encode_fim = encode_fim
        cls.get_vis_f2pts = get_vis_f2pts

    import iPERCore.tools.utils.morphology as morphology                   # B2': mask morphology of source_setup
    from iPERCore.tools.utils.morphology import morph_ops
    up_morph, up_soft = morph_ops.morph, morph_ops.soft_dilate

    def _ours(m, ks, kernel):
        return kernel is None and m.is_cuda and m.dim() == 4 and m.shape[1] == 1 and ks % 2 == 1 and ks <= 63

    def morph(src_bg_mask, ks, mode="erode", kernel=None):
        if not _ours(src_bg_mask, ks, kernel):
            return up_morph(src_bg_mask, ks, mode=mode, kernel=kernel)
        return ops.morph(src_bg_mask.float().contiguous(), ks, ops.MORPH_ERODE if mode == "erode" else ops.MORPH_DILATE)

    def soft_dilate(src_bg_mask, ks, kernel=None):
        if not _ours(src_bg_mask, ks, kernel):
            return up_soft(src_bg_mask, ks, kernel=kernel)
        return ops.morph(src_bg_mask.float().contiguous(), ks, ops.MORPH_SOFT_DILATE)

    morph_ops.morph, morph_ops.soft_dilate = morph, soft_dilate
    morphology.morph, morphology.soft_dilate = morph, soft_dilate
    for modname in ("iPERCore.models.flowcomposition", "iPERCore.tools.human_digitalizer.deformers.sil_deformer",
                    "iPERCore.tools.trainers.base"):      # modules that bound `morph` by name at import time
        mod = sys.modules.get(modname)
        if mod is not None and hasattr(mod, "morph"):
            mod.morph = morph

    from iPERCore.models import flowcomposition as fcm                     # B2'': one-time-per-source stage of source_setup
    from . import source_ops
    up_morph_image, up_uv_img = fcm.FlowComposition.make_morph_image, fcm.FlowComposition.make_uv_img

    def make_morph_image(self, src_img, src_info, erode_ks=3, dilate_ks=11):      # flowcomposition.py:335-386
        if not src_img.is_cuda:
            return up_morph_image(self, src_img, src_info, erode_ks=erode_ks, dilate_ks=dilate_ks)
        return source_ops.make_morph_image(src_img, src_info["confidant_sil"], src_info["outpad_sil"], erode_ks, dilate_ks)

    def make_uv_img(self, src_img, src_info):                                     # flowcomposition.py:87-137
        if not src_img.is_cuda:
            return up_uv_img(self, src_img, src_info)
        n = src_img.shape[0] * src_img.shape[1]
        return source_ops.make_uv_img(src_img, src_info["obj_f2pts"], src_info["only_vis_obj_f2pts"], self.uv_fim[0:n],
                                      self.uv_wim[0:n])

    fcm.FlowComposition.make_morph_image = make_morph_image
    fcm.FlowComposition.make_uv_img = make_uv_img

    from iPERCore.models.networks import NetworksFactory                   # B3
    from .generator import AttentionLWBGenerator
    upstream_get = NetworksFactory.get_by_name

    def get_by_name(network_name, *args, **kwargs):
        if network_name == "AttLWB-SPADE":
            net = AttentionLWBGenerator(*args, precision=precision, **kwargs)
            print("Network %s was created (ipercore_b200, %s)" % (network_name, precision))
            return net
        return upstream_get(network_name, *args, **kwargs)

    NetworksFactory.get_by_name = staticmethod(get_by_name)

    if patch_inference:                                                    # B4
        from iPERCore.models import imitator as imod
        upstream_inference = imod.Imitator.inference

        def inference(self, tgt_smpls, cam_strategy="smooth", output_dir="", prefix="pred_", use_selected_f2pts=False,
                      visualizer=None, verbose=True):
            ours = isinstance(self.generator, AttentionLWBGenerator)
            if not ours or self._opt.temporal or visualizer is not None or use_selected_f2pts or self._opt.only_vis:
                return upstream_inference(self, tgt_smpls, cam_strategy, output_dir, prefix, use_selected_f2pts,
                                          visualizer, verbose)
            return batched_inference(self, tgt_smpls, cam_strategy, output_dir, prefix, batch)

        imod.Imitator.inference = inference
    _INSTALLED = True


def _source_key(src):
    """Identity of the per-source cache the captured graph reads: same tensors (and versions) -> the graph is reused."""
    enc, res = src["feats"]
    ts = list(enc) + list(res) + [src["uv_img"], src["bg"], src["f2pts"]]
    return tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in ts)


class FrameWriter:
    """Output stage of Imitator.inference (imitator.py:365-379 -> cv_utils.save_cv2_img, cv_utils.py:100-116): PNG encodes
    run on a thread pool (cv2.imwrite releases the GIL) batch by batch as the frames land in pinned memory, instead of
    after the whole clip.  Without an output directory the reference returns float CHW arrays in [-1, 1]."""

    def __init__(self, T, output_dir, prefix, workers=None):
        import cv2
        from concurrent.futures import ThreadPoolExecutor
        self.cv2, self.dir, self.T = cv2, output_dir, T
        self.paths = [os.path.join(output_dir, prefix + "{:0>8}.png".format(t)) for t in range(T)] if output_dir else None
        self.arrays = None if output_dir else [None] * T
        self.pool = ThreadPoolExecutor(max_workers=workers or min(16, os.cpu_count() or 4))

    def _one(self, t, frame):
        if self.paths is not None:
            if not self.cv2.imwrite(self.paths[t], frame):
                raise IOError("cv2.imwrite failed for %s" % self.paths[t])
        else:
            f = frame[:, :, ::-1].astype(np.float32) / 255.0 * 2.0 - 1.0
            self.arrays[t] = np.ascontiguousarray(f.transpose(2, 0, 1))

    def sink(self, lo, hi, frames):
        arr = frames.numpy()                   # uint8 BGR HWC (cv_utils.save_cv2_img semantics), pinned ring slot
        return [self.pool.submit(self._one, lo + k, arr[k]) for k in range(hi - lo)]

    def close(self):
        self.pool.shutdown(wait=True)
        return self.paths if self.paths is not None else self.arrays


@torch.no_grad()
def batched_inference(imitator, tgt_smpls, cam_strategy="smooth", output_dir="", prefix="pred_", batch=16):
    """Imitator.inference (models/imitator.py:327-382) with the bs=1 loop replaced by the batched engine.

    Per-sequence host pre-pass exactly as upstream (:337-339, :298-305): stabilise, first_cam, cam swap; the SMPL body
    model produces the vertices in chunks of `batch` frames — on the device LBS kernels (ipercore_b200.smpl, built from
    the reference body model's buffers, hand-pose PCA included) unless install(device_lbs=False).  Every batch is
    issued on the engine's stream (cam swap -> LBS -> raster -> generator -> uint8 -> D2H into a bounded pinned ring) and
    its PNG files are encoded on a thread pool while the GPU runs the next batch.  The CUDA graph is captured once per
    source set, not per call.  Returns the list of paths (or of CHW float arrays when output_dir is empty)."""
    from .engine import FrameEngine
    dev, opt, src = imitator.device, imitator._opt, imitator.src_info
    tgt = torch.as_tensor(np.asarray(tgt_smpls), dtype=torch.float32).to(dev)
    if cam_strategy == "smooth":
        tgt = imitator.weak_cam_swapper.stabilize(tgt)
    imitator.first_cam = tgt[0:1, 0:3].clone() if cam_strategy == "smooth" else None
    eng = getattr(imitator, "_iper_engine", None)
    if eng is None or eng.B != batch or eng.gen is not imitator.generator:
        eng = FrameEngine(imitator.generator, _EngineRenderer(imitator.flow_comp.render), batch=batch, device=dev)
        eng._src_key = None
        imitator._iper_engine = eng
    key = _source_key(src)
    if eng._src_key != key:                       # new source set: new cache, re-capture on first use
        enc, res = src["feats"]
        eng.src = dict(enc=enc, res=res, uv_img=src["uv_img"].float().contiguous(),
                       bg=src["bg"].float().reshape(1, 3, opt.image_size, opt.image_size).contiguous(),
                       src_f2pts=src["f2pts"].float().contiguous())
        eng.graph = None
        eng._src_key = key
    from .smpl import SMPLHDevice
    body = imitator.body_rec
    if (_DEVICE_LBS and not isinstance(body, SMPLHDevice)
            and not (hasattr(src["links_ids"], "ndim") and src["links_ids"].ndim == 3)):
        if getattr(imitator, "_iper_smpl", None) is None:
            imitator._iper_smpl = SMPLHDevice.from_reference(body).to(dev)
        body = imitator._iper_smpl
    pinned = isinstance(body, SMPLHDevice)
    if pinned:                                    # every target frame is posed with the SOURCE shape (imitator.py:248-256)
        body.pin_shape(src["shape"][0], src["offsets"])
    T = tgt.shape[0]
    src_cam, src_shape, first_cam = src["cam"][0:1], src["shape"][0:1], imitator.first_cam

    def feed(lo, hi):
        # runs on the engine's stream: every tensor of the batch is created, consumed and freed on that one stream
        t = tgt[lo:hi]
        n = hi - lo
        cam = imitator.weak_cam_swapper.cam_swap(src_cam.expand(n, -1), t[:, 0:3],
                                                 first_cam.expand(n, -1) if first_cam is not None else None, cam_strategy)
        ref_smpl = torch.cat([cam, t[:, 3:-10], src_shape.expand(n, -1)], dim=1)
        info = body.get_details(ref_smpl, src["offsets"], links_ids=src["links_ids"])
        return info["cam"].float().contiguous(), info["verts"].float().contiguous()

    writer = FrameWriter(T, output_dir, prefix)
    try:
        eng.synthesize_stream(T, feed, writer.sink)
    finally:
        outputs = writer.close()
        if pinned:
            body.unpin_shape()
    return outputs


class _EngineRenderer:
    """Adapter giving FrameEngine the fused frame_inputs() on top of the REFERENCE's SMPLRenderer buffers."""

    def __init__(self, render):
        self.r = render
        self.image_size = render.image_size

    def frame_inputs(self, cam, vertices, uv_img, src_f2pts, want_fim=False):
        from . import ops
        r, S = self.r, self.image_size
        fused = dict(map_fn=r.map_fn.float().contiguous(), f_uvs2img=r.f_uvs2img.float().contiguous(),
                     uv_img=uv_img.reshape(-1, 3, S, S)[0].float().contiguous(), src_f2pts=src_f2pts.float().contiguous())
        return ops.raster_frames(vertices.float().contiguous(), cam.float().contiguous(), r.smpl_faces.int().contiguous(),
                                 S, want_fim=want_fim, want_f2pts=False, fused=fused)
