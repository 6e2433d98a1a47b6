"""GPU test of seam B4: patch.batched_inference (the replacement of Imitator.inference's bs=1 loop) driven through an
Imitator-shaped object — device LBS -> fused raster/flows -> generator -> uint8 frames / PNG files — against the same
frames produced step by step with the oracle-checked pieces."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import flow_ref, lbs_ref, synth, weights

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
CFG = dict(name="AttLWB-SPADE", BGNet=dict(cond_nc=4, n_res_block=6, num_filters=[64, 128, 128, 256]),
           SIDNet=dict(cond_nc=6, n_res_block=6, num_filters=[64, 128, 256]),
           TSFNet=dict(cond_nc=6, n_res_block=6, num_filters=[64, 128, 256]))


class _CamSwapper:
    """cam_pose_utils.WeakPerspectiveCamera surface used by inference: stabilize (identity here) + cam_swap('smooth')."""

    def stabilize(self, smpls):
        return smpls

    @staticmethod
    def cam_swap(src_cam, ref_cam, first_cam=None, strategy="smooth"):
        cam = src_cam.clone()
        cam[:, 1:] += ref_cam[:, 1:] - first_cam[:, 1:]
        cam[:, 0] = cam[:, 0] * ref_cam[:, 0] / first_cam[:, 0]
        return cam


def test_batched_inference_end_to_end(template, tmp_path):
    from ipercore_b200 import patch
    from ipercore_b200.generator import AttentionLWBGenerator
    from ipercore_b200.renders import SMPLRenderer
    from ipercore_b200.smpl import SMPLHDevice
    S, ns, T = 128, 2, 7
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    m = lbs_ref.synthetic_smplh(template=synth.base_verts(template).astype(np.float32) * 0.9)
    body = SMPLHDevice(m["v_template"], m["shapedirs"], m["posedirs"], m["J_regressor"], m["parents"], m["lbs_weights"],
                       m["hands_mean"]).to(DEV)
    gen = AttentionLWBGenerator(CFG); gen.load_state_dict(weights.synth_state_dict(0)); gen = gen.to(DEV)
    render = SMPLRenderer(image_size=S, tables=template).to(DEV)
    rng = np.random.Generator(np.random.PCG64(5))
    shape = (rng.standard_normal((1, 10)) * 0.5).astype(np.float32)
    src_smpl = np.concatenate([np.tile([[0.9, 0.0, -0.25]], (ns, 1)), rng.standard_normal((ns, 72)) * 0.15,
                               np.repeat(shape, ns, 0)], 1).astype(np.float32)
    src = body.get_details(t(src_smpl))
    src_f2pts, sfim, _ = render.render_fim_wim(src["cam"], src["verts"])
    scond, _ = render.encode_fim(fim=sfim)
    src_inputs = torch.cat([t(synth.smooth_image((ns, 3, S, S), seed=1)), scond], 1)[None]
    enc, res = gen.forward_src(src_inputs)
    im = types.SimpleNamespace(
        device=DEV, generator=gen, body_rec=body, weak_cam_swapper=_CamSwapper(), first_cam=None,
        flow_comp=types.SimpleNamespace(render=render),
        _opt=types.SimpleNamespace(image_size=S, temporal=False, only_vis=False),
        src_info=dict(feats=(enc, res), uv_img=t(synth.smooth_image((1, 3, S, S), seed=2)),
                      bg=t(synth.smooth_image((1, 3, S, S), seed=3)), f2pts=src_f2pts, cam=src["cam"],
                      shape=t(np.repeat(shape, ns, 0)), offsets=0, links_ids=None))
    tgt = np.concatenate([np.tile([[1.0, 0.05, -0.2]], (T, 1)) + rng.standard_normal((T, 3)) * 0.02,
                          rng.standard_normal((T, 72)) * 0.3, rng.standard_normal((T, 10))], 1).astype(np.float32)
    out_dir = str(tmp_path)
    paths = patch.batched_inference(im, tgt, "smooth", out_dir, "pred_", batch=3)       # 3 + 3 + 1 frames
    assert [os.path.basename(p) for p in paths] == ["pred_%08d.png" % i for i in range(T)]
    import cv2
    got = np.stack([cv2.imread(p, cv2.IMREAD_COLOR) for p in paths])
    arrs = patch.batched_inference(im, tgt, "smooth", "", "pred_", batch=4)
    assert len(arrs) == T and arrs[0].shape == (3, S, S)
    # step-by-step expectation with the individually verified pieces (same math, one frame at a time)
    first = t(tgt[0:1, 0:3])
    exp = []
    for i in range(T):
        ti = t(tgt[i:i + 1])
        cam = _CamSwapper.cam_swap(src["cam"][0:1], ti[:, 0:3], first)
        ref = body.get_details(torch.cat([cam, ti[:, 3:-10], t(shape)], 1))
        fi = render.frame_inputs(ref["cam"], ref["verts"], im.src_info["uv_img"], src_f2pts)
        _, _, pred = gen.forward_tsf(fi["tsf_inputs"], enc, res, fi["Tst"], bg_img=im.src_info["bg"], return_pred=True)
        exp.append(((pred[0] + 1) / 2.0 * 255).clamp(0, 255).byte().cpu().numpy()[::-1].transpose(1, 2, 0))
    exp = np.stack(exp)
    d = np.abs(got.astype(np.int32) - exp.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 0.01
    assert (sfim >= 0).sum() > 500, "the synthetic source must cover pixels for the test to mean anything"
