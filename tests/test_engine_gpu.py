"""GPU parity of the seams (B1/B2) and the batched engine: host-in/host-out frames equal the oracle's composite, the
CUDA-graph path equals the eager path, and partial last batches are handled."""
import numpy as np
import pytest
import torch

from oracle import flow_ref, generator_ref, synth, weights

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CFG = dict(name="AttLWB-SPADE", BGNet=dict(cond_nc=4, n_res_block=6, num_filters=[64, 128, 128, 256]),
           SIDNet=dict(cond_nc=6, n_res_block=6, num_filters=[64, 128, 256]),
           TSFNet=dict(cond_nc=6, n_res_block=6, num_filters=[64, 128, 256]))


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_neural_renderer_and_renderer_seams(template):
    """The reference's own call sequence (nmr.py:326-342) through the drop-in `nr` module == fused render_fim_wim."""
    from ipercore_b200 import neural_renderer as nr
    from ipercore_b200.renders import SMPLRenderer
    S = 96
    r = SMPLRenderer(image_size=S, tables=template, has_front=True, top_k=3).to(DEV)
    cams, verts = synth.pose_sweep(template, 3, total=9)
    cam, v = _t(cams), _t(verts)
    # --- reference glue, verbatim semantics ---
    faces = r.smpl_faces.repeat(3, 1, 1)
    proj = torch.cat((cam[:, 0].view(-1, 1, 1) * (v[:, :, :2] + cam[:, 1:3].view(3, 1, -1)), v[:, :, 2, None]), 2)
    proj[:, :, 1] *= -1
    fv = nr.vertices_to_faces(nr.look_at(proj, r.eye), faces)
    fim_a, wim_a = nr.rasterize_face_index_map_and_weight_map(fv, S, False)
    f2pts_a = fv[:, :, :, 0:2].clone(); f2pts_a[:, :, :, 1] *= -1
    # --- fused seam ---
    f2pts_b, fim_b, wim_b = r.render_fim_wim(cam, v, smpl_faces=True)
    assert fim_a.dtype == torch.int32 and wim_a.dtype == torch.float32
    assert torch.equal(fim_a, fim_b) and torch.equal(wim_a, wim_b) and torch.equal(f2pts_a, f2pts_b)
    # oracle
    f2o, fo, wo = flow_ref.render_fim_wim(cams, verts, template["faces"], S)
    np.testing.assert_array_equal(fim_b.cpu().numpy(), fo)
    # encode_fim / cal_bc_transform / get_vis_f2pts against the reference semantics (numpy restatement)
    cond, _ = r.encode_fim(fim=fim_b)
    np.testing.assert_array_equal(cond.cpu().numpy(), flow_ref.encode_fim(fo, template["map_fn"]))
    T = r.cal_bc_transform(r.get_f_uvs2img(3), fim_b, wim_b)
    np.testing.assert_allclose(T.cpu().numpy(), flow_ref.cal_bc_transform(np.repeat(template["f_uvs2img"][None], 3, 0), fo, wo),
                               atol=1e-6, rtol=0)
    vis = r.get_vis_f2pts(f2pts_b, fim_b).cpu().numpy()
    for i in range(3):
        ids = np.unique(fo[i])[1:]
        keep = np.unique(template["face_k_nearest"][ids])
        exp = np.full_like(f2o[i], -2.0); exp[keep] = f2o[i][keep]
        np.testing.assert_array_equal(vis[i], exp)
    uvf, uvw = r.render_uv_fim_wim(2)
    assert uvf.shape == (2, S, S) and (uvf[0] >= 0).sum() > 0.3 * S * S and torch.equal(uvf[0], uvf[1])


@pytest.mark.parametrize("use_graph", [False, True])
def test_engine_frames_match_oracle(use_graph, template):
    from ipercore_b200.engine import FrameEngine
    from ipercore_b200.generator import AttentionLWBGenerator
    from ipercore_b200.renders import SMPLRenderer
    S, ns, T, B = 128, 2, 5, 2                      # 5 frames in batches of 2 -> partial last batch
    sd = weights.synth_state_dict(0)
    gen = AttentionLWBGenerator(CFG); gen.load_state_dict(sd); gen = gen.to(DEV)
    r = SMPLRenderer(image_size=S, tables=template).to(DEV)
    cams, verts = synth.pose_sweep(template, T, total=11)
    scams, sverts = synth.source_views(template, ns)
    src_img = synth.smooth_image((ns, 3, S, S), seed=1); uv_img = synth.smooth_image((1, 3, S, S), seed=2)
    bg = synth.smooth_image((1, 3, S, S), seed=3)
    src_f2pts, sfim, _ = flow_ref.render_fim_wim(scams, sverts, template["faces"], S)
    src_inputs = np.concatenate([src_img, flow_ref.encode_fim(sfim, template["map_fn"])], 1)[None]
    eng = FrameEngine(gen, r, batch=B, use_graph=use_graph)
    eng.set_source(_t(src_inputs), _t(uv_img), _t(bg), _t(src_f2pts))
    out = eng.synthesize(torch.from_numpy(cams).pin_memory(), torch.from_numpy(verts).pin_memory())
    torch.cuda.synchronize()
    assert out.shape == (T, S, S, 3) and out.dtype == torch.uint8
    # oracle frames
    with torch.no_grad():
        se, sr = generator_ref.forward_src(sd, torch.from_numpy(src_inputs))
        fi = flow_ref.frame_inputs(cams, verts, template["faces"], template["map_fn"], template["f_uvs2img"], uv_img,
                                   src_f2pts, S)
        preds = []
        for i in range(T):     # the reference's forward_tsf pairs src features (bs*ns) with Tst (bs,ns): one frame at a time
            img, mask = generator_ref.forward_tsf(sd, torch.from_numpy(fi["tsf_inputs"][i:i + 1]), se, sr,
                                                  torch.from_numpy(fi["Tst"][i:i + 1]))
            preds.append(generator_ref.composite(img, mask, torch.from_numpy(bg)))
        pred = torch.cat(preds, 0)
    exp = ((pred + 1) / 2.0 * 255).clamp(0, 255).numpy().astype(np.uint8)[:, ::-1].transpose(0, 2, 3, 1)   # BGR, HWC
    diff = np.abs(out.numpy().astype(np.int32) - exp.astype(np.int32))
    assert diff.max() <= 1, "uint8 frames differ from the oracle by more than one code value"
    assert (diff > 0).mean() < 0.02
    assert eng.launches_per_batch > 50
