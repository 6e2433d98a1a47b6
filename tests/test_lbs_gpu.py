"""GPU parity of the device LBS (row (f) rank 1) against verts/joints produced by the REFERENCE's own lbs() + link()
(tests/golden/lbs.npz, synthetic SMPLH-shaped model) and against the oracle restatement on a larger batch."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(template):
    from ipercore_b200.smpl import SMPLHDevice
    from oracle import lbs_ref
    m = lbs_ref.synthetic_smplh(template=template["verts"])
    dev = SMPLHDevice(m["v_template"], m["shapedirs"], m["posedirs"], m["J_regressor"], m["parents"], m["lbs_weights"],
                      m["hands_mean"]).to("cuda:0")
    return m, dev


def test_lbs_matches_reference_golden(template, golden_dir):
    import make_golden
    g = np.load(os.path.join(golden_dir, "lbs.npz"))
    m, dev = _model(template)
    betas, pose, links, offsets = make_golden.lbs_inputs()
    verts, joints, _ = dev(torch.from_numpy(betas[:1]).cuda(), torch.from_numpy(pose).cuda(),
                           offsets=torch.from_numpy(offsets).cuda(), links_ids=links)
    # the fixture uses a different beta per sample; the device path shares one shape per batch -> compare sample 0, then
    # each sample separately
    np.testing.assert_allclose(verts[0].cpu().numpy(), g["verts"][0], atol=1e-5, rtol=0)
    np.testing.assert_allclose(joints[0].cpu().numpy(), g["joints"][0], atol=1e-5, rtol=0)
    for i in range(1, 3):
        v, j, _ = dev(torch.from_numpy(betas[i:i + 1]).cuda(), torch.from_numpy(pose[i:i + 1]).cuda(),
                      offsets=torch.from_numpy(offsets).cuda(), links_ids=links)
        np.testing.assert_allclose(v[0].cpu().numpy(), g["verts"][i], atol=1e-5, rtol=0)
        np.testing.assert_allclose(j[0].cpu().numpy(), g["joints"][i], atol=1e-5, rtol=0)


def test_lbs_batch_and_get_details(template):
    from oracle import lbs_ref
    m, dev = _model(template)
    rng = np.random.Generator(np.random.PCG64(11))
    B = 19                                          # not a multiple of the 8-frame chunk
    beta = rng.standard_normal((1, 10)).astype(np.float32)
    pose72 = (rng.standard_normal((B, 72)) * 0.5).astype(np.float32)
    cam = rng.uniform(0.5, 1.0, (B, 3)).astype(np.float32)
    theta = np.concatenate([cam, pose72, np.repeat(beta, B, 0)], 1)
    d = dev.get_details(torch.from_numpy(theta).cuda())
    vo, jo = lbs_ref.smplh_forward(m, np.repeat(beta, B, 0), pose72)
    np.testing.assert_allclose(d["verts"].cpu().numpy(), vo, atol=1e-5, rtol=0)
    np.testing.assert_allclose(d["j3d"].cpu().numpy(), jo, atol=1e-5, rtol=0)
    assert d["j2d"].shape == (B, 52, 2) and d["cam"].shape == (B, 3)
    # zero pose -> shaped template exactly (identity rotations, zero pose feature)
    z, _, _ = dev(torch.from_numpy(beta).cuda(), torch.zeros(1, 156).cuda())
    vs = m["v_template"] + np.einsum("l,mkl->mk", beta[0], m["shapedirs"])
    np.testing.assert_allclose(z[0].cpu().numpy(), vs, atol=2e-6, rtol=0)


def test_hand_pca_and_from_reference_cuda_module(template):
    """use_pca (batch_smplh.py:160-168): 78-dim poses expand through the PCA components exactly like the full 156-dim pose;
    from_reference() accepts a CUDA-resident module (every buffer, hands_mean included, is a device tensor there)."""
    import types
    from ipercore_b200.smpl import SMPLHDevice
    from oracle import lbs_ref, synth
    m = lbs_ref.synthetic_smplh(template=synth.base_verts(template).astype(np.float32))
    rng = np.random.Generator(np.random.PCG64(11))
    comps_l = np.linalg.qr(rng.standard_normal((45, 45)))[0][:6].astype(np.float32)
    comps_r = np.linalg.qr(rng.standard_normal((45, 45)))[0][:6].astype(np.float32)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ref_module = types.SimpleNamespace(v_template=t(m["v_template"]), shapedirs=t(m["shapedirs"]), posedirs=t(m["posedirs"]),
                                       J_regressor=t(m["J_regressor"]), parents=t(m["parents"]), lbs_weights=t(m["lbs_weights"]),
                                       hands_mean=t(m["hands_mean"]), left_hand_components=t(comps_l), right_hand_components=t(comps_r),
                                       use_pca=True)
    body = SMPLHDevice.from_reference(ref_module).to(dev)
    assert body.use_pca and body.hands_mean is not None
    betas = (rng.standard_normal((1, 10)) * 0.5).astype(np.float32)
    pose78 = (rng.standard_normal((3, 78)) * 0.3).astype(np.float32)
    v_pca, j_pca, full = body(t(betas), t(pose78))
    full_np = np.concatenate([pose78[:, :66], pose78[:, 66:72] @ comps_l, pose78[:, 72:78] @ comps_r], 1)
    np.testing.assert_allclose(full.cpu().numpy(), full_np, atol=1e-6)
    v_ref, j_ref = lbs_ref.lbs(m, np.repeat(betas, 3, 0), full_np)
    np.testing.assert_allclose(v_pca.cpu().numpy(), v_ref, atol=2e-5)
    np.testing.assert_allclose(j_pca.cpu().numpy(), j_ref, atol=2e-5)
