"""GPU parity: rasteriser + fused frame geometry (C ABI) against the CPU oracle — bit-exact for integer/index work."""
import os

import numpy as np
import pytest
import torch

from oracle import flow_ref, raster, synth

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _t(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x)).to(_dev())
    return t if dtype is None else t.to(dtype)


@pytest.mark.parametrize("S,n", [(64, 3), (128, 3), (512, 2), (100, 2)])
def test_raster_frames_bit_exact(S, n, template):
    from ipercore_b200 import ops
    cams, verts = synth.pose_sweep(template, n, total=7)
    f2pts_o, fim_o, wim_o = flow_ref.render_fim_wim(cams, verts, template["faces"], S)
    out = ops.raster_frames(_t(verts), _t(cams), _t(template["faces"]), S)
    torch.cuda.synchronize()
    assert (fim_o >= 0).sum() > 50
    np.testing.assert_array_equal(out["fim"].cpu().numpy(), fim_o)          # face indices: bit-exact
    np.testing.assert_array_equal(out["wim"].cpu().numpy(), wim_o)          # weights: same float sequence -> same bits
    np.testing.assert_array_equal(out["f2pts"].cpu().numpy(), f2pts_o)


def test_raster_contraction_switch_bit_exact(template):
    """iper_raster_set_contraction(1): the nvcc -fmad=true rounding model, bit-exact against the oracle's -DORACLE_FMA variant
    (and switching back restores the default model) — the re-validation switch for when the fork can be built."""
    from ipercore_b200 import _lib, ops
    S = 256
    cams, verts = synth.pose_sweep(template, 2, total=7)
    fv = flow_ref.vertices_to_faces(flow_ref.project(cams, verts), template["faces"])
    fim1, wim1 = raster.rasterize_fim_wim(fv, S, fma=True)
    fim0, wim0 = raster.rasterize_fim_wim(fv, S, fma=False)
    assert _lib.lib.iper_raster_get_contraction() == 0
    try:
        _lib.check(_lib.lib.iper_raster_set_contraction(1), "set_contraction")
        out = ops.raster_frames(_t(verts), _t(cams), _t(template["faces"]), S)
        np.testing.assert_array_equal(out["fim"].cpu().numpy(), fim1)
        np.testing.assert_array_equal(out["wim"].cpu().numpy(), wim1)
    finally:
        _lib.check(_lib.lib.iper_raster_set_contraction(0), "set_contraction")
    out = ops.raster_frames(_t(verts), _t(cams), _t(template["faces"]), S)
    np.testing.assert_array_equal(out["fim"].cpu().numpy(), fim0)
    np.testing.assert_array_equal(out["wim"].cpu().numpy(), wim0)
    assert (wim0 != wim1).any()


def test_vis_f2pts_matches_reference_golden(template, golden_dir):
    """iper_vis_f2pts vs keep-masks produced by the reference's own SMPLRenderer.get_vis_f2pts (tests/golden/vis.npz): maps
    with background (drops -1) and a map without background (drops the LOWEST visible face id, as `unique()[1:]` does)."""
    import make_golden
    from ipercore_b200 import ops
    v = np.load(os.path.join(golden_dir, "vis.npz"))
    g = np.load(os.path.join(golden_dir, "flow_S128.npz"))
    nb = _t(template["face_k_nearest"].astype(np.int64))
    for tag, f2pts, fim in (("s128", g["f2pts"], g["fim"]), ("nobg", g["f2pts"][:1], make_golden.vis_inputs())):
        want = np.where(v["keep_" + tag][:, :, None, None], f2pts, np.float32(-2.0))
        got = ops.vis_f2pts(_t(f2pts), _t(fim), nb)
        np.testing.assert_array_equal(got.cpu().numpy(), want)
        got1 = ops.vis_f2pts(_t(f2pts[0]), _t(fim[0]), nb)                 # the 3-D form of the reference signature
        np.testing.assert_array_equal(got1.cpu().numpy(), want[0])
    f3 = np.concatenate([g["f2pts"], np.ones_like(g["f2pts"][..., :1])], -1)   # (bs, nf, 3, 3) corners
    got3 = ops.vis_f2pts(_t(f3), _t(g["fim"]), nb)
    np.testing.assert_array_equal(got3.cpu().numpy(), np.where(v["keep_s128"][:, :, None, None], f3, np.float32(-2.0)))


def test_rasterize_faces_seam_any_batch(template):
    """Seam B1 incl. batch size 3 (the upstream bug the reference loops around, nmr.py:892-918) and empty batch."""
    from ipercore_b200 import ops
    S = 96
    cams, verts = synth.pose_sweep(template, 3, total=5)
    fv = flow_ref.vertices_to_faces(flow_ref.project(cams, verts), template["obj_faces"])
    fim_o, wim_o = raster.rasterize_fim_wim(fv, S)
    fim, wim = ops.rasterize_faces(_t(fv), S)
    np.testing.assert_array_equal(fim.cpu().numpy(), fim_o)
    np.testing.assert_array_equal(wim.cpu().numpy(), wim_o)
    fim0, wim0 = ops.rasterize_faces(_t(fv[:0]), S)
    assert fim0.shape == (0, S, S) and wim0.shape == (0, S, S, 3)


def test_raster_edge_cases():
    from ipercore_b200 import ops
    S = 8
    tri = np.array([[-1, -1, 1], [3, -1, 1], [-1, 3, 1]], np.float32)
    cases = np.stack([np.stack([tri, tri]),                                   # exact tie -> lowest index
                      np.stack([tri * [1, 1, 2], tri]),                       # nearer face second
                      np.stack([tri[[0, 2, 1]], tri * [1, 1, 200]]),          # back-facing + beyond far
                      np.stack([np.zeros((3, 3), np.float32), np.full((3, 3), np.nan, np.float32)])])  # degenerate / NaN
    cases = cases.astype(np.float32)
    fim_o, wim_o = raster.rasterize_fim_wim(cases, S, fast=False)
    fim, wim = ops.rasterize_faces(_t(cases), S)
    np.testing.assert_array_equal(fim.cpu().numpy(), fim_o)
    np.testing.assert_array_equal(wim.cpu().numpy(), wim_o)


@pytest.mark.parametrize("S", [64, 128])
def test_fused_frame_inputs_match_reference_golden(S, template, golden_dir):
    """tsf_inputs / Tst from the fused kernel vs the fixtures produced by the reference's own FlowComposition code."""
    from ipercore_b200 import ops
    g = np.load(os.path.join(golden_dir, "flow_S%d.npz" % S))
    n = g["fim"].shape[0]
    cams, verts = synth.pose_sweep(template, n, total=7)
    uv_img = synth.smooth_image((1, 3, S, S), seed=11)
    fused = dict(map_fn=_t(template["map_fn"]), f_uvs2img=_t(template["f_uvs2img"]), uv_img=_t(uv_img[0]),
                 src_f2pts=_t(g["src_f2pts"]))
    out = ops.raster_frames(_t(verts), _t(cams), _t(template["faces"]), S, fused=fused)
    np.testing.assert_array_equal(out["fim"].cpu().numpy(), g["fim"])
    np.testing.assert_array_equal(out["tsf_inputs"][:, 3:].cpu().numpy(), g["cond"])
    np.testing.assert_allclose(out["Tst"].cpu().numpy(), g["Tst"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(out["tsf_inputs"].cpu().numpy(), g["tsf_inputs"], atol=1e-5, rtol=0)
    # the stand-alone seam ops agree with the fused kernel
    T = ops.cal_bc_transform(_t(g["src_f2pts"]), out["fim"][:1].repeat(2, 1, 1), out["wim"][:1].repeat(2, 1, 1, 1))
    np.testing.assert_array_equal(T.cpu().numpy(), out["Tst"][0].cpu().numpy())
    cond = ops.encode_fim(out["fim"], _t(template["map_fn"]))
    np.testing.assert_array_equal(cond.cpu().numpy(), g["cond"])


def test_flow_resize_and_warp_match_torch(template):
    from ipercore_b200 import ops
    import torch.nn.functional as F
    S = 128
    cams, verts = synth.pose_sweep(template, 2, total=7)
    scams, sverts = synth.source_views(template, 2)
    src_f2pts, _, _ = flow_ref.render_fim_wim(scams, sverts, template["faces"], S)
    _, fim, wim = flow_ref.render_fim_wim(cams, verts, template["faces"], S)
    Tst = np.stack([flow_ref.make_trans_flow(src_f2pts, fim[i], wim[i]) for i in range(2)])      # (2,2,S,S,2)
    for h in (64, 32, 16):
        ref = F.interpolate(torch.from_numpy(Tst).reshape(4, S, S, 2).permute(0, 3, 1, 2), size=(h, h), mode="bilinear",
                            align_corners=True).permute(0, 2, 3, 1).reshape(2, 2, h, h, 2)
        got = ops.flow_resize(_t(Tst), h, h)
        np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), atol=2e-6, rtol=0)
        feat = synth.uniform((2, 32, h, h), seed=5)                                                # (ns,C,h,w)
        exp = torch.stack([F.grid_sample(torch.from_numpy(feat), ref[b], mode="bilinear", padding_mode="zeros",
                                         align_corners=False) for b in range(2)])                  # (B,ns,C,h,w)
        w = ops.warp_nhwc(_t(feat).permute(0, 2, 3, 1).contiguous(), got)
        np.testing.assert_allclose(w.permute(0, 1, 4, 2, 3).cpu().numpy(), exp.numpy(), atol=2e-5, rtol=0)


def test_raster_workspace_too_small_fails_loudly():
    """The caller owns the rasteriser scratch; an undersized one is an error, never a silent overrun."""
    from ipercore_b200 import _lib
    faces = torch.zeros((1, 4, 3, 3), device="cuda")
    fim = torch.empty((1, 32, 32), dtype=torch.int32, device="cuda")
    wim = torch.empty((1, 32, 32, 3), device="cuda")
    ws = torch.empty(8, dtype=torch.uint8, device="cuda")
    assert _lib.lib.iper_raster_workspace_bytes(1, 4, 0) == 16 and _lib.lib.iper_raster_workspace_bytes(2, 4, 1) == 32 + 288
    with pytest.raises(RuntimeError, match="workspace"):
        _lib.check(_lib.lib.iper_rasterize_faces(faces.data_ptr(), 1, 4, 32, 0.1, 100.0, fim.data_ptr(), wim.data_ptr(),
                                                 ws.data_ptr(), 8, None), "rasterize_faces")


@pytest.mark.parametrize("ks", [3, 11, 51])
def test_morph_matches_reference_golden(ks):
    """iper_morph vs the reference's morph()/soft_dilate() outputs (tests/golden/morph.npz) and the oracle: bit-exact."""
    from ipercore_b200 import ops
    from oracle import morph_ref
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    gold = np.load(os.path.join(here, "morph.npz"))
    m = synth.morph_masks()          # the same seeded masks the golden script fed to the reference's functions
    md = torch.from_numpy(m).cuda()
    for name, mode in (("erode", ops.MORPH_ERODE), ("dilate", ops.MORPH_DILATE), ("soft", ops.MORPH_SOFT_DILATE)):
        got = ops.morph(md, ks, mode).cpu().numpy()
        assert np.array_equal(got, gold["%s_%d" % (name, ks)].astype(np.float32)), (name, ks)
        assert np.array_equal(got, morph_ref.morph(m, ks, mode))
    with pytest.raises(RuntimeError, match="odd"):
        ops.morph(md, 4, ops.MORPH_ERODE)
