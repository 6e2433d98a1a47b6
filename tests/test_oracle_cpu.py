"""CPU tests: the oracle restatements against the golden fixtures produced by the REFERENCE's own code
(tests/golden/make_golden.py), and internal consistency of the rasteriser restatement (parity unpinned)."""
import os

import numpy as np
import pytest
import torch

from oracle import flow_ref, generator_ref, raster, synth, weights


def test_raster_fast_equals_definition(template):
    cams, verts = synth.pose_sweep(template, 2, total=7)
    fv = flow_ref.vertices_to_faces(flow_ref.project(cams, verts), template["faces"])
    fim_a, wim_a = raster.rasterize_fim_wim(fv, 48, fast=False)
    fim_b, wim_b = raster.rasterize_fim_wim(fv, 48, fast=True)
    assert (fim_a >= 0).sum() > 100
    np.testing.assert_array_equal(fim_a, fim_b)
    np.testing.assert_array_equal(wim_a, wim_b)


def test_raster_known_answers():
    """Hand-checkable cases of the upstream rules: winding, centre sampling, tie -> lowest index, near/far."""
    S = 4
    tri = lambda z, flip=False: np.array([[-1, -1, z], [-1, 3, z], [3, -1, z]] if not flip else
                                         [[-1, -1, z], [3, -1, z], [-1, 3, z]], np.float32)
    # CCW-in-NDC big triangle covers the whole image when front facing under the upstream backface rule
    for flip in (False, True):
        fim, wim = raster.rasterize_fim_wim(tri(1.0, flip)[None, None], S, fast=False)
        if (fim >= 0).all():
            front = flip
    fim, wim = raster.rasterize_fim_wim(tri(1.0, front)[None, None], S, fast=False)
    assert (fim == 0).all()
    np.testing.assert_allclose(wim.sum(-1), 1.0, atol=1e-6)
    back, _ = raster.rasterize_fim_wim(tri(1.0, not front)[None, None], S, fast=False)
    assert (back == -1).all(), "back-facing triangle must be culled"
    # two coincident faces: strict '<' keeps the lower index; a nearer face wins regardless of order
    two = np.stack([tri(1.0, front), tri(1.0, front)])[None]
    assert (raster.rasterize_fim_wim(two, S, fast=False)[0] == 0).all()
    near_first = np.stack([tri(2.0, front), tri(1.0, front)])[None]
    assert (raster.rasterize_fim_wim(near_first, S, fast=False)[0] == 1).all()
    # near / far rejection (0.1, 100)
    assert (raster.rasterize_fim_wim(tri(0.05, front)[None, None], S, fast=False)[0] == -1).all()
    assert (raster.rasterize_fim_wim(tri(150.0, front)[None, None], S, fast=False)[0] == -1).all()
    # vertical flip: a triangle in the upper NDC half (y>0) lands in the TOP image rows
    up = np.array([[-1, 0.01, 1], [-1, 3, 1], [3, 0.01, 1]], np.float32)
    up = up if not front else up[[0, 2, 1]]
    f, _ = raster.rasterize_fim_wim(up[None, None], S, fast=False)
    assert (f[0, :2] == 0).all() and (f[0, 2:] == -1).all()


def test_raster_empty_and_ragged():
    fim, wim = raster.rasterize_fim_wim(np.zeros((0, 5, 3, 3), np.float32), 8)
    assert fim.shape == (0, 8, 8) and wim.shape == (0, 8, 8, 3)
    # degenerate (zero-area) and NaN faces never win
    deg = np.array([[[0, 0, 1], [0, 0, 1], [0, 0, 1]], [[np.nan, 0, 1], [0, 1, 1], [1, 0, 1]]], np.float32)[None]
    fim, wim = raster.rasterize_fim_wim(deg, 8, fast=False)
    fim2, _ = raster.rasterize_fim_wim(deg, 8, fast=True)
    assert (fim == -1).all() and (fim2 == -1).all() and (wim == 0).all()


@pytest.mark.parametrize("S", [64, 128])
def test_flow_restatement_matches_reference(S, template, golden_dir):
    g = np.load(os.path.join(golden_dir, "flow_S%d.npz" % S))
    n = g["fim"].shape[0]
    cams, verts = synth.pose_sweep(template, n, total=7)
    scams, sverts = synth.source_views(template, 2)
    uv_img = synth.smooth_image((1, 3, S, S), seed=11)
    src_f2pts, _, _ = flow_ref.render_fim_wim(scams, sverts, template["faces"], S)
    out = flow_ref.frame_inputs(cams, verts, template["faces"], template["map_fn"], template["f_uvs2img"], uv_img,
                                src_f2pts, S)
    np.testing.assert_array_equal(out["fim"], g["fim"])
    np.testing.assert_array_equal(out["wim"], g["wim"])
    np.testing.assert_array_equal(out["f2pts"], g["f2pts"])         # projection glue is bit-exact
    np.testing.assert_array_equal(src_f2pts, g["src_f2pts"])
    np.testing.assert_array_equal(out["cond"], g["cond"])
    np.testing.assert_allclose(out["Tuv2t"], g["Tuv2t"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(out["Tst"], g["Tst"], atol=1e-6, rtol=0)
    np.testing.assert_allclose(out["tsf_inputs"], g["tsf_inputs"], atol=1e-5, rtol=0)


@pytest.mark.parametrize("S", [64, 256])
def test_generator_restatement_matches_reference(S, golden_dir):
    import make_golden
    g = np.load(os.path.join(golden_dir, "gen_S%d.npz" % S))
    inp = make_golden.gen_inputs(S)
    sd = weights.synth_state_dict(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        se, sr = generator_ref.forward_src(sd, torch.from_numpy(inp["src_inputs"]))
        img, mask = generator_ref.forward_tsf(sd, torch.from_numpy(inp["tsf_inputs"]), se, sr,
                                              torch.from_numpy(inp["Tst"]))
    np.testing.assert_allclose(se[2].numpy(), g["src_enc2"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(sr[5].numpy(), g["src_res5"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(img.numpy(), g["tsf_img"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(mask.numpy(), g["tsf_mask"], atol=2e-5, rtol=0)
    if "bg_img" in g.files:
        bg = generator_ref.forward_bg(sd, torch.from_numpy(inp["bg_inputs"]))
        np.testing.assert_allclose(bg.numpy(), g["bg_img"], atol=2e-5, rtol=0)


def test_state_dict_layout():
    shapes = weights.generator_param_shapes()
    assert len(shapes) == 221
    assert sum(int(np.prod(s)) for _, s in shapes) == 36276992


def test_lbs_restatement_matches_reference(template, golden_dir):
    """oracle/lbs_ref.py against verts/joints produced by the reference's own lbs() + link() (tests/golden/lbs.npz)."""
    import make_golden
    from oracle import lbs_ref
    g = np.load(os.path.join(golden_dir, "lbs.npz"))
    m = lbs_ref.synthetic_smplh(template=template["verts"])
    betas, pose, links, offsets = make_golden.lbs_inputs()
    v, j = lbs_ref.lbs(m, betas, pose, offsets=offsets, links=links)
    np.testing.assert_allclose(v, g["verts"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(j, g["joints"], atol=2e-6, rtol=0)
    # 72-dim body pose path pads the hands mean (batch_smplh.py:158-160)
    v72, _ = lbs_ref.smplh_forward(m, betas, pose[:, :72])
    full = np.concatenate([pose[:, :66], np.repeat(m["hands_mean"][None], 3, 0)], 1)
    np.testing.assert_array_equal(v72, lbs_ref.lbs(m, betas, full)[0])


def test_morph_oracle_matches_reference_golden():
    """oracle/morph_ref.py vs the outputs of the reference's morph()/soft_dilate() (tests/golden/morph.npz)."""
    import os
    import numpy as np
    from oracle import morph_ref
    from oracle import synth
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    m = synth.morph_masks()
    gold = np.load(os.path.join(here, "morph.npz"))
    for ks in (3, 11, 51):
        for name, mode in (("erode", morph_ref.ERODE), ("dilate", morph_ref.DILATE), ("soft", morph_ref.SOFT_DILATE)):
            assert np.array_equal(morph_ref.morph(m, ks, mode), gold["%s_%d" % (name, ks)].astype(np.float32)), (name, ks)


def test_raster_contraction_variant(template):
    """The -DORACLE_FMA variant (nvcc -fmad=true model of the upstream source, oracle/raster_ref.c header): same rules on the
    known answers, fast == definition, and on a real mesh it may differ from the separately-rounded variant only on pixels
    that sit on a face edge / depth tie (a handful), with weights equal to float rounding elsewhere."""
    S = 4
    tri = np.array([[-1, -1, 1], [3, -1, 1], [-1, 3, 1]], np.float32)
    for t in (tri, tri[[0, 2, 1]]):
        a, _ = raster.rasterize_fim_wim(t[None, None], S, fast=False)
        b, _ = raster.rasterize_fim_wim(t[None, None], S, fast=False, fma=True)
        np.testing.assert_array_equal(a, b)
    cams, verts = synth.pose_sweep(template, 2, total=7)
    fv = flow_ref.vertices_to_faces(flow_ref.project(cams, verts), template["faces"])
    fim_d, wim_d = raster.rasterize_fim_wim(fv, 48, fast=False, fma=True)
    fim_f, wim_f = raster.rasterize_fim_wim(fv, 48, fast=True, fma=True)
    np.testing.assert_array_equal(fim_d, fim_f)
    np.testing.assert_array_equal(wim_d, wim_f)
    fim0, wim0 = raster.rasterize_fim_wim(fv, 128)
    fim1, wim1 = raster.rasterize_fim_wim(fv, 128, fma=True)
    differ = (fim0 != fim1)
    assert differ.mean() < 1e-3, "contraction may only flip pixels that sit exactly on an edge or a depth tie"
    # upstream's pixel-space adjugate cancels catastrophically on small faces, so the WEIGHTS are sensitive to the rounding
    # model far above float epsilon (measured: median 8e-5, p99 2.5e-3, max 1e-1 at 512^2) — the reason the switch exists
    d = np.abs(wim0 - wim1).max(-1)[(fim0 >= 0) & ~differ]
    assert np.median(d) < 1e-3 and np.percentile(d, 99) < 2e-2
    assert (wim0 != wim1).any(), "the two rounding models must not be the same build"


def test_source_stage_restatement_matches_reference_golden(golden_dir):
    """oracle/source_ref.py vs outputs of the reference's own make_morph_image / make_uv_img / CannyFilter captured in a real
    Imitator.source_setup run (tests/golden/source_S96.npz)."""
    import importlib.util
    from oracle import source_ref
    g = np.load(os.path.join(golden_dir, "source_S96.npz"))
    # filter taps: the product's constant table (no CUDA needed to read it)
    src = open(os.path.join(os.path.dirname(golden_dir), "..", "ipercore_b200", "source_ops.py")).read()
    ns = {}
    exec("import numpy as np\n" + src[src.index("def canny_constants"):src.index("_CANNY = None")], ns)
    consts = ns["canny_constants"]()
    conf, outp, edges = (g[k].astype(np.float32) for k in ("confidant_sil", "outpad_sil", "thin_edges"))
    e = source_ref.canny_edges(conf[:, 0], consts)
    diff = e != edges[:, 0]
    assert diff.sum() <= 0.12 * edges.sum()          # NMS on exact float ties: the reference differs from itself by this much
    for i in range(conf.shape[0]):
        m = source_ref.morph_image(g["src_img"][i], conf[i, 0], outp[i, 0], edges[i, 0])
        d = np.abs(m - g["morph_img"][i]).max(0)
        unc = (outp[i, 0] * (1 - conf[i, 0])) != 0
        assert d[~unc].max() == 0 and d[g["unique3"][i, 0]].max() <= 1e-6
    uv = source_ref.make_uv_img(g["morph_img"], g["obj_f2pts"], g["only_vis_obj_f2pts"], g["uv_fim"], g["uv_wim"])
    assert np.abs(uv - g["uv_img"][0]).max() <= 1e-5


def test_generator_restatement_temporal_branch(golden_dir):
    """oracle/generator_ref.py's temporal attention branch vs the reference generator built with temporal=True."""
    import make_golden
    S = 64
    g = np.load(os.path.join(golden_dir, "gen_S%d_temporal.npz" % S))
    sd = weights.synth_state_dict(0)
    inp = make_golden.gen_inputs(S)
    prev, Ttt = make_golden.temporal_inputs(S)
    with torch.no_grad():
        se, sr = generator_ref.forward_src(sd, torch.from_numpy(inp["src_inputs"]))
        te, tr = generator_ref.forward_src(sd, torch.from_numpy(prev))
        img, mask = generator_ref.forward_tsf(sd, torch.from_numpy(inp["tsf_inputs"]), se, sr, torch.from_numpy(inp["Tst"]),
                                              temp_enc_outs=te, temp_res_outs=tr, Ttt=torch.from_numpy(Ttt))
    assert np.abs(img.numpy() - g["tsf_img"]).max() <= 1e-5 and np.abs(mask.numpy() - g["tsf_mask"]).max() <= 1e-5
