"""The stand-alone C generator handle (include/iper_b200.h iper_gen_*, csrc/generator.cu) driven through ctypes with raw
buffers only: same golden fixtures of the REFERENCE's own AttentionLWBGenerator as the Python-driven path, and bit-identical
to that path (same kernels, same packing) — i.e. the layer graph and the weight repacking really live in C."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
CFG = dict(name="AttLWB-SPADE", BGNet=dict(cond_nc=4, n_res_block=6, num_filters=[64, 128, 128, 256]),
           SIDNet=dict(cond_nc=6, n_res_block=6, num_filters=[64, 128, 256]),
           TSFNet=dict(cond_nc=6, n_res_block=6, num_filters=[64, 128, 256]))


@pytest.mark.parametrize("S,precision,tol", [(256, "fp16x2", 1e-3), (64, "fp16x2", 1e-3), (256, "fp16", 1e-2)])
def test_c_generator_matches_reference_golden(S, precision, tol, golden_dir):
    import make_golden
    from ipercore_b200.cgen import GeneratorHandle
    from ipercore_b200.generator import AttentionLWBGenerator
    from oracle.weights import synth_state_dict
    g = np.load(os.path.join(golden_dir, "gen_S%d.npz" % S))
    inp = {k: torch.from_numpy(v).to("cuda:0") for k, v in make_golden.gen_inputs(S).items()}
    sd = synth_state_dict(0)
    sd["module.bg_net.main.0.bias"] = sd.pop("bg_net.main.0.bias")         # a DDP-prefixed key loads like base_model.py:56-65
    h = GeneratorHandle(sd, precision=precision)
    h.forward_src(inp["src_inputs"])
    bg = torch.rand(1, 3, S, S, device="cuda:0") * 2 - 1
    img, mask, pred = h.forward_tsf(inp["tsf_inputs"], inp["Tst"], bg_img=bg, return_pred=True)
    torch.cuda.synchronize()
    e_img = np.abs(img.cpu().numpy() - g["tsf_img"]).max(); e_mask = np.abs(mask.cpu().numpy() - g["tsf_mask"]).max()
    print("C handle S=%d %s: tsf_img %.2e tsf_mask %.2e" % (S, precision, e_img, e_mask))
    assert e_img <= tol and e_mask <= tol
    torch.testing.assert_close(pred, mask * bg + (1 - mask) * img, atol=1e-6, rtol=0)
    # the Python-driven path on the same kernels
    net = AttentionLWBGenerator(CFG, precision=precision); net.load_state_dict(synth_state_dict(0)); net = net.to("cuda:0")
    enc, res = net.forward_src(inp["src_inputs"])
    img2, mask2 = net.forward_tsf(inp["tsf_inputs"], enc, res, inp["Tst"])
    d = max(float((img - img2).abs().max()), float((mask - mask2).abs().max()))
    print("C handle vs Python-driven graph: max diff %.2e" % d)
    assert d <= 2e-6       # identical kernels and packing; fp64 atomics of the statistics may differ in the last bits


def test_c_generator_batch_and_errors(golden_dir):
    import make_golden
    from ipercore_b200 import _lib
    from ipercore_b200.cgen import GeneratorHandle
    from oracle.weights import synth_state_dict
    S = 128
    inp = {k: torch.from_numpy(v).to("cuda:0") for k, v in make_golden.gen_inputs(S).items()}
    h = GeneratorHandle(synth_state_dict(0))
    h.forward_src(inp["src_inputs"])
    tsf = torch.cat([inp["tsf_inputs"], inp["tsf_inputs"].flip(-1), inp["tsf_inputs"] * 0.5], 0)
    Tst = torch.cat([inp["Tst"], inp["Tst"].flip(2), inp["Tst"]], 0).contiguous()
    img, mask = h.forward_tsf(tsf, Tst)
    for i in range(3):
        i1, m1 = h.forward_tsf(tsf[i:i + 1].contiguous(), Tst[i:i + 1].contiguous())
        assert float((i1[0] - img[i]).abs().max()) <= 1e-5 and float((m1[0] - mask[i]).abs().max()) <= 1e-5
    # an undersized workspace is an error, not an overrun
    small = torch.empty(1024, dtype=torch.uint8, device="cuda:0")
    with pytest.raises(RuntimeError, match="workspace"):
        h.forward_tsf(tsf, Tst, workspace=small)
    # a missing tensor is reported by name
    sd = synth_state_dict(0); sd.pop("res_blocks.3.main.2.weight")
    with pytest.raises(RuntimeError, match="res_blocks.3.main.2.weight"):
        GeneratorHandle(sd)
