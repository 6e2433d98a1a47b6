"""Training-step kernels (csrc/train.cu, ipercore_b200/train.py): the bf16 tcgen05 3x3 convolution — forward, data gradient,
weight gradient — against torch autograd in fp32 on the same bf16-rounded operands, and one whole G + D optimisation step
against the all-torch formulation of the same step."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CFG = dict(name="AttLWB-SPADE", BGNet=dict(cond_nc=4, n_res_block=6, num_filters=[64, 128, 128, 256]),
           SIDNet=dict(cond_nc=6, n_res_block=6, num_filters=[64, 128, 256]),
           TSFNet=dict(cond_nc=6, n_res_block=6, num_filters=[64, 128, 256]))


@pytest.mark.parametrize("N,ci,co,H,W", [(2, 128, 256, 64, 64), (1, 64, 128, 64, 64), (2, 128, 64, 16, 64), (1, 192, 128, 8, 128),
                                         (1, 256, 256, 64, 64)])
def test_conv3x3_bf16_forward_dgrad_wgrad(N, ci, co, H, W):
    from ipercore_b200 import train
    g = torch.Generator(device="cpu").manual_seed(ci * 7 + co)
    x = (torch.randn(N, ci, H, W, generator=g) * 0.5).to(DEV).bfloat16().float().requires_grad_(True)
    w = (torch.randn(co, ci, 3, 3, generator=g) * (1.0 / np.sqrt(9 * ci))).to(DEV).bfloat16().float().requires_grad_(True)
    b = (torch.randn(co, generator=g) * 0.1).to(DEV).requires_grad_(True)
    dy = (torch.randn(N, co, H, W, generator=g) * 0.5).to(DEV).bfloat16().float()
    y_ref = F.conv2d(x, w, b, padding=1)
    gx_ref, gw_ref, gb_ref = torch.autograd.grad(y_ref, (x, w, b), dy)
    assert train._eligible(x, w)
    x2, w2, b2 = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    y = train.conv3x3(x2, w2, b2)
    assert y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last)
    gx, gw, gb = torch.autograd.grad(y, (x2, w2, b2), dy.bfloat16())
    s = float(y_ref.abs().max())
    e_y = float((y.float() - y_ref).abs().max()) / s                 # bf16 output rounding: 2^-9 relative to the value
    e_x = float((gx.float() - gx_ref).abs().max()) / float(gx_ref.abs().max())
    e_w = float((gw.float() - gw_ref).abs().max()) / float(gw_ref.abs().max())
    e_b = float((gb.float() - gb_ref).abs().max()) / float(gb_ref.abs().max())
    print("conv3x3 bf16 %dx%d->%d %dx%d: rel err y %.2e dx %.2e dw %.2e db %.2e" % (N, ci, co, H, W, e_y, e_x, e_w, e_b))
    assert e_y <= 6e-3 and e_x <= 6e-3            # outputs are stored in bf16
    assert e_w <= 2e-4 and e_b <= 1e-5            # fp32 accumulation over all pixels, fp32 result


def test_conv3x3_falls_back_when_not_eligible():
    from ipercore_b200 import train
    x = torch.randn(1, 6, 32, 32, device=DEV).bfloat16()
    w = torch.randn(64, 6, 3, 3, device=DEV, requires_grad=True)
    assert not train._eligible(x, w)
    y = train.conv3x3(x, w)             # 6 input channels: the fp32 torch path (no bf16 cuDNN engine for its gradients)
    assert y.dtype == torch.float32
    torch.testing.assert_close(y, F.conv2d(x.float(), w, padding=1), atol=1e-4, rtol=1e-4)
    frozen = torch.randn(64, 64, 3, 3, device=DEV)                       # frozen layer (VGG): no weight gradient -> eligible at W = 32
    assert train._eligible(torch.randn(1, 64, 32, 32, device=DEV), frozen)


def _batch(S, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: (torch.rand(*s, generator=g) * 2 - 1).to(DEV)
    flow = r(1, 1, 2, S, S, 2)
    return dict(bg_inputs=torch.cat([r(1, 1, 3, S, S), (r(1, 1, 1, S, S) > 0).float()], 2), src_inputs=r(1, 2, 6, S, S),
                tsf_inputs=r(1, 1, 6, S, S), Tst=flow, real_src=r(1, 2, 3, S, S), real_tsf=r(1, 1, 3, S, S), real_bg=r(1, 3, S, S),
                body_mask=(r(1, 3, 1, S, S) > 0).float())


def test_train_step_matches_all_torch_formulation():
    """One G + D step at 512x512 (the size at which the 256^2/128^2/64^2 feature maps make every 3x3 layer eligible): losses and
    a sample of gradients with the tcgen05 convs vs the same step with F.conv2d everywhere (both bf16)."""
    from ipercore_b200 import train
    from ipercore_b200.generator import AttentionLWBGenerator
    from oracle.weights import synth_state_dict
    S = 512
    res = {}
    for use in (True, False):
        train.USE_KERNELS = use
        try:
            torch.manual_seed(0)
            net = AttentionLWBGenerator(CFG); net.load_state_dict(synth_state_dict(0))
            step = train.LWGTrainStep(net, torch.device(DEV))
            fake = step.G(**{k: v for k, v in _batch(S).items() if k in ("bg_inputs", "src_inputs", "tsf_inputs", "Tst")})
            loss = sum(t.float().mean() for t in fake[1:])
            names = ["net.res_blocks.2.main.0.weight", "net.res_attlwbs.1.spade.mlp_gamma.weight", "net.tsf_net_dec.skippers.1.0.weight",
                     "net.src_net.res_blocks.0.main.2.weight", "net.enc_attlwbs.0.spade.mlp_shared.0.weight"]
            params = dict(step.G.named_parameters())
            grads = torch.autograd.grad(loss, [params[n] for n in names])
            out = step.step(_batch(S))
            torch.cuda.synchronize()
            res[use] = (float(loss), [g.float().clone() for g in grads], {k: float(v) for k, v in out.items()})
        finally:
            train.USE_KERNELS = True
    (l1, g1, o1), (l0, g0, o0) = res[True], res[False]
    print("train step: probe loss %.5f vs %.5f; step losses %s vs %s" % (l1, l0, o1, o0))
    assert abs(l1 - l0) <= 2e-2 * max(1.0, abs(l0))
    for a, b in zip(g1, g0):
        cos = float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
        assert cos >= 0.95, cos                                   # two bf16 evaluation orders of a 70-layer network (measured 0.976-0.999)
    for k in o0:
        assert np.isfinite(o1[k]) and abs(o1[k] - o0[k]) <= 5e-2 * max(1.0, abs(o0[k])), (k, o1[k], o0[k])
