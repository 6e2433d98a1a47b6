"""Training-step kernels (csrc/train.cu, ipercore_b200/train.py): the bf16 tcgen05 stride-1 convolutions (1x1, 3x3, 5x5, 7x7, incl. the
zero-padded 1/3/4/6-channel ends) — forward with fused bias / residual / ReLU, data gradient, weight gradient (MN-major NHWC
operands), bias gradient — against torch autograd in fp32 on the same bf16-rounded operands; the fused Adam + repack pass against
torch.optim.Adam; one whole G + D optimisation step against the all-torch formulation, eager and as a CUDA graph."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.float() - b.float()).abs().max()) / max(float(b.float().abs().max()), 1e-20)
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


@pytest.mark.parametrize("N,ci,co,H,W,k,relu,add", [(1, 64, 64, 32, 48, 1, False, False), (2, 128, 256, 24, 40, 1, True, False),
                                                    (1, 64, 3, 64, 64, 5, False, False), (1, 64, 1, 40, 72, 5, False, False),
                                                    (1, 4, 64, 32, 32, 7, False, False), (1, 64, 3, 32, 64, 7, False, False),
                                                    (1, 6, 64, 16, 16, 3, True, False), (2, 256, 256, 16, 32, 3, False, True),
                                                    (1, 192, 128, 20, 36, 3, True, True)])
def test_conv_kxk_padded_ends_fused_epilogue(N, ci, co, H, W, k, relu, add):
    """Every kernel size of the step, ragged maps (tiles hanging over the border), zero-padded channel ends, fused residual / ReLU."""
    from ipercore_b200 import train
    g = torch.Generator(device="cpu").manual_seed(ci * 7 + co + k)
    rnd = lambda *s: torch.randn(*s, generator=g)
    x = (rnd(N, ci, H, W) * 0.5).to(DEV).bfloat16().float().requires_grad_(True)
    w = (rnd(co, ci, k, k) * (1.0 / np.sqrt(k * k * ci))).to(DEV).bfloat16().float().requires_grad_(True)
    b = (rnd(co) * 0.1).to(DEV).requires_grad_(True)
    r = (rnd(N, co, H, W) * 0.5).to(DEV).bfloat16().float().requires_grad_(True) if add else None
    dy = (rnd(N, co, H, W) * 0.5).to(DEV).bfloat16().float()
    y_ref = F.conv2d(x, w, b, padding=k // 2)
    if add:
        y_ref = y_ref + r
    if relu:
        y_ref = F.relu(y_ref)
    ins = (x, w, b) + ((r,) if add else ())
    g_ref = torch.autograd.grad(y_ref, ins, dy)
    assert train._eligible(x, w)
    ins2 = tuple(t.detach().clone().requires_grad_(True) for t in ins)
    y = train.conv(ins2[0], ins2[1], ins2[2], relu=relu, add=ins2[3] if add else None)
    assert y.dtype == torch.bfloat16 and y.shape == y_ref.shape
    g_got = torch.autograd.grad(y, ins2, dy.bfloat16())
    rel = lambda a, b_: float((a.float() - b_).abs().max()) / max(float(b_.abs().max()), 1e-20)
    errs = [rel(y, y_ref.detach())] + [rel(a, b_) for a, b_ in zip(g_got, g_ref)]
    print("conv %dx%d %d->%d %dx%d relu %d add %d: rel err y %.2e dx %.2e dw %.2e db %.2e" % ((k, k, ci, co, H, W, relu, add) + tuple(errs[:4])))
    assert errs[0] <= 6e-3 and errs[1] <= 6e-3              # bf16 outputs
    # dY passes through the ReLU mask / bf16 rounding of y only when relu is fused: the mask is taken from the bf16 y (ties at 0)
    assert errs[2] <= (2e-3 if relu else 2e-4) and errs[3] <= (2e-3 if relu else 1e-5)
    if add:
        assert errs[4] <= 6e-3


@pytest.mark.parametrize("mode,N,ci,co,H,W,k", [("s2", 1, 64, 128, 32, 64, 3), ("s2", 2, 6, 64, 48, 64, 3), ("s2", 1, 128, 256, 16, 32, 3),
                                                ("s2", 1, 6, 64, 32, 64, 4), ("s2", 1, 128, 256, 32, 32, 4),
                                                ("ct", 1, 256, 128, 16, 16, 4), ("ct", 2, 128, 64, 8, 24, 4), ("ct", 1, 64, 64, 24, 40, 4),
                                                ("k4", 1, 256, 256, 31, 31, 4), ("k4", 1, 128, 1, 20, 24, 4)])
def test_strided_transposed_and_non_same_convolutions(mode, N, ci, co, H, W, k):
    """Stride-2 convolutions (5-D parity view), ConvTranspose2d(4,2,1) (four phase convolutions) and the discriminator's 4x4 stride-1
    pad-1 layers: forward, data gradient, weight gradient, bias gradient against torch autograd in fp32."""
    from ipercore_b200 import train
    g = torch.Generator(device="cpu").manual_seed(ci * 3 + co + H)
    rnd = lambda *s: torch.randn(*s, generator=g)
    x = (rnd(N, ci, H, W) * 0.5).to(DEV).bfloat16().float().requires_grad_(True)
    wshape = (ci, co, k, k) if mode == "ct" else (co, ci, k, k)
    w = (rnd(*wshape) * (1.0 / np.sqrt(k * k * ci))).to(DEV).bfloat16().float().requires_grad_(True)
    b = (rnd(co) * 0.1).to(DEV).requires_grad_(True)
    kw = dict(stride=2, padding=1, transposed=True) if mode == "ct" else dict(stride=2 if mode == "s2" else 1, padding=1)
    y_ref = F.conv_transpose2d(x, w, b, stride=2, padding=1) if mode == "ct" else F.conv2d(x, w, b, stride=kw["stride"], padding=1)
    dy = (rnd(*y_ref.shape) * 0.5).to(DEV).bfloat16().float()
    g_ref = torch.autograd.grad(y_ref, (x, w, b), dy)
    assert train._kind(x, w, **kw) == {"s2": train.S2, "ct": train.CT, "k4": train.S1}[mode]
    x2, w2, b2 = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    y = train.conv(x2, w2, b2, **kw)
    assert y.dtype == torch.bfloat16 and y.shape == y_ref.shape
    g_got = torch.autograd.grad(y, (x2, w2, b2), dy.bfloat16())
    errs = [_rel(y, y_ref.detach())] + [_rel(a, b_) for a, b_ in zip(g_got, g_ref)]
    print("%s %dx%d %d->%d %dx%d: rel err y %.2e dx %.2e dw %.2e db %.2e" % ((mode, k, k, ci, co, H, W) + tuple(errs)))
    assert errs[0] <= 6e-3 and errs[1] <= 6e-3 and errs[2] <= 2e-4 and errs[3] <= 1e-5


def test_pad_to_channels_last_kernel():
    """The fused cast + zero-pad + channels_last copy of the tiny-channel ends: fp32 NCHW input, a bf16 channel slice of a
    channels_last tensor (the layout autograd hands the heads' gradients back in), and its backward."""
    from ipercore_b200 import train
    x = torch.randn(2, 6, 24, 40, device=DEV, requires_grad=True)
    y = train._to_cl(x, 64)
    ref = F.pad(x.detach().bfloat16(), (0, 0, 0, 0, 0, 58))
    assert y.dtype == torch.bfloat16 and y.is_contiguous(memory_format=torch.channels_last) and torch.equal(y, ref)
    g = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, g)
    assert gx.dtype == torch.float32 and torch.equal(gx, g[:, :6].float())
    big = torch.randn(1, 64, 16, 32, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last)
    sl = big[:, 5:8]                                                   # strided view: 3 of 64 channels
    assert torch.equal(train._to_cl(sl, 64), F.pad(sl, (0, 0, 0, 0, 0, 61)))
    assert torch.equal(train._to_cl(big, 64), big)                    # nothing to pad: no copy semantics change


@pytest.mark.parametrize("ci,co,H,W,k", [(64, 3, 40, 48, 5), (64, 1, 24, 40, 5), (4, 64, 32, 32, 7), (64, 3, 16, 32, 7)])
def test_thin_end_weight_gradient_kernel(ci, co, H, W, k):
    """The opt-in CUDA-core weight gradient of the layers with <= 4 channels on one side (train.USE_THIN) against torch autograd."""
    from ipercore_b200 import train
    g = torch.Generator(device="cpu").manual_seed(ci + co + k)
    x = (torch.randn(2, ci, H, W, generator=g) * 0.5).to(DEV).bfloat16().float().requires_grad_(True)
    w = (torch.randn(co, ci, k, k, generator=g) * (1.0 / np.sqrt(k * k * ci))).to(DEV).bfloat16().float().requires_grad_(True)
    dy = (torch.randn(2, co, H, W, generator=g) * 0.5).to(DEV).bfloat16().float()
    (gw_ref,) = torch.autograd.grad(F.conv2d(x, w, padding=k // 2), w, dy)
    w2 = w.detach().clone().requires_grad_(True)
    train.USE_THIN = True
    try:
        (gw,) = torch.autograd.grad(train.conv(x.detach(), w2), w2, dy.bfloat16())
    finally:
        train.USE_THIN = False
    assert _rel(gw, gw_ref) <= 2e-4


def test_conv_falls_back_when_not_eligible():
    from ipercore_b200 import train
    x = torch.randn(1, 64, 4, 8, device=DEV).bfloat16()               # map smaller than one 16x8 tile
    w = torch.randn(64, 64, 3, 3, device=DEV, requires_grad=True)
    assert not train._eligible(x, w)
    y = train.conv(x, w)
    torch.testing.assert_close(y.float(), F.conv2d(x.float(), w.bfloat16().float(), padding=1), atol=5e-2, rtol=5e-2)
    w4 = torch.randn(64, 64, 4, 4, device=DEV)                        # the discriminator's 4x4 / pad 1: output 15x15 is below one box
    assert not train._eligible(torch.randn(1, 64, 16, 16, device=DEV), w4, 1, 1)
    assert not train._eligible(torch.randn(1, 64, 16, 16, device=DEV), w, 2, 1)      # strided: the half-resolution map is below one box
    assert not train._eligible(torch.randn(1, 64, 32, 32, device=DEV), torch.randn(64, 64, 6, 6, device=DEV), 2, 2)
    frozen = torch.randn(64, 64, 3, 3, device=DEV)
    assert train._eligible(torch.randn(1, 64, 32, 32, device=DEV), frozen)


def test_pack_weight_matches_the_fused_adam_repack_and_adam_matches_torch():
    """ParamStore: parameters become views of the flat buffer, the kernel's repacking equals train.pack_weight, and three fused
    Adam steps equal torch.optim.Adam on the same gradients (fp32, 1e-6)."""
    from ipercore_b200 import train
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(6, 64, 3, padding=1), torch.nn.Conv2d(64, 3, 5, padding=2, bias=False),
                              torch.nn.Conv2d(64, 128, 1), torch.nn.Conv2d(128, 64, 4, 2, 1), torch.nn.ConvTranspose2d(64, 32, 4, 2, 1),
                              torch.nn.Conv2d(32, 8, 3, 2, 1), torch.nn.Conv2d(8, 8, 2)).to(DEV)
    kinds = {net[0].weight: (train.S1, 1), net[1].weight: (train.S1, 2), net[2].weight: (train.S1, 0), net[3].weight: (train.S2, 1),
             net[4].weight: (train.CT, 1), net[5].weight: (train.S2, 1)}
    ref = [p.detach().clone().requires_grad_(True) for p in net.parameters()]
    opt = torch.optim.Adam(ref, lr=1e-3, betas=(0.9, 0.999))
    bk = train.FlatGradBuckets(list(net.parameters()), n_buckets=2)
    store = train.ParamStore(list(net.named_parameters()), bk, native=lambda n, p: kinds.get(p), lr=1e-3)
    for p, r in zip(net.parameters(), ref):
        assert torch.equal(p.detach(), r.detach())
        assert store.p.data_ptr() <= p.data_ptr() < store.p.data_ptr() + store.p.numel() * 4
    for it in range(3):
        bk.zero()
        for p, r in zip(net.parameters(), ref):
            gr = torch.randn_like(r) * (0.1 + it)
            p.grad.copy_(gr); r.grad = gr.clone()
        store.step(); opt.step()
        for p, r in zip(net.parameters(), ref):
            assert float((p.detach() - r.detach()).abs().max()) <= 1e-6
    for m in list(net)[:6]:                      # the kernel's repacking after the last update == the torch formulation, all three kinds
        f, d = train.pack_weight(m.weight, *kinds[m.weight])
        assert torch.equal(m.weight._iper_pack[0], f.reshape(-1)) and torch.equal(m.weight._iper_pack[1], d.reshape(-1)), m
    assert not hasattr(net[6].weight, "_iper_pack")


@pytest.mark.parametrize("M,C,h,w", [(2, 64, 32, 48), (4, 256, 16, 16), (1, 128, 40, 24)])
def test_warp_bf16_forward_backward(M, C, h, w):
    """LWB.transform on NHWC bf16 vs F.grid_sample in fp32 (flows partly outside [-1, 1] and the -2 background value)."""
    from ipercore_b200 import train
    g = torch.Generator().manual_seed(M * 100 + C)
    src = torch.randn(M, C, h, w, generator=g).to(DEV).bfloat16().float().requires_grad_(True)
    T = (torch.rand(M, h, w, 2, generator=g) * 2.6 - 1.3).to(DEV)
    T[:, : h // 4] = -2.0
    dy = torch.randn(M, C, h, w, generator=g).to(DEV).bfloat16().float()
    ref = F.grid_sample(src, T, mode="bilinear", padding_mode="zeros", align_corners=False)
    (gref,) = torch.autograd.grad(ref, src, dy)
    s2 = src.detach().clone().requires_grad_(True)
    out = train._Warp.apply(s2, T)
    (gs,) = torch.autograd.grad(out, s2, dy.bfloat16())
    print("warp %dx%dx%dx%d: out %.2e dsrc %.2e" % (M, C, h, w, _rel(out, ref), _rel(gs, gref)))
    assert out.dtype == torch.bfloat16 and _rel(out, ref) <= 5e-3 and _rel(gs, gref) <= 5e-3       # bf16 rounding of the results


@pytest.mark.parametrize("bs,ns,C,h,w", [(1, 2, 64, 16, 32), (2, 3, 256, 8, 16), (1, 2, 128, 24, 24)])
def test_att_combine_forward_backward(bs, ns, C, h, w):
    from ipercore_b200 import train
    g = torch.Generator().manual_seed(bs + ns + C)
    mk = lambda *s: torch.randn(*s, generator=g).to(DEV).bfloat16().float().requires_grad_(True)
    k, v, q = mk(bs * ns, C, h, w), mk(bs * ns, C, h, w), mk(bs, C, h, w)
    da = torch.randn(bs, C, h, w, generator=g).to(DEV).bfloat16().float()
    kk, vv = k.view(bs, ns, C, h, w), v.view(bs, ns, C, h, w)
    logits = (kk * q.unsqueeze(1)).sum(dim=2, keepdim=True) / np.sqrt(C)
    ref = (torch.softmax(logits, dim=1) * vv).sum(dim=1)
    gref = torch.autograd.grad(ref, (k, v, q), da)
    ins = tuple(t.detach().clone().requires_grad_(True) for t in (k, v, q))
    a = train._AttCombine.apply(*ins, ns)
    got = torch.autograd.grad(a, ins, da.bfloat16())
    errs = [_rel(a, ref)] + [_rel(x, y) for x, y in zip(got, gref)]
    print("att_combine bs %d ns %d C %d: a %.2e dk %.2e dv %.2e dq %.2e" % ((bs, ns, C) + tuple(errs)))
    assert max(errs) <= 6e-3


@pytest.mark.parametrize("N,C,h,w,spade,act,slope", [(2, 64, 32, 32, True, False, 0.0), (1, 256, 16, 24, True, False, 0.0),
                                                     (2, 128, 16, 16, False, True, 0.0), (1, 512, 12, 12, False, True, 0.2),
                                                     (1, 64, 40, 40, False, False, 0.0)])
def test_norm_spade_forward_backward(N, C, h, w, spade, act, slope):
    from ipercore_b200 import train
    g = torch.Generator().manual_seed(N + C + h)
    mk = lambda sc=1.0: (torch.randn(N, C, h, w, generator=g) * sc + 0.3).to(DEV).bfloat16().float().requires_grad_(True)
    x, gm, bt = mk(2.0), mk(0.5), mk(0.5)
    dy = torch.randn(N, C, h, w, generator=g).to(DEV).bfloat16().float()
    ref = F.instance_norm(x, eps=1e-5)
    if spade:
        ref = ref * (1 + gm) + bt
    if act:
        ref = F.leaky_relu(ref, slope) if slope else F.relu(ref)
    ins = (x, gm, bt) if spade else (x,)
    gref = torch.autograd.grad(ref, ins, dy)
    ins2 = tuple(t.detach().clone().requires_grad_(True) for t in ins)
    y = train.inorm(ins2[0], act=act, slope=slope, gamma=ins2[1] if spade else None, beta=ins2[2] if spade else None)
    got = torch.autograd.grad(y, ins2, dy.bfloat16())
    errs = [_rel(y, ref)] + [_rel(a, b) for a, b in zip(got, gref)]
    print("norm N %d C %d %dx%d spade %d act %d: %s" % (N, C, h, w, spade, act, " ".join("%.2e" % e for e in errs)))
    assert y.dtype == torch.bfloat16 and max(errs) <= 8e-3          # bf16 stores; the activation mask is taken from the bf16 output


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
            step._zero(step.opt_G, step.bk_G)             # kernel mode: weight gradients land in the flat buffer behind p.grad
            loss.backward()
            grads = [params[n].grad.detach().clone() for n in names]
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


def test_train_step_cuda_graph_matches_eager():
    """The whole step captured in one CUDA graph (forward, both backward passes, both fused Adam passes) follows the eager run:
    same losses over 4 steps and the same parameters afterwards (bf16 atomics reorder sums: 1e-2 relative on the losses)."""
    from ipercore_b200 import train
    from ipercore_b200.generator import AttentionLWBGenerator
    from oracle.weights import synth_state_dict
    S = 256
    hist, params = {}, {}
    for graph in (False, True):
        torch.manual_seed(0)
        net = AttentionLWBGenerator(CFG); net.load_state_dict(synth_state_dict(0))
        step = train.LWGTrainStep(net, torch.device(DEV), graph=graph)
        batch = _batch(S)
        hist[graph] = [{k: float(v) for k, v in step.step(batch).items()} for _ in range(4)]
        torch.cuda.synchronize()
        params[graph] = step.G.net.res_blocks[1].main[0].weight.detach().float().clone()
        if graph:
            assert step._graph is not None and step.launches_per_step > 100
    for a, b in zip(hist[False], hist[True]):
        for k in a:
            assert np.isfinite(b[k]) and abs(a[k] - b[k]) <= 2e-2 * max(1.0, abs(a[k])), (k, a[k], b[k])
    moved = float((params[True] - synth_state_dict(0)["res_blocks.1.main.0.weight"].to(DEV)).abs().mean())
    assert moved > 0 and float((params[True] - params[False]).abs().mean()) <= 0.3 * moved


def test_personalize_loop_writes_a_checkpoint_the_inference_generator_loads(tmp_path):
    """Personalizer.run's schedule on LWGTrainStep: iterations = (no_decay + decay) * num_videos, G every 2nd batch, lr decay, and the
    product: personalized.pth under the reference's state_dict names, loaded strictly by the inference generator."""
    from ipercore_b200 import personalize, train
    from ipercore_b200.generator import AttentionLWBGenerator
    from oracle.weights import synth_state_dict
    sd0 = synth_state_dict(0)
    net = AttentionLWBGenerator(CFG); net.load_state_dict(sd0)
    step = train.LWGTrainStep(net, torch.device(DEV), graph=True)
    batches = [_batch(256, seed=s_) for s_ in (1, 2)]
    seen = []
    ckpt = str(tmp_path / "models" / "personalized.pth")
    hist = personalize.run(step, batches, num_videos=2, niters_no_decay=2, niters_decay=1, train_G_every_n_iterations=2, lr=1e-4,
                           ckpt_path=ckpt, on_iter=lambda i, out: seen.append(i))
    torch.cuda.synchronize()
    assert seen == [1, 2, 3, 4, 5, 6] and len(hist) == 6 and all(np.isfinite(float(h["G"])) for h in hist)
    assert step.st_G.lr < 1e-4                        # the decay phase ran
    sd = torch.load(ckpt)
    assert list(sd.keys()) == list(sd0.keys())
    moved = max(float((sd[k] - sd0[k]).abs().max()) for k in sd)
    assert 0 < moved < 1e-2                           # three G updates with Adam at lr 1e-4
    AttentionLWBGenerator(CFG).load_state_dict(sd, strict=True)
