"""Host-side logic of the training step on CPU: the bucketed flat gradient all-reduce (gloo, world_size 2)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ipercore_b200.train import FlatGradBuckets
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    net[3].weight.requires_grad_(False); net[3].bias.requires_grad_(False)
    bk = FlatGradBuckets(list(net.parameters()), n_buckets=3)
    assert sum(len(b) for b in bk.buckets) == 4
    for it in range(2):                       # two iterations: buffers are reused, hooks re-arm
        bk.zero()
        x = torch.full((5, 8), float(rank + 1 + it))
        net(x).sum().backward()
        bk.finish()
    # expectation: mean over ranks of the local gradients
    ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    ref.load_state_dict(net.state_dict())
    want = [torch.zeros_like(p) for p in list(ref.parameters())[:4]]
    for r in range(world):
        ref.zero_grad()
        ref(torch.full((5, 8), float(r + 2))).sum().backward()
        for w_, p in zip(want, list(ref.parameters())[:4]):
            w_ += p.grad / world
    ok = all(torch.allclose(p.grad, w_, atol=1e-6) for p, w_ in zip(list(net.parameters())[:4], want))
    views = all(p.grad.data_ptr() >= f.data_ptr() for b, f in zip(bk.buckets, bk.flat) for p in b)
    out[rank] = bool(ok and views)
    dist.destroy_process_group()


def test_flat_grad_buckets_gloo_world2():
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29541, out), nprocs=2, join=True)
    assert dict(out) == {0: True, 1: True}


def _unpack(rows_by_k, rows, taps, cols):
    return rows_by_k.float().reshape(rows, taps, cols)


def _emulate_transposed(x, phases, cout, k, pad):
    """csrc/train.cu iper_conv_transposed_bf16 in torch: four stride-1 phase convolutions over the input grid, tap offsets
    ((px + pad - kx) / 2, (py + pad - ky) / 2), weight blocks (rows cout_pad, K = (tap in phase, ci_pad)), interleaved stores."""
    import torch.nn.functional as F
    from ipercore_b200.train import _pad64, _phase_taps
    n, cin, h, w = x.shape
    cinp, coutp = _pad64(cin), _pad64(cout)
    out = torch.zeros(n, cout, 2 * h, 2 * w)
    off = 0
    for ph, taps in enumerate(_phase_taps(k, pad)):
        py, px = ph >> 1, ph & 1
        blk = phases[off:off + coutp * len(taps) * cinp].float().reshape(coutp, len(taps), cinp); off += blk.numel()
        acc = torch.zeros(n, cout, h, w)
        for t, (ky, kx) in enumerate(taps):
            dy, dx = (py + pad - ky) // 2, (px + pad - kx) // 2
            assert (py + pad - ky) % 2 == 0 and (px + pad - kx) % 2 == 0
            xs = F.pad(x, (2, 2, 2, 2))[:, :, 2 + dy:2 + dy + h, 2 + dx:2 + dx + w]          # x[y + dy, x + dx], zero outside
            acc += torch.einsum("nchw,oc->nohw", xs, blk[:cout, t, :cin])
        out[:, :, py::2, px::2] = acc
    return out


def test_phase_packing_reproduces_transposed_convolutions():
    """Index algebra of the stride-2 transposition (host side of csrc/train.cu + train.pack_weight): ConvTranspose2d(4,2,1) forward
    from the CT forward packing, and the data gradients of stride-2 convolutions (k = 3, 4) from the S2 dgrad packing."""
    import torch.nn.functional as F
    from ipercore_b200 import train
    torch.manual_seed(0)
    bf = lambda t: t.bfloat16().float()
    x = bf(torch.randn(2, 5, 6, 7))
    wt = bf(torch.randn(5, 3, 4, 4))                                   # ConvTranspose2d weight (ci, co, k, k)
    fwd, dg = train.pack_weight(wt, train.CT, 1)
    assert torch.allclose(_emulate_transposed(x, fwd, 3, 4, 1), F.conv_transpose2d(x, wt, stride=2, padding=1), atol=1e-5)
    # its data gradient = a stride-2 conv of dY with the plain (ci_T rows, K = (tap, co_T)) packing
    dy = bf(torch.randn(2, 3, 12, 14))
    plain = dg.float().reshape(64, 16, 64)[:5, :, :3]                  # (ci_T, tap, co_T)
    ref = F.conv2d(dy, wt.permute(0, 1, 2, 3), stride=2, padding=1)    # conv with weight (out = ci_T, in = co_T, k, k)
    assert torch.allclose(ref, F.conv2d(dy, plain.permute(0, 2, 1).reshape(5, 3, 4, 4), stride=2, padding=1), atol=1e-5)
    xg = x.clone().requires_grad_(True)
    (gx,) = torch.autograd.grad(F.conv_transpose2d(xg, wt, stride=2, padding=1), xg, dy)
    assert torch.allclose(gx, ref, atol=1e-4)
    for k in (3, 4):                                                    # stride-2 conv (co, ci, k, k): dX from the phase packing
        w = bf(torch.randn(4, 6, k, k))
        xin = bf(torch.randn(1, 6, 8, 10)).requires_grad_(True)
        y = F.conv2d(xin, w, stride=2, padding=1)
        dyy = bf(torch.randn_like(y))
        (gxin,) = torch.autograd.grad(y, xin, dyy)
        f2, d2 = train.pack_weight(w, train.S2, 1)
        assert torch.equal(f2.float().reshape(64, k * k, 64)[:4, :, :6], w.permute(0, 2, 3, 1).reshape(4, k * k, 6))
        assert torch.allclose(_emulate_transposed(dyy, d2, 6, k, 1), gxin, atol=1e-4), k


def test_rev_packing_gives_the_data_gradient_of_non_same_convolutions():
    """S1 with k = 4, pad 1 (the discriminator's last layers): dX = conv(dY, rev packing, padding k-1-pad)."""
    import torch.nn.functional as F
    from ipercore_b200 import train
    torch.manual_seed(1)
    bf = lambda t: t.bfloat16().float()
    for k, pad in ((4, 1), (3, 1), (5, 2), (1, 0)):
        w = bf(torch.randn(3, 5, k, k))
        x = bf(torch.randn(1, 5, 9, 11)).requires_grad_(True)
        y = F.conv2d(x, w, padding=pad)
        dy = bf(torch.randn_like(y))
        (gx,) = torch.autograd.grad(y, x, dy)
        _, rev = train.pack_weight(w, train.S1, pad)
        wr = rev.float().reshape(64, k * k, 64)[:5, :, :3].permute(0, 2, 1).reshape(5, 3, k, k)
        assert torch.allclose(F.conv2d(dy, wr, padding=k - 1 - pad), gx, atol=1e-4), (k, pad)


def test_native_weight_classification_covers_every_convolution_of_the_generator():
    """ParamStore keeps one forward + one dgrad packing per convolution weight; which layouts depends on how the layer is CALLED
    (TrainableGenerator.native_weight mirrors the layer code): 9 stride-2 encoder convs, 9 transposed decoder convs, the rest
    stride 1 with 'same' padding — and no 4-D parameter of the 221-tensor checkpoint is left unclassified."""
    from ipercore_b200 import train
    from ipercore_b200.generator import AttentionLWBGenerator
    cfg = dict(name="AttLWB-SPADE", BGNet=dict(cond_nc=4, n_res_block=6, num_filters=[64, 128, 128, 256]),
               SIDNet=dict(cond_nc=6, n_res_block=6, num_filters=[64, 128, 256]),
               TSFNet=dict(cond_nc=6, n_res_block=6, num_filters=[64, 128, 256]))
    G = train.TrainableGenerator(AttentionLWBGenerator(cfg))
    kinds = {n: train.TrainableGenerator.native_weight(n, p) for n, p in G.named_parameters() if p.dim() == 4}
    assert all(v is not None for v in kinds.values()), [n for n, v in kinds.items() if v is None]
    by = {k: [n for n, v in kinds.items() if v[0] == k] for k in (train.S1, train.S2, train.CT)}
    assert len(by[train.S2]) == 9 and all(G.net.get_parameter(n[4:]).shape[2] == 3 for n in by[train.S2])
    assert len(by[train.CT]) == 9 and all(G.net.get_parameter(n[4:]).shape[2] == 4 for n in by[train.CT])
    assert len(by[train.S1]) == len(kinds) - 18
    for n in by[train.S1]:
        k = G.net.get_parameter(n[4:]).shape[2]
        assert kinds[n] == (train.S1, k // 2) and k in (1, 3, 5, 7), n
    assert kinds["net.bg_net.main.0.weight"] == (train.S1, 3) and kinds["net.tsf_img_reg.0.weight"] == (train.S1, 2)
    assert kinds["net.tsf_net_dec.upconvs.0.0.weight"] == (train.CT, 1) and kinds["net.src_net.encoders.layers.1.0.weight"] == (train.S2, 1)
    # the discriminator: four stride-2 4x4 convs, two stride-1 4x4 / pad-1 convs
    dk = sorted(v for v in train.PatchDiscriminator().native_weight().values())
    assert dk == [(train.S1, 1)] * 2 + [(train.S2, 1)] * 4


def test_personalize_schedule_with_a_mock_step():
    """Host logic of ipercore_b200.personalize.run (Personalizer.run, personalization.py:95-151): iteration count, the
    train_G_every_n_iterations pattern over each pass of the loader, the linear decay phase, the checkpoint."""
    import tempfile
    from ipercore_b200 import personalize

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.ones(3))

    class G:
        net = Net()

    class Step:
        def __init__(self):
            self.G, self.calls, self.lrs = G(), [], []

        def set_lr(self, lr):
            self.lrs.append(lr)

        def step(self, batch, trainable=True):
            self.calls.append((batch, trainable))
            return {"G": torch.tensor(float(len(self.calls)))}

    st = Step()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "models", "m", "personalized.pth")
        hist = personalize.run(st, ["a", "b", "c"], num_videos=2, niters_no_decay=2, niters_decay=2, train_G_every_n_iterations=2, lr=1e-3,
                               final_lr=1e-5, ckpt_path=path)
        assert list(torch.load(path).keys()) == ["w"]
    assert len(hist) == 8 and [b for b, _ in st.calls] == ["a", "b", "c", "a", "b", "c", "a", "b"]
    assert [t for _, t in st.calls] == [False, True, False] * 2 + [False, True]          # G on every 2nd batch of each pass
    assert st.lrs[0] == 1e-3 and len(st.lrs) == 1 + 4                                     # constant for 4 iterations, then 4 decay steps
    assert abs(st.lrs[-1] - 1e-5) < 1e-9 and all(a > b for a, b in zip(st.lrs[1:], st.lrs[2:]))
    try:
        personalize.run(Step(), [], niters_no_decay=1)
        assert False, "an empty loader must raise"
    except ValueError:
        pass


def test_conv_falls_back_to_torch_on_cpu():
    """Without a GPU tensor nothing is routed to the kernels (_kind == 0) and train.conv is the plain bf16 torch formulation,
    tiny channel ends included (they are zero-padded to 8 for cuDNN's sake; on CPU that must not change the result)."""
    import torch.nn.functional as F
    from ipercore_b200 import train
    torch.manual_seed(0)
    x = torch.randn(1, 6, 20, 24)
    w = torch.randn(16, 6, 3, 3) * 0.2
    b = torch.randn(16) * 0.1
    assert train._kind(x, w, 2, 1) == 0 and train._kind(x, w) == 0
    for kw, ref in ((dict(stride=2, padding=1), F.conv2d(x, w, b, stride=2, padding=1)), (dict(), F.conv2d(x, w, b, padding=1))):
        y = train.conv(x, w, b, relu=True, **kw)
        assert y.dtype == torch.bfloat16 and float((y.float() - F.relu(ref)).abs().max()) <= 0.06 * float(ref.abs().max())
    wt = torch.randn(6, 8, 4, 4) * 0.2
    yt = train.conv(x, wt, None, stride=2, padding=1, transposed=True)
    rt = F.conv_transpose2d(x, wt, stride=2, padding=1)
    assert yt.shape == rt.shape and float((yt.float() - rt).abs().max()) <= 0.06 * float(rt.abs().max())
