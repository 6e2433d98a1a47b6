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
