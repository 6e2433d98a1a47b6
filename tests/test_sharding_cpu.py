"""CPU tests of the multi-GPU plumbing (no data-path collective): contiguous frame shards, per-rank workloads, and the
max-over-ranks timing reduction, exercised with world_size=2 over gloo on 127.0.0.1."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ipercore_b200.engine import shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 300, 304):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import argparse
    import bench
    args = argparse.Namespace(size=64, ns=2, frames=6)
    wl = bench.make_workload(args, rank)                    # each rank synthesizes its own clip (weak scaling)
    lo, hi = shard_range(12, rank, world)
    mine = torch.zeros(12); mine[lo:hi] = 1
    dist.all_reduce(mine)                                   # test-only: every frame owned exactly once
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                # the timing reduction bench.py uses
    q.put((rank, float(mine.min()), float(mine.max()), float(t), float(wl["cams"][0, 0]), wl["verts"].shape))
    dist.destroy_process_group()


def test_two_rank_gloo_plumbing():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert all(r[1] == 1.0 and r[2] == 1.0 for r in res)            # disjoint, complete cover
    assert all(r[3] == 11.0 for r in res)                           # max over ranks
    assert res[0][4] != res[1][4] and res[0][5] == (6, 6890, 3)     # different clips per rank, same shape
