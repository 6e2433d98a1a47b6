"""Micro-benchmark of one conv layer shape through ops.conv_gemm (GPU box): epilogue variants x precisions.
usage: python tools/bench_conv_layer.py [N H Cin Cout]   (3x3 stride-1, cta_pair=2)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from ipercore_b200 import ops
from ipercore_b200.ops import Planes

N, H, Cin, Cout = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (50, 64, 256, 256)
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
w = (torch.rand((Cout, Cin, 3, 3), generator=g) - 0.5) * 0.05
b = (torch.rand((Cout,), generator=g) - 0.5).to(dev)
x = torch.rand((N, Cin, H, H), generator=g).to(dev)
for P in (2, 3, 1):
    a = Planes.from_nchw(x, P)
    wp = ops.pack_conv_weight(w, P).to(dev)
    res = Planes.from_nchw(torch.rand((N, Cout, H, H), generator=g).to(dev), P)
    out = Planes.empty(P, N, H, H, Cout, dev)
    ws = ops.stats_workspace(N, Cout, dev)
    for name, kw in (("plain", {}), ("relu", dict(relu=True)), ("residual", dict(x=res)), ("stats", dict(stats_ws=ws)),
                     ("residual+stats", dict(x=res, stats_ws=ws))):
        for pair in (2, 1) if P != 3 else (2,):
            f = lambda: ops.conv_gemm(a, wp, 0, 3, Cout, 256 if Cout >= 256 else Cout, ops.IPER_EPI_PLANES, bias=b, out=out,
                                      cta_pair=pair, **kw)
            for _ in range(3):
                f()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                f()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            tf = 2.0 * N * H * H * Cout * Cin * 9 * {1: 1, 2: 3, 3: 2}[P] / (ms * 1e-3) / 1e12
            print("P=%d pair=%d %-15s %.3f ms  %5.0f TF/s (fp16-equivalent)" % (P, pair, name, ms, tf))
