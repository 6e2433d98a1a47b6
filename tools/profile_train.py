"""Kernel-time breakdown of one training step (bench.py --config train workload) with torch.profiler (CUPTI; no nsys in the image).
Run ON THE GPU BOX:  python tools/profile_train.py [out.txt]   -> table of device time per kernel name, ours vs library."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import CFG  # noqa: E402
from ipercore_b200 import train  # noqa: E402
from ipercore_b200.generator import AttentionLWBGenerator  # noqa: E402
from oracle import weights  # noqa: E402


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "gpurun_out/train_profile.txt"
    S, ns = 512, 2
    dev = torch.device("cuda", 0)
    net = AttentionLWBGenerator(CFG)
    net.load_state_dict(weights.synth_state_dict(0), strict=True)
    step = train.LWGTrainStep(net, dev, graph="--graph" in sys.argv)
    g = torch.Generator().manual_seed(100)
    r = lambda *s: (torch.rand(*s, generator=g) * 2 - 1).to(dev)
    batch = dict(bg_inputs=torch.cat([r(1, 1, 3, S, S), (r(1, 1, 1, S, S) > 0).float()], 2), src_inputs=r(1, ns, 6, S, S),
                 tsf_inputs=r(1, 1, 6, S, S), Tst=r(1, 1, ns, S, S, 2), real_src=r(1, ns, 3, S, S), real_tsf=r(1, 1, 3, S, S),
                 real_bg=r(1, 3, S, S), body_mask=(r(1, ns + 1, 1, S, S) > 0).float())
    for _ in range(5):
        step.step(batch)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            step.step(batch)
        torch.cuda.synchronize()
    rows = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            t = rows.setdefault(ev.name, [0, 0.0])
            t[0] += 1; t[1] += ev.device_time
    tot = sum(v[1] for v in rows.values())
    ours = sum(v[1] for k, v in rows.items() if "iper" in k)
    with open(out_path, "w") as f:
        f.write("2 training steps: device time %.2f ms total, %.2f ms in iper:: kernels (%.0f%%), %d launches\n"
                % (tot / 1e3, ours / 1e3, 100 * ours / tot, sum(v[0] for v in rows.values())))
        for k, v in sorted(rows.items(), key=lambda kv: -kv[1][1])[:70]:
            f.write("%9.1f us %6d x  %5.1f%%  %s\n" % (v[1], v[0], 100 * v[1] / tot, k[:150] + (" ... " + k[-260:] if len(k) > 420 else "")))
        f.write("\nCPU-side (self CPU time, top 25):\n")
        f.write(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
    print(open(out_path).read()[:6000])


if __name__ == "__main__":
    main()
