#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of one steady-state batch.
usage: launch_summary.py launches.csv [first_kernel_substring]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    anchor = sys.argv[2] if len(sys.argv) > 2 else "raster_kernel"
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    names = [re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("iper::", "") for r in rows]
    vals = [float(r["Metric Value"].replace(",", "")) / 1000.0 for r in rows]
    starts = [i for i, n in enumerate(names) if anchor in n]
    if len(starts) < 2:
        lo, hi = 0, len(names)
    else:
        lo, hi = starts[-2], starts[-1]
    batch = list(zip(names[lo:hi], vals[lo:hi]))
    tot = sum(v for _, v in batch)
    print("one batch: %d launches, %.1f us (cold-cache, serialised: compare shares)" % (len(batch), tot))
    agg = collections.OrderedDict()
    for k, v in batch:
        agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += v
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("  %-42s n=%3d %10.1f us %5.1f%%" % (k[:42], c, v, 100 * v / tot))
    return batch


if __name__ == "__main__":
    b = main()
    if "-v" in sys.argv:
        for i, (k, v) in enumerate(b):
            print(i, k[:40], "%.1f" % v)
