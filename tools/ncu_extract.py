#!/usr/bin/env python
"""Trim an .ncu-rep (ncu --set full) to the metrics the roofline claims rest on -> CSV for profiles/.
usage: ncu_extract.py report.ncu-rep out.csv"""
import csv
import subprocess
import sys

KEEP = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("sm__cycles_elapsed.max", "sm_cycles"),
]
UNIT_SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3,
              "ns": 1e-3, "us": 1.0, "ms": 1e3, "second": 1e6}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in raw.splitlines() if not l.startswith("==")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as fp:
        w = csv.writer(fp)
        w.writerow([n for _, n in KEEP])
        for r in data:
            line = []
            for key, name in KEEP:
                i = idx.get(key)
                if i is None:
                    line.append(""); continue
                v = r[i]
                if name in ("time_us", "dram_read_MB", "dram_write_MB"):
                    v = "%.3f" % (float(v.replace(",", "")) * UNIT_SCALE.get(units[i], 1.0))
                elif name not in ("kernel", "grid", "block"):
                    try:
                        v = "%.2f" % float(v.replace(",", ""))
                    except ValueError:
                        pass
                line.append(v)
            w.writerow(line)
    print("wrote", out, len(data), "launches")


if __name__ == "__main__":
    main()
