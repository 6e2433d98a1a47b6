"""Print the per-launch conv timing table(s) written by `bench.py --dump-layers` side by side."""
import json
import os
import sys

tabs = [json.load(open(p)) for p in sys.argv[1:]]
MODE = {0: "s1", 1: "s2", 2: "convT", 3: "row5"}
for i, rows in enumerate(zip(*tabs)):
    a = rows[0]
    print("%2d %-5s k%d rows %3d K %4d %3dx%-3d epi %d pair %d |" % (i, MODE[a["mode"]], a["ksize"], a["rows"], a["K"], a["H"], a["W"],
                                                                a["epi"], a["cta_pair"]),
          "  ".join("%.3f ms %4.0f TF" % (r["ms"], r["exec_tflops"]) for r in rows))
print("total ms:", "  ".join("%.2f" % sum(r["ms"] for r in t) for t in tabs))
for p in sys.argv[1:]:
    if os.path.exists(p + ".other.json"):
        o = json.load(open(p + ".other.json"))
        print("non-conv launches of the same batch (%s):" % os.path.basename(p),
              "  ".join("%s x%d %.3f ms" % (k, v["launches"], v["ms"]) for k, v in o.items()),
              " | total %.2f ms" % sum(v["ms"] for v in o.values()))
