#!/usr/bin/env python
"""Development aid: per-step totals of an ncu launch list (gpu__time_duration.sum, --csv) of lfcuda_build_blas."""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
t = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
    k = r[ki]
    if "k_blas_step" in k:
        k = k.split("k_blas_step<")[-1].split(">")[0].split("::")[-1]
    t[k[:40]][0] += 1
    t[k[:40]][1] += v
tot = sum(us for _, us in t.values())
for k, (c, us) in sorted(t.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} {c:5d} launches {us / 1e3:9.3f} ms  {100 * us / tot:5.1f} %")
print(f"{'total':40s} {sum(c for c, _ in t.values()):5d} launches {tot / 1e3:9.3f} ms")
