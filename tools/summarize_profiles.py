#!/usr/bin/env python
"""Turn the ncu CSVs of tools/gpu_job_final.sh into the summaries kept under profiles/ (development aid).

    python tools/summarize_profiles.py <tag> [--write]

  * <tag>_traffic_<workload>.csv  -> profiles/extend_traffic.json (DRAM bytes per closest-hit / shadow launch; bench.py's roofline.traffic)
  * <tag>_c2_launches.csv         -> kernel shares of one step (printed)
  * <tag>_c2_ncu_raw.csv          -> key counters of the --set full capture (printed)
Instrumented launches (template argument COUNT = 1, the visit-count pass outside the timed region) are excluded.
"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")


def rows_of(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    return rows[0], rows[1:]


def short(name):
    m = re.match(r"(?:void )?(?:lf::)?(\w+)(<[^>]*>)?", name)
    return m.group(1), [a.strip() for a in (m.group(2) or "<>")[1:-1].split(",") if a.strip()]


def traffic(tag, workload):
    h, rows = rows_of(os.path.join(OUT, f"{tag}_traffic_{workload}.csv"))
    ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = defaultdict(dict)
    for r in rows:
        per[r[ii]]["name"] = r[ki]
        per[r[ii]][r[mi]] = float(r[vi].replace(",", ""))
    out = {}
    for kind, any_flag in (("extend", "0"), ("shadow", "1")):
        sel = []
        for v in per.values():
            k, args = short(v["name"])
            if k == "k_trace" and args[0] == any_flag and args[2] == "0":
                sel.append(v)
        n = len(sel)
        out[kind] = {"launches": n,
                     "dram_read_bytes_per_launch": sum(v["dram__bytes_read.sum"] for v in sel) / n,
                     "dram_write_bytes_per_launch": sum(v["dram__bytes_write.sum"] for v in sel) / n,
                     "ncu_ms_per_launch": sum(v["gpu__time_duration.sum"] for v in sel) / n / 1e6}
    out["dram_bytes_per_launch"] = out["extend"]["dram_read_bytes_per_launch"] + out["extend"]["dram_write_bytes_per_launch"]
    out["source"] = (f"profiles/{tag}_traffic_{workload}.csv: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                     f"--clock-control none -k regex:k_trace over bench.py --steps 1 --warmup 1 --workload {workload}; mean over the "
                     "non-instrumented closest-hit launches (warm-up and timed step)")
    return out


def shares(tag):
    h, rows = rows_of(os.path.join(OUT, f"{tag}_c2_launches.csv"))
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    tot = defaultdict(float)
    for r in rows:
        k, args = short(r[ki])
        v = float(r[vi].replace(",", "")) / 1e6
        if k == "k_trace":
            if args[2] == "1":
                continue
            k = "k_trace<ANY=%s>" % args[0]
        elif k == "k_shade":
            if args[0] == "1":
                continue
        elif k == "k_generate":
            if args and args[0] == "1":
                continue
        elif k in ("k_read_probe", "k_node_probe"):
            continue
        tot[k] += v
    T = sum(tot.values())
    print("kernel shares (ncu, cold and serialised):")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"  {k:22s} {v:9.2f} ms  {v / T:6.3f}")


def raw(tag):
    h, rows = rows_of(os.path.join(OUT, f"{tag}_c2_ncu_raw.csv"))
    want = ["gpu__time_duration.sum", "launch__registers_per_thread", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            "sm__warps_active.avg.pct_of_peak_sustained_active"]
    ki = h.index("Kernel Name")
    for r in rows[1:]:
        print(short(r[ki]))
        for w in want:
            if w in h:
                print(f"    {w:80s} {r[h.index(w)]}")


if __name__ == "__main__":
    tag = sys.argv[1]
    t = {w: traffic(tag, w) for w in ("c2_full", "c4_stress") if os.path.exists(os.path.join(OUT, f"{tag}_traffic_{w}.csv"))}
    print(json.dumps(t, indent=1))
    if "--write" in sys.argv:
        json.dump(t, open(os.path.join(ROOT, "profiles", "extend_traffic.json"), "w"), indent=1)
    if os.path.exists(os.path.join(OUT, f"{tag}_c2_launches.csv")):
        shares(tag)
    if os.path.exists(os.path.join(OUT, f"{tag}_c2_ncu_raw.csv")):
        raw(tag)
