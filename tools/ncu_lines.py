#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump per source line, and key raw metrics."""
import csv
import sys


def f(x):
    try:
        return float(x)
    except Exception:
        return 0.0


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    for w in ['Kernel Name', 'gpu__time_duration.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
              'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
              'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
              'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
              'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']:
        if w in hdr:
            i = hdr.index(w)
            print(w, rows[1][i], [r[i] for r in rows[2:]])


def lines(path, top=30):
    rows = list(csv.reader(open(path)))
    out, cur, hdr, ci = [], None, None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] == 'Function Name':
            continue
        if r[0] == 'Line No':
            hdr, ci = r, {}
            for i, n in enumerate(r):
                ci.setdefault(n, i)
            continue
        if r[0] and hdr:
            out.append((cur, int(r[0]), r[1].strip()[:80], f(r[ci['Instructions Executed']]), f(r[ci['Thread Instructions Executed']]), f(r[ci['# Samples']])))
    T = sum(o[3] for o in out)
    S = sum(o[5] for o in out)
    print('total warp-inst %.3g thread-inst %.3g lanes %.2f' % (T, sum(o[4] for o in out), sum(o[4] for o in out) / T))
    for o in sorted(out, key=lambda o: -o[3])[:top]:
        print(f"{o[0]:14s}:{o[1]:4d} inst {o[3] / T * 100:5.1f}% lanes {o[4] / max(o[3], 1):5.1f} stall {o[5] / S * 100:5.1f}% | {o[2]}")


if __name__ == '__main__':
    raw(sys.argv[1])
    lines(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
