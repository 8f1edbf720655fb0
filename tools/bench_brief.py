#!/usr/bin/env python
"""Print the headline fields of bench.py JSON lines read from stdin (development aid)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    j = json.loads(line)
    r = j.get("roofline", {})
    print(f"value {j['value']/1e6:.1f} M samples/s | e2e {j['e2e']['value']/1e6:.1f} | {j.get('mrays_per_s', 0):.0f} Mrays/s | ms/step {j['ms_per_step']:.2f} | "
          f"stage_ms {r.get('stage_ms')} | extend {r.get('achieved', 0):.0f} GB/s frac_hbm {r.get('frac', 0):.3f} frac_l2 {r.get('frac_of_l2', 0):.3f} | launches {j.get('gpu_launches')} | clocks {j.get('clocks')}")
