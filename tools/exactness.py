#!/usr/bin/env python
"""Report how many pixels of the CUDA path agree bit for bit with the CPU oracle (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lavaframe_b200 as lf
from oracle_api import Oracle
from parity_metrics import radiance_agreement
pt = lf.PathTracer(0)
for name in ["cornell", "c2mini", "c3mini"]:
    pack = lf.ScenePack(os.path.join(ROOT, "tests", "golden", f"{name}.lfpack"))
    pt.upload_pack(pack)
    o = Oracle(pack.path)
    for first, n in ((2, 1), (3, 8)):
        pt.clear(); pt.render_frames(first, n); img = pt.read_accum(); ref = o.render_frames(first, n)
        neq = np.argwhere((img != ref).any(axis=2))
        print(name, f"frames {first}+{n}: bit-exact pixels {1 - len(neq) / (img.shape[0] * img.shape[1]):.6f} ({len(neq)} differ), within 1e-3: {radiance_agreement(img, ref):.6f}")
        for y, x in neq[:3]:
            print("   ", (x, y), img[y, x], ref[y, x])
    o.close()
