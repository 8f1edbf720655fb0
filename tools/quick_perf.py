#!/usr/bin/env python
"""Ad-hoc device timing of the path tracer on a pack (development aid; bench.py is the contract)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lavaframe_b200 as lf
from lavaframe_b200.pathtracer import algorithmic_bytes, algorithmic_bytes_total

ap = argparse.ArgumentParser()
ap.add_argument("pack")
ap.add_argument("--res", type=int, nargs=2)
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--no-cull", action="store_true")
ap.add_argument("--fif", type=int, default=0)
ap.add_argument("--depth", type=int, default=0)
a = ap.parse_args()

pack = lf.ScenePack(a.pack)
pt = lf.PathTracer(0)
ov = dict(kernel_mode=a.mode, no_cull=int(a.no_cull), frames_in_flight=a.fif)
if a.res:
    ov.update(width=a.res[0], height=a.res[1], tile_width=a.res[0], tile_height=a.res[1])
if a.depth:
    ov.update(max_depth=a.depth)
pt.upload_pack(pack, **ov)
W, H = pt.params.width, pt.params.height
stream = torch.cuda.Stream()      # a non-default stream: handle 0 would select the context's own stream
torch.cuda.set_stream(stream)
pt.set_stream(stream.cuda_stream)
pt.clear(); pt.render_frames(2, a.spp); torch.cuda.synchronize()
times = []
for r in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pt.clear()
    e0.record(stream); pt.render_frames(2, a.spp); e1.record(stream); torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
ms = min(times)
# counters + stage times in a separate instrumented pass
pt.update_params(count_work=1); pt.reset_counters(); pt.clear(); pt.render_frames(2, a.spp); c = pt.counters(); pt.update_params(count_work=0)
pt.set_profiling(True); pt.clear(); pt.render_frames(2, a.spp); st = pt.stage_stats(); pt.set_profiling(False)
rays = c["rays_closest"] + c["rays_shadow"]
out = dict(pack=os.path.basename(a.pack), res=[W, H], spp=a.spp, mode=a.mode, cull=not a.no_cull, ms=ms, samples_per_s=W * H * a.spp / ms * 1e3,
           mrays_per_s=rays / ms / 1e3, rays_per_sample=rays / c["samples"], bytes_per_ray=algorithmic_bytes(c) / rays,
           trav_GBps_alg=algorithmic_bytes(c) / ((st["extend"]["ms"] + st["shadow"]["ms"] + st["megakernel"]["ms"]) * 1e6),
           stages={k: round(v["ms"], 3) for k, v in st.items()}, counters=c)
print(json.dumps(out))
