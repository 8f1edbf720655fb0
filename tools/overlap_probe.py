#!/usr/bin/env python
"""Feasibility probe: do two half-size wavefront pipelines on two streams overlap usefully on one GPU? (development aid)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lavaframe_b200 as lf
pack = lf.ScenePack(sys.argv[1])
spp = 16
def run(nctx, fif, reps=6):
    pts = []
    for k in range(nctx):
        pt = lf.PathTracer(0); pt.upload_pack(pack, frames_in_flight=fif); pts.append(pt)
    W, H = pts[0].params.width, pts[0].params.height
    def step(i):
        for k, pt in enumerate(pts):
            n = spp // nctx
            pt.render_frames(2 + i * spp + k * n, n)
    for i in range(3): step(i)
    for pt in pts: pt.synchronize()
    t0 = time.perf_counter()
    for i in range(reps): step(3 + i)
    for pt in pts: pt.synchronize()
    dt = time.perf_counter() - t0
    for pt in pts: pt.close()
    return W * H * spp * reps / dt / 1e6
print("1 ctx, 16 frames in flight:", run(1, 16))
os.environ["LF_CTAS_PER_SM"] = "4"
print("2 ctx x 8 frames, 4 CTAs/SM each:", run(2, 8))
os.environ["LF_CTAS_PER_SM"] = "8"
print("2 ctx x 8 frames, 8 CTAs/SM each:", run(2, 8))
os.environ["LF_CTAS_PER_SM"] = "3"
print("3 ctx x 5 frames, 3 CTAs/SM each:", run(3, 5))
