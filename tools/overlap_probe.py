#!/usr/bin/env python
"""Development aid (run through gpurun): do two half-size wavefront pipelines on two streams of ONE GPU overlap usefully?
Two contexts (own stream, own path state) render alternate halves of every step's frames; the launches of a step are enqueued back to back,
so whatever one pipeline leaves idle (the tails of its persistent traversal kernels, its small deep bounces) the other can fill.
usage: overlap_probe.py <workload> [spp per step]"""
import json
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (ensure_pack)
import lavaframe_b200 as lf  # noqa: E402

name = sys.argv[1]
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
pack = lf.ScenePack(bench.ensure_pack(name))


def run(nctx, ctas=None, reps=6):
    if ctas is None:
        os.environ.pop("LF_CTAS_PER_SM", None)
    else:
        os.environ["LF_CTAS_PER_SM"] = str(ctas)
    pts = []
    for k in range(nctx):
        pt = lf.PathTracer(0); pt.upload_pack(pack); pts.append(pt)
    W, H = pts[0].params.width, pts[0].params.height
    n = spp // nctx

    def step(i):
        for k, pt in enumerate(pts):
            pt.render_frames(2 + i * spp + k * n, n)
    for i in range(3):
        step(i)
    for pt in pts:
        pt.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        step(3 + i)
    for pt in pts:
        pt.synchronize()
    dt = time.perf_counter() - t0
    for pt in pts:
        pt.close()
    return round(W * H * n * nctx * reps / dt / 1e6, 1)


res = {"workload": name, "spp_per_step": spp}
res["1ctx"] = run(1)
res["2ctx"] = run(2)
res["2ctx_5ctas"] = run(2, 5)
res["1ctx_again"] = run(1)
res["3ctx"] = run(3)
print(json.dumps(res))
