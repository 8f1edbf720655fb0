#!/usr/bin/env python
"""Development aid: price the bit-exact arithmetic.  Runs whatever liblfcuda.so is selected (LF_LFCUDA_SO=ab/fmad.so = the same source built
with -fmad=true, i.e. multiply-adds contracted) against the goldens of the UNMODIFIED reference on llvmpipe and prints north_star's three bars:
primary-hit IDs (>= 99.99 %, t within 1e-5), 1-spp radiance within 1e-3 (>= 99.9 %), converged RMSE / mean luminance (< 0.5 %)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import lavaframe_b200 as lf  # noqa: E402
from parity_metrics import hits_agreement, radiance_agreement, rmse_over_mean_luminance  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
pt = lf.PathTracer(0)
print("library:", os.environ.get("LF_LFCUDA_SO", "in-tree liblfcuda.so"))
for name in ("cornell", "c2mini", "c3mini", "c4gold"):
    g = np.load(os.path.join(G, f"{name}_llvmpipe.npz"))
    pt.upload_pack(lf.ScenePack(os.path.join(G, f"{name}.lfpack")))
    t, tri, mat, em = pt.primary_hits(2)
    surf = em == 0
    ids, ids_t = hits_agreement(t[surf], tri[surf], mat[surf], g["hits_t"][surf], g["hits_tri"][surf], g["hits_mat"][surf])
    pt.clear(); pt.render_frames(2, 1)
    one = pt.read_accum()
    n = int(g["nspp"])
    pt.clear(); pt.render_frames(2, n)
    many = pt.read_accum() / np.float32(n)
    print(f"{name:8s} primary ids {ids:.6f} (ids + t within 1e-5: {ids_t:.6f}) | 1 spp within 1e-3: {radiance_agreement(one, g['spp1']):.6f}, bit-identical "
          f"{float((one == g['spp1']).all(axis=2).mean()):.6f} | {n} spp: RMSE / mean luminance {rmse_over_mean_luminance(many, g['sppN']):.3e}, within 1e-3 {radiance_agreement(many, g['sppN']):.6f}")
g = np.load(os.path.join(G, "cornell_llvmpipe_4096spp.npz"))
pt.upload_pack(lf.ScenePack(os.path.join(G, "cornell.lfpack")))
pt.clear(); pt.render_frames(2, 4096)
img = pt.read_output(1.0 / 4096, 0)
print(f"cornell 4096 spp vs the reference's own 4096-spp image: RMSE / mean luminance {rmse_over_mean_luminance(img, g['spp4096']):.3e} (bar 5e-3), "
      f"within 1e-3 {radiance_agreement(img, g['spp4096']):.6f}, bit-identical {float((img == g['spp4096']).all(axis=2).mean()):.6f}")
pt.close()
