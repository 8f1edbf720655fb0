"""Development aid: how closely the CPU oracle reproduces the reference-on-llvmpipe golden images (tests/golden/*.npz).

    python tools/oracle_vs_llvmpipe.py [cornell c2mini c3mini]

Prints, per scene: primary-hit t bit-equality, and for the 1-spp image the fraction of pixels that are bit-identical and
within 1e-3 (north_star check 2).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_api import Oracle  # noqa: E402
from parity_metrics import radiance_agreement  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    for name in (sys.argv[1:] or ["cornell", "c2mini", "c3mini"]):
        g = np.load(os.path.join(GOLD, f"{name}_llvmpipe.npz"))
        o = Oracle(os.path.join(GOLD, f"{name}.lfpack"))
        t, tri, mat, em = o.primary_hits(2)
        s1 = o.render_frames(2, 1)
        n = int(g["nspp"])
        sN = o.render_frames(2, n) / np.float32(n)
        o.close()
        surf = (em == 0) & (g["hits_emitter"] == 0)
        print(f"{name}: hit t bit-equal {np.mean(t[surf] == g['hits_t'][surf]):.6f} | 1 spp: bit-identical "
              f"{np.mean((s1.reshape(-1, 3) == g['spp1'].reshape(-1, 3)).all(axis=1)):.6f}, within 1e-3 {radiance_agreement(s1, g['spp1']):.6f}"
              f" | {n} spp within 1e-3 {radiance_agreement(sN, g['sppN']):.6f}")


if __name__ == "__main__":
    main()
