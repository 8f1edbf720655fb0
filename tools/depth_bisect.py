"""Development aid: compare the CPU oracle with the reference-on-llvmpipe on a golden scene at a chosen maxDepth / options,
to localise which stage of the path first departs bit for bit.

    python tools/depth_bisect.py c2mini 1 [2 3 ...]        (maxDepth values)
Needs the authoring container (oracle/_ref, lavaframe_b200/bin/lf_scenepack).
"""
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from scenes import gen_scenes  # noqa: E402
from oracle_api import Oracle  # noqa: E402

REFBIN = os.path.join(ROOT, "oracle", "_ref", "lf_ref_llvmpipe")
PACKBIN = os.path.join(ROOT, "lavaframe_b200", "bin", "lf_scenepack")
BUILDERS = {"cornell": gen_scenes.cornell_256, "c2mini": gen_scenes.c2_mini, "c3mini": gen_scenes.c3_mini}


def compare(name, depth, shaders=None, keep=None):
    with tempfile.TemporaryDirectory() as tmp:
        scene = BUILDERS[name](os.path.join(tmp, "assets"))
        text = open(scene).read()
        text2 = re.sub(r"maxDepth \d+", f"maxDepth {depth}", text, flags=re.I)
        assert text2 != text or f"maxDepth {depth}" in text
        open(scene, "w").write(text2)
        pack = os.path.join(tmp, "s.lfpack")
        subprocess.run([PACKBIN, scene, pack], check=True, capture_output=True)
        cmd = [REFBIN, "--scene", scene, "--spp", "1", "--out", os.path.join(tmp, "s1.f32"), "--timing-json"]
        if shaders:
            cmd += ["--shaders", shaders]
        res = subprocess.run(cmd, env=gen_scenes.llvmpipe_env(), check=True, capture_output=True, text=True)
        info = json.loads(res.stdout.strip().splitlines()[-1])
        W, H = info["width"], info["height"]
        ref = np.fromfile(os.path.join(tmp, "s1.f32"), np.float32).reshape(H, W, 3)
        o = Oracle(pack)
        t, tri, mat, em = o.primary_hits(2)
        img = o.render_frames(2, 1)
        o.close()
    a, b = img.reshape(-1, 3), ref.reshape(-1, 3)
    exact = (a == b).all(axis=1)
    close = (np.abs(a.astype(np.float64) - b) <= 1e-3 * np.abs(b) + 1e-6).all(axis=1)
    key = np.where(em.reshape(-1) > 0, -2, np.where(t.reshape(-1) >= 1e6, -1, mat.reshape(-1)))
    print(f"{name} maxDepth {depth}: bit-identical {exact.mean():.6f}, within 1e-3 {close.mean():.6f}")
    for k in np.unique(key):
        m = key == k
        print(f"   primary mat {k:3d}: n={m.sum():6d} exact {exact[m].mean():.4f} close {close[m].mean():.4f}")
    if keep:
        np.savez(keep, oracle=img, ref=ref, mat=key.reshape(H, W))
    return exact, close


if __name__ == "__main__":
    for d in sys.argv[2:]:
        compare(sys.argv[1], int(d))
