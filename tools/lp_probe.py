"""Development aid: evaluate a GLSL expression on Mesa llvmpipe (oracle/_ref/lp_probe) over numpy inputs.

    from tools.lp_probe import glsl
    out = glsl("vec4(sin(a.x), cos(a.x), 0, 0)", a)      # a: (N,4) float32 -> (N,4) float32
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scenes import gen_scenes  # noqa: E402

PROBE = os.path.join(ROOT, "oracle", "_ref", "lp_probe")


def glsl(expr, a, pre="", tex8=None, texf=None):
    """a: (N,) or (N, 4) -> `a`; or (N, K, 4): K vec4 per item, read in the shader with A(0) .. A(K-1)."""
    a = np.ascontiguousarray(a, np.float32)
    if a.ndim == 1:
        a = np.stack([a, np.zeros_like(a), np.zeros_like(a), np.zeros_like(a)], axis=1)
    n = a.shape[0]
    stride = a.shape[1] if a.ndim == 3 else 1
    with tempfile.TemporaryDirectory() as tmp:
        a.tofile(os.path.join(tmp, "in.f32"))
        cmd = [PROBE, "--in", os.path.join(tmp, "in.f32"), "--n", str(n), "--expr", expr, "--out", os.path.join(tmp, "out.f32")]
        if pre:
            cmd += ["--pre", pre]
        if stride > 1:
            cmd += ["--stride", str(stride)]
        if tex8 is not None:                         # (L, H, W, 4) uint8
            t = np.ascontiguousarray(tex8, np.uint8)
            t.tofile(os.path.join(tmp, "t8.bin"))
            cmd += ["--tex8", str(t.shape[2]), str(t.shape[1]), str(t.shape[0]), os.path.join(tmp, "t8.bin")]
        if texf is not None:                         # (H, W, 3) float32
            t = np.ascontiguousarray(texf, np.float32)
            t.tofile(os.path.join(tmp, "tf.bin"))
            cmd += ["--texf", str(t.shape[1]), str(t.shape[0]), os.path.join(tmp, "tf.bin")]
        subprocess.run(cmd, env=gen_scenes.llvmpipe_env(), check=True)
        return np.fromfile(os.path.join(tmp, "out.f32"), np.float32).reshape(n, 4)
