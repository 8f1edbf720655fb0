#!/usr/bin/env python
"""Regenerates tests/golden/llvmpipe_builtins.npz: GLSL built-ins EXECUTED ON THE TIER-1 LLVMPIPE (oracle/_ref/lp_probe), the
fixture that pins the oracle's restatement of sin/cos/tan/exp/log/pow/acos/atan, `/`, mix, refract, reflect, normalize,
length, inversesqrt and the RGBA8 LINEAR/REPEAT texture filter bit for bit (tests/test_oracle_golden.py::test_builtins_*).

Runs only in the authoring container (needs oracle/_ref/lp_probe, built by `make -C oracle ref`, and the Mesa libGL that
ships with Nsight Compute).  Inputs are seeded; the expressions are the GROUPS below, mirrored by Oracle::BuiltinKat.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tools.lp_probe import glsl  # noqa: E402

N = 4096
GROUPS = {
    0: "vec4(sin(a.x), cos(a.x), tan(a.y), sqrt(abs(a.z)))",
    1: "vec4(exp(a.x), log(a.y), pow(a.y, a.z), acos(a.w))",
    2: "vec4(atan(a.x, a.y), a.x / a.y, mix(a.x, a.y, a.z), inversesqrt(abs(a.w)))",
    3: "vec4(refract(normalize(a.xyz), normalize(a.zxy * vec3(1, -1, 1) + 0.1), a.w), 0)",
    4: "vec4(reflect(normalize(a.xyz), normalize(a.zxy * vec3(1, -1, 1) + 0.1)), length(a.xyz))",
    5: "texture(tex8, a.xyz)",
}


def inputs(rng):
    f = np.float32
    u = lambda lo, hi: rng.uniform(lo, hi, N).astype(f)  # noqa: E731
    a = {}
    a[0] = np.stack([u(-7, 7), u(0.02, 1.5), u(0, 4), u(0, 1)], 1)                       # angles of the path: [0, 2 pi] and fov / 2
    x1 = np.concatenate([u(-30, 5)[: N // 2], -np.exp(rng.uniform(-8, 5, N - N // 2)).astype(f)])
    y1 = np.concatenate([u(0, 1)[: N // 2], np.exp(rng.uniform(-14, 3, N - N // 2)).astype(f)])  # colours, roughness^2 ... wide range
    z1 = rng.choice(np.array([2.2, 1 / 2.2, 0.5, 5.0], f), N)
    z1[::3] = u(0, 1)[::3]                                                             # pow(a2, 1 - r1)
    a[1] = np.stack([x1, y1, z1, u(-1, 1)], 1)
    a[1][:8, 3] = [1, -1, 0, 0.99999994, -0.99999994, 0.5, -0.5, 1e-8]
    a[2] = np.stack([u(-1, 1), u(-1, 1), u(0, 1), np.exp(rng.uniform(-10, 10, N)).astype(f)], 1)
    a[2][:6, :2] = [[0, 1], [1, 0], [0, -1], [-1, 0], [0.5, 0.5], [-0.5, 0.5]]
    v = rng.normal(size=(N, 3)).astype(f)
    a[3] = np.concatenate([v, rng.choice(np.array([1.45, 1 / 1.45, 1.0, 1.33, 1 / 1.33], f), N)[:, None]], 1)
    a[4] = np.concatenate([rng.normal(size=(N, 3)).astype(f), np.zeros((N, 1), f)], 1)
    a[5] = np.stack([u(-1.5, 2.5), u(-1.5, 2.5), rng.integers(0, 2, N).astype(f), np.zeros(N, f)], 1)
    return a


def main():
    rng = np.random.default_rng(20261017)
    a = inputs(rng)
    tex = rng.integers(0, 256, (2, 8, 16, 4), dtype=np.uint8)                          # 2 layers of 16 x 8 RGBA8
    out = {"tex": tex, "expr": np.array([GROUPS[k] for k in sorted(GROUPS)])}
    for k, expr in GROUPS.items():
        res = glsl(expr, a[k], tex8=tex if k == 5 else None)
        out[f"in{k}"] = a[k]
        out[f"out{k}"] = res
        print(k, expr, "->", res[:2].tolist())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "llvmpipe_builtins.npz"), **out)


if __name__ == "__main__":
    main()
