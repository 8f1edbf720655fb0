#!/usr/bin/env python
"""Regenerates tests/golden/llvmpipe_bsdf.npz: the reference's OWN BSDF functions (shaders/common/disney.glsl, sampling.glsl, with
globals.glsl and uniforms.glsl in front, text unmodified) EXECUTED ON THE TIER-1 LLVMPIPE on seeded random arguments - DisneyEval,
DisneySample (direction, pdf, value and the RNG state it leaves), GTR1 / GTR2 / SmithG_GGX / DielectricFresnel, the three importance
samplers.  This catches what whole-image goldens can miss: what Mesa's compiler does to an expression (e.g. it folds the two
constants of `PI * log(a2)` into one, because log(x) is lowered to log2(x) * ln 2 first) for EVERY parameter combination.

Runs only in the authoring container (needs /root/reference and oracle/_ref/lp_probe).  Item layout: Oracle::BsdfKat.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tools.lp_probe import glsl  # noqa: E402

REF = "/root/reference/shaders/common/"
N = 2048

GLUE = """
State mkState() {
    State s;
    s.mat.albedo = A(3).xyz; s.mat.specular = A(3).w;
    s.mat.metallic = A(4).x; s.mat.roughness = A(4).y; s.mat.subsurface = A(4).z; s.mat.specularTint = A(4).w;
    s.mat.sheen = A(5).x; s.mat.sheenTint = A(5).y; s.mat.clearcoat = A(5).z; s.mat.clearcoatRoughness = A(5).w;
    s.mat.specTrans = A(6).x; s.eta = A(6).y;
    s.normal = A(1).xyz; s.ffnormal = A(1).xyz; s.tangent = A(7).xyz; s.bitangent = A(8).xyz;
    seed = uvec4(uint(A(7).w), uint(A(8).w), uint(A(6).z), uint(A(6).w));
    return s;
}
vec4 kat(int op) {
    State s = mkState();
    float pdf = 0.0;
    vec3 L = vec3(0.0);
    if (op == 0) { vec3 f = DisneyEval(s, A(0).xyz, A(1).xyz, A(2).xyz, pdf); return vec4(f, pdf); }
    if (op == 1) { DisneySample(s, A(0).xyz, A(1).xyz, L, pdf); return vec4(L, pdf); }
    if (op == 2) { vec3 f = DisneySample(s, A(0).xyz, A(1).xyz, L, pdf); return vec4(f, rand()); }
    if (op == 3) return vec4(GTR1(A(0).x, A(0).y), GTR2(A(0).x, A(0).y), SmithG_GGX(A(0).x, A(0).y), DielectricFresnel(A(0).x, A(6).y));
    vec3 h1 = ImportanceSampleGTR1(A(4).y, A(0).x, A(0).y), h2 = ImportanceSampleGTR2(A(4).y, A(0).x, A(0).y), c = CosineSampleHemisphere(A(0).x, A(0).y);
    return vec4(h1.x + h1.z, h2.x + h2.z, c.x + c.z, h1.y + h2.y + c.y);
}
"""


def reference_glsl():
    text = ""
    for f in ("globals.glsl", "uniforms.glsl", "sampling.glsl", "disney.glsl"):
        text += "\n".join(ln for ln in open(REF + f).read().splitlines() if not ln.strip().startswith("#include")) + "\n"
    return text + GLUE


def items(rng, n, unit_args):
    f = np.float32
    a = np.zeros((n, 9, 4), f)

    def unit(v):
        return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(f)

    Nn = unit(rng.normal(size=(n, 3)))
    up = np.where(np.abs(Nn[:, 2:3]) < 0.999, np.array([[0, 0, 1.0]]), np.array([[1.0, 0, 0]]))
    T = unit(np.cross(up, Nn))
    B = np.cross(Nn, T).astype(f)
    V = unit(Nn + rng.normal(size=(n, 3)) * 0.9)
    flip = (V * Nn).sum(axis=1) < 0
    V[flip] = -V[flip]                                   # V on the ffnormal side, as on the path
    L = unit(rng.normal(size=(n, 3)))
    if unit_args:                                        # ops 3 / 4: a.x, a.y in [0, 1] (cosines / random numbers), a.y bounded away from 0
        a[:, 0, 0] = rng.uniform(0, 1, n)
        a[:, 0, 1] = rng.uniform(0.001, 1, n)
    else:
        a[:, 0, :3] = V
    a[:, 1, :3] = Nn
    a[:, 2, :3] = L
    a[:, 3, :3] = rng.uniform(0.02, 1, (n, 3))
    a[:, 3, 3] = rng.uniform(0, 1, n)
    a[:, 4] = rng.uniform(0, 1, (n, 4))
    a[:, 4, 0] = rng.choice([0, 0, 0.3, 1], n)
    a[:, 4, 1] = np.maximum(rng.uniform(0, 1, n) ** 2, 0.001)
    a[:, 4, 2] *= rng.uniform(0, 1, n) < 0.5
    a[:, 5] = rng.uniform(0, 1, (n, 4))
    a[:, 5, 0] *= rng.uniform(0, 1, n) < 0.5
    a[:, 5, 2] *= rng.uniform(0, 1, n) < 0.6
    a[:, 6, 0] = rng.choice([0, 0, 0.5, 1], n)
    ior = rng.uniform(1.05, 2.2, n)
    a[:, 6, 1] = np.where(rng.uniform(0, 1, n) < 0.5, ior, 1.0 / ior)
    a[:, 7, :3] = T
    a[:, 8, :3] = B
    seeds = rng.integers(0, 1 << 20, (n, 4))
    a[:, 7, 3], a[:, 8, 3], a[:, 6, 2], a[:, 6, 3] = seeds[:, 0], seeds[:, 1], seeds[:, 2], seeds[:, 3]
    return a


def main():
    rng = np.random.default_rng(20261018)
    pre = reference_glsl()
    out = {}
    for op in range(5):
        a = items(rng, N, unit_args=op >= 3)
        res = glsl(f"kat({op})", a, pre=pre)
        out[f"in{op}"], out[f"out{op}"] = a, res
        print(op, res[:2].tolist())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "llvmpipe_bsdf.npz"), **out)


if __name__ == "__main__":
    main()
