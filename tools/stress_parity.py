"""Development aid: held-out parity stress.  Generates scenes the restatement was never fitted on - grids of spheres with RANDOM Disney
parameters (every lobe and parameter of Material.h, emission, all four texture kinds, HDR env, quad + sphere lights, thin lens) - renders
them with the unmodified reference on llvmpipe and with the oracle, and reports bit identity.

    python tools/stress_parity.py [seed ...]          (authoring container only: needs oracle/_ref and lavaframe_b200/bin/lf_scenepack)
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from scenes import gen_scenes as g  # noqa: E402
from oracle_api import Oracle  # noqa: E402

REFBIN = os.path.join(ROOT, "oracle", "_ref", "lf_ref_llvmpipe")
PACKBIN = os.path.join(ROOT, "lavaframe_b200", "bin", "lf_scenepack")


def variant(seed):
    """Seeds >= 100 also vary the shader's #defines: lights / env on or off, Russian roulette off or from another depth, constant background."""
    rng = np.random.RandomState(seed + 7919)
    v = dict(lights=True, env=True, rr=None, bg=None)
    if seed >= 100:
        v["lights"] = rng.rand() < 0.6
        v["env"] = (rng.rand() < 0.6) or not v["lights"]
        v["rr"] = [None, "off", 0, 1, 3][rng.randint(5)]
        if rng.rand() < 0.35:
            v["bg"] = tuple(np.round(rng.uniform(0.1, 0.9, 3), 3))
    return v


def scene(outdir, seed):
    rng = np.random.RandomState(seed)
    var = variant(seed)
    assets = g.reference_assets(outdir)
    v, n, t = g.displaced_sphere(16, 24, 1.0, amp=0.2 * rng.rand(), seed=seed)
    g.write_obj(os.path.join(assets, "s_ball.obj"), v, n, t)
    g.write_floor(os.path.join(assets, "s_floor.obj"), half=8.0, uvscale=3.0)
    for k, kind in enumerate(("albedo", "mr", "normal", "albedo")):
        g.write_png_blocks(os.path.join(assets, f"s_tex{k}.png"), 64, int(rng.choice([1, 2, 4, 8])), seed * 10 + k, kind)
    g.write_hdr(os.path.join(assets, "s_sky.hdr"), g.sky_image(64, 32, sun_radiance=float(rng.choice([50.0, 5e3]))))
    extra = []
    if var["rr"] == "off":
        extra.append("\tenableRR False")
    elif var["rr"] is not None:
        extra.append(f"\tRRDepth {var['rr']}")
    s = g._renderer(256, 128, int(rng.choice([3, 5, 8])), hdr="s_sky.hdr" if var["env"] else None, extra=extra)
    s += g._camera((0, 3.0 + rng.rand(), -8), (0, 0.5, 0), 40, aperture=float(rng.choice([0.0, 0.03])), focal=7.5)
    s += g._material("ground", albedo=(1, 1, 1), roughness=0.6, albedoTexture="s_tex0.png", normalTexture="s_tex2.png", emissionTexture="s_tex3.png"
                     if rng.rand() < 0.5 else "s_tex0.png")
    names = []
    for i in range(12):
        u = rng.rand
        kw = dict(albedo=tuple(np.round(rng.uniform(0.05, 1.0, 3), 3)), metallic=float(rng.choice([0, 0, 0.5, 1])), roughness=round(float(u() ** 2), 3),
                  subsurface=round(float(u() * (u() < 0.4)), 3), specular=round(float(u()), 3), specularTint=round(float(u() * (u() < 0.5)), 3),
                  sheen=round(float(u() * (u() < 0.5)), 3), sheenTint=round(float(u()), 3), clearcoat=round(float(u() * (u() < 0.5)), 3),
                  clearcoatRoughness=round(float(u()), 3), transmission=float(rng.choice([0, 0, 0.5, 1])), ior=round(float(rng.uniform(1.05, 2.2)), 3),
                  extinction=tuple(np.round(rng.uniform(0.3, 1.0, 3), 3)))
        if u() < 0.15:
            kw["emission"] = tuple(np.round(rng.uniform(0, 3.0, 3), 3))
        if u() < 0.25:
            kw["albedoTexture"] = "s_tex0.png"
        if u() < 0.2:
            kw["metallicRoughnessTexture"] = "s_tex1.png"
        if u() < 0.2:
            kw["normalTexture"] = "s_tex2.png"
        names.append(f"m{i}")
        s += g._material(f"m{i}", **kw)
    s += g._mesh("s_floor.obj", "ground", (0, 0, 0), (1, 1, 1))
    for gz in range(5):
        for gx in range(6):
            sc = rng.uniform(0.35, 0.7, 3) if rng.rand() < 0.5 else np.full(3, rng.uniform(0.4, 0.7))
            s += g._mesh("s_ball.obj", names[rng.randint(len(names))], ((gx - 2.5) * 1.5, float(sc[1]) + 0.01, (gz - 2) * 1.5), tuple(sc))
    if var["lights"]:
        s += g._quad_light((-2, 5, -1), (-2, 5, 1), (0.5, 5, -1), (18, 17, 15))
        s += g._sphere_light((3, 3, -3), 0.35, (20, 22, 25))
    path = os.path.join(assets, "stress.scene")
    open(path, "w").write(s)
    return path


def main():
    for seed in [int(a) for a in sys.argv[1:]] or [1, 2, 3]:
        with tempfile.TemporaryDirectory() as tmp:
            sc = scene(os.path.join(tmp, "assets"), seed)
            pack = os.path.join(tmp, "s.lfpack")
            subprocess.run([PACKBIN, sc, pack], check=True, capture_output=True)
            res = {}
            var = variant(seed)
            bgargs = ["--bg"] + [str(x) for x in var["bg"]] if var["bg"] else []
            for spp in (1, 4):
                out = os.path.join(tmp, f"s{spp}.f32")
                r = subprocess.run([REFBIN, "--scene", sc, "--spp", str(spp), "--out", out, "--timing-json"] + bgargs, env=g.llvmpipe_env(), check=True, capture_output=True, text=True)
                info = json.loads(r.stdout.strip().splitlines()[-1])
                res[spp] = np.fromfile(out, np.float32).reshape(info["height"], info["width"], 3)
            o = Oracle(pack)
            if var["bg"]:
                o.params.use_constant_bg = 1
                o.params.bg_color[0], o.params.bg_color[1], o.params.bg_color[2] = var["bg"]
                o.update_params()
            t, tri, mat, em = o.primary_hits(2)
            depth = o.params.max_depth
            a1 = o.render_frames(2, 1)
            a4 = o.render_frames(2, 4) / np.float32(4)
            o.close()
        surf = em == 0
        e1 = (a1 == res[1]).all(axis=2)
        e4 = (a4 == res[4]).all(axis=2)
        print(f"seed {seed} (depth {depth}, {' '.join(f'{k}={v}' for k, v in var.items())}): 1 spp bit-identical {e1.mean():.6f} (non-emitter pixels {e1[surf].mean():.6f}) | 4 spp {e4.mean():.6f} "
              f"(non-emitter {e4[surf].mean():.6f}) | NaN pixels ref {int(np.isnan(res[1]).any(axis=2).sum())} oracle {int(np.isnan(a1).any(axis=2).sum())}", flush=True)


if __name__ == "__main__":
    main()
