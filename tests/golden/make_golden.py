#!/usr/bin/env python
"""Regenerates the golden fixtures of tests/golden/ from THE REFERENCE ITSELF.

Runs only in the authoring container (needs /root/reference, built into oracle/_ref by `make -C oracle ref`
and lavaframe_b200/bin/lf_scenepack by CMake):

  * <name>.lfpack            the flattened scene arrays, produced by the reference's unchanged loader + BVH builder
  * <name>_llvmpipe.npz      outputs of the reference's unmodified TiledRenderer + GLSL on Mesa llvmpipe:
        hits_t / hits_tri / hits_mat / hits_emitter   first camera ray of frame 2 ("--probe hits")
        spp1                                           1-spp radiance (frame 2), W*H*3 float32, rows bottom-up
        sppN                                           N-spp mean (frames 2..N+1)

  * cornell_llvmpipe_4096spp.npz   the converged reference image for the north_star's third check (20 min of llvmpipe):
        spp4096 = mean of frames 2..4097

  * <name>_llvmpipe_preview.npz    the preview engine's image (shaders/preview_flareon.glsl drawn by TiledRenderer::Render
        while camera->isMoving, read back from previewFBO): half = previewScale 0.5, maxDepth 2;
        full_dof = previewScale 1.0 with "#define USE_DOF"

  * cornell64_llvmpipe_post.npz    the post-process pass: every tonemapper, vignette, chromatic aberration (see post_goldens)

Usage:  python tests/golden/make_golden.py [cornell] [c2mini] [c3mini] [c4gold] [cornell4096] [preview] [tiled] [post] ...
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from scenes import gen_scenes  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "lf_ref_llvmpipe")
PACKBIN = os.path.join(ROOT, "lavaframe_b200", "bin", "lf_scenepack")

# name -> (scene builder, N for the N-spp mean)
SCENES = {
    "cornell": (gen_scenes.cornell_256, 64),
    "c2mini": (gen_scenes.c2_mini, 16),
    "c3mini": (gen_scenes.c3_mini, 16),
    "c4gold": (gen_scenes.c4_gold, 8),
}


def run_ref(scene, spp, out, probe=None):
    cmd = [REFBIN, "--scene", scene, "--spp", str(spp), "--out", out, "--timing-json"]
    if probe:
        cmd += ["--probe", probe]
    res = subprocess.run(cmd, env=gen_scenes.llvmpipe_env(), check=True, capture_output=True, text=True)
    return json.loads(res.stdout.strip().splitlines()[-1])


def converged_cornell(spp=4096):
    with tempfile.TemporaryDirectory() as tmp:
        scene = gen_scenes.cornell_256(os.path.join(tmp, "assets"))
        info = run_ref(scene, spp, os.path.join(tmp, "conv.f32"))
        img = np.fromfile(os.path.join(tmp, "conv.f32"), np.float32).reshape(info["height"], info["width"], 3)
        np.savez_compressed(os.path.join(GOLD, f"cornell_llvmpipe_{spp}spp.npz"), **{f"spp{spp}": img, "nspp": np.int32(spp)})
        print("cornell", spp, "spp mean", img.mean(axis=(0, 1)), "render_s", info["render_s"])


def tiled_cornell(spp=4, tile=64):
    """cornell_tiled_llvmpipe.npz: the Cornell scene with `tileWidth 64 / tileHeight 64` (Loader.cpp:256-261), `spp` samples =
    spp * 16 tile steps; every tile step advances `frame` (TiledRenderer.cpp:485-487), so the RNG seeds differ from the untiled run."""
    with tempfile.TemporaryDirectory() as tmp:
        scene = gen_scenes.cornell_256(os.path.join(tmp, "assets"))
        text = open(scene).read()
        tiled = os.path.join(os.path.dirname(scene), "cornell_tiled.scene")
        open(tiled, "w").write(text.replace("resolution 256 256", f"resolution 256 256\n\ttileWidth {tile}\n\ttileHeight {tile}"))
        out = os.path.join(tmp, "tiled.f32")
        info = run_ref(tiled, spp, out)
        img = np.fromfile(out, np.float32).reshape(info["height"], info["width"], 3)
        np.savez_compressed(os.path.join(GOLD, "cornell_tiled_llvmpipe.npz"), sppN=img, nspp=np.int32(spp), tile=np.int32(tile))
        print("cornell tiled", spp, "spp mean", img.mean(axis=(0, 1)), info)


POST_CASES = {   # name -> (tonemapIndex, vignette (intensity, power) or None, chromatic aberration (distortion, distance, p1, p2, p3) or None)
    "tm1": (1, None, None), "tm2": (2, None, None), "tm3": (3, None, None), "tm4": (4, None, None), "tm5": (5, None, None),
    "tm6": (6, None, None), "tm2_vig": (2, (0.6, 1.5), None), "ca0": (0, None, (0, 0.05, 5.0, -0.5, 0.5)),
    "ca1": (0, None, (1, 0.05, 5.0, -0.5, 0.5)), "tm3_ca1_vig": (3, (0.4, 2.0), (1, 0.08, 3.0, -0.7, 0.45)),
}


def post_goldens(spp=4):
    """cornell64_llvmpipe_post.npz: GetOutputBufferHDR of the reference for every tonemapper, the vignette and both chromatic
    aberration modes (RenderOptions the UI sets, Main.cpp:470-500; postprocess.glsl) on the Cornell box at 64x64, `spp` samples.
    tm0 (identity) * spp is the accumulation buffer the other images were made from (the scale is a power of two: exact)."""
    with tempfile.TemporaryDirectory() as tmp:
        scene = gen_scenes.cornell_256(os.path.join(tmp, "assets"), res=(64, 64))
        arrays = {"nspp": np.int32(spp)}

        def run(extra, u8=None):
            out = os.path.join(tmp, "o.f32")
            cmd = [REFBIN, "--scene", scene, "--spp", str(spp), "--out", out, "--timing-json"] + extra
            if u8:
                cmd += ["--out8", os.path.join(tmp, "o.u8")]
            res = subprocess.run(cmd, env=gen_scenes.llvmpipe_env(), check=True, capture_output=True, text=True)
            info = json.loads(res.stdout.strip().splitlines()[-1])
            if u8:   # GetOutputBuffer: the 8-bit RGB image SaveFrame / SaveFrameJPG / ... of Export.h write
                arrays[u8] = np.fromfile(os.path.join(tmp, "o.u8"), np.uint8).reshape(info["height"], info["width"], 3)
            return np.fromfile(out, np.float32).reshape(info["height"], info["width"], 3)

        arrays["tm0"] = run([])
        for name, (tm, vig, ca) in POST_CASES.items():
            extra = ["--tonemap", str(tm)]
            if vig:
                extra += ["--vignette", str(vig[0]), str(vig[1])]
            if ca:
                extra += ["--ca"] + [str(v) for v in ca]
            arrays[name] = run(extra, u8=f"{name}_u8" if name in ("tm2", "tm5") else None)
            print("post", name, arrays[name].mean(axis=(0, 1)))
        np.savez_compressed(os.path.join(GOLD, "cornell64_llvmpipe_post.npz"), **arrays)


def preview_goldens(only=None):
    for name, (builder, _) in SCENES.items():
        if only and name not in only:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            scene = builder(os.path.join(tmp, "assets"))
            arrays = {}
            for key, scale, dof in (("half", 0.5, False), ("full_dof", 1.0, True)):
                out = os.path.join(tmp, key + ".f32")
                cmd = [REFBIN, "--scene", scene, "--out", out, "--preview", str(scale)] + (["--preview-dof"] if dof else [])
                res = subprocess.run(cmd, env=gen_scenes.llvmpipe_env(), check=True, capture_output=True, text=True)
                line = [l for l in res.stdout.splitlines() if l.startswith("preview ")][-1]
                pw, ph = (int(v) for v in line.split()[1].split("x"))
                arrays[key] = np.fromfile(out, np.float32).reshape(ph, pw, 3)
                print(name, line, "mean", arrays[key].mean(axis=(0, 1)))
            np.savez_compressed(os.path.join(GOLD, f"{name}_llvmpipe_preview.npz"), **arrays)


def main(names):
    if "preview" in names:                      # `preview` alone = every scene; `preview c4gold` = that scene's preview only
        rest = [n for n in names if n != "preview"]
        preview_goldens(rest or None)
        if rest:
            return
    if "post" in names:
        post_goldens()
        names = [n for n in names if n != "post"]
    if "tiled" in names:
        tiled_cornell()
        names = [n for n in names if n != "tiled"]
    if "cornell4096" in names:
        converged_cornell(4096)
        names = [n for n in names if n != "cornell4096"]
    for name in names:
        builder, nspp = SCENES[name]
        with tempfile.TemporaryDirectory() as tmp:
            scene = builder(os.path.join(tmp, "assets"))
            pack = os.path.join(GOLD, f"{name}.lfpack")
            subprocess.run([PACKBIN, scene, pack], check=True)
            info = run_ref(scene, 1, os.path.join(tmp, "hits.f32"), probe="hits")
            W, H = info["width"], info["height"]
            hits = np.fromfile(os.path.join(tmp, "hits.f32"), np.float32).reshape(H, W, 3)
            run_ref(scene, 1, os.path.join(tmp, "s1.f32"))
            s1 = np.fromfile(os.path.join(tmp, "s1.f32"), np.float32).reshape(H, W, 3)
            tN = run_ref(scene, nspp, os.path.join(tmp, "sN.f32"))
            sN = np.fromfile(os.path.join(tmp, "sN.f32"), np.float32).reshape(H, W, 3)
            assert W & (W - 1) == 0 and H & (H - 1) == 0, "golden scenes must have power-of-two sizes (exact LINEAR reads)"
            tri_f = hits[..., 1]
            emitter = (np.abs(tri_f - np.floor(tri_f)) == 0.5)
            assert np.all(emitter | (tri_f == np.floor(tri_f)))
            tri = np.where(emitter, -1, np.floor(tri_f)).astype(np.int32)
            np.savez_compressed(os.path.join(GOLD, f"{name}_llvmpipe.npz"),
                                hits_t=hits[..., 0], hits_tri=tri, hits_mat=hits[..., 2].astype(np.int32),
                                hits_emitter=emitter.astype(np.int8), spp1=s1, sppN=sN, nspp=np.int32(nspp),
                                gl_renderer=np.bytes_(tN["gl_renderer"]), gl_version=np.bytes_(tN["gl_version"]))
            print(name, "->", pack, f"{W}x{H}", "mean", s1.mean(axis=(0, 1)), "steady samples/s", tN["samples_per_s_steady"])


if __name__ == "__main__":
    main(sys.argv[1:] or ["cornell"])
