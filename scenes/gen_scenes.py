#!/usr/bin/env python
"""Deterministic generators of the benchmark / parity scenes (BASELINE.json `configs`).

Every scene is a LavaFrame scene file (grammar: LavaFrame/Loader.cpp:45-386) plus the OBJ /
PNG / Radiance .hdr assets it names, written under a caller-given directory.  Nothing is downloaded; the only
inputs not generated here are the reference's own Cornell-box assets (build_include/assets), which are unpacked
from the oracle/_ref harness binary (`lf_ref_llvmpipe --extract-assets`) because /root/reference does not exist
on the GPU box.

    C1  cornell_256   the reference's cornell_box.lfs with `resolution 256 256`
    C2  c2_full       "dragon-class" closed displaced surface (~0.87 M triangles), metal + glass instances, ground
                      quad, procedural 2048x1024 HDR env with a small sun, 1280x720
        c2_mini       the same ingredients inside llvmpipe's limits (for parity against the real reference)
    C3  c3_full       instanced multi-mesh scene, textured materials, 2 quad lights + 1 sphere light, 1920x1080
        c3_mini       the same at 512x256
    C4  c4_stress     36x36 instances of glass_sphere.obj (15 872 tris -> 20.57 M instanced), depth 8, 3840x2160
        c4_mini       the same geometry at 480x270 (llvmpipe-comparable)

    python scenes/gen_scenes.py <name> <outdir> [--pack]      # writes the scene, optionally the .lfpack
"""
import glob
import math
import os
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFBIN = os.path.join(ROOT, "oracle", "_ref", "lf_ref_llvmpipe")
PACKBIN = os.path.join(ROOT, "lavaframe_b200", "bin", "lf_scenepack")


# ------------------------------------------------------------------------------------------------ environment
def mesa_dir():
    """The software GL bundled with Nsight Compute."""
    hits = sorted(glob.glob("/opt/nvidia/nsight-compute/*/host/linux-desktop-glibc_2_11_3-x64/Mesa/libGL.so.1"))
    return os.path.dirname(hits[-1]) if hits else None


def llvmpipe_env(threads=None):
    env = dict(os.environ)
    parts = [os.path.join(ROOT, "oracle", "_ref")]
    if mesa_dir():
        parts.append(mesa_dir())
    if env.get("LD_LIBRARY_PATH"):
        parts.append(env["LD_LIBRARY_PATH"])
    env["LD_LIBRARY_PATH"] = ":".join(parts)
    if threads:
        env["LP_NUM_THREADS"] = str(threads)
    return env


def reference_assets(outdir):
    """Unpack the reference's build_include/assets tree under outdir; returns outdir/assets."""
    dst = os.path.join(outdir, "assets")
    if not os.path.exists(os.path.join(dst, "cornell_box.lfs")):
        os.makedirs(outdir, exist_ok=True)
        if os.path.exists("/root/reference/build_include/assets/cornell_box.lfs"):
            subprocess.run(["cp", "-r", "/root/reference/build_include/assets", outdir], check=True)
            subprocess.run(["chmod", "-R", "u+w", dst], check=True)
        else:
            subprocess.run([REFBIN, "--extract-assets", outdir], check=True, env=llvmpipe_env())
    return dst


def write_pack(scene_path, pack_path):
    # the tool's summary line goes to OUR stderr: callers such as bench.py own stdout (one JSON line)
    subprocess.run([PACKBIN, scene_path, pack_path], check=True, stdout=sys.stderr)
    return pack_path


# ------------------------------------------------------------------------------------------------ asset writers
def write_obj(path, verts, normals, tris, uvs=None):
    """OBJ with one normal (and optionally one vt) per vertex: f a/a/a or a//a (Mesh.cpp:49-88 needs vn)."""
    with open(path, "w") as f:
        np.savetxt(f, verts, fmt="v %.7g %.7g %.7g")
        if uvs is not None:
            np.savetxt(f, uvs, fmt="vt %.7g %.7g")
        np.savetxt(f, normals, fmt="vn %.7g %.7g %.7g")
        t = np.asarray(tris, np.int64) + 1
        if uvs is not None:
            np.savetxt(f, np.repeat(t, 3, axis=1), fmt="f %d/%d/%d %d/%d/%d %d/%d/%d")
        else:
            np.savetxt(f, np.repeat(t, 2, axis=1), fmt="f %d//%d %d//%d %d//%d")


def displaced_sphere(stacks, slices, radius=1.0, amp=0.0, seed=7):
    """Closed UV-sphere-topology surface; with amp > 0 the radius is modulated by a fixed sum of spherical
    waves (a "dragon-class" bumpy closed mesh).  Triangles = 2 * slices * (stacks - 1)."""
    th = np.linspace(0.0, math.pi, stacks + 1)[1:-1]                  # interior rings
    ph = np.linspace(0.0, 2 * math.pi, slices, endpoint=False)
    T, P = np.meshgrid(th, ph, indexing="ij")

    def rad(t, p):
        if amp == 0.0:
            return np.full_like(t, radius)
        rng = np.random.RandomState(seed)
        r = np.full_like(t, radius)
        for k in range(10):
            a, b = rng.randint(2, 18), rng.randint(1, 14)
            ph0, th0 = rng.uniform(0, 2 * math.pi, 2)
            w = amp * rng.uniform(0.3, 1.0) / (1 + 0.15 * (a + b))
            r = r + w * np.sin(a * t + th0) * np.cos(b * p + ph0) * np.sin(t) ** 2
        return r

    def pos(t, p):
        r = rad(t, p)
        return np.stack([r * np.sin(t) * np.cos(p), r * np.cos(t), r * np.sin(t) * np.sin(p)], axis=-1)

    V = pos(T, P)
    eps = 1e-4
    dt = pos(T + eps, P) - pos(T - eps, P)
    dp = pos(T, P + eps) - pos(T, P - eps)
    N = np.cross(dp, dt)
    N /= np.linalg.norm(N, axis=-1, keepdims=True)
    verts = np.concatenate([V.reshape(-1, 3), [[0, rad(np.array(0.0), np.array(0.0)), 0]], [[0, -rad(np.array(math.pi), np.array(0.0)), 0]]])
    norms = np.concatenate([N.reshape(-1, 3), [[0, 1, 0]], [[0, -1, 0]]])
    R = stacks - 1
    idx = np.arange(R * slices).reshape(R, slices)
    nxt = np.roll(idx, -1, axis=1)
    a, b, c, d = idx[:-1], nxt[:-1], idx[1:], nxt[1:]
    quads = np.concatenate([np.stack([a, c, b], -1).reshape(-1, 3), np.stack([b, c, d], -1).reshape(-1, 3)])
    top, bot = R * slices, R * slices + 1
    capt = np.stack([np.full(slices, top), idx[0], nxt[0]], -1)
    capb = np.stack([np.full(slices, bot), nxt[-1], idx[-1]], -1)
    return verts.astype(np.float32), norms.astype(np.float32), np.concatenate([quads, capt, capb])


def write_floor(path, half=6.0, uvscale=1.0):
    v = np.array([[-half, 0, -half], [half, 0, -half], [half, 0, half], [-half, 0, half]], np.float32)
    n = np.tile(np.array([[0, 1, 0]], np.float32), (4, 1))
    uv = np.array([[0, 0], [uvscale, 0], [uvscale, uvscale], [0, uvscale]], np.float32)
    write_obj(path, v, n, [[0, 2, 1], [0, 3, 2]], uv)


def write_png_blocks(path, size, blocks, seed, kind="albedo"):
    """size x size RGBA8 patchwork of `blocks` x `blocks` flat tiles: bilinear footprints almost always see four
    equal texels, so fp32, hardware and llvmpipe's 8-bit filtering agree exactly."""
    from PIL import Image
    rng = np.random.RandomState(seed)
    if kind == "albedo":
        cols = rng.randint(40, 250, size=(blocks, blocks, 3))
    elif kind == "mr":                                # metallic in .x, roughness = y^2 (pathtrace.glsl:87-93)
        cols = np.zeros((blocks, blocks, 3), int)
        cols[..., 0] = rng.choice([0, 255], size=(blocks, blocks), p=[0.6, 0.4])
        cols[..., 1] = rng.randint(60, 230, size=(blocks, blocks))
    elif kind == "normal":                            # flat tangent-space normal
        cols = np.tile(np.array([128, 128, 255]), (blocks, blocks, 1))
    img = np.repeat(np.repeat(cols, size // blocks, axis=0), size // blocks, axis=1).astype(np.uint8)
    Image.fromarray(img, "RGB").save(path)


def write_hdr(path, rgb):
    """Radiance RGBE with new-style scanlines made of literal (non-run) packets, as hdrloader.cpp:225-266 decodes
    them; header = '#?RADIANCE\\n' + one line + empty line + '-Y h +X w\\n' (hdrloader.cpp:145-177)."""
    h, w, _ = rgb.shape
    m = rgb.max(axis=2)
    with np.errstate(divide="ignore", invalid="ignore"):
        f, e = np.frexp(m)
        scale = np.where(m > 1e-32, f * 256.0 / m, 0.0)
    rgbe = np.zeros((h, w, 4), np.uint8)
    rgbe[..., :3] = np.clip(rgb * scale[..., None], 0, 255).astype(np.uint8)
    rgbe[..., 3] = np.where(m > 1e-32, e + 128, 0).astype(np.uint8)
    with open(path, "wb") as fo:
        fo.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n")
        fo.write(f"-Y {h} +X {w}\n".encode())
        for y in range(h):
            if w < 8 or w > 0x7fff:
                fo.write(rgbe[y].tobytes())
                continue
            fo.write(bytes([2, 2, (w >> 8) & 0xff, w & 0xff]))
            for c in range(4):
                row = rgbe[y, :, c].tobytes()
                for s in range(0, w, 128):
                    chunk = row[s:s + 128]
                    fo.write(bytes([len(chunk)]) + chunk)


def sky_image(w, h, sun_radiance=5.0e4, sun_dir=(0.45, 0.55, -0.70), sun_cos=0.9995):
    """Sky gradient + ground + small sun disc (importance sampling matters).  Row 0 = +Y (theta = 0);
    direction convention of EnvSample: (-sin t cos p, cos t, -sin t sin p), p = 2 pi u (sampling.glsl:258-264)."""
    v = (np.arange(h) + 0.5) / h
    u = (np.arange(w) + 0.5) / w
    T, P = np.meshgrid(v * math.pi, u * 2 * math.pi, indexing="ij")
    d = np.stack([-np.sin(T) * np.cos(P), np.cos(T), -np.sin(T) * np.sin(P)], -1)
    up = d[..., 1]
    sky = np.where(up[..., None] > 0,
                   (1 - up[..., None]) * np.array([0.9, 0.9, 1.0]) + up[..., None] * np.array([0.25, 0.45, 0.95]),
                   np.array([0.18, 0.16, 0.14]) * np.ones_like(d))
    s = np.array(sun_dir) / np.linalg.norm(sun_dir)
    sun = (d @ s) > sun_cos
    img = sky * 1.0
    img[sun] = sun_radiance * np.array([1.0, 0.95, 0.85])
    return img.astype(np.float32)


# ------------------------------------------------------------------------------------------------ scene text helpers
def _material(name, **kw):
    lines = [f"material {name}", "{"]
    for k, v in kw.items():
        if isinstance(v, (tuple, list)):
            v = " ".join(f"{x:g}" for x in v)
        lines.append(f"\t{k} {v}")
    return "\n".join(lines + ["}", ""])


def _mesh(file, material, pos=(0, 0, 0), scale=(1, 1, 1)):
    return "\n".join(["mesh", "{", f"\tfile {file}", f"\tmaterial {material}",
                      "\tposition " + " ".join(f"{x:.6g}" for x in pos), "\tscale " + " ".join(f"{x:.6g}" for x in scale), "}", ""])


def _renderer(w, h, depth, hdr=None, extra=()):
    lines = ["Renderer", "{", f"\tresolution {w} {h}", f"\tmaxDepth {depth}"]
    if hdr:
        lines += [f"\thdriMap {hdr}", "\thdriMultiplier 1.0"]
    return "\n".join(lines + list(extra) + ["}", ""])


def _camera(pos, look, fov, aperture=None, focal=None):
    lines = ["Camera", "{", "\tposition " + " ".join(f"{x:g}" for x in pos), "\tlookAt " + " ".join(f"{x:g}" for x in look), f"\tfov {fov:g}"]
    if aperture is not None:
        lines += [f"\taperture {aperture:g}", f"\tfocalDistance {focal:g}"]
    return "\n".join(lines + ["}", ""])


def _quad_light(pos, v1, v2, emission):
    return "\n".join(["light", "{", "\ttype Quad", "\tposition " + " ".join(f"{x:g}" for x in pos), "\tv1 " + " ".join(f"{x:g}" for x in v1),
                      "\tv2 " + " ".join(f"{x:g}" for x in v2), "\temission " + " ".join(f"{x:g}" for x in emission), "}", ""])


def _sphere_light(pos, radius, emission):
    return "\n".join(["light", "{", "\ttype Sphere", "\tposition " + " ".join(f"{x:g}" for x in pos), f"\tradius {radius:g}",
                      "\temission " + " ".join(f"{x:g}" for x in emission), "}", ""])


# ------------------------------------------------------------------------------------------------ scenes
def cornell_256(outdir, res=(256, 256)):
    """C1: text copy of the reference's cornell_box.lfs with only the resolution line changed."""
    assets = reference_assets(outdir)
    src = open(os.path.join(assets, "cornell_box.lfs")).read().splitlines()
    out = [(f"\tresolution {res[0]} {res[1]}" if ln.strip().startswith("resolution") else ln) for ln in src]
    path = os.path.join(assets, "cornell_box.scene")
    open(path, "w").write("\n".join(out) + "\n")
    return path


def _c2(outdir, name, res, mesh_res, hdr_res, depth, with_light):
    assets = reference_assets(outdir)
    v, n, t = displaced_sphere(mesh_res[0], mesh_res[1], 1.0, amp=0.35)
    write_obj(os.path.join(assets, f"{name}_mesh.obj"), v, n, t)
    write_floor(os.path.join(assets, f"{name}_floor.obj"), half=8.0, uvscale=4.0)
    write_png_blocks(os.path.join(assets, f"{name}_checker.png"), 1024 if hdr_res[0] > 256 else 256, 8, 11)
    write_hdr(os.path.join(assets, f"{name}_sky.hdr"), sky_image(*hdr_res))
    s = _renderer(res[0], res[1], depth, hdr=f"{name}_sky.hdr")
    s += _camera((0, 1.3, -5.2), (0, 0.75, 0), 40)
    s += _material("glass", albedo=(1, 1, 1), transmission=1.0, ior=1.45, roughness=0.05, extinction=(0.8, 0.9, 0.8))
    s += _material("gold", albedo=(0.9, 0.7, 0.3), metallic=1.0, roughness=0.2)
    s += _material("ground", albedo=(1, 1, 1), roughness=0.6, albedoTexture=f"{name}_checker.png")
    s += _mesh(f"{name}_mesh.obj", "glass", (-1.15, 0.85, 0.0), (0.8, 0.8, 0.8))
    s += _mesh(f"{name}_mesh.obj", "gold", (1.15, 0.85, 0.35), (0.8, 0.8, 0.8))
    s += _mesh(f"{name}_floor.obj", "ground", (0, -0.2, 0), (1, 1, 1))
    if with_light:
        s += _sphere_light((0, 4, -1), 0.3, (30, 30, 30))
    path = os.path.join(assets, f"{name}.scene")
    open(path, "w").write(s)
    return path


def c2_mini(outdir):
    """C2 ingredients inside llvmpipe's limits: 2 x 3 968-triangle bumpy spheres (glass, gold), textured ground,
    64x32 env + sphere light, 256x128, depth 6.  Power-of-two resolution on purpose: the reference reads its
    accumulation / output textures through LINEAR samplers at texel centres, and llvmpipe evaluates those coordinates in
    fp32, so a non-power-of-two frame gets a ~1e-5 neighbour blend per pass that would blur the
    golden hit IDs; with 256x128 every weight is exactly 0 or 1."""
    return _c2(outdir, "c2mini", (256, 128), (32, 63), (64, 32), 6, True)


def c2_full(outdir):
    """C2: 660x660 grid -> 869 880-triangle closed displaced surface, instanced as metal and glass, 2048x1024 env
    with a 5e4 sun, 1280x720, depth 6."""
    return _c2(outdir, "c2full", (1280, 720), (660, 660), (2048, 1024), 6, False)


def _c3(outdir, name, res, grid, tex):
    assets = reference_assets(outdir)
    v, n, t = displaced_sphere(24, 48, 1.0, amp=0.0)
    write_obj(os.path.join(assets, f"{name}_ball.obj"), v, n, t)
    write_floor(os.path.join(assets, f"{name}_floor.obj"), half=10.0, uvscale=5.0)
    write_png_blocks(os.path.join(assets, f"{name}_albedo0.png"), tex, 8, 1, "albedo")
    write_png_blocks(os.path.join(assets, f"{name}_albedo1.png"), tex, 4, 2, "albedo")
    write_png_blocks(os.path.join(assets, f"{name}_mr.png"), tex, 8, 3, "mr")
    write_png_blocks(os.path.join(assets, f"{name}_nrm.png"), tex, 1, 4, "normal")
    s = _renderer(res[0], res[1], 4)
    s += _camera((0, 4.5, -11), (0, 0.8, 0), 38)
    s += _material("ground", albedo=(1, 1, 1), roughness=0.7, albedoTexture=f"{name}_albedo0.png", normalTexture=f"{name}_nrm.png")
    s += _material("painted", albedo=(1, 1, 1), roughness=0.4, albedoTexture=f"{name}_albedo1.png")
    s += _material("patch_metal", albedo=(0.95, 0.9, 0.8), metallicRoughnessTexture=f"{name}_mr.png")
    s += _material("textured_coat", albedo=(1, 1, 1), clearcoat=1.0, clearcoatRoughness=0.1, roughness=0.5, albedoTexture=f"{name}_albedo0.png",
                   metallicRoughnessTexture=f"{name}_mr.png")
    s += _material("red", albedo=(0.63, 0.065, 0.05))
    s += _material("green", albedo=(0.14, 0.45, 0.091))
    s += _material("mirror", albedo=(0.9, 0.9, 0.9), metallic=1.0, roughness=0.05)
    s += _material("glass", albedo=(1, 1, 1), transmission=1.0, ior=1.5, roughness=0.02, extinction=(0.9, 0.95, 0.9))
    s += _material("sheen", albedo=(0.2, 0.3, 0.8), sheen=1.0, sheenTint=0.5, subsurface=0.3, specularTint=0.5)
    s += _mesh(f"{name}_floor.obj", "ground", (0, 0, 0), (1, 1, 1))
    mats = ["painted", "patch_metal", "textured_coat", "red", "green", "mirror", "glass", "sheen"]
    props = [f"{name}_ball.obj", "cornell_box/cbox_smallbox.obj", "cornell_box/cbox_largebox.obj"]
    rng = np.random.RandomState(5)
    k = 0
    for gz in range(grid):
        for gx in range(grid):
            prop = props[k % len(props)]
            mat = mats[k % len(mats)]
            x = (gx - (grid - 1) / 2) * 2.2
            z = (gz - (grid - 1) / 2) * 2.2
            if prop.endswith("ball.obj"):
                sx, sy, sz = rng.uniform(0.5, 0.9), rng.uniform(0.5, 0.9), rng.uniform(0.5, 0.9)   # non-uniform: exercises the normal matrix
                s += _mesh(prop, mat, (x, sy, z), (sx, sy, sz))
            else:
                sc = rng.uniform(0.006, 0.009)
                s += _mesh(prop, mat, (x, 165 * sc if "large" in prop else 82.5 * sc, z), (sc, sc * rng.uniform(0.8, 1.2), sc))
            k += 1
    s += _quad_light((-3, 6, -2), (-3, 6, 1), (0, 6, -2), (20, 19, 17))
    s += _quad_light((2, 5, 2), (2, 5, 4), (4.5, 5.5, 2), (12, 14, 18))
    s += _sphere_light((0, 3.5, -6), 0.4, (25, 22, 18))
    path = os.path.join(assets, f"{name}.scene")
    open(path, "w").write(s)
    return path


def c3_mini(outdir):
    """C3 at 512x256 (power of two, see c2_mini) with a 4x4 prop grid and 256^2 textures (parity against llvmpipe)."""
    return _c3(outdir, "c3mini", (512, 256), 4, 256)


def c3_small(outdir):
    """The C3 scene (same 8x8 grid, meshes, materials, 1024^2 textures, lights) at 480x270: what the reference on llvmpipe is timed on
    for the C3 baseline (a 1920x1080 frame costs it minutes per sample)."""
    return _c3(outdir, "c3small", (480, 270), 8, 1024)


def c3_full(outdir):
    """C3: 8x8 = 64 instances over 3 meshes + ground, 9 materials (4 textured, 1024^2 PNGs), 2 quad + 1 sphere light, 1920x1080."""
    return _c3(outdir, "c3full", (1920, 1080), 8, 1024)


def _c4(outdir, name, res, depth=8, grid=36):
    assets = reference_assets(outdir)
    s = _renderer(res[0], res[1], depth)
    s += _camera((0, 9, -26), (0, 0, -2), 42)
    s += _material("diffuse_a", albedo=(0.75, 0.7, 0.65))
    s += _material("diffuse_b", albedo=(0.3, 0.5, 0.75), roughness=0.8)
    s += _material("metal", albedo=(0.95, 0.85, 0.6), metallic=1.0, roughness=0.15)
    s += _material("glass", albedo=(1, 1, 1), transmission=1.0, ior=1.45, roughness=0.03)
    mats = ["diffuse_a", "diffuse_b", "metal", "glass"]
    k = 0
    for gz in range(grid):
        for gx in range(grid):
            x = (gx - (grid - 1) / 2) * 1.0
            z = (gz - (grid - 1) / 2) * 1.0
            y = 0.25 * math.sin(0.7 * gx) * math.cos(0.5 * gz)
            s += _mesh("cornell_box/glass_sphere.obj", mats[(gx * 7 + gz * 3 + k) % 4], (x, y, z), (0.4, 0.4, 0.4))
            k += 1
    s += _sphere_light((0, 14, -6), 1.5, (40, 38, 35))
    path = os.path.join(assets, f"{name}.scene")
    open(path, "w").write(s)
    return path


def c4_gold(outdir, grid=14):
    """The C4 recipe inside llvmpipe's limits and a fixture's size: grid x grid instances (196) of one 1 984-triangle sphere,
    the same four materials (2 diffuse, metal, glass) in the same pattern, one sphere light, depth 8, 512x256 (power of two,
    see c2_mini).  tests/golden/c4gold_llvmpipe.npz pins the instanced deep-bounce path against the reference."""
    assets = reference_assets(outdir)
    v, n, t = displaced_sphere(32, 32, 1.0)
    write_obj(os.path.join(assets, "c4gold_sphere.obj"), v, n, t)
    s = _renderer(512, 256, 8)
    s += _camera((0, 4.5, -12.5), (0, 0, -1), 42)
    s += _material("diffuse_a", albedo=(0.75, 0.7, 0.65))
    s += _material("diffuse_b", albedo=(0.3, 0.5, 0.75), roughness=0.8)
    s += _material("metal", albedo=(0.95, 0.85, 0.6), metallic=1.0, roughness=0.15)
    s += _material("glass", albedo=(1, 1, 1), transmission=1.0, ior=1.45, roughness=0.03)
    mats = ["diffuse_a", "diffuse_b", "metal", "glass"]
    k = 0
    for gz in range(grid):
        for gx in range(grid):
            x = (gx - (grid - 1) / 2) * 1.0
            z = (gz - (grid - 1) / 2) * 1.0
            y = 0.25 * math.sin(0.7 * gx) * math.cos(0.5 * gz)
            s += _mesh("c4gold_sphere.obj", mats[(gx * 7 + gz * 3 + k) % 4], (x, y, z), (0.4, 0.4, 0.4))
            k += 1
    s += _sphere_light((0, 7, -3), 0.8, (40, 38, 35))
    path = os.path.join(assets, "c4gold.scene")
    open(path, "w").write(s)
    return path


def c4_stress(outdir):
    """C4/C5: 1296 x glass_sphere.obj = 20.57 M instanced triangles, 1 sphere light, depth 8, 3840x2160."""
    return _c4(outdir, "c4stress", (3840, 2160))


def c4_mini(outdir):
    """The C4 geometry at 480x270 (the size the survey timed on llvmpipe)."""
    return _c4(outdir, "c4mini", (480, 270))


SCENES = {"cornell_256": cornell_256, "c2_mini": c2_mini, "c2_full": c2_full, "c3_mini": c3_mini, "c3_full": c3_full,
          "c4_stress": c4_stress, "c4_mini": c4_mini, "c4_gold": c4_gold, "c3_small": c3_small}


def build_pack(name, outdir):
    """Generate scene `name` under outdir and convert it with the reference's loader/BVH builder; returns the .lfpack path."""
    scene = SCENES[name](outdir)
    pack = os.path.join(outdir, f"{name}.lfpack")
    if not os.path.exists(pack) or os.path.getmtime(pack) < os.path.getmtime(scene):
        write_pack(scene, pack)
    return pack


if __name__ == "__main__":
    nm, out = sys.argv[1], sys.argv[2]
    sc = SCENES[nm](out)
    print(sc)
    if "--pack" in sys.argv:
        print(write_pack(sc, os.path.join(out, f"{nm}.lfpack")))
