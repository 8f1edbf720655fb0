"""GPU tests (-m gpu) at BASELINE.json's FULL sizes: C2 (869 880-triangle mesh x2 + HDR env, 1280x720, depth 6), C3
(65 instances, textures, 3 lights, 1920x1080) and C4/C5 (1296 x 15 872 = 20.57 M instanced triangles, depth 8, 3840x2160).

The scenes are generated on the spot (scenes/gen_scenes.py -> the reference's unchanged loader + BVH builder, a few seconds)
; parity is checked
  * directly, bit for bit, on the WHOLE frame at one sample per pixel and on single tiles of the full-resolution frame at more
    (the tile uniforms of renderer.glsl:27-33 make a tile an exact crop: same pixel coordinates, same RNG seeds), and
  * through size-independent properties over the whole frame: wavefront == megakernel, cull == no cull, frame-strided
    subsets (the multi-GPU spp split) add up to the whole, and a tiled render equals the untiled one.
Everything goes through the C ABI (liblfcuda.so)."""
import os

import numpy as np
import pytest

from oracle_api import Oracle
from parity_metrics import radiance_agreement
import lavaframe_b200 as lf
from lavaframe_b200.capi import lib_path

pytestmark = pytest.mark.gpu

FULL = ["c2_full", "c3_full", "c4_stress"]
# (tile width, tile height) dividing the frame, and the tiles (tileX, tileY) compared against the oracle
TILES = {
    "c2_full": ((80, 45), [(8, 8), (5, 3), (0, 15)]),       # 1280x720 -> 16 x 16 tiles: centre (mesh), ground + mesh edge, top-left sky
    "c3_full": ((120, 60), [(8, 9), (3, 4), (15, 17)]),     # 1920x1080 -> 16 x 18 tiles
    "c4_stress": ((120, 60), [(16, 18), (9, 11), (31, 35)]),  # 3840x2160 -> 32 x 36 tiles
}
TILED_MIN = {"c2_full": 0.9, "c3_full": 0.99, "c4_stress": 0.99}


@pytest.fixture(scope="module")
def full_packs(gpu, tmp_path_factory):
    if not os.path.exists(lib_path("liblfhost.so")):
        pytest.fail("liblfhost.so missing: __graft_entry__.build() must run where /root/reference exists")
    from scenes import gen_scenes
    cache = {}

    def get(name):
        if name not in cache:
            out = str(tmp_path_factory.mktemp(name))
            gen_scenes.build_pack(name, out)
            cache[name] = lf.ScenePack(os.path.join(out, f"{name}.lfpack"))
        return cache[name]
    return get


@pytest.fixture(scope="module")
def tracer(gpu):
    pt = lf.PathTracer(gpu)
    yield pt
    pt.close()


def crop(img, tw, th, tx, ty):
    return img[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]


@pytest.mark.parametrize("name", FULL)
def test_full_scene_tiles_vs_oracle(tracer, full_packs, oracle_lib, name):
    """Tiles of the full-resolution frame, 2 spp at the config's full depth: CUDA == oracle on every bit, and nothing is
    written outside the tile."""
    pack = full_packs(name)
    (tw, th), tiles = TILES[name]
    assert pack.width % tw == 0 and pack.height % th == 0
    tracer.upload_pack(pack, tile_width=tw, tile_height=th)
    o = Oracle(pack.path)
    o.update_params(tile_width=tw, tile_height=th)
    for tx, ty in tiles:
        tracer.clear()
        tracer.render_frames(2, 2, 1, tx, ty)
        img = tracer.read_accum()
        ref = o.render_frames(2, 2, 1, tx, ty)
        a, b = crop(img, tw, th, tx, ty), crop(ref, tw, th, tx, ty)
        if (tx, ty) == tiles[0]:
            assert b.any(), f"{name} tile {tx},{ty}: the oracle rendered nothing"
        differ = int((a != b).any(axis=2).sum())
        assert differ == 0, f"{name} tile {tx},{ty}: {differ} of {tw * th} pixels differ from the oracle"
        outside = img.copy()
        crop(outside, tw, th, tx, ty)[:] = 0
        assert not outside.any()
    o.close()


@pytest.mark.parametrize("name", FULL)
def test_full_frame_radiance_vs_oracle(tracer, full_packs, oracle_lib, name):
    """The WHOLE frame at BASELINE's resolution and the config's full depth, one sample per pixel (frame 2, the reference's first
    sample): 0.9 M / 2.1 M / 8.3 M paths, every pixel's radiance equal to the oracle's bit for bit (north_star check 2 asks for
    1e-3 on 99.9 % of the pixels).  The oracle needs about 0.5 / 1 / 10 s on the box's host cores."""
    pack = full_packs(name)
    tracer.upload_pack(pack)
    tracer.clear(); tracer.render_frames(2, 1)
    img = tracer.read_accum()
    o = Oracle(pack.path)
    ref = o.render_frames(2, 1)
    o.close()
    assert ref.any() and np.isfinite(ref).all()
    differ = int((img != ref).any(axis=2).sum())
    print(f"{name}: {pack.width}x{pack.height} 1-spp radiance, {differ} pixels differ from the oracle; within 1e-3 on {radiance_agreement(img, ref):.6f}")
    assert differ == 0, f"{name}: {differ} of {pack.width * pack.height} pixels differ from the oracle"


@pytest.mark.parametrize("name", FULL)
def test_full_frame_primary_hits(tracer, full_packs, oracle_lib, name):
    """Whole frame at full resolution (0.9 M / 2.1 M / 8.3 M primary rays): hit t, triangle, material and emitter flag equal
    the oracle's ClosestHit bit for bit, and the distance cull changes none of them."""
    pack = full_packs(name)
    res = []
    for no_cull in (0, 1):
        tracer.upload_pack(pack, no_cull=no_cull)
        res.append(tracer.primary_hits(2))
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    t, tri, mat, em = res[0]
    assert (t < 1e6).mean() > 0.2                               # the camera does look at the scene
    o = Oracle(pack.path)
    ot, otri, omat, oem = o.primary_hits(2)
    o.close()
    assert np.array_equal(t, ot), f"{name}: t differs on {int((t != ot).sum())} pixels"
    assert np.array_equal(em > 0, oem > 0)
    assert np.array_equal(np.where(em > 0, -1, tri), np.where(oem > 0, -1, otri))
    assert np.array_equal(np.where(em > 0, -1, mat), np.where(oem > 0, -1, omat))


@pytest.mark.parametrize("name", FULL)
def test_full_frame_properties(tracer, full_packs, name):
    """Whole frame, full depth: wavefront == megakernel bitwise; frames 2..5 in one call == two calls; the stride-2 split
    (what two GPUs render) adds up to the whole within fp32 summation order; a tile-by-tile render equals the untiled one."""
    pack = full_packs(name)
    tracer.upload_pack(pack, kernel_mode=0)
    tracer.clear(); tracer.render_frames(2, 4); whole = tracer.read_accum()
    assert np.isfinite(whole).all() and whole.mean() > 0
    tracer.clear(); tracer.render_frames(2, 1); tracer.render_frames(3, 3); split = tracer.read_accum()
    assert np.array_equal(whole, split)
    tracer.clear(); tracer.render_frames(2, 2, 2); even = tracer.read_accum()
    tracer.clear(); tracer.render_frames(3, 2, 2); odd = tracer.read_accum()
    np.testing.assert_allclose(even + odd, whole, rtol=4e-6, atol=1e-6)
    tracer.clear(); tracer.render_frames(2, 1); one = tracer.read_accum()

    tracer.upload_pack(pack, kernel_mode=1)
    tracer.clear(); tracer.render_frames(2, 1); mega = tracer.read_accum()
    assert np.array_equal(one, mega), f"{int((one != mega).any(axis=2).sum())} pixels differ between wavefront and megakernel"

    # tiled == untiled: 4 x 4 tiles of frame 2 (power-of-two splits keep the tile uniforms exact in fp32)
    if pack.width % 4 == 0 and pack.height % 4 == 0:
        tw, th = pack.width // 4, pack.height // 4
        tracer.upload_pack(pack, kernel_mode=0, tile_width=tw, tile_height=th)
        tracer.clear()
        for ty in range(4):
            for tx in range(4):
                tracer.render_frames(2, 1, 1, tx, ty)
        tiled = tracer.read_accum()
        # the tile uniforms (renderer.glsl:27-33) round the pixel coordinate differently in the last ulp; diffuse pixels absorb
        # that, glass / rough-metal chains amplify it (C2), so the bar is per scene
        same = radiance_agreement(tiled, one, rel=1e-4)
        print(f"{name}: tiled == untiled within 1e-4 on {same:.6f} of the pixels")
        assert same >= TILED_MIN[name], f"tiled render equals the untiled one on {same:.6f} of the pixels"
