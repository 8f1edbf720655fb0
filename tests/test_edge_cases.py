"""GPU tests (-m gpu): edge cases of the path, CUDA (through the C ABI) against the CPU oracle, bit for bit.

  * ragged frames: resolutions that are not multiples of the 8x4 pixel blocks the kernels work in, 1x1 and 1-row frames,
    tiles that do not divide the frame (the last tile column / row hangs over the screen, TiledRenderer.cpp:57-58,342)
  * uniform extremes: maxDepth 1, Russian roulette off / from depth 0, constant background, thin lens on
  * maximum sizes: a hand-built 40-level chain BVH that needs the 64-entry traversal stack (closest_hit.glsl:70), and a
    70-level one that must be refused, not mis-rendered
  * empty work: zero frames, a camera that sees nothing
  * bad arguments: every entry point reports instead of crashing
"""
import ctypes as C
import os

import numpy as np
import pytest

import synth_pack
from oracle_api import Oracle
import lavaframe_b200 as lf
from lavaframe_b200.capi import LfParams, LfCamera, LfSceneView

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tracer(gpu):
    pt = lf.PathTracer(gpu)
    yield pt
    pt.close()


def both(tracer, pack, first=2, n=2, tile=(0, 0), cam=None, **params):
    """Render frames [first, first + n) of one tile on both sides with the same parameter overrides."""
    tracer.upload_pack(pack, **params)
    o = Oracle(pack.path)
    o.update_params(**params)
    if cam is not None:
        c = pack.camera()
        for k, v in cam.items():
            setattr(c, k, v)
        tracer.set_camera(c)
        o.lib.lforacle_set_params(o.h, None, C.byref(c))
    tracer.clear()
    tracer.render_frames(first, n, 1, *tile)
    img = tracer.read_accum()
    ref = o.render_frames(first, n, 1, *tile)
    o.close()
    return img, ref


def assert_same(img, ref, what):
    assert img.shape == ref.shape
    differ = int((img != ref).any(axis=2).sum())
    assert differ == 0, f"{what}: {differ} of {img.shape[0] * img.shape[1]} pixels differ from the oracle"


@pytest.mark.parametrize("res", [(1, 1), (7, 3), (37, 23), (129, 1), (1, 65), (250, 130)])
def test_ragged_resolutions(tracer, golden_dir, res):
    pack = lf.ScenePack(os.path.join(golden_dir, "cornell.lfpack"))
    w, h = res
    img, ref = both(tracer, pack, width=w, height=h, tile_width=w, tile_height=h)
    assert ref.any()
    assert_same(img, ref, f"cornell {w}x{h}")


@pytest.mark.parametrize("name", ["cornell", "c3mini"])
def test_tiles_that_do_not_divide_the_frame(tracer, golden_dir, name):
    """numTiles = ceil(screen / tile) (TiledRenderer.cpp:57-58): the last tile column and row hang over the screen edge; the
    copy into the accumulation target clips them (TiledRenderer.cpp:341-344)."""
    pack = lf.ScenePack(os.path.join(golden_dir, f"{name}.lfpack"))
    tw, th = 96, 80
    ntx, nty = -(-pack.width // tw), -(-pack.height // th)
    for tile in [(0, 0), (ntx - 1, 0), (0, nty - 1), (ntx - 1, nty - 1)]:
        img, ref = both(tracer, pack, n=1, tile=tile, tile_width=tw, tile_height=th)
        assert_same(img, ref, f"{name} tile {tile}")


@pytest.mark.parametrize("params", [
    dict(max_depth=0), dict(max_depth=1), dict(max_depth=2, enable_rr=0), dict(enable_rr=1, rr_depth=0), dict(max_depth=8, enable_rr=0), dict(use_envmap=0),
])
@pytest.mark.parametrize("name", ["c2mini", "c3mini"])
def test_uniform_extremes(tracer, golden_dir, name, params):
    pack = lf.ScenePack(os.path.join(golden_dir, f"{name}.lfpack"))
    img, ref = both(tracer, pack, **params)
    assert_same(img, ref, f"{name} {params}")


@pytest.mark.parametrize("name", ["cornell", "c2mini"])
def test_constant_background(tracer, golden_dir, name):
    """renderOptions.useConstantBg: misses add bgColor and the env-map NEE is compiled out (#define CONSTANT_BG,
    TiledRenderer.cpp:86-87; pathtrace.glsl:226-227,141)."""
    pack = lf.ScenePack(os.path.join(golden_dir, f"{name}.lfpack"))
    tracer.upload_pack(pack)
    o = Oracle(pack.path)
    for q in (tracer.params, o.params):
        q.use_constant_bg = 1
        q.bg_color[0], q.bg_color[1], q.bg_color[2] = 0.25, 0.5, 0.75
    tracer.set_params(tracer.params)
    o.update_params()
    tracer.clear(); tracer.render_frames(2, 2)
    img = tracer.read_accum()
    ref = o.render_frames(2, 2)
    o.close()
    assert_same(img, ref, f"{name} constant background")


@pytest.mark.parametrize("name", ["cornell", "c2mini"])
def test_thin_lens(tracer, golden_dir, name):
    """aperture > 0: the lens sample of renderer.glsl:57-62 moves the ray origin; focal distance in front of the geometry."""
    pack = lf.ScenePack(os.path.join(golden_dir, f"{name}.lfpack"))
    img, ref = both(tracer, pack, cam=dict(aperture=0.05, focal_dist=0.8))
    assert_same(img, ref, f"{name} aperture 0.05")


def test_deep_chain_bvh_uses_the_64_entry_stack(tracer, tmp_path, oracle_lib):
    path = synth_pack.chain_scene(str(tmp_path / "chain40.lfpack"), 40)
    pack = lf.ScenePack(path)
    tracer.upload_pack(pack)
    o = Oracle(path)
    hits, ohits = tracer.primary_hits(2), o.primary_hits(2)
    for a, b in zip(hits, ohits):
        assert np.array_equal(a, b)
    assert (hits[0] < 1e6).mean() > 0.3
    for mode in (0, 1):                                           # wavefront and megakernel
        tracer.upload_pack(pack, kernel_mode=mode)
        tracer.clear(); tracer.render_frames(2, 4)
        assert_same(tracer.read_accum(), o.render_frames(2, 4), f"chain40 mode {mode}")
    # the walk really is deep: without the cull every closest-hit ray visits dozens of inner nodes
    tracer.upload_pack(pack, no_cull=1, count_work=1)
    tracer.reset_counters(); tracer.clear(); tracer.render_frames(2, 1)
    c = tracer.counters()
    assert c["inner_visits"] / c["rays_closest"] > 15
    tracer.update_params(count_work=0)
    o.close()


def test_distant_light_branch(tracer, tmp_path, oracle_lib):
    """sampleDistantLight (sampling.glsl:208-216, light type 2): the loader never writes one, the shader and both restatements
    carry the branch.  A hand-built pack with a distant light next to a quad light: CUDA == oracle on every bit."""
    zc = 12 + 6.0
    lights = [synth_pack.distant_light((0.3, 0.5, 1.0), (2.0, 1.5, 1.0)),
              synth_pack.quad_light((-1.0, -1.0, zc), (0.0, 3.0, 0.0), (6.0, 0.0, 0.0), (20.0, 20.0, 20.0))]
    path = synth_pack.chain_scene(str(tmp_path / "distant.lfpack"), 12, lights=lights)
    pack = lf.ScenePack(path)
    o = Oracle(path)
    ref = o.render_frames(2, 8)
    only_quad = Oracle(synth_pack.chain_scene(str(tmp_path / "quad.lfpack"), 12)).render_frames(2, 8)
    assert not np.array_equal(ref, only_quad)                    # the distant light does light the scene
    for mode in (0, 1):
        tracer.upload_pack(pack, kernel_mode=mode)
        tracer.clear(); tracer.render_frames(2, 8)
        assert_same(tracer.read_accum(), ref, f"distant light, mode {mode}")
    o.close()


def test_axis_ray_on_flat_box_plane(tracer, tmp_path, oracle_lib):
    """KAT for AABBIntersect's NaN semantics (synth_pack.axis_ray_scene): d.z == 0 exactly, origin on the
    plane of a z-flat box.  llvmpipe's MINPS / MAXPS make t1 NaN and the box a miss; the GPU's FMNMX would make it a hit."""
    path = synth_pack.axis_ray_scene(str(tmp_path / "axis.lfpack"))
    pack = lf.ScenePack(path)
    o = Oracle(path)
    ohits = o.primary_hits(2)
    assert (ohits[1] == 3).all() and (ohits[0] > 9.9).all()      # the reference walks past the flat box to the far triangle
    for no_cull in (0, 1):
        tracer.upload_pack(pack, no_cull=no_cull)
        for a, b in zip(tracer.primary_hits(2), ohits):
            assert np.array_equal(a, b)
        for mode in (0, 1):
            tracer.update_params(kernel_mode=mode)
            tracer.clear(); tracer.render_frames(2, 2)
            assert_same(tracer.read_accum(), o.render_frames(2, 2), f"axis ray, no_cull {no_cull}, mode {mode}")
    o.close()


def test_bvh_deeper_than_the_reference_stack_is_refused(tracer, tmp_path):
    pack = lf.ScenePack(synth_pack.chain_scene(str(tmp_path / "chain70.lfpack"), 70))
    with pytest.raises(lf.LfCudaError, match="deeper than 64"):
        tracer.upload_pack(pack)


def test_empty_work(tracer, golden_dir):
    pack = lf.ScenePack(os.path.join(golden_dir, "cornell.lfpack"))
    tracer.upload_pack(pack)
    tracer.clear()
    tracer.render_frames(2, 0)                                    # zero frames: nothing happens
    assert not tracer.read_accum().any()
    # a camera that looks away from everything: every path misses at depth 0, black without an env map
    img, ref = both(tracer, pack, cam=dict(forward=(C.c_float * 3)(0.0, 0.0, -1.0)))
    assert not ref.any()
    assert_same(img, ref, "camera looking away")


def test_uniform_changes_keep_the_accumulation(tracer, golden_dir, oracle_lib):
    """The reference changes maxDepth / tile uniforms without touching accumTexture (TiledRenderer.cpp:505-521).  Here: the
    primary-hit probe with tile != resolution, a new max_depth and a new batch size all leave the accumulated image and its
    device address (lfcuda_accum_device_ptr, what NCCL reduces) alone; only a new resolution re-creates the buffer."""
    pack = lf.ScenePack(os.path.join(golden_dir, "cornell.lfpack"))
    tracer.upload_pack(pack, tile_width=64, tile_height=64)
    tracer.clear(); tracer.render_frames(2, 2, 1, 1, 2)
    img0 = tracer.read_accum()
    ptr0 = tracer.accum_device_ptr()
    assert img0.any()
    hits = tracer.primary_hits(2)
    assert np.array_equal(tracer.read_accum(), img0) and tracer.accum_device_ptr() == ptr0
    tracer.update_params(max_depth=2)
    assert np.array_equal(tracer.read_accum(), img0) and tracer.accum_device_ptr() == ptr0
    tracer.update_params(max_depth=6, frames_in_flight=3)
    assert np.array_equal(tracer.read_accum(), img0) and tracer.accum_device_ptr() == ptr0
    tracer.update_params(tile_width=128, tile_height=32, frames_in_flight=0)
    assert np.array_equal(tracer.read_accum(), img0) and tracer.accum_device_ptr() == ptr0
    # ... and the probe's answer does not depend on the tile size it was called under
    tracer.update_params(tile_width=pack.width, tile_height=pack.height)
    for a, b in zip(hits, tracer.primary_hits(2)):
        assert np.array_equal(a, b)
    # the state still renders correctly after all of that: tile (0, 0) of 128 x 32 at depth 6 on top of the old image
    tracer.update_params(tile_width=128, tile_height=32)
    tracer.render_frames(5, 1, 1, 0, 0)
    o = Oracle(pack.path)
    o.update_params(tile_width=64, tile_height=64)
    ref = o.render_frames(2, 2, 1, 1, 2)
    o.update_params(tile_width=128, tile_height=32, max_depth=6)
    ref = o.render_frames(5, 1, 1, 0, 0, accum=ref)
    o.close()
    assert_same(tracer.read_accum(), ref, "accumulation across uniform changes")
    tracer.update_params(width=128, height=128, tile_width=128, tile_height=128)      # a new resolution starts from black
    assert not tracer.read_accum().any()


def test_bad_arguments_are_reported(gpu, golden_dir):
    pack = lf.ScenePack(os.path.join(golden_dir, "cornell.lfpack"))
    pt = lf.PathTracer(gpu)
    lib = pt.lib
    # parameters before a scene, rendering before parameters
    with pytest.raises(lf.LfCudaError):
        pt.render_frames(2, 1)
    pt.upload_pack(pack)
    with pytest.raises(lf.LfCudaError):
        pt.render_frames(2, -1)
    bad = LfParams()
    C.memmove(C.byref(bad), C.byref(pt.params), C.sizeof(LfParams))
    bad.width = 0
    assert lib.lfcuda_set_params(pt.h, C.byref(bad)) != 0
    assert b"" != lib.lfcuda_last_error(pt.h)
    assert lib.lfcuda_set_params(pt.h, None) != 0
    assert lib.lfcuda_set_camera(pt.h, None) != 0
    assert lib.lfcuda_read_accum(pt.h, None) != 0
    assert lib.lfcuda_upload_scene(pt.h, None) != 0
    # a scene whose vertex indices point outside the vertex array
    view = pack.view()
    idx = pack.vert_indices.copy()
    idx[5] = pack.num_vertices + 7
    view.vert_indices = idx.ctypes.data_as(C.POINTER(C.c_int32))
    assert lib.lfcuda_upload_scene(pt.h, C.byref(view)) != 0
    assert b"out of range" in lib.lfcuda_last_error(pt.h)
    # a scene view with no instances
    view = pack.view()
    view.num_instances = 0
    assert lib.lfcuda_upload_scene(pt.h, C.byref(view)) != 0
    # the context survives all of that
    pt.upload_pack(pack)
    pt.clear(); pt.render_frames(2, 1)
    assert pt.read_accum().any()
    pt.close()
