"""CPU tests: the SOURCE TEXT of the CUDA device functions, compiled for the host, against the oracle - bit for bit.

tests/hostcheck/lf_hostcheck.cpp compiles lavaframe_b200/csrc/{lf_math,lf_device,lf_shade}.cuh with g++ (-ffp-contract=off
mirrors nvcc -fmad=false), supplies the handful of device intrinsics, and runs k_megakernel's per-sample loop on the CPU over
the arrays lf_repack.cpp makes for the GPU.  Any statement of the kernels that departs from the oracle (and, through the
oracle's own golden tests, from the reference on llvmpipe) shows up here without a GPU.  It is a checker: nothing in the
product links or calls it, and the `-m gpu` tests remain the parity tests proper (they exercise the real kernels, the
warp-cooperative traversal, the texture objects and the device's arithmetic).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle_api import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HC_DIR = os.path.join(ROOT, "tests", "hostcheck")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")

SCENES = ["cornell", "c2mini", "c3mini", "c4gold"]


@pytest.fixture(scope="session")
def hostcheck():
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    subprocess.run(["make", "-s", "-C", HC_DIR], check=True)
    lib = C.CDLL(os.path.join(HC_DIR, "liblfhostcheck.so"))
    lib.lfhc_open_pack.restype = C.c_void_p
    lib.lfhc_open_pack.argtypes = [C.c_char_p]
    lib.lfhc_close.argtypes = [C.c_void_p]
    lib.lfhc_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.lfhc_render_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.lfhc_render_preview.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.lfhc_set_params.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lfhc_get_params.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lfhc_bsdf_kat.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.lfhc_post_process.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    return lib


def _open(lib, golden_dir, name):
    pack = os.path.join(golden_dir, f"{name}.lfpack")
    h = lib.lfhc_open_pack(pack.encode())
    assert h, f"hostcheck cannot open {pack}"
    w, hh = C.c_int(), C.c_int()
    lib.lfhc_size(h, C.byref(w), C.byref(hh))
    return pack, h, w.value, hh.value


@pytest.mark.parametrize("name", SCENES)
def test_device_source_matches_oracle(golden_dir, oracle_lib, hostcheck, name):
    pack, h, W, H = _open(hostcheck, golden_dir, name)
    n = 4
    img = np.zeros((H, W, 3), np.float32)
    culled = np.zeros((H, W, 3), np.float32)
    assert hostcheck.lfhc_render_frames(h, 2, n, 1, 0, img.ctypes.data) == 0
    assert hostcheck.lfhc_render_frames(h, 2, n, 1, 1, culled.ctypes.data) == 0
    strided = np.zeros((H, W, 3), np.float32)
    assert hostcheck.lfhc_render_frames(h, 3, 2, 2, 1, strided.ctypes.data) == 0      # frames 3, 5: a rank's share of an spp split
    hostcheck.lfhc_close(h)
    o = Oracle(pack)
    ref = o.render_frames(2, n)
    ref_strided = o.render_frames(3, 2, frame_stride=2)
    o.close()
    assert np.array_equal(img, ref), f"{name}: device source differs from the oracle on {np.mean((img != ref).any(axis=2)):.6f} of pixels"
    assert np.array_equal(culled, img), f"{name}: the distance cull changes the image"
    assert np.array_equal(strided, ref_strided)


@pytest.mark.parametrize("name", SCENES)
def test_device_source_matches_llvmpipe(golden_dir, hostcheck, name):
    """... and therefore the reference itself: the 1-spp golden image of the unmodified GLSL on llvmpipe, bit for bit."""
    g = np.load(os.path.join(golden_dir, f"{name}_llvmpipe.npz"))
    pack, h, W, H = _open(hostcheck, golden_dir, name)
    img = np.zeros((H, W, 3), np.float32)
    assert hostcheck.lfhc_render_frames(h, 2, 1, 1, 1, img.ctypes.data) == 0
    hostcheck.lfhc_close(h)
    bits = float(np.mean((img == g["spp1"]).all(axis=2)))
    assert bits >= 0.9999, f"{name}: bit-identical to llvmpipe on {bits:.6f} of pixels"


@pytest.mark.parametrize("name", SCENES)
def test_device_source_preview_matches_oracle(golden_dir, oracle_lib, hostcheck, name):
    pack, h, W, H = _open(hostcheck, golden_dir, name)
    o = Oracle(pack)
    for pw, ph, dof in ((W // 2, H // 2, 0), (W, H, 1)):
        img = np.zeros((ph, pw, 3), np.float32)
        assert hostcheck.lfhc_render_preview(h, pw, ph, 2, dof, img.ctypes.data) == 0
        ref = o.render_preview(pw, ph, 2, bool(dof))
        assert np.array_equal(img, ref), f"{name} preview {pw}x{ph} dof={dof}: differs on {np.mean((img != ref).any(axis=2)):.6f} of pixels"
    o.close()
    hostcheck.lfhc_close(h)


def test_device_source_postprocess_matches_oracle(golden_dir, oracle_lib, hostcheck):
    """lf_post.cuh (k_post's per-pixel function) against the oracle's PostProcess, and through it the reference's output."""
    from oracle_api import post_process
    from test_oracle_golden import _post_cases
    accum, inv, cases = _post_cases(golden_dir)
    H, W, _ = accum.shape
    for name, tm, pp, ref in cases:
        out = np.empty_like(accum)
        hostcheck.lfhc_post_process(accum.ctypes.data, W, H, inv, tm, C.cast(C.byref(pp), C.c_void_p), out.ctypes.data)
        assert np.array_equal(out, post_process(accum, inv, tm, pp)), name


EDGE_CASES = [   # (scene, LfParams overrides, LfCamera overrides): the CPU mirror of tests/test_edge_cases.py (single-tile cases)
    ("cornell", dict(width=1, height=1, tile_width=1, tile_height=1), {}),
    ("cornell", dict(width=37, height=23, tile_width=37, tile_height=23), {}),
    ("cornell", dict(width=129, height=1, tile_width=129, tile_height=1), {}),
    ("c2mini", dict(max_depth=1), {}),
    ("c2mini", dict(max_depth=2, enable_rr=0), {}),
    ("c3mini", dict(enable_rr=1, rr_depth=0), {}),
    ("c3mini", dict(max_depth=8, enable_rr=0), {}),
    ("c2mini", dict(use_envmap=0), {}),
    ("c2mini", dict(use_constant_bg=1), {}),
    ("cornell", {}, dict(aperture=0.02, focal_dist=0.6)),
    ("c2mini", {}, dict(aperture=0.05, focal_dist=3.0)),
]


@pytest.mark.parametrize("name,params,cam", EDGE_CASES)
def test_device_source_edge_cases(golden_dir, oracle_lib, hostcheck, name, params, cam):
    import lavaframe_b200 as lf
    pack, h, W, H = _open(hostcheck, golden_dir, name)
    p, c = lf.LfParams(), lf.LfCamera()
    hostcheck.lfhc_get_params(h, C.byref(p), C.byref(c))
    for k, v in params.items():
        setattr(p, k, v)
    if params.get("use_constant_bg"):
        p.bg_color[0], p.bg_color[1], p.bg_color[2] = 0.25, 0.5, 0.75
    for k, v in cam.items():
        setattr(c, k, v)
    hostcheck.lfhc_set_params(h, C.byref(p), C.byref(c))
    o = Oracle(pack)
    o.update_params(**params)
    if params.get("use_constant_bg"):
        o.params.bg_color[0], o.params.bg_color[1], o.params.bg_color[2] = 0.25, 0.5, 0.75
        o.update_params()
    o.lib.lforacle_set_params(o.h, None, C.byref(c))
    img = np.zeros((p.height, p.width, 3), np.float32)
    assert hostcheck.lfhc_render_frames(h, 2, 2, 1, 1, img.ctypes.data) == 0
    ref = o.render_frames(2, 2)
    o.close()
    hostcheck.lfhc_close(h)
    assert ref.any() or params.get("width") == 1
    assert np.array_equal(img, ref), f"{name} {params} {cam}: {int((img != ref).any(axis=2).sum())} pixels differ"


BSDF_OPS = ["DisneyEval (f, pdf)", "DisneySample (L, pdf)", "DisneySample (f, next rand())", "GTR1 GTR2 SmithG_GGX DielectricFresnel", "importance samplers"]


@pytest.mark.parametrize("op", range(5))
def test_device_source_bsdf_vs_llvmpipe(golden_dir, hostcheck, op):
    """The kernels' BSDF functions against the reference's OWN disney.glsl / sampling.glsl functions executed on llvmpipe
    (tests/golden/make_bsdf_golden.py): 2048 seeded random parameter sets per group, every value bit for bit."""
    g = np.load(os.path.join(golden_dir, "llvmpipe_bsdf.npz"))
    a = np.ascontiguousarray(g[f"in{op}"], np.float32)
    ref = g[f"out{op}"]
    out = np.empty_like(ref)
    hostcheck.lfhc_bsdf_kat(op, a.ctypes.data, a.shape[0], out.ctypes.data)
    same = (out.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(out) & np.isnan(ref)) | ((out == 0) & (ref == 0))
    assert same.all(), f"{BSDF_OPS[op]}: {int((~same.all(axis=1)).sum())} of {same.shape[0]} items differ from llvmpipe"


def test_axis_ray_on_flat_box_plane(tmp_path, oracle_lib, hostcheck):
    """NaN semantics of AABBIntersect (intersection.glsl:53-67 with llvmpipe's MINPS / MAXPS): a ray with d.z == 0 whose origin
    lies on the plane of a z-flat box misses that box in the reference (0 * inf = NaN slabs); the device source must too."""
    import synth_pack
    path = synth_pack.axis_ray_scene(str(tmp_path / "axis.lfpack"))
    o = Oracle(path)
    t, tri, _, _ = o.primary_hits(2)
    ref = o.render_frames(2, 2)
    o.close()
    assert (tri == 3).all() and (t > 9.9).all()          # the far triangle: the flat box in front of it was missed
    h = hostcheck.lfhc_open_pack(path.encode())
    assert h
    for cull in (0, 1):
        img = np.zeros_like(ref)
        assert hostcheck.lfhc_render_frames(h, 2, 2, 1, cull, img.ctypes.data) == 0
        assert np.array_equal(img, ref)
    hostcheck.lfhc_close(h)


def test_distant_light_branch(tmp_path, oracle_lib, hostcheck):
    """sampleDistantLight (sampling.glsl:208-216; light type 2, never written by the loader): device source == oracle."""
    import synth_pack
    lights = [synth_pack.distant_light((0.3, 0.5, 1.0), (2.0, 1.5, 1.0)),
              synth_pack.quad_light((-1.0, -1.0, 18.0), (0.0, 3.0, 0.0), (6.0, 0.0, 0.0), (20.0, 20.0, 20.0))]
    path = synth_pack.chain_scene(str(tmp_path / "distant.lfpack"), 12, lights=lights)
    o = Oracle(path)
    ref = o.render_frames(2, 8)
    o.close()
    h = hostcheck.lfhc_open_pack(path.encode())
    img = np.zeros_like(ref)
    assert hostcheck.lfhc_render_frames(h, 2, 8, 1, 1, img.ctypes.data) == 0
    hostcheck.lfhc_close(h)
    assert ref.any() and np.array_equal(img, ref)
