"""CPU tests: the oracle (oracle/lf_oracle.cpp) against the reference's own outputs.

The golden files were produced by the UNMODIFIED reference renderer + GLSL on Mesa llvmpipe
(tests/golden/make_golden.py).  The bars are BASELINE.json's north_star: hit IDs >= 99.99 % with t within 1e-5
relative, 1-spp radiance within 1e-3 relative on >= 99.9 % of pixels.
"""
import os

import numpy as np
import pytest

from oracle_api import Oracle, rand_kat
from parity_metrics import hits_agreement, radiance_agreement, rmse_over_mean_luminance
from lavaframe_b200 import ScenePack

# seed.x and rand() for the first four draws after InitRNG((px+.5, py+.5), frame),
# confirmed digit for digit by executing the reference's globals.glsl on llvmpipe.
RAND_KAT = {
    (0, 0, 2): ([0xEAB98F60, 0xF182283F, 0xB52B0237, 0xBF437CAE], [0.916893899, 0.943392277, 0.707687497, 0.747123539]),
    (128, 64, 2): ([0xC8492D99, 0x8222E53D, 0xB4916DB2, 0x056006AB], [0.782366633, 0.508344948, 0.705344081, 0.0209964905]),
    (255, 255, 65): ([0x91D6F7C4, 0x65A2743E, 0x5710FEE0, 0xC25A1331], [0.569686413, 0.397010088, 0.34010309, 0.759186924]),
    (1919, 1079, 2): ([0x6BF019D2, 0xA20F1902, 0x36F3FC5A, 0xA7E94E82], [0.421632409, 0.633042872, 0.214660421, 0.655903757]),
    (255, 255, 2): (None, [0.127037153, 0.390085995, 0.898335993, 0.526637554]),
    (0, 0, 65): (None, [0.976941526, 0.753729939, 0.993238211, 0.842056215]),
    (128, 64, 65): (None, [0.332470775, 0.9059605, 0.251751006, 0.606678545]),
    (1919, 1079, 65): (None, [0.216778189, 0.417842865, 0.309214175, 0.419827402]),
}

# primary hit of frame 2 and 1-spp radiance for 11 Cornell pixels, from the real reference
CORNELL_KAT = [
    ((0, 0), 0.8559915, 12, 3, (0.1289219, 0.1251144, 0.1176499)),
    ((10, 10), 0.9096831, 12, 3, (0.09328944, 0.09143415, 0.08772359)),
    ((128, 128), 1.043115, 84, 5, (0, 0, 0)),
    ((200, 50), 1.305775, 15, 3, (0.08850549, 0.08672039, 0.08315022)),
    ((64, 200), 1.357808, 18, 2, (0.01706175, 0.04305766, 0.01201389)),
    ((128, 250), 0.8370075, 0, 1, (0, 0, 0)),
    ((255, 255), 0.849043, 6, 1, (0.0786657, 0.07484642, 0.06756029)),
    ((30, 128), 1.034486, 102, 6, (0.1843958, 0.04996428, 0.03754358)),
    ((225, 128), 1.033656, 99, 7, (0.02668874, 0.07729597, 0.01868953)),
    ((90, 90), 1.020428, 87, 5, (0.04019721, 0.03937477, 0.03773256)),
    ((170, 60), 0.850257, 36, 4, (0, 0, 0)),
]

SCENES = ["cornell", "c2mini", "c3mini", "c4gold"]

# Fraction of pixels whose 1-spp radiance must be within 1e-3 of the reference-on-llvmpipe (north_star: 99.9 %), and of the
# N-spp mean.  Since the oracle restates llvmpipe's own evaluation of every GLSL built-in (test_builtins_bit_exact below),
# its x * (1 / y) division and its 8-bit texture filter, the 1-spp images are BIT-IDENTICAL on every pixel of all three
# scenes, glass, rough metal, clearcoat and textures included.  The one thing no implementation can reproduce is the
# reference's read of an unwritten material when the nearest hit is an analytic light (pathtrace.glsl:246-253 runs
# GetMaterialsAndTextures on a State whose matID was never set; llvmpipe leaves the temporaries uninitialised, and the same
# binary returns different values from run to run): those pixels are counted in the bars below, and excluded only from the
# bit-identity check of the N-spp mean.
MIN_SPP1 = {"cornell": 0.9999, "c2mini": 0.9999, "c3mini": 0.9999, "c4gold": 0.9999}
MIN_SPP1_BITS = {"cornell": 0.9999, "c2mini": 0.9999, "c3mini": 0.9999, "c4gold": 0.9999}
MIN_SPPN = {"cornell": 0.999, "c2mini": 0.999, "c3mini": 0.97, "c4gold": 0.999}     # c3mini: 7 % of its pixels look straight at a light
MIN_SPPN_BITS_NO_EMITTER = 0.999


def _pack(golden_dir, name):
    p = os.path.join(golden_dir, f"{name}.lfpack")
    if not os.path.exists(p):
        pytest.skip(f"{p} not generated")
    return p


def test_rand_kat(oracle_lib):
    for (px, py, frame), (seeds, vals) in RAND_KAT.items():
        s, v = rand_kat(px, py, frame, 4)
        if seeds is not None:
            assert [int(x) for x in s] == seeds
        np.testing.assert_allclose(v, np.array(vals, np.float32), rtol=0, atol=1e-9)


def _same_bits(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b)) | ((a == 0) & (b == 0))


@pytest.mark.parametrize("group", range(6))
def test_builtins_bit_exact(golden_dir, oracle_lib, group):
    """The oracle's GLSL built-ins against llvmpipe EXECUTING them (tests/golden/make_builtin_golden.py): sin cos tan sqrt |
    exp log pow acos | atan(y,x) `/` mix inversesqrt | refract | reflect length | RGBA8 LINEAR/REPEAT texture filter."""
    from oracle_api import builtin_kat
    g = np.load(os.path.join(golden_dir, "llvmpipe_builtins.npz"))
    out = builtin_kat(group, g[f"in{group}"], tex=g["tex"] if group == 5 else None)
    same = _same_bits(out, g[f"out{group}"])
    assert same.all(), f"{g['expr'][group]}: {int((~same).sum())} of {same.size} values differ from llvmpipe"


@pytest.mark.parametrize("op", range(5))
def test_bsdf_functions_bit_exact(golden_dir, oracle_lib, op):
    """The oracle's DisneyEval / DisneySample / microfacet helpers / importance samplers against the reference's OWN GLSL functions
    executed on llvmpipe with random parameters (tests/golden/make_bsdf_golden.py): pins what Mesa's compiler does to each expression
    (constant folding across `PI * log(a2)`, the rebalanced `FH * sheen` product ...) for every lobe and parameter, not only those
    a golden scene happens to use."""
    from oracle_api import bsdf_kat
    g = np.load(os.path.join(golden_dir, "llvmpipe_bsdf.npz"))
    out = bsdf_kat(op, g[f"in{op}"])
    same = _same_bits(out, g[f"out{op}"])
    assert same.all(), f"op {op}: {int((~same.all(axis=1)).sum())} of {same.shape[0]} items differ from llvmpipe"


def test_cornell_pack_matches_reference_dump(golden_dir):
    """SURVEY.md Appendix A KAT: the flattened Cornell arrays as the reference's own host code builds them."""
    p = ScenePack(_pack(golden_dir, "cornell"))
    assert (p.num_nodes, p.top_index, p.num_tri_refs, p.num_vertices, p.num_instances, p.num_materials, p.num_lights) == (43, 29, 36, 108, 7, 8, 1)
    nodes = p.nodes.reshape(-1, 9)
    assert list(nodes[0, 6:]) == [1, 2, 0] and list(nodes[1, 6:]) == [2, 2, 1] and list(nodes[2, 6:]) == [0, 2, 1]
    assert list(nodes[29, 6:]) == [30, 35, 0] and list(nodes[31, 6:]) == [27, 7, -6] and list(nodes[41, 6:]) == [0, 1, -1]
    assert (p.width, p.height, p.max_depth, p.enable_rr, p.rr_depth, p.use_envmap) == (256, 256, 4, 1, 2, 0)
    np.testing.assert_allclose(p.fhdr[4:7], [0.276, 0.275, -0.75], atol=1e-6)
    assert abs(p.camera().fov - 0.698132) < 1e-6
    light = p.lights.reshape(-1, 15)[0]
    np.testing.assert_allclose(light[6:12], [0, 0, 0.105, -0.13, 0, 0], atol=1e-6)
    assert abs(light[13] - 0.01365) < 1e-6 and light[14] == 0.0


def test_cornell_known_pixels(golden_dir, oracle_lib):
    o = Oracle(_pack(golden_dir, "cornell"))
    t, tri, mat, em = o.primary_hits(2)
    img = o.render_frames(2, 1)
    for (x, y), kt, ktri, kmat, rgb in CORNELL_KAT:
        assert tri[y, x] == ktri and mat[y, x] == kmat
        assert abs(t[y, x] - kt) <= 2e-6 * kt
        np.testing.assert_allclose(img[y, x], rgb, rtol=1e-3, atol=1e-7)
    # whole-image facts of the reference's frame-2 primary rays: 425 pixels see the quad light, exactly one pixel slips
    # through a crack of the epsilon-free triangle test
    assert int(em.sum()) == 425 and int((t == 1e6).sum()) == 1
    o.close()


@pytest.mark.parametrize("name", SCENES)
def test_oracle_vs_llvmpipe(golden_dir, oracle_lib, name):
    gold = os.path.join(golden_dir, f"{name}_llvmpipe.npz")
    if not os.path.exists(gold):
        pytest.skip(f"{gold} not generated")
    g = np.load(gold)
    o = Oracle(_pack(golden_dir, name))
    t, tri, mat, em = o.primary_hits(2)
    g_tri = np.where(g["hits_emitter"] > 0, -1, g["hits_tri"])
    tri = np.where(em > 0, -1, tri)
    mat_o = np.where(em > 0, -1, mat)
    same, both = hits_agreement(t, tri, mat_o, g["hits_t"], g_tri, np.where(g["hits_emitter"] > 0, -1, g["hits_mat"]))
    assert same >= 0.9999, f"{name}: hit IDs agree on {same:.6f}"
    assert both >= 0.9999, f"{name}: hit IDs + t(1e-5) agree on {both:.6f}"
    assert np.array_equal(em > 0, g["hits_emitter"] > 0) or (np.mean((em > 0) == (g["hits_emitter"] > 0)) >= 0.9999)
    s1 = o.render_frames(2, 1)
    frac = radiance_agreement(s1, g["spp1"])
    assert frac >= MIN_SPP1[name], f"{name}: 1-spp radiance within 1e-3 on {frac:.6f} of pixels"
    bits = float(np.mean((s1 == g["spp1"]).all(axis=2)))
    assert bits >= MIN_SPP1_BITS[name], f"{name}: 1-spp radiance bit-identical to llvmpipe on {bits:.6f} of pixels"
    n = int(g["nspp"])
    sN = o.render_frames(2, n) / np.float32(n)
    assert radiance_agreement(sN, g["sppN"], rel=1e-3) >= MIN_SPPN[name]
    surface = em == 0                                   # first hit is not an analytic light (see the note above)
    bitsN = float(np.mean((sN == g["sppN"]).all(axis=2)[surface]))
    assert bitsN >= MIN_SPPN_BITS_NO_EMITTER, f"{name}: {n}-spp mean bit-identical on {bitsN:.6f} of the non-emitter pixels"
    assert rmse_over_mean_luminance(sN, g["sppN"]) < 0.05
    o.close()


def test_converged_64spp_window_vs_llvmpipe_4096(golden_dir, oracle_lib):
    """A CPU-sized slice of the converged check: the oracle's 4096-spp mean over a 32x32 window of the Cornell frame
    against the reference's 4096-spp image (the full-frame check runs on the GPU)."""
    g = np.load(os.path.join(golden_dir, "cornell_llvmpipe_4096spp.npz"))["spp4096"]
    o = Oracle(_pack(golden_dir, "cornell"))
    o.update_params(tile_width=32, tile_height=32)
    acc = o.render_frames(2, 4096, 1, 3, 4)             # tile (3, 4): pixels x 96..127, y 128..159
    o.close()
    win = acc[128:160, 96:128] / np.float32(4096)
    ref = g[128:160, 96:128]
    assert rmse_over_mean_luminance(win, ref) < 0.005
    assert radiance_agreement(win, ref, rel=1e-3) >= 0.99
    # 4096 samples per pixel accumulated in frame order like the reference: the converged window is bit-identical to the reference's
    # (the whole 256x256 frame is too, profiles/r1_oracle_vs_llvmpipe.txt; 100 s of CPU, so not part of this suite)
    assert np.array_equal(win, ref)


@pytest.mark.parametrize("name", SCENES)
def test_cull_preserves_results(golden_dir, oracle_lib, name):
    """The distance cull (not in the reference) must not change a single hit or radiance value."""
    a = Oracle(_pack(golden_dir, name), cull=False)
    b = Oracle(_pack(golden_dir, name), cull=True)
    for x, y in zip(a.primary_hits(2), b.primary_hits(2)):
        assert np.array_equal(x, y)
    assert np.array_equal(a.render_frames(2, 2), b.render_frames(2, 2))
    a.close(); b.close()


# Same reasoning as MIN_SPP1: the Cornell box is pinned at the north_star bar, glass / metal / textures cannot be.
MIN_PREVIEW = {"cornell": 0.9999, "c2mini": 0.9999, "c3mini": 0.99, "c4gold": 0.9999}   # c3mini: emitter pixels, see above


@pytest.mark.parametrize("name", SCENES)
def test_preview_engine_vs_llvmpipe(golden_dir, oracle_lib, name):
    """shaders/preview_flareon.glsl as TiledRenderer::Render draws it while the camera moves
    (read back from previewFBO of the unmodified reference on llvmpipe) against the oracle's RenderPreview."""
    gold = os.path.join(golden_dir, f"{name}_llvmpipe_preview.npz")
    if not os.path.exists(gold):
        pytest.skip(f"{gold} not generated")
    g = np.load(gold)
    o = Oracle(_pack(golden_dir, name))
    for key, dof in (("half", False), ("full_dof", True)):
        ref = g[key]
        h, w, _ = ref.shape
        img = o.render_preview(w, h, 2, dof)
        frac = radiance_agreement(img, ref)
        assert frac >= MIN_PREVIEW[name], f"{name}/{key}: preview radiance within 1e-3 on {frac:.6f} of pixels"
        assert rmse_over_mean_luminance(img, ref) < 0.05
    o.close()


def tile_sequence(n_samples, ntx, nty, first_frame=2):
    """(frame, tileX, tileY) of every Render() of TiledRenderer for n_samples samples: tileX runs fastest, tileY counts DOWN from
    the top row, and `frame` advances with every tile step (TiledRenderer.cpp:55-64,485-501)."""
    frame = first_frame
    for _ in range(n_samples):
        for ty in range(nty - 1, -1, -1):
            for tx in range(ntx):
                yield frame, tx, ty
                frame += 1


def test_tiled_render_vs_llvmpipe(golden_dir, oracle_lib):
    """The reference run with tileWidth = tileHeight = 64 (16 tile steps per sample, a new `frame` for each): pins the tile
    order, the tile uniforms of renderer.glsl:27-33 and the frame numbering the RNG is seeded with."""
    g = np.load(os.path.join(golden_dir, "cornell_tiled_llvmpipe.npz"))
    n, tile = int(g["nspp"]), int(g["tile"])
    o = Oracle(_pack(golden_dir, "cornell"))
    o.update_params(tile_width=tile, tile_height=tile)
    acc = np.zeros((256, 256, 3), np.float32)
    for frame, tx, ty in tile_sequence(n, 256 // tile, 256 // tile):
        o.render_frames(frame, 1, 1, tx, ty, acc)
    o.close()
    img = acc / np.float32(n)
    assert radiance_agreement(img, g["sppN"]) >= 0.999
    assert rmse_over_mean_luminance(img, g["sppN"]) < 0.005
    bits = float(np.mean((img == g["sppN"]).all(axis=2)))
    assert bits >= 0.9999, f"tiled render bit-identical to the reference on {bits:.6f} of pixels"


def _post_cases(golden_dir):
    """(name, tonemapIndex, LfPostParams or None, reference image) of tests/golden/cornell64_llvmpipe_post.npz + the accumulation buffer."""
    import importlib.util
    import lavaframe_b200 as lf
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(golden_dir, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = np.load(os.path.join(golden_dir, "cornell64_llvmpipe_post.npz"))
    n = int(g["nspp"])
    accum = g["tm0"] * np.float32(n)            # exact: n is a power of two
    cases = []
    for name, (tm, vig, ca) in mg.POST_CASES.items():
        pp = lf.LfPostParams()
        if vig:
            pp.use_vignette, pp.vignette_intensity, pp.vignette_power = 1, vig[0], vig[1]
        if ca:
            pp.use_ca, pp.use_ca_distortion, pp.ca_distance, pp.ca_p1, pp.ca_p2, pp.ca_p3 = 1, ca[0], ca[1], ca[2], ca[3], ca[4]
        cases.append((name, tm, pp, g[name]))
    return accum, 1.0 / n, cases


def test_postprocess_vs_llvmpipe(golden_dir, oracle_lib):
    """the oracle's postprocess.glsl (6 tonemappers, vignette, both chromatic-aberration modes) against the
    reference's own GetOutputBufferHDR on llvmpipe, bit for bit.  One case samples the accumulation texture outside [0, 1]
    (MIRRORED_REPEAT): llvmpipe mirrors the coordinate in floating point before scaling, which moves 2 of 12 288 values by
    a few ulp; everything else is identical."""
    from oracle_api import post_process
    accum, inv, cases = _post_cases(golden_dir)
    for name, tm, pp, ref in cases:
        out = post_process(accum, inv, tm, pp)
        same = float(np.mean(out == ref))
        assert same >= (0.999 if name == "tm3_ca1_vig" else 1.0), f"{name}: {same:.6f} of the values bit-identical to llvmpipe"
        np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-7)
    # GetOutputBuffer (the 8-bit image the exporters of Export.h write) = round-to-nearest-even of clamp(colour, 0, 1) * 255
    g = np.load(os.path.join(golden_dir, "cornell64_llvmpipe_post.npz"))
    for name, tm, pp, ref in cases:
        if f"{name}_u8" in g:
            u8 = np.rint(np.clip(post_process(accum, inv, tm, pp), 0, 1) * np.float32(255)).astype(np.uint8)
            assert np.array_equal(u8, g[f"{name}_u8"]), f"{name}: 8-bit output differs from the reference's GetOutputBuffer"


@pytest.mark.parametrize("name", SCENES)
def test_no_denormals_on_the_path(golden_dir, oracle_lib, name):
    """llvmpipe runs with denormals flushed to zero; the oracle and the kernels keep them.  That is the one known arithmetic
    difference left, and it cannot matter if no denormal is ever consumed or produced: the MXCSR denormal-operand and underflow
    flags of every OpenMP worker stay clear over a multi-sample render (also checked by hand on tiles of the full-size C2 / C4)."""
    o = Oracle(_pack(golden_dir, name))
    oracle_lib.lforacle_fp_flags(1)
    o.render_frames(2, 8)
    flags = oracle_lib.lforacle_fp_flags(0)
    o.close()
    assert not flags & 0x02, "a denormal operand was consumed"
    assert not flags & 0x10, "a result underflowed into the denormal range"
