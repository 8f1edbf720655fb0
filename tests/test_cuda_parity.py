"""GPU tests (-m gpu): the CUDA path through the C ABI against the CPU oracle, the llvmpipe goldens, and
size-independent properties.  Every call goes through liblfcuda.so's extern "C" entry points."""
import os

import numpy as np
import pytest

from oracle_api import Oracle
from parity_metrics import hits_agreement, radiance_agreement, rmse_over_mean_luminance
import lavaframe_b200 as lf

pytestmark = pytest.mark.gpu

SCENES = ["cornell", "c2mini", "c3mini", "c4gold"]


def pack_path(golden_dir, name):
    return os.path.join(golden_dir, f"{name}.lfpack")


@pytest.fixture(scope="module")
def tracer(gpu):
    pt = lf.PathTracer(gpu)
    yield pt
    pt.close()


def masked(tri, mat, em):
    return np.where(em > 0, -1, tri), np.where(em > 0, -1, mat)


@pytest.mark.parametrize("name", SCENES)
def test_primary_hits_vs_oracle(tracer, golden_dir, oracle_lib, name):
    """Check 1 of the north_star, CUDA vs oracle: the arithmetic of a camera ray + traversal uses only + - * / sqrt,
    so it must be identical bit for bit."""
    pack = lf.ScenePack(pack_path(golden_dir, name))
    tracer.upload_pack(pack)
    t, tri, mat, em = tracer.primary_hits(2)
    o = Oracle(pack.path)
    ot, otri, omat, oem = o.primary_hits(2)
    o.close()
    assert np.array_equal(em > 0, oem > 0)
    tri, mat = masked(tri, mat, em)
    otri, omat = masked(otri, omat, oem)
    same, both = hits_agreement(t, tri, mat, ot, otri, omat)
    assert same == 1.0 and both == 1.0, f"{name}: ids {same:.6f}, ids+t {both:.6f}"
    assert np.array_equal(t, ot), f"{name}: t differs in {(t != ot).sum()} pixels, max rel {np.max(np.abs(t - ot) / ot):.3e}"


@pytest.mark.parametrize("name", SCENES)
def test_primary_hits_vs_llvmpipe(tracer, golden_dir, name):
    """Check 1 against the reference itself (golden from the unmodified GLSL on llvmpipe)."""
    g = np.load(os.path.join(golden_dir, f"{name}_llvmpipe.npz"))
    tracer.upload_pack(lf.ScenePack(pack_path(golden_dir, name)))
    t, tri, mat, em = tracer.primary_hits(2)
    tri, mat = masked(tri, mat, em)
    gtri, gmat = masked(g["hits_tri"], g["hits_mat"], g["hits_emitter"])
    same, both = hits_agreement(t, tri, mat, g["hits_t"], gtri, gmat)
    assert same >= 0.9999 and both >= 0.9999, f"{name}: ids {same:.6f}, ids+t(1e-5) {both:.6f}"


@pytest.mark.parametrize("name", SCENES)
@pytest.mark.parametrize("mode", [0, 1])
def test_radiance_vs_oracle(tracer, golden_dir, oracle_lib, name, mode):
    """Check 2, CUDA vs oracle, wavefront (mode 0) and megakernel (mode 1): 1-spp radiance of frame 2 and an 8-frame sum
    must be IDENTICAL BIT FOR BIT on every pixel - diffuse, glass, rough metal, clearcoat, textures, env map alike.
    Both sides evaluate the same fp32 expression trees without FMA contraction and the same plain-fp32 restatement of
    llvmpipe's transcendental kernels (lf_math.cuh / lf_math_oracle.h), so there is no tolerance to hide a bug in."""
    pack = lf.ScenePack(pack_path(golden_dir, name))
    tracer.upload_pack(pack, kernel_mode=mode)
    o = Oracle(pack.path)
    for first, n in ((2, 1), (3, 8)):
        tracer.clear()
        tracer.render_frames(first, n)
        img = tracer.read_accum()
        ref = o.render_frames(first, n)
        differ = int((img != ref).any(axis=2).sum())
        assert differ == 0, f"{name} mode {mode} frames {first}+{n}: {differ} pixels differ (within 1e-3 on {radiance_agreement(img, ref):.6f})"
    o.close()


# CUDA vs the reference on llvmpipe.  The kernels restate llvmpipe's own evaluation of every GLSL built-in, its x * (1 / y)
# division and its 8-bit texture filter, so the 1-spp image is bit-identical to the reference's on every pixel
# (tests/test_oracle_golden.py explains the one exception in multi-sample images: the reference's own undefined read of
# an unwritten material on pixels that look straight at an analytic light).
MIN_VS_LLVMPIPE = {"cornell": 0.9999, "c2mini": 0.9999, "c3mini": 0.9999, "c4gold": 0.9999}


@pytest.mark.parametrize("name", SCENES)
def test_radiance_vs_llvmpipe(tracer, golden_dir, name):
    """Checks 2 and 3 against the reference itself."""
    g = np.load(os.path.join(golden_dir, f"{name}_llvmpipe.npz"))
    tracer.upload_pack(lf.ScenePack(pack_path(golden_dir, name)))
    tracer.clear()
    tracer.render_frames(2, 1)
    img = tracer.read_accum()
    frac = radiance_agreement(img, g["spp1"])
    assert frac >= MIN_VS_LLVMPIPE[name], f"{name}: 1-spp within 1e-3 on {frac:.6f}"
    bits = float(np.mean((img == g["spp1"]).all(axis=2)))
    assert bits >= MIN_VS_LLVMPIPE[name], f"{name}: 1-spp bit-identical to the reference on llvmpipe on {bits:.6f} of pixels"
    n = int(g["nspp"])
    tracer.clear()
    tracer.render_frames(2, n)
    mean = tracer.read_output(1.0 / n, 0)
    assert rmse_over_mean_luminance(mean, g["sppN"]) < 0.05


def test_converged_4096spp_vs_llvmpipe(tracer, golden_dir):
    """Check 3 of the north_star: the converged 4096-spp Cornell image against the reference's own 4096-spp image
    (20 minutes of llvmpipe, frames 2..4097): RMSE below 0.5 % of the mean luminance."""
    g = np.load(os.path.join(golden_dir, "cornell_llvmpipe_4096spp.npz"))
    tracer.upload_pack(lf.ScenePack(pack_path(golden_dir, "cornell")))
    tracer.clear()
    tracer.render_frames(2, 4096)
    img = tracer.read_output(1.0 / 4096, 0)
    rel = rmse_over_mean_luminance(img, g["spp4096"])
    assert rel < 0.005, f"RMSE / mean luminance = {rel:.5f}"
    assert radiance_agreement(img, g["spp4096"]) >= 0.99


@pytest.mark.parametrize("name", SCENES)
def test_preview_engine(tracer, golden_dir, oracle_lib, name):
    """SURVEY 8(f) row 3, shaders/preview_flareon.glsl: lfcuda_render_preview against the oracle (bit for bit: the same
    wavefront kernels at another launch shape) and against the preview the unmodified reference drew on llvmpipe.  Also
    an odd size that ends in partial 8x4 blocks, at the scene's full depth, and that a preview leaves the accumulation
    buffer untouched."""
    pack = lf.ScenePack(pack_path(golden_dir, name))
    tracer.upload_pack(pack)
    tracer.clear()
    tracer.render_frames(2, 2)
    accum = tracer.read_accum()
    g = np.load(os.path.join(golden_dir, f"{name}_llvmpipe_preview.npz"))
    o = Oracle(pack.path)
    for key, dof in (("half", False), ("full_dof", True)):
        h, w, _ = g[key].shape
        img = tracer.render_preview(w, h, 2, dof)
        ref = o.render_preview(w, h, 2, dof)
        assert np.array_equal(img, ref), f"{name}/{key}: {int((img != ref).any(axis=2).sum())} pixels differ from the oracle"
        frac = radiance_agreement(img, g[key])
        assert frac >= (0.99 if name == "c3mini" else 0.9999), f"{name}/{key}: within 1e-3 of llvmpipe on {frac:.6f}"   # c3mini: emitter pixels
    img = tracer.render_preview(77, 45, pack.max_depth, False)
    assert np.array_equal(img, o.render_preview(77, 45, pack.max_depth, False))
    o.close()
    assert np.array_equal(tracer.read_accum(), accum)


@pytest.mark.parametrize("name", SCENES)
def test_wavefront_equals_megakernel(tracer, golden_dir, name):
    """Two kernel organisations of the same device functions must agree bit for bit."""
    pack = lf.ScenePack(pack_path(golden_dir, name))
    imgs = []
    for mode in (0, 1):
        tracer.upload_pack(pack, kernel_mode=mode)
        tracer.clear()
        tracer.render_frames(2, 3)
        imgs.append(tracer.read_accum())
    assert np.array_equal(imgs[0], imgs[1]), f"{(imgs[0] != imgs[1]).any(axis=2).sum()} pixels differ"


@pytest.mark.parametrize("name", SCENES)
def test_cull_does_not_change_results(tracer, golden_dir, name):
    """The distance cull is the one deliberate departure from the reference's traversal; it must be invisible."""
    pack = lf.ScenePack(pack_path(golden_dir, name))
    res = []
    for no_cull in (0, 1):
        tracer.upload_pack(pack, no_cull=no_cull)
        hits = tracer.primary_hits(2)
        tracer.clear()
        tracer.render_frames(2, 2)
        res.append((hits, tracer.read_accum()))
    for a, b in zip(res[0][0], res[1][0]):
        assert np.array_equal(a, b)
    assert np.array_equal(res[0][1], res[1][1])


def test_counters_match_oracle_without_cull(tracer, golden_dir, oracle_lib):
    """With the cull off the kernel walks the BVH exactly like the reference: every closest-hit ray makes the same
    visits.  Shadow rays whose NEE contribution is rejected whatever the visibility (bsdf pdf <= 0, MIS weight 0) are
    not traced by the kernel (the reference traces, then discards them), so the shadow-side counts are upper-bounded."""
    pack = lf.ScenePack(pack_path(golden_dir, "cornell"))
    tracer.upload_pack(pack, no_cull=1, count_work=1)
    tracer.reset_counters()
    tracer.clear()
    tracer.render_frames(2, 1)
    c = tracer.counters()
    o = Oracle(pack.path, cull=False, count=True)
    o.render_frames(2, 1)
    oc = o.counters()
    o.close()
    for k in ("samples", "rays_closest"):
        assert c[k] == oc[k], (k, c[k], oc[k])
    # shaded_hits: the reference re-fetches the (stale) material on emitter hits too (pathtrace.glsl:246-247); the kernel carries it
    for k in ("shaded_hits", "rays_shadow", "inner_visits", "leaf_visits", "tlas_visits", "tri_tests", "light_tests"):
        assert 0.5 * oc[k] <= c[k] <= oc[k], (k, c[k], oc[k])
    # closest-hit-only comparison: a depth-1 render traces no rays the reference would not
    tracer.update_params(max_depth=1)
    tracer.reset_counters(); tracer.clear(); tracer.render_frames(2, 1)
    c1 = tracer.counters()
    o = Oracle(pack.path, cull=False, count=True)
    o.update_params(max_depth=1)
    o.render_frames(2, 1)
    o1 = o.counters()
    o.close()
    assert c1["rays_closest"] == o1["rays_closest"] == 65536
    closest_only = {k: c1[k] for k in ("inner_visits", "leaf_visits", "tlas_visits", "tri_tests")}
    assert all(closest_only[k] <= o1[k] for k in closest_only)
    tracer.update_params(count_work=0)


def test_linearity_and_determinism(tracer, golden_dir):
    """Accumulation is a plain per-pixel sum in frame order: rendering frames 2..9 in one call, in two calls, or
    twice, gives identical bits; frame-strided subsets partition the sum (the multi-GPU split)."""
    pack = lf.ScenePack(pack_path(golden_dir, "cornell"))
    tracer.upload_pack(pack)
    tracer.clear(); tracer.render_frames(2, 8); a = tracer.read_accum()
    tracer.clear(); tracer.render_frames(2, 3); tracer.render_frames(5, 5); b = tracer.read_accum()
    tracer.clear(); tracer.render_frames(2, 8); c = tracer.read_accum()
    assert np.array_equal(a, b) and np.array_equal(a, c)
    tracer.clear(); tracer.render_frames(2, 4, 2); even = tracer.read_accum()
    tracer.clear(); tracer.render_frames(3, 4, 2); odd = tracer.read_accum()
    np.testing.assert_allclose(even + odd, a, rtol=2e-6, atol=1e-7)


def test_tiles_equal_full_frame(tracer, golden_dir):
    """A dividing tile size renders the same pixel-samples: tile steps with the frame numbers TiledRenderer would
    use for them reproduce per-tile what a single-tile run of those frames gives."""
    pack = lf.ScenePack(pack_path(golden_dir, "cornell"))
    tracer.upload_pack(pack)
    tracer.clear(); tracer.render_frames(7, 1); full = tracer.read_accum()
    tracer.upload_pack(pack, tile_width=128, tile_height=64)
    tracer.clear()
    for ty in range(4):
        for tx in range(2):
            tracer.render_frames(7, 1, 1, tx, ty)
    tiled = tracer.read_accum()
    frac = radiance_agreement(tiled, full, rel=1e-4)
    assert frac >= 0.999, frac


def test_postprocess_and_u8(tracer, golden_dir):
    pack = lf.ScenePack(pack_path(golden_dir, "cornell"))
    tracer.upload_pack(pack)
    tracer.clear(); tracer.render_frames(2, 4)
    acc = tracer.read_accum()
    lin = tracer.read_output(0.25, 0)
    np.testing.assert_array_equal(lin, acc * np.float32(0.25))
    u8 = tracer.read_output_u8(0.25, 0)
    np.testing.assert_array_equal(u8, np.rint(np.clip(lin, 0, 1) * 255).astype(np.uint8))
    aces = tracer.read_output(0.25, 2)
    c = lin.astype(np.float64)
    ref = np.clip((c * (2.51 * c + 0.03)) / (c * (2.43 * c + 0.59) + 0.14), 0, 1) ** (1 / 2.2)
    np.testing.assert_allclose(aces, ref, rtol=2e-5, atol=2e-6)


def _post_reference(acc, inv, tonemap, pp):
    """numpy restatement of shaders/postprocess.glsl:96-172 (float64) for the checks below."""
    H, W, _ = acc.shape
    a = acc.astype(np.float64)
    ys, xs = np.mgrid[0:H, 0:W]
    tu, tv = (xs + 0.5) / W, (ys + 0.5) / H

    def mirror(i, n):
        m = np.mod(i, 2 * n)
        return np.where(m < n, m, 2 * n - 1 - m)

    def linear(u, v, ch):
        x, y = u * W - 0.5, v * H - 0.5
        fx, fy = np.floor(x), np.floor(y)
        wx, wy = x - fx, y - fy
        x0, x1 = mirror(fx.astype(int), W), mirror(fx.astype(int) + 1, W)
        y0, y1 = mirror(fy.astype(int), H), mirror(fy.astype(int) + 1, H)
        top = a[y0, x0, ch] + (a[y0, x1, ch] - a[y0, x0, ch]) * wx
        bot = a[y1, x0, ch] + (a[y1, x1, ch] - a[y1, x0, ch]) * wx
        return top + (bot - top) * wy

    c = a * inv
    if pp.use_ca:
        dist = np.hypot(tu - pp.ca_p3, tv - pp.ca_p3) ** pp.ca_p1 * pp.ca_p2
        o = pp.ca_distance * dist if pp.use_ca_distortion else pp.ca_distance * 0.025 * pp.ca_p2
        c[..., 0] = linear(tu + o, tv + o, 0) * inv
        c[..., 2] = linear(tu - o, tv - o, 2) * inv
    if tonemap == 2:
        c = np.clip((c * (2.51 * c + 0.03)) / (c * (2.43 * c + 0.59) + 0.14), 0, 1) ** (1 / 2.2)
    elif tonemap == 3:
        c = np.clip(c / (c + 1), 0, 1) ** (1 / 2.2)
    if pp.use_vignette:
        c = c * (1.0 - np.hypot(tu - 0.5, tv - 0.5) ** pp.vignette_power * pp.vignette_intensity)[..., None]
    return c


def test_postprocess_ca_and_vignette(tracer, golden_dir):
    """chromatic aberration (both modes) and vignette of postprocess.glsl."""
    pack = lf.ScenePack(pack_path(golden_dir, "cornell"))
    tracer.upload_pack(pack)
    tracer.clear(); tracer.render_frames(2, 4)
    acc = tracer.read_accum()
    for use_dist, tonemap in ((0, 0), (1, 2), (0, 3)):
        pp = lf.LfPostParams()
        pp.use_ca, pp.use_ca_distortion, pp.ca_distance, pp.ca_p1, pp.ca_p2, pp.ca_p3 = 1, use_dist, 0.05, 5.0, -0.5, 0.5
        pp.use_vignette, pp.vignette_intensity, pp.vignette_power = 1, 0.8, 1.5
        tracer.set_post(pp)
        out = tracer.read_output(0.25, tonemap)
        ref = _post_reference(acc, 0.25, tonemap, pp)
        np.testing.assert_allclose(out, ref, rtol=2e-4, atol=1e-3)   # fp32 vs float64 restatement; pow(x, 1/2.2) is steep at 0
    tracer.set_post(None)
    np.testing.assert_array_equal(tracer.read_output(0.25, 0), acc * np.float32(0.25))


def test_postprocess_vs_oracle_and_llvmpipe(tracer, golden_dir, oracle_lib):
    """SURVEY 8(f) row 1 at the parity bar of the hot path: k_post (6 tonemappers, vignette, both chromatic-aberration modes)
    bit-identical to the oracle's PostProcess, and to the reference's own GetOutputBufferHDR on llvmpipe
    (tests/golden/cornell64_llvmpipe_post.npz: Cornell at 64x64, 4 spp); the 4-spp image itself must match the reference's."""
    from oracle_api import post_process
    from test_oracle_golden import _post_cases
    accum, inv, cases = _post_cases(golden_dir)
    H, W, _ = accum.shape
    pack = lf.ScenePack(pack_path(golden_dir, "cornell"))
    tracer.upload_pack(pack, width=W, height=H, tile_width=W, tile_height=H)
    tracer.clear(); tracer.render_frames(2, int(round(1 / inv)))
    acc = tracer.read_accum()
    assert np.array_equal(acc, accum), f"the {int(round(1 / inv))}-spp accumulation differs from the reference's on {(acc != accum).any(axis=2).mean():.6f} of pixels"
    for name, tm, pp, ref in cases:
        tracer.set_post(pp)
        out = tracer.read_output(inv, tm)
        assert np.array_equal(out, post_process(acc, inv, tm, pp)), f"{name}: k_post differs from the oracle"
        same = float(np.mean(out == ref))
        assert same >= (0.999 if name == "tm3_ca1_vig" else 1.0), f"{name}: {same:.6f} of the values bit-identical to llvmpipe"
    tracer.set_post(None)


def test_errors_are_reported(gpu):
    pt = lf.PathTracer(gpu)
    with pytest.raises(lf.LfCudaError, match="no scene"):
        pt.render_frames(2, 1)
    pt.close()
