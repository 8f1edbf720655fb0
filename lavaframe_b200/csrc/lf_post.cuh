// lf_post.cuh — the post-process pass (shaders/postprocess.glsl:26-172) for one pixel, as a device function shared by k_post
// and tests/hostcheck.  Arithmetic as llvmpipe evaluates the GLSL (lf_math.cuh, fdiv): pinned bit for bit against the
// reference's own tonemapped / vignetted / chromatic-aberration output (tests/golden/cornell64_llvmpipe_post.npz).
#pragma once

#include "lf_device.cuh"

namespace lf {

LFD float tm_aces(float c) { return clampf(fdiv(c * (2.51f * c + 0.03f), c * (2.43f * c + 0.59f) + 0.14f), 0.0f, 1.0f); }   // :34-43
LFD float tm_kanjero(float c) {                                                                                             // :69-84
    float v = lf_pow(fdiv(c * (c * (1.2295f * c + 0.3135f) + 1.1935f * 0.4655f), c * (1.1935f * c + 0.4655f) + 0.073f), 1.7f);
    v = lf_pow(v, 1.0f / 0.8f);
    v *= 0.8f;
    return clampf(v, 0.0f, 1.0f);
}
LFD float tm_hejl(float c) { c = gmax(0.0f, c - 0.004f); return fdiv(c * (6.2f * c + .5f), c * (6.2f * c + 1.7f) + 0.06f); }   // :52-56
LFD float tm_uncharted(float c) {                                                                                           // :87-96
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return fdiv(c * (A * c + C * B) + D * E, c * (A * c + B) + D * F) - fdiv(E, F);
}
// Where the accumulated sum comes from.  AccumOne: one buffer (a single context).  AccumSum: the buffers of the contexts of a
// multi-GPU group (spp split: every device holds the sum of ITS frames); the pass reads all of them - peers over NVLink by plain
// loads through peer access - and adds them in device order, so the all-device image is formed inside the post-process pass itself
// and never stored: the reduce and the tonemap are one kernel.
constexpr int kMaxGroup = 16;
struct AccumOne {
    const float* p;
    LFD float operator[](size_t i) const { return p[i]; }
};
struct AccumSum {
    const float* p[kMaxGroup];
    int n;
    LFD float operator[](size_t i) const {
        float s = p[0][i];
        for (int k = 1; k < n; k++) s = s + p[k][i];
        return s;
    }
};
// accumTexture is LINEAR / MIRRORED_REPEAT (TiledRenderer.cpp:165-173): bilinear fetch at a normalised coordinate
LFD int mirrori(int i, int n) { int m = i % (2 * n); if (m < 0) m += 2 * n; return m < n ? m : 2 * n - 1 - m; }
template <class A>
LFD float accum_linear(const A& accum, int W, int H, float u, float v, int ch) {
    float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float wx = x - fx, wy = y - fy;
    int x0 = mirrori((int)fx, W), x1 = mirrori((int)fx + 1, W), y0 = mirrori((int)fy, H), y1 = mirrori((int)fy + 1, H);
    float a = accum[3 * ((size_t)y0 * W + x0) + ch], b = accum[3 * ((size_t)y0 * W + x1) + ch];
    float c = accum[3 * ((size_t)y1 * W + x0) + ch], e = accum[3 * ((size_t)y1 * W + x1) + ch];
    float top = a + (b - a) * wx, bot = c + (e - c) * wx;
    return top + (bot - top) * wy;
}

// main() of postprocess.glsl for pixel i of a W x H image (rows bottom-up): c = the colour written to the output texture
template <class A>
LFD void post_pixel(const A& accum, int W, int H, int i, float inv, int tonemap, const LfPostParams& pp, float c[3]) {
    const int px = i % W, py = i / W;
    const float tu = ((float)px + 0.5f) / (float)W, tv = ((float)py + 0.5f) / (float)H;   // TexCoords of the fullscreen quad
    c[0] = accum[3 * i] * inv; c[1] = accum[3 * i + 1] * inv; c[2] = accum[3 * i + 2] * inv;
    if (pp.use_ca) {   // chromaticAberration(), :98-118: red and blue fetched at +/- an offset
        float offset = pp.ca_distance;
        float dx = tu - pp.ca_p3, dy = tv - pp.ca_p3;
        float dist = 0.f + (lf_pow(sqrtf(dx * dx + dy * dy), pp.ca_p1) * pp.ca_p2);
        float o = pp.use_ca_distortion ? offset * dist : (offset * 0.025f) * pp.ca_p2;
        c[0] = accum_linear(accum, W, H, tu + o, tv + o, 0) * inv;
        c[2] = accum_linear(accum, W, H, tu - o, tv - o, 2) * inv;
    }
    const float g = 1.0f / 2.2f;
    if (tonemap == 1) {        // pow(tonemap(color, 2), 1 / 2.2): c * 1.0 / (1.0 + luminance / limit), :26-31
        float lum = (0.3f * c[0] + 0.6f * c[1]) + 0.1f * c[2];
        float r = 1.0f / (1.0f + fdiv(lum, 2.f));
        for (int k = 0; k < 3; k++) c[k] = lf_pow((c[k] * 1.0f) * r, g);
    } else if (tonemap == 2) {
        for (int k = 0; k < 3; k++) c[k] = lf_pow(tm_aces(c[k]), g);
    } else if (tonemap == 3) { // Reinhard, :46-49
        for (int k = 0; k < 3; k++) c[k] = lf_pow(clampf(fdiv(c[k], c[k] + 1.f), 0.0f, 1.0f), g);
    } else if (tonemap == 4) {
        for (int k = 0; k < 3; k++) c[k] = lf_pow(tm_kanjero(c[k]), g);
    } else if (tonemap == 5) {
        for (int k = 0; k < 3; k++) c[k] = tm_hejl(c[k]);
    } else if (tonemap == 6) {
        for (int k = 0; k < 3; k++) c[k] = lf_pow(tm_uncharted(c[k]), g) * 1.75f;
    }
    if (pp.use_vignette) {     // vignette(), :121-124
        float dx = tu - 0.5f, dy = tv - 0.5f;
        float d = 1.0f - lf_pow(sqrtf(dx * dx + dy * dy), pp.vignette_power) * pp.vignette_intensity;
        for (int k = 0; k < 3; k++) c[k] *= d;
    }
}

LFD void post_pixel(const float* accum, int W, int H, int i, float inv, int tonemap, const LfPostParams& pp, float c[3]) {
    AccumOne a; a.p = accum;
    post_pixel(a, W, H, i, inv, tonemap, pp, c);
}

}  // namespace lf
