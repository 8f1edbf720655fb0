// lf_oracle.cpp — TEST INFRASTRUCTURE: CPU restatement of the reference's path-tracing shader.
// See lf_oracle.h.  Citations are file:line under /root/reference/shaders/ unless a directory is given.
//
// GLSL built-ins are restated the way Mesa's GLSL front end + llvmpipe evaluate them (the only runnable
// reference implementation): normalize(v) = v * (1/sqrt(dot(v,v))), min/max with
// x86 MINPS/MAXPS operand semantics, dot() summed left to right, mat*vec summed column by column,
// inverse() by cofactors times 1/det, x / y = x * (1 / y) (Mesa lowers every float division to MUL(x, RCP(y)) and
// gallivm's RCP is the correctly rounded 1 / y), mix(a, b, t) = a * (1 - t) + b * t, refract() with
// k = 1 - eta * (eta * (1 - d * d)).  Transcendentals: lf_math_oracle.h restates llvmpipe's own polynomial
// evaluations.  All of these are pinned bit for bit by executing the GLSL on llvmpipe (oracle/_ref/lp_probe,
// tests/golden/llvmpipe_builtins.npz).
#include "lf_oracle.h"
#include "lf_math_oracle.h"

#include <cmath>
#include <cstring>
#include <omp.h>

namespace lforacle {

// ------------------------------------------------------------------------------------------------
// GLSL-like value types
// ------------------------------------------------------------------------------------------------
struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };

static inline vec3 V3(float a) { return {a, a, a}; }
static inline vec3 V3(float a, float b, float c) { return {a, b, c}; }
static inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
// GLSL x / y on Mesa: x * (1 / y), two roundings (FDIV_TO_MUL_RCP; pinned with lp_probe)
static inline float fdiv(float a, float b) { return a * (1.0f / b); }
static inline vec3 operator/(vec3 a, vec3 b) { return {fdiv(a.x, b.x), fdiv(a.y, b.y), fdiv(a.z, b.z)}; }
static inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline vec3 operator*(float s, vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline vec3 operator/(vec3 a, float s) { float r = 1.0f / s; return {a.x * r, a.y * r, a.z * r}; }
static inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
static inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
static inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
static inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
static inline vec3& operator/=(vec3& a, float s) { a = a / s; return a; }
static inline vec2 operator+(vec2 a, vec2 b) { return {a.x + b.x, a.y + b.y}; }
static inline vec2 operator*(vec2 a, float s) { return {a.x * s, a.y * s}; }

static inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
static inline float inversesqrt(float x) { return 1.0f / sqrtf(x); }
static inline vec3 normalize(vec3 v) { return v * inversesqrt(dot(v, v)); }
static inline float length(vec3 v) { return sqrtf(dot(v, v)); }
// x86 MAXPS/MINPS(a, b): second operand when either is NaN
static inline float gmax(float a, float b) { return a > b ? a : b; }
static inline float gmin(float a, float b) { return a < b ? a : b; }
static inline vec3 gmax(vec3 a, vec3 b) { return {gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z)}; }
static inline vec3 gmin(vec3 a, vec3 b) { return {gmin(a.x, b.x), gmin(a.y, b.y), gmin(a.z, b.z)}; }
static inline float clampf(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
static inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }   // as Mesa lowers mix() for llvmpipe (pinned with lp_probe)
static inline vec3 mix3(vec3 a, vec3 b, float t) { return {mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)}; }
static inline vec3 reflect(vec3 I, vec3 N) { return I - (2.0f * dot(N, I)) * N; }
static inline vec3 refract(vec3 I, vec3 N, float eta) {
    float ndi = dot(N, I);
    float k = 1.0f - eta * (eta * (1.0f - ndi * ndi));   // Mesa's builtin: mul(eta, mul(eta, ...))
    if (k < 0.0f) return V3(0.0f);
    return eta * I - (eta * ndi + sqrtf(k)) * N;
}
static inline vec3 pow3(vec3 a, float e) { return {lfom::pow(a.x, e), lfom::pow(a.y, e), lfom::pow(a.z, e)}; }
static inline vec3 exp3(vec3 a) { return {lfom::exp(a.x), lfom::exp(a.y), lfom::exp(a.z)}; }
static inline vec3 log3(vec3 a) { return {lfom::log(a.x), lfom::log(a.y), lfom::log(a.z)}; }

struct mat4 { vec4 c[4]; };   // columns
struct mat3 { vec3 c[3]; };

static inline vec4 mul(const mat4& m, vec4 v) {
    vec4 r;
    r.x = ((m.c[0].x * v.x + m.c[1].x * v.y) + m.c[2].x * v.z) + m.c[3].x * v.w;
    r.y = ((m.c[0].y * v.x + m.c[1].y * v.y) + m.c[2].y * v.z) + m.c[3].y * v.w;
    r.z = ((m.c[0].z * v.x + m.c[1].z * v.y) + m.c[2].z * v.z) + m.c[3].z * v.w;
    r.w = ((m.c[0].w * v.x + m.c[1].w * v.y) + m.c[2].w * v.z) + m.c[3].w * v.w;
    return r;
}
static inline vec3 mul(const mat3& m, vec3 v) {
    return {(m.c[0].x * v.x + m.c[1].x * v.y) + m.c[2].x * v.z,
            (m.c[0].y * v.x + m.c[1].y * v.y) + m.c[2].y * v.z,
            (m.c[0].z * v.x + m.c[1].z * v.y) + m.c[2].z * v.z};
}

// inverse(mat4): adjugate by 2x2 sub-factors, times 1/det (GLM / Mesa builtin formulation).
static mat4 inverse(const mat4& mm) {
    float m[4][4];
    for (int c = 0; c < 4; c++) { m[c][0] = mm.c[c].x; m[c][1] = mm.c[c].y; m[c][2] = mm.c[c].z; m[c][3] = mm.c[c].w; }
    float s00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
    float s01 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    float s02 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
    float s03 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    float s04 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
    float s05 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    float s06 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
    float s07 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
    float s08 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
    float s09 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
    float s10 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
    float s11 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
    float s12 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    float s13 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    float s14 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    float s15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    float s16 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    float s17 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    float a[4][4];
    a[0][0] = +((m[1][1] * s00 - m[1][2] * s01) + m[1][3] * s02);
    a[0][1] = -((m[0][1] * s00 - m[0][2] * s01) + m[0][3] * s02);
    a[0][2] = +((m[0][1] * s06 - m[0][2] * s07) + m[0][3] * s08);
    a[0][3] = -((m[0][1] * s12 - m[0][2] * s13) + m[0][3] * s14);
    a[1][0] = -((m[1][0] * s00 - m[1][2] * s03) + m[1][3] * s04);
    a[1][1] = +((m[0][0] * s00 - m[0][2] * s03) + m[0][3] * s04);
    a[1][2] = -((m[0][0] * s06 - m[0][2] * s09) + m[0][3] * s10);
    a[1][3] = +((m[0][0] * s12 - m[0][2] * s15) + m[0][3] * s16);
    a[2][0] = +((m[1][0] * s01 - m[1][1] * s03) + m[1][3] * s05);
    a[2][1] = -((m[0][0] * s01 - m[0][1] * s03) + m[0][3] * s05);
    a[2][2] = +((m[0][0] * s07 - m[0][1] * s09) + m[0][3] * s11);
    a[2][3] = -((m[0][0] * s13 - m[0][1] * s15) + m[0][3] * s17);
    a[3][0] = -((m[1][0] * s02 - m[1][1] * s04) + m[1][2] * s05);
    a[3][1] = +((m[0][0] * s02 - m[0][1] * s04) + m[0][2] * s05);
    a[3][2] = -((m[0][0] * s08 - m[0][1] * s10) + m[0][2] * s11);
    a[3][3] = +((m[0][0] * s14 - m[0][1] * s16) + m[0][2] * s17);
    float det = ((m[0][0] * a[0][0] + m[0][1] * a[1][0]) + m[0][2] * a[2][0]) + m[0][3] * a[3][0];
    float inv = 1.0f / det;
    mat4 r;
    for (int c = 0; c < 4; c++) r.c[c] = {a[c][0] * inv, a[c][1] * inv, a[c][2] * inv, a[c][3] * inv};
    return r;
}

static mat3 inverse(const mat3& mm) {
    float m[3][3] = {{mm.c[0].x, mm.c[0].y, mm.c[0].z}, {mm.c[1].x, mm.c[1].y, mm.c[1].z}, {mm.c[2].x, mm.c[2].y, mm.c[2].z}};
    float a[3][3];
    a[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]);
    a[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]);
    a[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]);
    a[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]);
    a[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]);
    a[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]);
    a[0][2] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]);
    a[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]);
    a[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]);
    float det = (m[0][0] * a[0][0] + m[0][1] * a[1][0]) + m[0][2] * a[2][0];
    float inv = 1.0f / det;
    mat3 r;
    for (int c = 0; c < 3; c++) r.c[c] = {a[c][0] * inv, a[c][1] * inv, a[c][2] * inv};
    return r;
}
static inline mat3 transpose(const mat3& m) {
    return {{{m.c[0].x, m.c[1].x, m.c[2].x}, {m.c[0].y, m.c[1].y, m.c[2].y}, {m.c[0].z, m.c[1].z, m.c[2].z}}};
}

// ------------------------------------------------------------------------------------------------
// common/globals.glsl
// ------------------------------------------------------------------------------------------------
static const float PI = 3.14159265358979323f;       // globals.glsl:6
static const float TWO_PI = 6.28318530717958648f;   // :7
static const float INFINITY_ = 1000000.0f;          // :8
static const float EPS = 0.001f;                    // :9

struct Ray { vec3 origin, direction; };               // :18-22
struct Material {                                     // :24-46
    vec3 albedo; float specular; vec3 emission; float anisotropic;
    float metallic, roughness, subsurface, specularTint, sheen, sheenTint, clearcoat, clearcoatRoughness;
    float specTrans, ior, atDistance; vec3 extinction; vec4 texIDs;
};
struct Light { vec3 position, emission, u, v; float radius, area, type; };   // :60-69
struct State {                                        // :71-90
    int depth; float eta, hitDist;
    vec3 fhp, normal, ffnormal, tangent, bitangent;
    bool isEmitter; vec2 texCoord; vec3 bary; int triID[3]; int matID; Material mat;
};
struct BsdfSampleRec { vec3 L, f; float pdf; };       // :92-97
struct LightSampleRec { vec3 normal, emission, direction; float dist, pdf; };   // :99-106

// Per-invocation globals of the shader (globals.glsl:15-16,113-114) + the bound "textures".
struct Inv {
    const Oracle* o;
    mat4 transform;
    vec3 tempTexCoords;
    uint32_t seed[4];
    LfCounters* cnt;
};

// globals.glsl:122-128
static inline void pcg4d(uint32_t v[4]) {
    for (int i = 0; i < 4; i++) v[i] = v[i] * 1664525u + 1013904223u;
    v[0] += v[1] * v[3]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]; v[3] += v[1] * v[2];
    for (int i = 0; i < 4; i++) v[i] ^= v[i] >> 16u;
    v[0] += v[1] * v[3]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]; v[3] += v[1] * v[2];
}
// globals.glsl:130-133: float(seed.x) / float(0xffffffffu); float(0xffffffffu) rounds to 2^32.
static inline float rnd(Inv& g) {
    pcg4d(g.seed);
    return (float)g.seed[0] / 4294967296.0f;
}
// globals.glsl:116-120
static inline void InitRNG(Inv& g, vec2 p, int frame) {
    g.seed[0] = (uint32_t)p.x; g.seed[1] = (uint32_t)p.y; g.seed[2] = (uint32_t)frame;
    g.seed[3] = (uint32_t)p.x + (uint32_t)p.y;
}

// ------------------------------------------------------------------------------------------------
// "texture units" (LavaFrame/Renderer.cpp:87-185)
// ------------------------------------------------------------------------------------------------
static inline vec3 bvhTexel(const Oracle* o, int texel) { const float* p = o->scene.bvh_nodes + 3 * (size_t)texel; return {p[0], p[1], p[2]}; }
static inline vec4 vtx(const Oracle* o, int i) { const float* p = o->scene.vertices_uvx + 4 * (size_t)i; return {p[0], p[1], p[2], p[3]}; }
static inline vec4 nrm(const Oracle* o, int i) { const float* p = o->scene.normals_uvy + 4 * (size_t)i; return {p[0], p[1], p[2], p[3]}; }
static inline Light fetchLight(const Oracle* o, int i) {
    // texelFetch outside the texture returns zeros on llvmpipe (reachable only through rand() == 1.0, pathtrace.glsl:168)
    Light l;
    std::memset(&l, 0, sizeof l);
    if (i < 0 || i >= o->scene.num_lights) return l;
    const float* p = o->scene.lights + 15 * (size_t)i;
    l.position = {p[0], p[1], p[2]}; l.emission = {p[3], p[4], p[5]}; l.u = {p[6], p[7], p[8]}; l.v = {p[9], p[10], p[11]};
    l.radius = p[12]; l.area = p[13]; l.type = p[14];
    return l;
}
static inline int wrapi(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }
// GL_NEAREST, GL_REPEAT: texel floor(u*W) mod W
static inline int nearestIdx(float u, int n) { return wrapi((int)floorf(u * (float)n), n); }
// GL_LINEAR, GL_REPEAT on an RGB32F texture: exact fp32 lerp (measured on llvmpipe)
static vec3 hdrLinear(const Oracle* o, float u, float v) {
    int W = o->scene.hdr_width, H = o->scene.hdr_height;
    float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float wx = x - fx, wy = y - fy;
    int x0 = wrapi((int)fx, W), x1 = wrapi((int)fx + 1, W), y0 = wrapi((int)fy, H), y1 = wrapi((int)fy + 1, H);
    const float* c = o->scene.hdr_cols;
    auto T = [&](int xx, int yy) { const float* p = c + 3 * ((size_t)yy * W + xx); return vec3{p[0], p[1], p[2]}; };
    vec3 a = T(x0, y0), b = T(x1, y0), cc = T(x0, y1), d = T(x1, y1);
    vec3 top = a + (b - a) * wx, bot = cc + (d - cc) * wx;
    return top + (bot - top) * wy;
}
static inline vec2 marginal(const Oracle* o, float u) { const float* p = o->scene.hdr_marginal + 2 * (size_t)nearestIdx(u, o->scene.hdr_height); return {p[0], p[1]}; }
static inline vec2 conditional(const Oracle* o, float u, float v) {
    int W = o->scene.hdr_width, H = o->scene.hdr_height;
    const float* p = o->scene.hdr_conditional + 2 * ((size_t)nearestIdx(v, H) * W + nearestIdx(u, W));
    return {p[0], p[1]};
}
// GL_RGBA8 2D array, GL_LINEAR, GL_REPEAT the way llvmpipe filters 8-bit formats (its AoS path, pinned bit for bit with
// oracle/_ref/lp_probe): texel coordinates in 24.8 fixed point, X = round_to_nearest_even(u * W * 256) - 128, texel
// floor(X / 256) and its +1 neighbour (wrapped), weight X & 255; each lerp is v0 + ((w * (v1 - v0)) >> 8) on the 8-bit
// channel values (floor, no rounding), x first, then y; the 8-bit result times the float constant 1/255.
static inline int lerp8(int w, int v0, int v1) { return (v0 + ((w * (v1 - v0)) >> 8)) & 255; }
static vec4 tex8Linear(const uint8_t* base, int W, int H, float u, float v) {
    int X = (int)nearbyintf(u * (float)W * 256.0f) - 128, Y = (int)nearbyintf(v * (float)H * 256.0f) - 128;
    int wx = X & 255, wy = Y & 255;
    int x0 = wrapi(X >> 8, W), x1 = wrapi((X >> 8) + 1, W), y0 = wrapi(Y >> 8, H), y1 = wrapi((Y >> 8) + 1, H);
    const uint8_t* a = base + 4 * ((size_t)y0 * W + x0);
    const uint8_t* b = base + 4 * ((size_t)y0 * W + x1);
    const uint8_t* c = base + 4 * ((size_t)y1 * W + x0);
    const uint8_t* d = base + 4 * ((size_t)y1 * W + x1);
    float r[4];
    for (int k = 0; k < 4; k++) r[k] = (float)lerp8(wy, lerp8(wx, a[k], b[k]), lerp8(wx, c[k], d[k])) * (float)(1.0 / 255.0);
    return {r[0], r[1], r[2], r[3]};
}
static vec4 texArrayLinear(const Oracle* o, float u, float v, int layer, LfCounters* cnt) {
    int W = o->scene.tex_width, H = o->scene.tex_height;
    if (layer < 0) layer = 0;
    if (layer >= o->scene.num_textures) layer = o->scene.num_textures - 1;
    if (cnt) cnt->tex_samples++;
    return tex8Linear(o->scene.texture_maps + (size_t)4 * W * H * layer, W, H, u, v);
}

// ------------------------------------------------------------------------------------------------
// common/intersection.glsl
// ------------------------------------------------------------------------------------------------
static float SphereIntersect(float rad, vec3 pos, const Ray& r) {   // intersection.glsl:7-27
    vec3 op = pos - r.origin;
    float eps = 0.001f;
    float b = dot(op, r.direction);
    float det = b * b - dot(op, op) + rad * rad;
    if (det < 0.0f) return INFINITY_;
    det = sqrtf(det);
    float t1 = b - det;
    if (t1 > eps) return t1;
    float t2 = b + det;
    if (t2 > eps) return t2;
    return INFINITY_;
}
static float RectIntersect(vec3 pos, vec3 u, vec3 v, vec4 plane, const Ray& r) {   // intersection.glsl:30-50
    vec3 n = {plane.x, plane.y, plane.z};
    float dt = dot(r.direction, n);
    float t = fdiv(plane.w - dot(n, r.origin), dt);
    if (t > EPS) {
        vec3 p = r.origin + r.direction * t;
        vec3 vi = p - pos;
        float a1 = dot(u, vi);
        if (a1 >= 0.f && a1 <= 1.f) {
            float a2 = dot(v, vi);
            if (a2 >= 0.f && a2 <= 1.f) return t;
        }
    }
    return INFINITY_;
}
static float AABBIntersect(vec3 minCorner, vec3 maxCorner, const Ray& r, float* entry) {   // intersection.glsl:53-67
    vec3 invdir = V3(1.0f) / r.direction;
    vec3 f = (maxCorner - r.origin) * invdir;
    vec3 n = (minCorner - r.origin) * invdir;
    vec3 tmax = gmax(f, n);
    vec3 tmin = gmin(f, n);
    float t1 = gmin(tmax.x, gmin(tmax.y, tmax.z));
    float t0 = gmax(tmin.x, gmax(tmin.y, tmin.z));
    *entry = t0;
    return (t1 >= t0) ? (t0 > 0.f ? t0 : t1) : -1.0f;
}

// ------------------------------------------------------------------------------------------------
// common/closest_hit.glsl + common/anyhit.glsl — one traversal, two modes
// ------------------------------------------------------------------------------------------------
// Conservative distance cull used only when Oracle::cull is set (the reference never culls): a child box
// whose entry distance exceeds the best hit by this relative margin cannot contain a closer hit.
static const float CULL_SLACK = 1.0001f;

template <bool ANY>
static float Traverse(Inv& g, const Ray& r, State& state, LightSampleRec& lightSampleRec, float maxDist, bool* anyHit) {
    const Oracle* o = g.o;
    float t = INFINITY_;
    float d;
    if (g.cnt) { if (ANY) g.cnt->rays_shadow++; else g.cnt->rays_closest++; }

    if (o->scene.num_lights > 0) {   // #ifdef LIGHTS  (closest_hit.glsl:13-67, anyhit.glsl:11-46)
        for (int i = 0; i < o->scene.num_lights; i++) {
            Light L = fetchLight(o, i);
            if (g.cnt) { g.cnt->light_tests++; if (ANY) g.cnt->light_tests_shadow++; }
            vec3 u = L.u, v = L.v;
            if (L.type == 0.f) {
                vec3 normal = normalize(cross(u, v));
                if (!ANY && dot(normal, r.direction) > 0.f) continue;   // closest_hit.glsl:31-32 (any-hit is two-sided)
                vec4 plane = {normal.x, normal.y, normal.z, dot(normal, L.position)};
                u *= 1.0f / dot(u, u);
                v *= 1.0f / dot(v, v);
                d = RectIntersect(L.position, u, v, plane, r);
                if (ANY) {
                    if (d > 0.0f && d < maxDist) { *anyHit = true; return d; }
                } else {
                    if (d < 0.f) d = INFINITY_;
                    if (d < t) {
                        t = d;
                        float cosTheta = dot(-r.direction, normal);
                        float pdf = fdiv(t * t, L.area * cosTheta);
                        lightSampleRec.emission = L.emission;
                        lightSampleRec.pdf = pdf;
                        state.isEmitter = true;
                    }
                }
            }
            if (L.type == 1.f) {
                d = SphereIntersect(L.radius, L.position, r);
                if (ANY) {
                    if (d > 0.0f && d < maxDist) { *anyHit = true; return d; }
                } else {
                    if (d < 0.f) d = INFINITY_;
                    if (d < t) {
                        t = d;
                        float pdf = fdiv(t * t, L.area);
                        lightSampleRec.emission = L.emission;
                        lightSampleRec.pdf = pdf;
                        state.isEmitter = true;
                    }
                }
            }
        }
    }

    int stack[64];
    int ptr = 0;
    stack[ptr++] = -1;
    int idx = o->scene.top_bvh_index;
    float leftHit = 0.0f, rightHit = 0.0f;
    int currMatID = 0;
    bool meshBVH = false;
    Ray r_trans = r;
    mat4 temp_transform;
    std::memset(&temp_transform, 0, sizeof temp_transform);
    const bool cull = o->cull;
    const float limit0 = ANY ? maxDist : 0.f;

    while (idx > -1 || meshBVH) {
        int n = idx;
        if (meshBVH && idx < 0) {
            meshBVH = false;
            idx = stack[--ptr];
            r_trans = r;
            continue;
        }
        vec3 lr = bvhTexel(o, n * 3 + 2);
        int leftIndex = (int)lr.x, rightIndex = (int)lr.y, leaf = (int)lr.z;

        if (leaf > 0) {   // BLAS leaf: closest_hit.glsl:108-147 / anyhit.glsl:87-116
            if (g.cnt) { g.cnt->leaf_visits++; if (ANY) g.cnt->leaf_visits_shadow++; }
            for (int i = 0; i < rightIndex; i++) {
                int index = leftIndex + i;
                const int32_t* vi = o->scene.vert_indices + 3 * (size_t)index;
                vec4 v0 = vtx(o, vi[0]), v1 = vtx(o, vi[1]), v2 = vtx(o, vi[2]);
                if (g.cnt) { g.cnt->tri_tests++; if (ANY) g.cnt->tri_tests_shadow++; }
                vec3 p0 = {v0.x, v0.y, v0.z};
                vec3 e0 = vec3{v1.x, v1.y, v1.z} - p0;
                vec3 e1 = vec3{v2.x, v2.y, v2.z} - p0;
                vec3 pv = cross(r_trans.direction, e1);
                float det = dot(e0, pv);
                vec3 tv = r_trans.origin - p0;
                vec3 qv = cross(tv, e0);
                vec4 uvt;
                uvt.x = dot(tv, pv);
                uvt.y = dot(r_trans.direction, qv);
                uvt.z = dot(e1, qv);
                float rdet = 1.0f / det;   // uvt.xyz / det = uvt.xyz * rcp(det)
                uvt.x = uvt.x * rdet; uvt.y = uvt.y * rdet; uvt.z = uvt.z * rdet;
                uvt.w = 1.0f - uvt.x - uvt.y;
                bool inside = uvt.x >= 0.f && uvt.y >= 0.f && uvt.z >= 0.f && uvt.w >= 0.f;
                if (ANY) {
                    if (inside && uvt.z < maxDist) { *anyHit = true; return uvt.z; }
                } else if (inside && uvt.z < t) {
                    t = uvt.z;
                    state.isEmitter = false;
                    state.triID[0] = vi[0]; state.triID[1] = vi[1]; state.triID[2] = vi[2];
                    state.matID = currMatID;
                    state.fhp = r_trans.origin + r_trans.direction * t;
                    state.bary = {uvt.w, uvt.x, uvt.y};
                    g.tempTexCoords = {v0.w, v1.w, v2.w};
                    vec4 w = mul(temp_transform, vec4{state.fhp.x, state.fhp.y, state.fhp.z, 1.0f});
                    state.fhp = {w.x, w.y, w.z};
                    g.transform = temp_transform;
                }
            }
        } else if (leaf < 0) {   // TLAS leaf: closest_hit.glsl:148-166
            if (g.cnt) { g.cnt->tlas_visits++; if (ANY) g.cnt->tlas_visits_shadow++; }
            idx = leftIndex;
            const float* m = o->scene.transforms + 16 * (size_t)(-leaf - 1);
            for (int c = 0; c < 4; c++) temp_transform.c[c] = {m[4 * c + 0], m[4 * c + 1], m[4 * c + 2], m[4 * c + 3]};
            mat4 invm = inverse(temp_transform);
            vec4 oo = mul(invm, vec4{r.origin.x, r.origin.y, r.origin.z, 1.0f});
            vec4 dd = mul(invm, vec4{r.direction.x, r.direction.y, r.direction.z, 0.0f});
            r_trans.origin = {oo.x, oo.y, oo.z};
            r_trans.direction = {dd.x, dd.y, dd.z};
            stack[ptr++] = -1;
            meshBVH = true;
            currMatID = rightIndex;
            continue;
        } else {   // inner node: closest_hit.glsl:167-199
            if (g.cnt) { g.cnt->inner_visits++; if (ANY) g.cnt->inner_visits_shadow++; }
            float le, re;
            leftHit = AABBIntersect(bvhTexel(o, leftIndex * 3 + 0), bvhTexel(o, leftIndex * 3 + 1), r_trans, &le);
            rightHit = AABBIntersect(bvhTexel(o, rightIndex * 3 + 0), bvhTexel(o, rightIndex * 3 + 1), r_trans, &re);
            bool lok = leftHit > 0.0f, rok = rightHit > 0.0f;
            if (cull) {   // not in the reference; result-preserving (see CULL_SLACK)
                float best = ANY ? limit0 : t;
                if (lok && le > best * CULL_SLACK) lok = false;
                if (rok && re > best * CULL_SLACK) rok = false;
            }
            if (lok && rok) {
                int deferred;
                if (leftHit > rightHit) { idx = rightIndex; deferred = leftIndex; }
                else { idx = leftIndex; deferred = rightIndex; }
                stack[ptr++] = deferred;
                continue;
            } else if (lok) { idx = leftIndex; continue; }
            else if (rok) { idx = rightIndex; continue; }
        }
        idx = stack[--ptr];
    }
    if (!ANY) state.hitDist = t;
    return t;
}

static float ClosestHit(Inv& g, const Ray& r, State& state, LightSampleRec& lrec) {   // closest_hit.glsl:7-205
    bool dummy = false;
    return Traverse<false>(g, r, state, lrec, 0.f, &dummy);
}
static bool AnyHit(Inv& g, const Ray& r, float maxDist) {   // anyhit.glsl:7-173
    bool hit = false;
    State s; LightSampleRec l;
    Traverse<true>(g, r, s, l, maxDist, &hit);
    return hit;
}

// ------------------------------------------------------------------------------------------------
// common/sampling.glsl
// ------------------------------------------------------------------------------------------------
static vec3 ImportanceSampleGTR1(float rgh, float r1, float r2) {   // sampling.glsl:7-21
    float a = gmax(0.001f, rgh);
    float a2 = a * a;
    float phi = r1 * TWO_PI;
    float cosTheta = sqrtf(fdiv(1.0f - lfom::pow(a2, 1.0f - r1), 1.0f - a2));
    float sinTheta = clampf(sqrtf(1.0f - (cosTheta * cosTheta)), 0.0f, 1.0f);
    float sinPhi, cosPhi;
    lfom::sincos(phi, sinPhi, cosPhi);
    (void)r2;
    return {sinTheta * cosPhi, sinTheta * sinPhi, cosTheta};
}
static vec3 ImportanceSampleGTR2(float rgh, float r1, float r2) {   // sampling.glsl:37-49
    float a = gmax(0.001f, rgh);
    float phi = r1 * TWO_PI;
    float cosTheta = sqrtf(fdiv(1.0f - r2, 1.0f + (a * a - 1.0f) * r2));
    float sinTheta = clampf(sqrtf(1.0f - (cosTheta * cosTheta)), 0.0f, 1.0f);
    float sinPhi, cosPhi;
    lfom::sincos(phi, sinPhi, cosPhi);
    return {sinTheta * cosPhi, sinTheta * sinPhi, cosTheta};
}
static float SchlickFresnel(float u) {   // sampling.glsl:52-58
    float m = clampf(1.0f - u, 0.0f, 1.0f);
    float m2 = m * m;
    return m2 * m2 * m;
}
static float DielectricFresnel(float cos_theta_i, float eta) {   // sampling.glsl:61-76
    float sinThetaTSq = eta * eta * (1.0f - cos_theta_i * cos_theta_i);
    if (sinThetaTSq > 1.0f) return 1.0f;
    float cos_theta_t = sqrtf(gmax(1.0f - sinThetaTSq, 0.0f));
    float rs = fdiv(eta * cos_theta_t - cos_theta_i, eta * cos_theta_t + cos_theta_i);
    float rp = fdiv(eta * cos_theta_i - cos_theta_t, eta * cos_theta_i + cos_theta_t);
    return 0.5f * (rs * rs + rp * rp);
}
static float GTR1(float NDotH, float a) {   // sampling.glsl:79-87
    if (a >= 1.0f) return (1.0f / PI);
    float a2 = a * a;
    float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
    // PI * log(a2): Mesa lowers log(x) to log2(x) * ln 2 and then folds the two constants of the product chain
    // (opt_algebraic's reassociate_constant): (PI * ln 2) * log2(a2).  Pinned with lp_probe on the reference's own GTR1.
    return fdiv(a2 - 1.0f, ((PI * 0.693147180559945309417f) * lfom::log2(a2)) * t);
}
static float GTR2(float NDotH, float a) {   // sampling.glsl:90-96
    float a2 = a * a;
    float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
    return fdiv(a2, PI * t * t);
}
static float SmithG_GGX(float NDotV, float alphaG) {   // sampling.glsl:110-116
    float a = alphaG * alphaG;
    float b = NDotV * NDotV;
    return 1.0f / (NDotV + sqrtf(a + b - a * b));
}
static vec3 CosineSampleHemisphere(float r1, float r2) {   // sampling.glsl:129-140
    vec3 dir;
    float r = sqrtf(r1);
    float phi = TWO_PI * r2;
    float sp, cp;
    lfom::sincos(phi, sp, cp);
    dir.x = r * cp;
    dir.y = r * sp;
    dir.z = sqrtf(gmax(0.0f, 1.0f - dir.x * dir.x - dir.y * dir.y));
    return dir;
}
static vec3 UniformSampleSphere(float r1, float r2) {   // sampling.glsl:153-160
    float z = 1.0f - 2.0f * r1;
    float r = sqrtf(gmax(0.0f, 1.0f - z * z));
    float phi = TWO_PI * r2;
    float sp, cp;
    lfom::sincos(phi, sp, cp);
    return {r * cp, r * sp, z};
}
static float powerHeuristic(float a, float b) {   // sampling.glsl:163-169
    float t = a * a;
    return fdiv(t, b * b + t);
}
static void sampleSphereLight(Inv& g, const Light& light, vec3 surfacePos, LightSampleRec& rec) {   // sampling.glsl:172-189
    float r1 = rnd(g), r2 = rnd(g);
    vec3 lightSurfacePos = light.position + UniformSampleSphere(r1, r2) * light.radius;
    rec.direction = lightSurfacePos - surfacePos;
    rec.dist = length(rec.direction);
    float distSq = rec.dist * rec.dist;
    rec.direction /= rec.dist;
    rec.normal = normalize(lightSurfacePos - light.position);
    rec.emission = light.emission * (float)g.o->scene.num_lights;
    rec.pdf = fdiv(distSq, light.area * fabsf(dot(rec.normal, rec.direction)));
}
static void sampleRectLight(Inv& g, const Light& light, vec3 surfacePos, LightSampleRec& rec) {   // sampling.glsl:192-206
    float r1 = rnd(g), r2 = rnd(g);
    vec3 lightSurfacePos = light.position + light.u * r1 + light.v * r2;
    rec.direction = lightSurfacePos - surfacePos;
    rec.dist = length(rec.direction);
    float distSq = rec.dist * rec.dist;
    rec.direction /= rec.dist;
    rec.normal = normalize(cross(light.u, light.v));
    rec.emission = light.emission * (float)g.o->scene.num_lights;
    rec.pdf = fdiv(distSq, light.area * fabsf(dot(rec.normal, rec.direction)));
}
static void sampleDistantLight(Inv& g, const Light& light, vec3 surfacePos, LightSampleRec& rec) {   // sampling.glsl:209-216
    rec.direction = normalize(light.position - V3(0.0f));
    rec.normal = normalize(surfacePos - light.position);
    rec.emission = light.emission * (float)g.o->scene.num_lights;
    rec.dist = INFINITY_;
    rec.pdf = 1.0f;
}
static void sampleOneLight(Inv& g, const Light& light, vec3 surfacePos, LightSampleRec& rec) {   // sampling.glsl:219-230
    int type = (int)light.type;
    if (type == 0) sampleRectLight(g, light, surfacePos, rec);
    else if (type == 1) sampleSphereLight(g, light, surfacePos, rec);
    else sampleDistantLight(g, light, surfacePos, rec);
}
static float EnvPdf(Inv& g, const Ray& r) {   // sampling.glsl:236-243
    float theta = lfom::acos(clampf(r.direction.y, -1.0f, 1.0f));
    vec2 uv = {(PI + lfom::atan2(r.direction.z, r.direction.x)) * (1.0f / TWO_PI), theta * (1.0f / PI)};
    float pdf = conditional(g.o, uv.x, uv.y).y * marginal(g.o, uv.y).y;
    float st, ct;
    lfom::sincos(theta, st, ct);
    return fdiv(pdf * (float)(g.o->scene.hdr_width * g.o->scene.hdr_height), 2.0f * PI * PI * st);
}
static vec4 EnvSample(Inv& g, vec3& color) {   // sampling.glsl:246-265
    float r1 = rnd(g), r2 = rnd(g);
    if (g.cnt) g.cnt->env_nee++;
    float v = marginal(g.o, r1).x;
    float u = conditional(g.o, r2, v).x;
    color = hdrLinear(g.o, u, v) * g.o->params.hdr_multiplier;
    float pdf = conditional(g.o, u, v).y * marginal(g.o, v).y;
    float phi = u * TWO_PI;
    float theta = v * PI;
    float st, ct, sph, cph;
    lfom::sincos(theta, st, ct);
    lfom::sincos(phi, sph, cph);
    if (st == 0.0f) pdf = 0.0f;
    float hdrResolution = (float)(g.o->scene.hdr_width * g.o->scene.hdr_height);
    return {-st * cph, ct, -st * sph, fdiv(pdf * hdrResolution, 2.0f * PI * PI * st)};
}
static vec3 EmitterSample(const State& state, const LightSampleRec& lrec, const BsdfSampleRec& brec) {   // sampling.glsl:271-282
    if (state.depth == 0) return lrec.emission;
    return powerHeuristic(brec.pdf, lrec.pdf) * lrec.emission;
}

// ------------------------------------------------------------------------------------------------
// common/disney.glsl
// ------------------------------------------------------------------------------------------------
static vec3 EvalDielectricReflection(const State& s, vec3 V, vec3 N, vec3 L, vec3 H, float& pdf) {   // disney.glsl:19-33
    pdf = 0.0f;
    if (dot(N, L) <= 0.0f) return V3(0.0f);
    float F = DielectricFresnel(dot(V, H), s.eta);
    float D = GTR2(dot(N, H), s.mat.roughness);
    pdf = fdiv(D * dot(N, H) * F, 4.0f * fabsf(dot(V, H)));
    float G = SmithG_GGX(fabsf(dot(N, L)), s.mat.roughness) * SmithG_GGX(fabsf(dot(N, V)), s.mat.roughness);
    return s.mat.albedo * F * D * G;
}
static vec3 EvalDielectricRefraction(const State& s, vec3 V, vec3 N, vec3 L, vec3 H, float& pdf) {   // disney.glsl:36-52
    pdf = 0.0f;
    if (dot(N, L) >= 0.0f) return V3(0.0f);
    float F = DielectricFresnel(fabsf(dot(V, H)), s.eta);
    float D = GTR2(dot(N, H), s.mat.roughness);
    float denomSqrt = dot(L, H) + dot(V, H) * s.eta;
    pdf = fdiv(D * dot(N, H) * (1.0f - F) * fabsf(dot(L, H)), denomSqrt * denomSqrt);
    float G = SmithG_GGX(fabsf(dot(N, L)), s.mat.roughness) * SmithG_GGX(fabsf(dot(N, V)), s.mat.roughness);
    return s.mat.albedo * (1.0f - F) * D * G * fabsf(dot(V, H)) * fabsf(dot(L, H)) * 4.0f * s.eta * s.eta / (denomSqrt * denomSqrt);
}
static vec3 EvalSpecular(const State& s, vec3 Cspec0, vec3 V, vec3 N, vec3 L, vec3 H, float& pdf) {   // disney.glsl:55-69
    pdf = 0.0f;
    if (dot(N, L) <= 0.0f) return V3(0.0f);
    float D = GTR2(dot(N, H), s.mat.roughness);
    pdf = fdiv(D * dot(N, H), 4.0f * dot(V, H));
    float FH = SchlickFresnel(dot(L, H));
    vec3 F = mix3(Cspec0, V3(1.0f), FH);
    float G = SmithG_GGX(fabsf(dot(N, L)), s.mat.roughness) * SmithG_GGX(fabsf(dot(N, V)), s.mat.roughness);
    return F * D * G;
}
static vec3 EvalClearcoat(const State& s, vec3 V, vec3 N, vec3 L, vec3 H, float& pdf) {   // disney.glsl:72-86
    pdf = 0.0f;
    if (dot(N, L) <= 0.0f) return V3(0.0f);
    float D = GTR1(dot(N, H), mixf(0.1f, 0.001f, s.mat.clearcoatRoughness));
    pdf = fdiv(D * dot(N, H), 4.0f * dot(V, H));
    float FH = SchlickFresnel(dot(L, H));
    float F = mixf(0.04f, 1.0f, FH);
    float G = SmithG_GGX(dot(N, L), 0.25f) * SmithG_GGX(dot(N, V), 0.25f);
    return V3(0.25f * s.mat.clearcoat * F * D * G);
}
static vec3 EvalDiffuse(const State& s, vec3 Csheen, vec3 V, vec3 N, vec3 L, vec3 H, float& pdf) {   // disney.glsl:89-113
    pdf = 0.0f;
    if (dot(N, L) <= 0.0f) return V3(0.0f);
    pdf = dot(N, L) * (1.0f / PI);
    float FL = SchlickFresnel(dot(N, L));
    float FV = SchlickFresnel(dot(N, V));
    float FH = SchlickFresnel(dot(L, H));
    float Fss90 = dot(L, H) * dot(L, H) * s.mat.roughness;
    float Fss = mixf(1.0f, Fss90, FL) * mixf(1.0f, Fss90, FV);
    float ss = 1.f * (Fss * (1.0f / (dot(N, L) + dot(N, V)) - 0.5f) + 0.5f);
    // FH * sheen * Csheen with FH = m2 * m2 * m inlined: Mesa's opt_rebalance_tree turns the scalar product chain
    // ((m2 * m2) * m) * sheen into (m2 * m2) * (m * sheen).  Pinned with the reference's own EvalDiffuse on llvmpipe (llvmpipe_bsdf.npz).
    float mH = clampf(1.0f - dot(L, H), 0.0f, 1.0f);
    float mH2 = mH * mH;
    vec3 Fsheen = ((mH2 * mH2) * (mH * s.mat.sheen)) * Csheen;
    return ((1.0f / PI) * (ss + s.mat.subsurface) * s.mat.albedo + Fsheen) * (1.0f - s.mat.metallic);   // :112 (fork's variant)
}

static void disneyTints(const State& s, vec3& Cspec0, vec3& Csheen) {   // disney.glsl:140-145 == :266-273
    vec3 Cdlin = s.mat.albedo;
    float Cdlum = 0.3f * Cdlin.x + 0.6f * Cdlin.y + 0.1f * Cdlin.z;
    vec3 Ctint = Cdlum > 0.0f ? Cdlin / Cdlum : V3(1.0f);
    Cspec0 = mix3(s.mat.specular * 0.08f * mix3(V3(1.0f), Ctint, s.mat.specularTint), Cdlin, s.mat.metallic);
    Csheen = mix3(V3(1.0f), Ctint, s.mat.sheenTint);
}

static vec3 DisneySample(Inv& g, State& state, vec3 V, vec3 N, vec3& L, float& pdf) {   // disney.glsl:128-225
    pdf = 0.0f;
    vec3 f = V3(0.0f);
    float r1 = rnd(g), r2 = rnd(g);
    float diffuseRatio = 0.5f * (1.0f - state.mat.metallic);
    float transWeight = (1.0f - state.mat.metallic) * state.mat.specTrans;
    vec3 Cspec0, Csheen;
    disneyTints(state, Cspec0, Csheen);

    if (rnd(g) < transWeight) {
        vec3 H = ImportanceSampleGTR2(state.mat.roughness, r1, r2);
        H = state.tangent * H.x + state.bitangent * H.y + N * H.z;
        if (dot(V, H) < 0.0f) H = -H;
        vec3 R = reflect(-V, H);
        float F = DielectricFresnel(fabsf(dot(R, H)), state.eta);
        if (rnd(g) < F) {
            L = normalize(R);
            f = EvalDielectricReflection(state, V, N, L, H, pdf);
        } else {
            L = normalize(refract(-V, H, state.eta));
            f = EvalDielectricRefraction(state, V, N, L, H, pdf);
        }
        f *= transWeight;
        pdf *= transWeight;
    } else {
        if (rnd(g) < diffuseRatio) {
            L = CosineSampleHemisphere(r1, r2);
            L = state.tangent * L.x + state.bitangent * L.y + N * L.z;
            vec3 H = normalize(L + V);
            f = EvalDiffuse(state, Csheen, V, N, L, H, pdf);
            pdf *= diffuseRatio;
        } else {
            float primarySpecRatio = 1.0f / (1.0f + state.mat.clearcoat);
            if (rnd(g) < primarySpecRatio) {
                vec3 H = ImportanceSampleGTR2(state.mat.roughness, r1, r2);
                H = state.tangent * H.x + state.bitangent * H.y + N * H.z;
                if (dot(V, H) < 0.0f) H = -H;
                L = normalize(reflect(-V, H));
                f = EvalSpecular(state, Cspec0, V, N, L, H, pdf);
                pdf *= primarySpecRatio * (1.0f - diffuseRatio);
            } else {
                vec3 H = ImportanceSampleGTR1(mixf(0.1f, 0.001f, state.mat.clearcoatRoughness), r1, r2);
                H = state.tangent * H.x + state.bitangent * H.y + N * H.z;
                if (dot(V, H) < 0.0f) H = -H;
                L = normalize(reflect(-V, H));
                f = EvalClearcoat(state, V, N, L, H, pdf);
                pdf *= (1.0f - primarySpecRatio) * (1.0f - diffuseRatio);
            }
        }
        f *= (1.0f - transWeight);
        pdf *= (1.0f - transWeight);
    }
    return f;
}

static vec3 DisneyEval(const State& state, vec3 V, vec3 N, vec3 L, float& pdf) {   // disney.glsl:228-291
    vec3 H;
    bool refl = dot(N, L) > 0.0f;
    if (refl) H = normalize(L + V);
    else H = normalize(L + V * state.eta);
    if (dot(V, H) < 0.0f) H = -H;

    float diffuseRatio = 0.5f * (1.0f - state.mat.metallic);
    float primarySpecRatio = 1.0f / (1.0f + state.mat.clearcoat);
    float transWeight = (1.0f - state.mat.metallic) * state.mat.specTrans;
    vec3 brdf = V3(0.0f), bsdf = V3(0.0f);
    float brdfPdf = 0.0f, bsdfPdf = 0.0f;

    if (transWeight > 0.0f) {
        if (refl) bsdf = EvalDielectricReflection(state, V, N, L, H, bsdfPdf);
        else bsdf = EvalDielectricRefraction(state, V, N, L, H, bsdfPdf);
    }
    float m_pdf;
    if (transWeight < 1.0f) {
        vec3 Cspec0, Csheen;
        disneyTints(state, Cspec0, Csheen);
        brdf += EvalDiffuse(state, Csheen, V, N, L, H, m_pdf);
        brdfPdf += m_pdf * diffuseRatio;
        brdf += EvalSpecular(state, Cspec0, V, N, L, H, m_pdf);
        brdfPdf += m_pdf * primarySpecRatio * (1.0f - diffuseRatio);
        brdf += EvalClearcoat(state, V, N, L, H, m_pdf);
        brdfPdf += m_pdf * (1.0f - primarySpecRatio) * (1.0f - diffuseRatio);
    }
    pdf = mixf(brdfPdf, bsdfPdf, transWeight);
    return mix3(brdf, bsdf, transWeight);
}

// ------------------------------------------------------------------------------------------------
// common/pathtrace.glsl
// ------------------------------------------------------------------------------------------------
static void Onb(vec3 N, vec3& T, vec3& B) {   // pathtrace.glsl:7-13
    vec3 UpVector = fabsf(N.z) < 0.999f ? V3(0, 0, 1) : V3(1, 0, 0);
    T = normalize(cross(UpVector, N));
    B = cross(N, T);
}

static void GetNormalsAndTexCoord(Inv& g, State& state, const Ray& r) {   // pathtrace.glsl:16-37
    vec4 n1 = nrm(g.o, state.triID[0]), n2 = nrm(g.o, state.triID[1]), n3 = nrm(g.o, state.triID[2]);
    vec2 t1 = {g.tempTexCoords.x, n1.w}, t2 = {g.tempTexCoords.y, n2.w}, t3 = {g.tempTexCoords.z, n3.w};
    state.texCoord = t1 * state.bary.x + t2 * state.bary.y + t3 * state.bary.z;
    vec3 normal = normalize(vec3{n1.x, n1.y, n1.z} * state.bary.x + vec3{n2.x, n2.y, n2.z} * state.bary.y + vec3{n3.x, n3.y, n3.z} * state.bary.z);
    mat3 m3 = {{{g.transform.c[0].x, g.transform.c[0].y, g.transform.c[0].z},
                {g.transform.c[1].x, g.transform.c[1].y, g.transform.c[1].z},
                {g.transform.c[2].x, g.transform.c[2].y, g.transform.c[2].z}}};
    mat3 normalMatrix = transpose(inverse(m3));
    normal = normalize(mul(normalMatrix, normal));
    state.normal = normal;
    state.ffnormal = dot(normal, r.direction) <= 0.0f ? normal : normal * -1.0f;
    Onb(state.normal, state.tangent, state.bitangent);
}

static void GetMaterialsAndTextures(Inv& g, State& state, const Ray& r) {   // pathtrace.glsl:40-123
    const float* p = g.o->scene.materials + 28 * (size_t)state.matID;
    Material mat;
    mat.albedo = {p[0], p[1], p[2]}; mat.specular = p[3];
    mat.emission = {p[4], p[5], p[6]}; mat.anisotropic = p[7];
    mat.metallic = p[8]; mat.roughness = gmax(p[9], 0.001f);
    mat.subsurface = p[10]; mat.specularTint = p[11];
    mat.sheen = p[12]; mat.sheenTint = p[13]; mat.clearcoat = p[14]; mat.clearcoatRoughness = p[15];
    mat.specTrans = p[16]; mat.ior = p[17]; mat.atDistance = p[18];
    mat.extinction = {p[20], p[21], p[22]};
    mat.texIDs = {p[24], p[25], p[26], p[27]};
    if (g.cnt) g.cnt->shaded_hits++;

    vec2 texUV = state.texCoord;
    texUV.y = 1.0f - texUV.y;

    if ((int)mat.texIDs.x >= 0) {   // :83-84
        vec4 c = texArrayLinear(g.o, texUV.x, texUV.y, (int)mat.texIDs.x, g.cnt);
        mat.albedo *= pow3(vec3{c.x, c.y, c.z}, 2.2f);
    }
    if ((int)mat.texIDs.y >= 0) {   // :87-93  (float layer: GL rounds to nearest)
        vec4 c = texArrayLinear(g.o, texUV.x, texUV.y, (int)floorf(mat.texIDs.y + 0.5f), g.cnt);
        mat.metallic = c.x;
        mat.roughness = gmax(c.y * c.y, 0.001f);
    }
    if ((int)mat.texIDs.z >= 0) {   // :96-109
        vec4 c = texArrayLinear(g.o, texUV.x, texUV.y, (int)mat.texIDs.z, g.cnt);
        vec3 n = normalize(vec3{c.x, c.y, c.z} * 2.0f - V3(1.0f));
        vec3 T, B;
        Onb(state.normal, T, B);
        n = T * n.x + B * n.y + state.normal * n.z;
        state.normal = normalize(n);
        state.ffnormal = dot(state.normal, r.direction) <= 0.0f ? state.normal : state.normal * -1.0f;
        Onb(state.normal, state.tangent, state.bitangent);
    }
    if (mat.texIDs.w >= 0) {   // :112-113
        vec4 c = texArrayLinear(g.o, texUV.x, texUV.y, (int)floorf(mat.texIDs.w + 0.5f), g.cnt);
        mat.emission = pow3(vec3{c.x, c.y, c.z}, 2.2f);
    }
    state.mat = mat;
    state.eta = dot(r.direction, state.normal) < 0.0f ? (1.0f / mat.ior) : mat.ior;   // :122
}

static vec3 DirectLight(Inv& g, const Ray& r, const State& state) {   // pathtrace.glsl:126-204
    const Oracle* o = g.o;
    vec3 Li = V3(0.0f);
    vec3 surfacePos = state.fhp + state.normal * EPS;
    BsdfSampleRec bsdfSampleRec;
    std::memset(&bsdfSampleRec, 0, sizeof bsdfSampleRec);

    if (o->params.use_envmap && !o->params.use_constant_bg) {   // #ifdef ENVMAP / #ifndef CONSTANT_BG, :134-160
        vec3 color;
        vec4 dirPdf = EnvSample(g, color);
        vec3 lightDir = {dirPdf.x, dirPdf.y, dirPdf.z};
        float lightPdf = dirPdf.w;
        Ray shadowRay = {surfacePos, lightDir};
        bool inShadow = AnyHit(g, shadowRay, INFINITY_ - EPS);
        if (!inShadow) {
            bsdfSampleRec.f = DisneyEval(state, -r.direction, state.ffnormal, lightDir, bsdfSampleRec.pdf);
            if (bsdfSampleRec.pdf > 0.0f) {
                float misWeight = powerHeuristic(lightPdf, bsdfSampleRec.pdf);
                if (misWeight > 0.0f)
                    Li += misWeight * bsdfSampleRec.f * fabsf(dot(lightDir, state.ffnormal)) * color / lightPdf;
            }
        }
    }
    if (o->scene.num_lights > 0) {   // #ifdef LIGHTS, :163-201
        LightSampleRec lightSampleRec;
        std::memset(&lightSampleRec, 0, sizeof lightSampleRec);
        int index = (int)(rnd(g) * (float)o->scene.num_lights);
        Light light = fetchLight(o, index);
        sampleOneLight(g, light, surfacePos, lightSampleRec);
        if (dot(lightSampleRec.direction, lightSampleRec.normal) < 0.0f) {
            Ray shadowRay = {surfacePos, lightSampleRec.direction};
            bool inShadow = AnyHit(g, shadowRay, lightSampleRec.dist - EPS);
            if (!inShadow) {
                bsdfSampleRec.f = DisneyEval(state, -r.direction, state.ffnormal, lightSampleRec.direction, bsdfSampleRec.pdf);
                float weight = 1.0f;
                if (light.area > 0.0f) weight = powerHeuristic(lightSampleRec.pdf, bsdfSampleRec.pdf);
                if (bsdfSampleRec.pdf > 0.0f)
                    Li += weight * bsdfSampleRec.f * fabsf(dot(state.ffnormal, lightSampleRec.direction)) * lightSampleRec.emission / lightSampleRec.pdf;
            }
        }
    }
    return Li;
}

static vec3 PathTrace(Inv& g, Ray r) {   // pathtrace.glsl:208-295
    const Oracle* o = g.o;
    vec3 radiance = V3(0.0f);
    vec3 throughput = V3(1.0f);
    State state;                 // declared outside the loop and never reset (:213-215); GLSL leaves it undefined,
    LightSampleRec lightSampleRec;   // llvmpipe zero-fills: zero-init
    BsdfSampleRec bsdfSampleRec;
    std::memset(&state, 0, sizeof state);
    std::memset(&lightSampleRec, 0, sizeof lightSampleRec);
    std::memset(&bsdfSampleRec, 0, sizeof bsdfSampleRec);
    vec3 absorption = V3(0.0f);

    for (int depth = 0; depth < o->params.max_depth; depth++) {
        state.depth = depth;
        float t = ClosestHit(g, r, state, lightSampleRec);

        if (t == INFINITY_) {
            if (o->params.use_constant_bg) {
                radiance += vec3{o->params.bg_color[0], o->params.bg_color[1], o->params.bg_color[2]} * throughput;
            } else if (o->params.use_envmap) {
                float misWeight = 1.0f;
                vec2 uv = {(PI + lfom::atan2(r.direction.z, r.direction.x)) * (1.0f / TWO_PI), lfom::acos(r.direction.y) * (1.0f / PI)};
                if (depth > 0) {
                    float lightPdf = EnvPdf(g, r);
                    misWeight = powerHeuristic(bsdfSampleRec.pdf, lightPdf);
                }
                if (g.cnt) g.cnt->env_miss++;
                radiance += misWeight * hdrLinear(o, uv.x, uv.y) * throughput * o->params.hdr_multiplier;
            }
            return radiance;
        }

        GetNormalsAndTexCoord(g, state, r);
        GetMaterialsAndTextures(g, state, r);

        if (dot(state.normal, state.ffnormal) > 0.0f) absorption = V3(0.0f);   // :250-251

        radiance += state.mat.emission * throughput;   // :253

        if (o->scene.num_lights > 0 && state.isEmitter) {   // #ifdef LIGHTS :255-261
            radiance += EmitterSample(state, lightSampleRec, bsdfSampleRec) * throughput;
            break;
        }

        throughput *= exp3(-absorption * t);   // :264

        radiance += DirectLight(g, r, state) * throughput;   // :266

        bsdfSampleRec.f = DisneySample(g, state, -r.direction, state.ffnormal, bsdfSampleRec.L, bsdfSampleRec.pdf);   // :268

        if (dot(state.ffnormal, bsdfSampleRec.L) < 0.0f)   // :271-272
            absorption = -log3(state.mat.extinction) / state.mat.atDistance;

        if (bsdfSampleRec.pdf > 0.0f)   // :274-277
            throughput *= bsdfSampleRec.f * fabsf(dot(state.ffnormal, bsdfSampleRec.L)) / bsdfSampleRec.pdf;
        else
            break;

        if (o->params.enable_rr && depth >= o->params.rr_depth) {   // #ifdef RR :279-288
            float q = gmin(gmax(throughput.x, gmax(throughput.y, throughput.z)) + 0.001f, 0.95f);
            if (rnd(g) > q) break;
            throughput /= q;
        }

        r.direction = bsdfSampleRec.L;   // :290-291
        r.origin = state.fhp + r.direction * EPS;
    }
    return radiance;
}

// ------------------------------------------------------------------------------------------------
// renderer.glsl
// ------------------------------------------------------------------------------------------------
static inline float mapf(float value, float low1, float high1, float low2, float high2) {   // renderer.glsl:20-23
    return low2 + fdiv((value - low1) * (high2 - low2), high1 - low1);
}

// renderer.glsl:27-62: pixel mapping, RNG init, jitter, thin-lens camera ray.  Returns the full-frame pixel.
static Ray CameraRay(Inv& g, int lx, int ly, int tileX, int tileY, int frame, int* px, int* py) {
    const Oracle* o = g.o;
    const LfParams& P = o->params;
    const LfCamera& C = o->camera;
    vec2 screenResolution = {(float)P.width, (float)P.height};
    float invNumTilesX = 1.0f / ((float)P.width / P.tile_width);     // TiledRenderer.cpp:226-227
    float invNumTilesY = 1.0f / ((float)P.height / P.tile_height);
    // the fullscreen quad's TexCoords at the fragment centre of a tileWidth x tileHeight viewport
    vec2 TexCoords = {((float)lx + 0.5f) / (float)P.tile_width, ((float)ly + 0.5f) / (float)P.tile_height};
    vec2 coordsTile = TexCoords, coordsFS;
    float xoffset = -1.0f + 2.0f * invNumTilesX * (float)tileX;
    float yoffset = -1.0f + 2.0f * invNumTilesY * (float)tileY;
    coordsTile.x = mapf(coordsTile.x, 0.0f, 1.0f, xoffset, xoffset + 2.0f * invNumTilesX);
    coordsTile.y = mapf(coordsTile.y, 0.0f, 1.0f, yoffset, yoffset + 2.0f * invNumTilesY);
    coordsFS.x = mapf(TexCoords.x, 0.0f, 1.0f, invNumTilesX * (float)tileX, invNumTilesX * (float)tileX + invNumTilesX);
    coordsFS.y = mapf(TexCoords.y, 0.0f, 1.0f, invNumTilesY * (float)tileY, invNumTilesY * (float)tileY + invNumTilesY);

    vec2 p = {coordsFS.x * screenResolution.x, coordsFS.y * screenResolution.y};
    InitRNG(g, p, frame);
    // the tile is copied into the accumulation texture at viewport offset (tileW*tileX, tileH*tileY) (TiledRenderer.cpp:342)
    *px = P.tile_width * tileX + lx;
    *py = P.tile_height * tileY + ly;

    float r1 = 2.0f * rnd(g);
    float r2 = 2.0f * rnd(g);
    vec2 jitter;
    jitter.x = r1 < 1.0f ? sqrtf(r1) - 1.0f : 1.0f - sqrtf(2.0f - r1);
    jitter.y = r2 < 1.0f ? sqrtf(r2) - 1.0f : 1.0f - sqrtf(2.0f - r2);
    jitter.x = fdiv(jitter.x, screenResolution.x * 0.5f);
    jitter.y = fdiv(jitter.y, screenResolution.y * 0.5f);
    vec2 d = coordsTile + jitter;

    float scale = lfom::tan(C.fov * 0.5f);
    d.y *= fdiv(screenResolution.y, screenResolution.x) * scale;
    d.x *= scale;
    vec3 right = {C.right[0], C.right[1], C.right[2]}, up = {C.up[0], C.up[1], C.up[2]}, fwd = {C.forward[0], C.forward[1], C.forward[2]};
    vec3 pos = {C.position[0], C.position[1], C.position[2]};
    vec3 rayDir = normalize(d.x * right + d.y * up + fwd);
    vec3 focalPoint = C.focal_dist * rayDir;
    float cam_r1 = rnd(g) * TWO_PI;
    float cam_r2 = rnd(g) * C.aperture;
    float sl, cl;
    lfom::sincos(cam_r1, sl, cl);
    vec3 randomAperturePos = (cl * right + sl * up) * sqrtf(cam_r2);
    vec3 finalRayDir = normalize(focalPoint - randomAperturePos);
    return {pos + randomAperturePos, finalRayDir};
}

// preview_flareon.glsl:22-61: the camera ray of fragment (x, y) of the pvW x pvH preview viewport.  Differences from
// renderer.glsl: RNG seeded with gl_FragCoord.xy and frame 1 (:24), jitter divided by the full screenResolution (:33),
// d = 2 * TexCoords - 1 (no tile mapping, :34), and the thin lens only under #define USE_DOF (:44-57) although its two
// rand() are always drawn (:42-43).
static Ray PreviewRay(Inv& g, int x, int y, int pvW, int pvH, bool useDof) {
    const Oracle* o = g.o;
    const LfParams& P = o->params;
    const LfCamera& C = o->camera;
    vec2 screenResolution = {(float)P.width, (float)P.height};
    vec2 TexCoords = {((float)x + 0.5f) / (float)pvW, ((float)y + 0.5f) / (float)pvH};
    InitRNG(g, vec2{(float)x + 0.5f, (float)y + 0.5f}, 1);
    float r1 = 2.0f * rnd(g);
    float r2 = 2.0f * rnd(g);
    vec2 jitter;
    jitter.x = r1 < 1.0f ? sqrtf(r1) - 1.0f : 1.0f - sqrtf(2.0f - r1);
    jitter.y = r2 < 1.0f ? sqrtf(r2) - 1.0f : 1.0f - sqrtf(2.0f - r2);
    jitter.x = fdiv(jitter.x, screenResolution.x);
    jitter.y = fdiv(jitter.y, screenResolution.y);
    vec2 d = {(2.0f * TexCoords.x - 1.0f) + jitter.x, (2.0f * TexCoords.y - 1.0f) + jitter.y};
    float scale = lfom::tan(C.fov * 0.5f);
    d.y *= fdiv(screenResolution.y, screenResolution.x) * scale;
    d.x *= scale;
    vec3 right = {C.right[0], C.right[1], C.right[2]}, up = {C.up[0], C.up[1], C.up[2]}, fwd = {C.forward[0], C.forward[1], C.forward[2]};
    vec3 pos = {C.position[0], C.position[1], C.position[2]};
    vec3 rayDir = normalize(d.x * right + d.y * up + fwd);
    vec3 focalPoint = C.focal_dist * rayDir;
    float cam_r1 = rnd(g) * TWO_PI;
    float cam_r2 = rnd(g) * C.aperture;
    if (!useDof) return {pos, normalize(focalPoint)};
    float sl, cl;
    lfom::sincos(cam_r1, sl, cl);
    vec3 randomAperturePos = (cl * right + sl * up) * sqrtf(cam_r2);
    return {pos + randomAperturePos, normalize(focalPoint - randomAperturePos)};
}

// ------------------------------------------------------------------------------------------------
// Oracle
// ------------------------------------------------------------------------------------------------
Oracle::Oracle(const LfSceneView& s, const LfParams& p, const LfCamera& c) : scene(s), params(p), camera(c) {
    std::memset(&counters, 0, sizeof counters);
}

void Oracle::Sample(int lx, int ly, int tileX, int tileY, int frame, float rgb[3], int* px, int* py, LfCounters* cnt) const {
    Inv g;
    std::memset(&g, 0, sizeof g);
    g.o = this;
    g.cnt = cnt;
    Ray ray = CameraRay(g, lx, ly, tileX, tileY, frame, px, py);
    if (cnt) cnt->samples++;
    vec3 c = PathTrace(g, ray);
    rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
}

static void addCounters(LfCounters& a, const LfCounters& b) {
    uint64_t* pa = reinterpret_cast<uint64_t*>(&a);
    const uint64_t* pb = reinterpret_cast<const uint64_t*>(&b);
    for (size_t i = 0; i < sizeof(LfCounters) / sizeof(uint64_t); i++) pa[i] += pb[i];
}

void Oracle::RenderFrames(int firstFrame, int nframes, int frameStride, int tileX, int tileY, float* accum) {
    const int W = params.width, H = params.height, TW = params.tile_width, TH = params.tile_height;
    LfCounters total;
    std::memset(&total, 0, sizeof total);
#pragma omp parallel
    {
        LfCounters local;
        std::memset(&local, 0, sizeof local);
#pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < TW * TH; i++) {
            int lx = i % TW, ly = i / TW;
            for (int k = 0; k < nframes; k++) {   // frame order per pixel == the reference's sequential sum (renderer.glsl:68)
                float rgb[3];
                int px, py;
                Sample(lx, ly, tileX, tileY, firstFrame + k * frameStride, rgb, &px, &py, count ? &local : nullptr);
                if (px < 0 || py < 0 || px >= W || py >= H) continue;
                float* a = accum + 3 * ((size_t)py * W + px);
                a[0] = rgb[0] + a[0]; a[1] = rgb[1] + a[1]; a[2] = rgb[2] + a[2];
            }
        }
#pragma omp critical
        addCounters(total, local);
    }
    addCounters(counters, total);
}

// One draw of previewEngineShader into the pvW x pvH preview target (TiledRenderer.cpp:327-333) with uniform maxDepth
// (2 while the camera moves, :532).  out = pvW * pvH * 3 floats, rows bottom-up; nothing is accumulated.
void Oracle::RenderPreview(int pvW, int pvH, int maxDepth, bool useDof, float* out) {
    LfParams saved = params;
    params.max_depth = maxDepth;
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < pvW * pvH; i++) {
        Inv g;
        std::memset(&g, 0, sizeof g);
        g.o = this;
        Ray ray = PreviewRay(g, i % pvW, i / pvW, pvW, pvH, useDof);
        vec3 c = PathTrace(g, ray);
        out[3 * (size_t)i] = c.x; out[3 * (size_t)i + 1] = c.y; out[3 * (size_t)i + 2] = c.z;
    }
    params = saved;
}

void Oracle::PrimaryHits(int frame, float* t, int32_t* triX, int32_t* matID, int32_t* emitter) {
    const int W = params.width, H = params.height;
    LfParams saved = params;
    params.tile_width = W; params.tile_height = H;   // single tile covering the frame
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < W * H; i++) {
        Inv g;
        std::memset(&g, 0, sizeof g);
        g.o = this;
        int px, py;
        Ray ray = CameraRay(g, i % W, i / W, 0, 0, frame, &px, &py);
        State state; LightSampleRec lrec;
        std::memset(&state, 0, sizeof state);
        std::memset(&lrec, 0, sizeof lrec);
        state.triID[0] = -1; state.matID = -1;
        t[i] = ClosestHit(g, ray, state, lrec);
        triX[i] = state.triID[0]; matID[i] = state.matID; emitter[i] = state.isEmitter ? 1 : 0;
    }
    params = saved;
}

void Oracle::RandKat(int px, int py, int frame, int n, uint32_t* seedx, float* values) {
    Inv g;
    std::memset(&g, 0, sizeof g);
    InitRNG(g, vec2{(float)px + 0.5f, (float)py + 0.5f}, frame);
    for (int i = 0; i < n; i++) { values[i] = rnd(g); seedx[i] = g.seed[0]; }
}

// ------------------------------------------------------------------------------------------------
// postprocess.glsl
// ------------------------------------------------------------------------------------------------
static inline float tmACES(float c) { return clampf(fdiv(c * (2.51f * c + 0.03f), c * (2.43f * c + 0.59f) + 0.14f), 0.0f, 1.0f); }   // postprocess.glsl:34-43
static inline float tmKanjero(float c) {   // :69-84 (the alpha channel skips the second pow; it is not part of the RGB output)
    float v = lfom::pow(fdiv(c * (c * (1.2295f * c + 0.3135f) + 1.1935f * 0.4655f), c * (1.1935f * c + 0.4655f) + 0.073f), 1.7f);
    v = lfom::pow(v, 1.0f / 0.8f);
    v *= 0.8f;
    return clampf(v, 0.0f, 1.0f);
}
static inline float tmHejl(float c) { c = gmax(0.0f, c - 0.004f); return fdiv(c * (6.2f * c + .5f), c * (6.2f * c + 1.7f) + 0.06f); }   // :52-56
static inline float tmUncharted(float c) {   // :87-96
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return fdiv(c * (A * c + C * B) + D * E, c * (A * c + B) + D * F) - fdiv(E, F);
}
static inline int mirrori(int i, int n) { int m = i % (2 * n); if (m < 0) m += 2 * n; return m < n ? m : 2 * n - 1 - m; }
// texture(pathTraceTexture, uv).ch: RGB32F, LINEAR, MIRRORED_REPEAT (TiledRenderer.cpp:165-173)
static float accumLinear(const float* accum, int W, int H, float u, float v, int ch) {
    float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float wx = x - fx, wy = y - fy;
    int x0 = mirrori((int)fx, W), x1 = mirrori((int)fx + 1, W), y0 = mirrori((int)fy, H), y1 = mirrori((int)fy + 1, H);
    float a = accum[3 * ((size_t)y0 * W + x0) + ch], b = accum[3 * ((size_t)y0 * W + x1) + ch];
    float c = accum[3 * ((size_t)y1 * W + x0) + ch], e = accum[3 * ((size_t)y1 * W + x1) + ch];
    float top = a + (b - a) * wx, bot = c + (e - c) * wx;
    return top + (bot - top) * wy;
}
void Oracle::PostProcess(const float* accum, int W, int H, float inv, int tonemapIndex, const LfPostParams& pp, float* out) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < W * H; i++) {
        const int px = i % W, py = i / W;
        const float tu = ((float)px + 0.5f) / (float)W, tv = ((float)py + 0.5f) / (float)H;   // TexCoords at the fragment centre
        float c[3] = {accum[3 * (size_t)i] * inv, accum[3 * (size_t)i + 1] * inv, accum[3 * (size_t)i + 2] * inv};   // :129-133
        if (pp.use_ca) {   // chromaticAberration(), :98-118
            float offset = pp.ca_distance;
            float dx = tu - pp.ca_p3, dy = tv - pp.ca_p3;
            float dist = 0.f + (lfom::pow(sqrtf(dx * dx + dy * dy), pp.ca_p1) * pp.ca_p2);
            float o = pp.use_ca_distortion ? offset * dist : (offset * 0.025f) * pp.ca_p2;
            c[0] = accumLinear(accum, W, H, tu + o, tv + o, 0) * inv;
            c[2] = accumLinear(accum, W, H, tu - o, tv - o, 2) * inv;
        }
        const float g = 1.0f / 2.2f;
        switch (tonemapIndex) {   // :135-165
        case 1: {   // pow(tonemap(color, 2), 1 / 2.2), tonemap = c * 1.0 / (1.0 + luminance / limit) (:26-31)
            float lum = (0.3f * c[0] + 0.6f * c[1]) + 0.1f * c[2];
            float r = 1.0f / (1.0f + fdiv(lum, 2.f));
            for (int k = 0; k < 3; k++) c[k] = lfom::pow((c[k] * 1.0f) * r, g);
            break;
        }
        case 2: for (int k = 0; k < 3; k++) c[k] = lfom::pow(tmACES(c[k]), g); break;
        case 3: for (int k = 0; k < 3; k++) c[k] = lfom::pow(clampf(fdiv(c[k], c[k] + 1.f), 0.0f, 1.0f), g); break;   // Reinhard, :46-49
        case 4: for (int k = 0; k < 3; k++) c[k] = lfom::pow(tmKanjero(c[k]), g); break;
        case 5: for (int k = 0; k < 3; k++) c[k] = tmHejl(c[k]); break;
        case 6: for (int k = 0; k < 3; k++) c[k] = lfom::pow(tmUncharted(c[k]), g) * 1.75f; break;
        default: break;
        }
        if (pp.use_vignette) {   // vignette(), :121-124
            float dx = tu - 0.5f, dy = tv - 0.5f;
            float d = 1.0f - lfom::pow(sqrtf(dx * dx + dy * dy), pp.vignette_power) * pp.vignette_intensity;
            for (int k = 0; k < 3; k++) c[k] *= d;
        }
        out[3 * (size_t)i] = c[0]; out[3 * (size_t)i + 1] = c[1]; out[3 * (size_t)i + 2] = c[2];
    }
}

// The reference's BSDF functions on explicit arguments (tests/golden/make_bsdf_golden.py runs the same calls on llvmpipe with the
// reference's GLSL text).  Item layout, 9 vec4: [0] V | [1] N (= normal = ffnormal) | [2] L | [3] albedo, specular |
// [4] metallic, roughness, subsurface, specularTint | [5] sheen, sheenTint, clearcoat, clearcoatRoughness |
// [6] specTrans, eta, seed.z, seed.w | [7] tangent, seed.x | [8] bitangent, seed.y   (seeds are small integers stored as floats)
void Oracle::BsdfKat(int op, const float* in, int n, float* out4) {
    for (int i = 0; i < n; i++) {
        const float* a = in + 36 * (size_t)i;
        float* r = out4 + 4 * (size_t)i;
        State s;
        std::memset(&s, 0, sizeof s);
        vec3 V = {a[0], a[1], a[2]}, N = {a[4], a[5], a[6]}, L = {a[8], a[9], a[10]};
        s.mat.albedo = {a[12], a[13], a[14]}; s.mat.specular = a[15];
        s.mat.metallic = a[16]; s.mat.roughness = a[17]; s.mat.subsurface = a[18]; s.mat.specularTint = a[19];
        s.mat.sheen = a[20]; s.mat.sheenTint = a[21]; s.mat.clearcoat = a[22]; s.mat.clearcoatRoughness = a[23];
        s.mat.specTrans = a[24]; s.eta = a[25];
        s.normal = N; s.ffnormal = N;
        s.tangent = {a[28], a[29], a[30]}; s.bitangent = {a[32], a[33], a[34]};
        Inv g;
        std::memset(&g, 0, sizeof g);
        g.seed[0] = (uint32_t)a[31]; g.seed[1] = (uint32_t)a[35]; g.seed[2] = (uint32_t)a[26]; g.seed[3] = (uint32_t)a[27];
        float pdf = 0.0f;
        switch (op) {
        case 0: { vec3 f = DisneyEval(s, V, N, L, pdf); r[0] = f.x; r[1] = f.y; r[2] = f.z; r[3] = pdf; break; }
        case 1: { vec3 Ls = V3(0.0f); DisneySample(g, s, V, N, Ls, pdf); r[0] = Ls.x; r[1] = Ls.y; r[2] = Ls.z; r[3] = pdf; break; }
        case 2: { vec3 Ls = V3(0.0f); vec3 f = DisneySample(g, s, V, N, Ls, pdf); r[0] = f.x; r[1] = f.y; r[2] = f.z; r[3] = rnd(g); break; }
        case 3: r[0] = GTR1(a[0], a[1]); r[1] = GTR2(a[0], a[1]); r[2] = SmithG_GGX(a[0], a[1]); r[3] = DielectricFresnel(a[0], a[25]); break;
        case 4: {
            vec3 h1 = ImportanceSampleGTR1(a[17], a[0], a[1]), h2 = ImportanceSampleGTR2(a[17], a[0], a[1]), c = CosineSampleHemisphere(a[0], a[1]);
            r[0] = h1.x + h1.z; r[1] = h2.x + h2.z; r[2] = c.x + c.z; r[3] = h1.y + h2.y + c.y;
            break;
        }
        default: r[0] = r[1] = r[2] = r[3] = 0.0f; break;
        }
    }
}

// The expression groups of tests/golden/make_builtin_golden.py, evaluated with this file's restatement of the GLSL built-ins.
void Oracle::BuiltinKat(int op, const float* in4, int n, float* out4, const uint8_t* tex, int texW, int texH, int texL) {
    for (int i = 0; i < n; i++) {
        const float* a = in4 + 4 * (size_t)i;
        float* r = out4 + 4 * (size_t)i;
        r[0] = r[1] = r[2] = r[3] = 0.0f;
        switch (op) {
        case 0: { float s, c; lfom::sincos(a[0], s, c); r[0] = s; r[1] = c; r[2] = lfom::tan(a[1]); r[3] = sqrtf(fabsf(a[2])); break; }   // vec4(sin(a.x), cos(a.x), tan(a.y), sqrt(abs(a.z)))
        case 1: r[0] = lfom::exp(a[0]); r[1] = lfom::log(a[1]); r[2] = lfom::pow(a[1], a[2]); r[3] = lfom::acos(a[3]); break;          // vec4(exp(a.x), log(a.y), pow(a.y, a.z), acos(a.w))
        case 2: r[0] = lfom::atan2(a[0], a[1]); r[1] = fdiv(a[0], a[1]); r[2] = mixf(a[0], a[1], a[2]); r[3] = inversesqrt(fabsf(a[3])); break;   // vec4(atan(a.x, a.y), a.x / a.y, mix(a.x, a.y, a.z), inversesqrt(abs(a.w)))
        case 3: {   // vec4(refract(normalize(a.xyz), normalize(a.zxy * vec3(1, -1, 1) + 0.1), a.w), 0)
            vec3 I = normalize(vec3{a[0], a[1], a[2]});
            vec3 N = normalize(vec3{a[2], a[0], a[1]} * vec3{1.0f, -1.0f, 1.0f} + V3(0.1f));
            vec3 q = refract(I, N, a[3]);
            r[0] = q.x; r[1] = q.y; r[2] = q.z;
            break;
        }
        case 4: {   // vec4(reflect(normalize(a.xyz), normalize(a.zxy * vec3(1, -1, 1) + 0.1)), length(a.xyz))
            vec3 I = normalize(vec3{a[0], a[1], a[2]});
            vec3 N = normalize(vec3{a[2], a[0], a[1]} * vec3{1.0f, -1.0f, 1.0f} + V3(0.1f));
            vec3 q = reflect(I, N);
            r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = length(vec3{a[0], a[1], a[2]});
            break;
        }
        case 5: {   // texture(tex8, a.xyz): RGBA8 array, LINEAR, REPEAT
            int layer = (int)floorf(a[2] + 0.5f);
            if (layer < 0) layer = 0;
            if (layer >= texL) layer = texL - 1;
            vec4 t = tex8Linear(tex + (size_t)4 * texW * texH * layer, texW, texH, a[0], a[1]);
            r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
            break;
        }
        default: break;
        }
    }
}

}  // namespace lforacle
