// lf_device.cuh — device functions of the CUDA path tracer: fp32 vector math, RNG, ray/box/triangle/light
// intersection, two-level BVH traversal (closest-hit and any-hit), Disney BSDF sample/eval, light and
// environment sampling.  Each function cites the GLSL it implements (paths under /root/reference/shaders/).
//
// Arithmetic contract: IEEE fp32, compiled with -fmad=false so that no multiply-add is contracted (the
// reference GLSL on llvmpipe does not contract either); square roots and reciprocals are the correctly rounded
// ones, and a GLSL division x / y is x * (1 / y) (fdiv below) because that is what Mesa's compiler emits.  Operation ORDER follows the GLSL expression trees, because the parity bar (1e-3 per pixel on 99.9 %
// of pixels with a seeded RNG) needs identical branch decisions almost everywhere.
#pragma once

#include "lf_types.h"
#include "lf_math.cuh"

namespace lf {

// ---------------------------------------------------------------------------------------------- math
struct f3 { float x, y, z; };

#define LFD __device__ __forceinline__
// The shade kernel is instruction-cache bound (`no_instruction` is its top stall).  Two remedies were measured: the
// transcendental functions out of line (lf_math.cuh, kept: -5 %) and the BSDF lobes / DisneyEval out of line (rejected).
#ifdef LF_OUTLINE_EVAL   // measured on C2: outlining the BSDF lobes / DisneyEval costs 40 % in the shade kernel (struct traffic through the stack)
#define LFN __device__ __noinline__
#else
#define LFN __device__ __forceinline__
#endif

LFD f3 mk3(float a, float b, float c) { f3 r; r.x = a; r.y = b; r.z = c; return r; }
LFD f3 mk3(float a) { return mk3(a, a, a); }
LFD f3 xyz(float4 v) { return mk3(v.x, v.y, v.z); }
LFD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
LFD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
LFD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
// GLSL x / y as Mesa lowers it (MUL by RCP; two roundings).  Pinned against llvmpipe by the oracle's golden tests.
LFD float fdiv(float a, float b) { return a * (1.0f / b); }
LFD f3 operator/(f3 a, f3 b) { return mk3(fdiv(a.x, b.x), fdiv(a.y, b.y), fdiv(a.z, b.z)); }
LFD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
LFD f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
LFD f3 operator/(f3 a, float s) { float r = 1.0f / s; return mk3(a.x * r, a.y * r, a.z * r); }
LFD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
LFD float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
LFD f3 cross(f3 a, f3 b) { return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
LFD f3 normalize(f3 v) { return v * (1.0f / sqrtf(dot(v, v))); }     // v * inversesqrt(dot(v, v))
LFD float length(f3 v) { return sqrtf(dot(v, v)); }
LFD float gmax(float a, float b) { return a > b ? a : b; }          // MAXPS(a, b) operand semantics
LFD float gmin(float a, float b) { return a < b ? a : b; }
LFD float clampf(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
LFD float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }   // mix() as Mesa lowers it for llvmpipe
LFD f3 mix3(f3 a, f3 b, float t) { return mk3(mixf(a.x, b.x, t), mixf(a.y, b.y, t), mixf(a.z, b.z, t)); }
LFD f3 reflect3(f3 I, f3 N) { return I - (2.0f * dot(N, I)) * N; }
LFD f3 refract3(f3 I, f3 N, float eta) {
    float ndi = dot(N, I);
    float k = 1.0f - eta * (eta * (1.0f - ndi * ndi));   // Mesa's builtin: mul(eta, mul(eta, ...))
    if (k < 0.0f) return mk3(0.0f);
    return eta * I - (eta * ndi + sqrtf(k)) * N;
}
LFD f3 pow3(f3 a, float e) { return mk3(lf_pow(a.x, e), lf_pow(a.y, e), lf_pow(a.z, e)); }

constexpr float kPI = 3.14159265358979323f;        // globals.glsl:6-9
constexpr float kTWO_PI = 6.28318530717958648f;
constexpr float kINF = 1000000.0f;
constexpr float kEPS = 0.001f;
// Distance cull (not in the reference, which enters every pierced box): a child whose slab ENTRY distance
// exceeds the best accepted t by this relative margin cannot hold a closer hit; the margin covers the
// rounding gap between the slab test and the triangle test.  LfParams.no_cull switches it off.
constexpr float kCullSlack = 1.0001f;

LFD float4 ldg4(const float4* p) { return __ldg(p); }
// 256-bit read-only load (LDG.E.ENL2.256.CONSTANT, sm_100+): a 64-byte node costs 2 L1 lookups per lane instead of 4.
// Used for the nodes only: on the 48-byte triangle records (256 + 128 bit) it measured slower than 3 x LDG.E.128.
struct f8 { float4 lo, hi; };
#ifdef LF_HOST_CHECK   // tests/hostcheck compiles this header for the host: no PTX there
LFD f8 ldg8(const float4* p) { f8 r; r.lo = p[0]; r.hi = p[1]; return r; }
#else
LFD f8 ldg8(const float4* p) {
    f8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p));
    return r;
}
#endif

// ---------------------------------------------------------------------------------------------- RNG
// globals.glsl:122-133.  State lives in registers inside a kernel, in PathSoA::rng between kernels.
struct Rng { unsigned x, y, z, w; };
LFD void pcg4d(Rng& v) {
    v.x = v.x * 1664525u + 1013904223u; v.y = v.y * 1664525u + 1013904223u;
    v.z = v.z * 1664525u + 1013904223u; v.w = v.w * 1664525u + 1013904223u;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    v.x ^= v.x >> 16u; v.y ^= v.y >> 16u; v.z ^= v.z >> 16u; v.w ^= v.w >> 16u;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
}
LFD float rnd(Rng& s) {
    pcg4d(s);
    return __uint2float_rn(s.x) * 2.3283064365386963e-10f;   // float(seed.x) / float(0xffffffffu): the divisor rounds to 2^32
}

// ---------------------------------------------------------------------------------------------- records
struct Ray { f3 o, d; };
struct Hit {
    float t, u, v, lpdf;       // uvt.z, uvt.x, uvt.y; emitter pdf
    int tri, inst, light, mat; // triangle ref, instance, analytic light index (-1 = surface or miss), material
    f3 fhp;                    // state.fhp (world)
};
struct Mat {                   // globals.glsl:24-46 (fields the path reads)
    f3 albedo; float specular; f3 emission;
    float metallic, roughness, subsurface, specularTint, sheen, sheenTint, clearcoat, clearcoatRoughness;
    float specTrans, ior, atDistance; f3 extinction;
    float texA, texMR, texN, texE;
};
struct Surf {                  // the part of `State` (globals.glsl:71-90) DisneySample/DisneyEval read
    f3 normal, ffnormal, tangent, bitangent;
    float eta;
    Mat mat;
};

template <bool COUNT> LFD void bump(DevCounters* c, int which, unsigned n = 1) {
    if (COUNT) atomicAdd(&c->v[which], (unsigned long long)n);
}

// ---------------------------------------------------------------------------------------------- intersection.glsl
LFD float SphereIntersect(float rad, f3 pos, const Ray& r) {   // intersection.glsl:7-27
    f3 op = pos - r.o;
    float b = dot(op, r.d);
    float det = b * b - dot(op, op) + rad * rad;
    if (det < 0.0f) return kINF;
    det = sqrtf(det);
    float t1 = b - det;
    if (t1 > 0.001f) return t1;
    float t2 = b + det;
    if (t2 > 0.001f) return t2;
    return kINF;
}
LFD float RectIntersect(f3 pos, f3 u, f3 v, f3 n, float planeW, const Ray& r) {   // intersection.glsl:30-50
    float dt = dot(r.d, n);
    float t = fdiv(planeW - dot(n, r.o), dt);
    if (t > kEPS) {
        f3 p = r.o + r.d * t;
        f3 vi = p - pos;
        float a1 = dot(u, vi);
        if (a1 >= 0.f && a1 <= 1.f) {
            float a2 = dot(v, vi);
            if (a2 >= 0.f && a2 <= 1.f) return t;
        }
    }
    return kINF;
}
// intersection.glsl:53-67 with invdir hoisted out (same quotient every call); also returns the entry distance.
// GLSL min / max on llvmpipe are MINPS / MAXPS: (a < b ? a : b), i.e. the SECOND operand whenever one is NaN, whereas the
// GPU's FMNMX (fminf / fmaxf) returns the non-NaN one.  A slab product is NaN only as 0 * inf: a direction component that is
// exactly 0 (1/d = inf) with the origin exactly on that face's plane.  `axisRay` (any component of invdir infinite, kept per
// ray in Walk::axis) selects the operand-exact form; every other ray has no NaN and the two forms agree on every value the
// caller looks at (they differ only in the sign of a zero, which no comparison here sees), so it keeps the 1-instruction
// min / max.  KAT: tests/test_edge_cases.py::test_axis_ray_on_flat_box_plane.
LFD float AABBIntersect(f3 mn, f3 mx, f3 o, f3 invdir, bool axisRay, float& entry) {
    f3 f = (mx - o) * invdir;
    f3 n = (mn - o) * invdir;
    float t1, t0;
    if (axisRay) {
        t1 = gmin(gmax(f.x, n.x), gmin(gmax(f.y, n.y), gmax(f.z, n.z)));
        t0 = gmax(gmin(f.x, n.x), gmax(gmin(f.y, n.y), gmin(f.z, n.z)));
    } else {
        t1 = fminf(fmaxf(f.x, n.x), fminf(fmaxf(f.y, n.y), fmaxf(f.z, n.z)));
        t0 = fmaxf(fminf(f.x, n.x), fmaxf(fminf(f.y, n.y), fminf(f.z, n.z)));
    }
    entry = t0;
    return (t1 >= t0) ? (t0 > 0.f ? t0 : t1) : -1.0f;
}
LFD bool has_inf(f3 v) { return fabsf(v.x) == __int_as_float(0x7f800000) || fabsf(v.y) == __int_as_float(0x7f800000) || fabsf(v.z) == __int_as_float(0x7f800000); }

struct LightRec {              // 7 float4, see lf_repack.cpp
    f3 position, emission, u, v, normal, uu, vv;
    float radius, area, type, planeW;
};
LFD LightRec load_light(const DevScene& S, int i) {
    LightRec L;
    if (i < 0 || i >= S.num_lights) {   // out-of-range texelFetch reads zeros (reachable via rand() == 1.0, pathtrace.glsl:168)
        L.position = L.emission = L.u = L.v = L.uu = L.vv = mk3(0.f);
        L.normal = mk3(__int_as_float(0x7fc00000));   // normalize(cross(0,0)) = NaN
        L.radius = L.area = L.type = 0.f; L.planeW = __int_as_float(0x7fc00000);
        return L;
    }
    const float4* p = S.lights + (size_t)kLightStride * i;
    float4 a = ldg4(p), b = ldg4(p + 1), c = ldg4(p + 2), d = ldg4(p + 3), e = ldg4(p + 4), f = ldg4(p + 5), g = ldg4(p + 6);
    L.position = mk3(a.x, a.y, a.z); L.emission = mk3(a.w, b.x, b.y); L.u = mk3(b.z, b.w, c.x); L.v = mk3(c.y, c.z, c.w);
    L.radius = d.x; L.area = d.y; L.type = d.z; L.planeW = d.w;
    L.normal = mk3(e.x, e.y, e.z); L.uu = mk3(e.w, f.x, f.y); L.vv = mk3(f.z, f.w, g.x);
    return L;
}

// ---------------------------------------------------------------------------------------------- traversal
// closest_hit.glsl:7-205 (ANY = false) and anyhit.glsl:7-173 (ANY = true) over the re-packed arrays, cut into the
// steps both kernel organisations share: per-thread walk (trace(), megakernel) and warp-cooperative walk (k_trace).
// `stk` is this thread's column of the CTA's shared-memory stack (stride = kBlockThreads ints).

struct Walk {                  // traversal registers of one ray
    f3 o, d, idir;             // ray in the current space (world or instance-local), 1/d
    int ref, sp;               // current node reference, stack pointer
    int curInst, curMat;
    bool inBlas;
    bool axis;                 // some component of idir is infinite: AABBIntersect takes its NaN-exact form
};

LFD void hit_clear(Hit& hit) { hit.light = -1; hit.tri = -1; hit.inst = -1; hit.mat = -1; hit.u = hit.v = 0.f; hit.lpdf = 0.f; hit.t = kINF; }

// Analytic lights first (closest_hit.glsl:13-67, anyhit.glsl:11-46).  ANY: returns true when a light blocks the ray.
template <bool ANY, bool COUNT>
LFD bool test_lights(const DevScene& S, const Ray& r, float maxDist, Hit& hit, DevCounters* cnt) {
    for (int i = 0; i < S.num_lights; i++) {
        LightRec L = load_light(S, i);
        bump<COUNT>(cnt, C_LIGHT); if (ANY) bump<COUNT>(cnt, C_LIGHT_SH);
        if (L.type == 0.f) {
            if (!ANY && dot(L.normal, r.d) > 0.f) continue;        // back-facing quad hidden from closest-hit only
            float d = RectIntersect(L.position, L.uu, L.vv, L.normal, L.planeW, r);
            if (ANY) { if (d > 0.0f && d < maxDist) return true; }
            else {
                if (d < 0.f) d = kINF;
                if (d < hit.t) {
                    hit.t = d;
                    float cosTheta = dot(-r.d, L.normal);
                    hit.lpdf = fdiv(d * d, L.area * cosTheta);
                    hit.light = i;
                }
            }
        }
        if (L.type == 1.f) {
            float d = SphereIntersect(L.radius, L.position, r);
            if (ANY) { if (d > 0.0f && d < maxDist) return true; }
            else {
                if (d < 0.f) d = kINF;
                if (d < hit.t) { hit.t = d; hit.lpdf = fdiv(d * d, L.area); hit.light = i; }
            }
        }
    }
    return false;
}

// A thread's traversal stack (the reference's `int stack[64]`, closest_hit.glsl:70).
//   PlainStk       a column of a [depth][kBlockThreads] int array (megakernel: shared memory; tests/hostcheck: host memory)
struct PlainStk {
    int* col;
    LFD void push(int& sp, int v) const { col[sp * kBlockThreads] = v; sp++; }
    LFD int pop(int& sp) const { --sp; return col[sp * kBlockThreads]; }
};
template <class ST>
LFD void walk_begin(const DevScene& S, const Ray& r, Walk& w, const ST& stk) {
    w.sp = 0;
    stk.push(w.sp, kRefSentinel);                 // stack[ptr++] = -1
    w.ref = S.top_ref;
    w.inBlas = false;
    w.curInst = -1; w.curMat = 0;
    w.o = r.o; w.d = r.d;
    w.idir = mk3(1.0f) / r.d;
    w.axis = has_inf(w.idir);
}

// inverse(M) * vec4(origin, 1) and * vec4(direction, 0), summed column by column like the GLSL mat4*vec4 (closest_hit.glsl:159-160)
LFD void to_instance(const float4 r0, const float4 r1, const float4 r2, const Ray& r, f3& o, f3& d) {
    o = mk3(((r0.x * r.o.x + r0.y * r.o.y) + r0.z * r.o.z) + r0.w * 1.0f,
            ((r1.x * r.o.x + r1.y * r.o.y) + r1.z * r.o.z) + r1.w * 1.0f,
            ((r2.x * r.o.x + r2.y * r.o.y) + r2.z * r.o.z) + r2.w * 1.0f);
    d = mk3(((r0.x * r.d.x + r0.y * r.d.y) + r0.z * r.d.z) + r0.w * 0.0f,
            ((r1.x * r.d.x + r1.y * r.d.y) + r1.z * r.d.z) + r1.w * 0.0f,
            ((r2.x * r.d.x + r2.y * r.d.y) + r2.z * r.d.z) + r2.w * 0.0f);
}

// One step of the walk for a reference that is not a triangle leaf: inner node, instance entry, or stack marker.
// Returns false when the walk is over (marker popped at world level).  `limit` = current best t (closest) or maxDist (any).
// AXM: how the slab test treats Walk::axis.  0 = look at the flag (per-thread walks), 1 = the caller guarantees !w.axis (k_trace's
// inner loop: no branch per node), 2 = the caller guarantees w.axis (k_trace steps those rare lanes apart, one node at a time).
template <bool ANY, bool CULL, bool COUNT, int AXM = 0, class ST>
LFD bool walk_step(const DevScene& S, const Ray& r, Walk& w, float limit, const ST& stk, DevCounters* cnt) {
    if (w.ref >= 0) {                             // inner node (closest_hit.glsl:167-199)
        bump<COUNT>(cnt, C_INNER); if (ANY) bump<COUNT>(cnt, C_INNER_SH);
        const float4* n = S.nodes + (size_t)4 * w.ref;
#if defined(LF_NODE_TEX) && LF_NODE_TEX == 1   // experiment: the whole node through the texture unit (4 x TLD4-style float4 fetches)
        float4 n0 = tex1Dfetch<float4>(S.nodes_tex, 4 * w.ref), n1 = tex1Dfetch<float4>(S.nodes_tex, 4 * w.ref + 1);
        float4 n2 = tex1Dfetch<float4>(S.nodes_tex, 4 * w.ref + 2), n3 = tex1Dfetch<float4>(S.nodes_tex, 4 * w.ref + 3);
#elif defined(LF_NODE_TEX) && LF_NODE_TEX == 2   // experiment: half through the LSU (one 256-bit load), half through the texture unit
        f8 na = ldg8(n);
        float4 n0 = na.lo, n1 = na.hi;
        float4 n2 = tex1Dfetch<float4>(S.nodes_tex, 4 * w.ref + 2), n3 = tex1Dfetch<float4>(S.nodes_tex, 4 * w.ref + 3);
#elif !defined(LF_NODE_LDG128)   // 256-bit node loads: extend -4.6 %, shadow -3 % on C2 against 4 x LDG.E.128 (A/B on one box)
        f8 na = ldg8(n), nb = ldg8(n + 2);
        float4 n0 = na.lo, n1 = na.hi, n2 = nb.lo, n3 = nb.hi;
#else
        float4 n0 = ldg4(n), n1 = ldg4(n + 1), n2 = ldg4(n + 2), n3 = ldg4(n + 3);
#endif
        float le, re;
        const bool axisRay = AXM == 0 ? w.axis : AXM == 2;
        float leftHit = AABBIntersect(mk3(n0.x, n0.y, n0.z), mk3(n0.w, n1.x, n1.y), w.o, w.idir, axisRay, le);
        float rightHit = AABBIntersect(mk3(n1.z, n1.w, n2.x), mk3(n2.y, n2.z, n2.w), w.o, w.idir, axisRay, re);
        int leftRef = __float_as_int(n3.x), rightRef = __float_as_int(n3.y);
        bool lok = leftHit > 0.0f, rok = rightHit > 0.0f;
        if (CULL) {
            float lim = limit * kCullSlack;
            lok = lok && !(le > lim);
            rok = rok && !(re > lim);
        }
        if (lok && rok) {
            bool swap = leftHit > rightHit;       // near child first, far child deferred (:172-184)
            w.ref = swap ? rightRef : leftRef;
            stk.push(w.sp, swap ? leftRef : rightRef);
        } else if (lok) w.ref = leftRef;
        else if (rok) w.ref = rightRef;
        else w.ref = stk.pop(w.sp);
        return true;
    }
    if (w.ref == kRefSentinel) {                  // `idx < 0`: end of a BLAS (restore the world ray) or of the walk
        if (!w.inBlas) return false;
        w.inBlas = false;
        w.o = r.o; w.d = r.d; w.idir = mk3(1.0f) / r.d;
        w.axis = has_inf(w.idir);
        w.ref = stk.pop(w.sp);
        return true;
    }
    // TLAS leaf (closest_hit.glsl:148-166)
    bump<COUNT>(cnt, C_TLAS); if (ANY) bump<COUNT>(cnt, C_TLAS_SH);
    w.curInst = ref_instance(w.ref);
    const float4* ip = S.inst + (size_t)kInstStride * w.curInst;
    float4 r0 = ldg4(ip), r1 = ldg4(ip + 1), r2 = ldg4(ip + 2), meta = ldg4(ip + 3);
    to_instance(r0, r1, r2, r, w.o, w.d);
    w.idir = mk3(1.0f) / w.d;
    w.axis = has_inf(w.idir);
    stk.push(w.sp, kRefSentinel);
    w.inBlas = true;
    w.curMat = __float_as_int(meta.y);
    w.ref = __float_as_int(meta.x);
    return true;
}

// BLAS leaf (closest_hit.glsl:108-147, anyhit.glsl:87-116).  ANY: returns true at the first hit below maxDist.
template <bool ANY, bool COUNT>
LFD bool walk_leaf(const DevScene& S, const Walk& w, float maxDist, Hit& hit, DevCounters* cnt) {
    bump<COUNT>(cnt, C_LEAF); if (ANY) bump<COUNT>(cnt, C_LEAF_SH);
    const int first = ref_leaf_first(w.ref), count = ref_leaf_count(w.ref);
    for (int i = 0; i < count; i++) {
#ifdef LF_TRI_TEX   // experiment: triangle records through the texture unit
        const int ti = kTriStride * (first + i);
        float4 q0 = tex1Dfetch<float4>(S.tris_tex, ti), q1 = tex1Dfetch<float4>(S.tris_tex, ti + 1), q2 = tex1Dfetch<float4>(S.tris_tex, ti + 2);
#else
        const float4* tp = S.tris + (size_t)kTriStride * (first + i);
        float4 q0 = ldg4(tp), q1 = ldg4(tp + 1), q2 = ldg4(tp + 2);
#endif
        bump<COUNT>(cnt, C_TRI); if (ANY) bump<COUNT>(cnt, C_TRI_SH);
        f3 e0 = mk3(q1.x, q1.y, q1.z), e1 = mk3(q2.x, q2.y, q2.z);
        f3 pv = cross(w.d, e1);
        float det = dot(e0, pv);
        f3 tv = w.o - mk3(q0.x, q0.y, q0.z);
        // The shader accepts iff u = a/det, v = b/det, t = c/det and w = 1-u-v are all >= 0 (:121-131).  A quotient of
        // two non-zero floats of strictly opposite sign is negative, so those cases are rejected from the sign of the
        // product, before any division; everything else takes the shader's exact arithmetic.
        float a = dot(tv, pv);
#ifdef LF_TRI_BRANCHLESS   // experiment: one combined sign test after all three products (same values, no early exits)
        f3 qv = cross(tv, e0);
        float b = dot(w.d, qv);
        float c = dot(e1, qv);
        if (a * det < 0.f || b * det < 0.f || c * det < 0.f) continue;
#else
        if (a * det < 0.f) continue;
        f3 qv = cross(tv, e0);
        float b = dot(w.d, qv);
        if (b * det < 0.f) continue;
        float c = dot(e1, qv);
        if (c * det < 0.f) continue;
#endif
        float rdet = 1.0f / det;                  // uvt.xyz / det = uvt.xyz * rcp(det)
        float uu = a * rdet;
        float vv = b * rdet;
        float ww = 1.0f - uu - vv;
        if (!(uu >= 0.f) || !(vv >= 0.f) || !(ww >= 0.f)) continue;
        float tt = c * rdet;
        if (!(tt >= 0.f)) continue;
        if (ANY) { if (tt < maxDist) return true; }
        else if (tt < hit.t) { hit.t = tt; hit.u = uu; hit.v = vv; hit.tri = first + i; hit.inst = w.curInst; hit.mat = w.curMat; hit.light = -1; }
    }
    return false;
}

// state.fhp = vec3(M * vec4(r_trans.origin + r_trans.direction * t, 1)) (closest_hit.glsl:139,143) for the final hit
LFD void hit_point(const DevScene& S, const Ray& r, Hit& hit) {
    if (hit.light < 0 && hit.tri >= 0) {
        const float4* ip = S.inst + (size_t)kInstStride * hit.inst;
        float4 r0 = ldg4(ip), r1 = ldg4(ip + 1), r2 = ldg4(ip + 2);
        f3 oo, dd;
        to_instance(r0, r1, r2, r, oo, dd);
        f3 ph = oo + dd * hit.t;
        float4 m0 = ldg4(ip + 4), m1 = ldg4(ip + 5), m2 = ldg4(ip + 6);
        hit.fhp = mk3(((m0.x * ph.x + m0.y * ph.y) + m0.z * ph.z) + m0.w * 1.0f,
                      ((m1.x * ph.x + m1.y * ph.y) + m1.z * ph.z) + m1.w * 1.0f,
                      ((m2.x * ph.x + m2.y * ph.y) + m2.z * ph.z) + m2.w * 1.0f);
    } else {
        hit.fhp = mk3(0.f);
    }
}

// Per-thread walk (megakernel).  Closest: fills `hit`, returns t != INFINITY.  Any: returns true when occluded.
template <bool ANY, bool CULL, bool COUNT>
LFD bool trace(const DevScene& S, const Ray& r, float maxDist, Hit& hit, int* stkcol, DevCounters* cnt) {
    PlainStk stk; stk.col = stkcol;
    if (!ANY) hit_clear(hit);
    bump<COUNT>(cnt, ANY ? C_RAYS_SHADOW : C_RAYS_CLOSEST);
    if (test_lights<ANY, COUNT>(S, r, maxDist, hit, cnt)) return true;
    Walk w;
    walk_begin(S, r, w, stk);
    for (;;) {
        bool more = true;
        while (w.ref >= 0 || (w.ref & kRefTlasBit)) {
            more = walk_step<ANY, CULL, COUNT>(S, r, w, ANY ? maxDist : hit.t, stk, cnt);
            if (!more) break;
        }
        if (!more) break;
        if (walk_leaf<ANY, COUNT>(S, w, maxDist, hit, cnt)) return true;
        w.ref = stk.pop(w.sp);
    }
    if (ANY) return false;
    hit_point(S, r, hit);
    return hit.t != kINF;
}

// ---------------------------------------------------------------------------------------------- sampling.glsl
LFD f3 ImportanceSampleGTR1(float rgh, float r1) {   // sampling.glsl:7-21
    float a = gmax(0.001f, rgh);
    float a2 = a * a;
    float phi = r1 * kTWO_PI;
    float cosTheta = sqrtf(fdiv(1.0f - lf_pow(a2, 1.0f - r1), 1.0f - a2));
    float sinTheta = clampf(sqrtf(1.0f - (cosTheta * cosTheta)), 0.0f, 1.0f);
    float sinPhi, cosPhi;
    lf_sincos(phi, sinPhi, cosPhi);
    return mk3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
}
LFD f3 ImportanceSampleGTR2(float rgh, float r1, float r2) {   // sampling.glsl:37-49
    float a = gmax(0.001f, rgh);
    float phi = r1 * kTWO_PI;
    float cosTheta = sqrtf(fdiv(1.0f - r2, 1.0f + (a * a - 1.0f) * r2));
    float sinTheta = clampf(sqrtf(1.0f - (cosTheta * cosTheta)), 0.0f, 1.0f);
    float sinPhi, cosPhi;
    lf_sincos(phi, sinPhi, cosPhi);
    return mk3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
}
LFD float SchlickFresnel(float u) {   // sampling.glsl:52-58
    float m = clampf(1.0f - u, 0.0f, 1.0f);
    float m2 = m * m;
    return m2 * m2 * m;
}
LFD float DielectricFresnel(float cos_theta_i, float eta) {   // sampling.glsl:61-76
    float sinThetaTSq = eta * eta * (1.0f - cos_theta_i * cos_theta_i);
    if (sinThetaTSq > 1.0f) return 1.0f;
    float cos_theta_t = sqrtf(gmax(1.0f - sinThetaTSq, 0.0f));
    float rs = fdiv(eta * cos_theta_t - cos_theta_i, eta * cos_theta_t + cos_theta_i);
    float rp = fdiv(eta * cos_theta_i - cos_theta_t, eta * cos_theta_i + cos_theta_t);
    return 0.5f * (rs * rs + rp * rp);
}
LFD float GTR1(float NDotH, float a) {   // sampling.glsl:79-87
    if (a >= 1.0f) return (1.0f / kPI);
    float a2 = a * a;
    float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
    // PI * log(a2) as Mesa compiles it: log(x) = log2(x) * ln 2, then the two constants of the product chain are folded
    return fdiv(a2 - 1.0f, ((kPI * 0.693147180559945309417f) * lf_log2(a2)) * t);
}
LFD float GTR2(float NDotH, float a) {   // sampling.glsl:90-96
    float a2 = a * a;
    float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
    return fdiv(a2, kPI * t * t);
}
LFD float SmithG_GGX(float NDotV, float alphaG) {   // sampling.glsl:110-116
    float a = alphaG * alphaG;
    float b = NDotV * NDotV;
    return 1.0f / (NDotV + sqrtf(a + b - a * b));
}
LFD f3 CosineSampleHemisphere(float r1, float r2) {   // sampling.glsl:129-140
    f3 dir;
    float r = sqrtf(r1);
    float phi = kTWO_PI * r2;
    float sp, cp;
    lf_sincos(phi, sp, cp);
    dir.x = r * cp;
    dir.y = r * sp;
    dir.z = sqrtf(gmax(0.0f, 1.0f - dir.x * dir.x - dir.y * dir.y));
    return dir;
}
LFD f3 UniformSampleSphere(float r1, float r2) {   // sampling.glsl:153-160
    float z = 1.0f - 2.0f * r1;
    float r = sqrtf(gmax(0.0f, 1.0f - z * z));
    float phi = kTWO_PI * r2;
    float sp, cp;
    lf_sincos(phi, sp, cp);
    return mk3(r * cp, r * sp, z);
}
LFD float powerHeuristic(float a, float b) {   // sampling.glsl:163-169
    float t = a * a;
    return fdiv(t, b * b + t);
}

struct LightSample { f3 normal, emission, direction; float dist, pdf; };   // LightSampleRec, globals.glsl:99-106

// sampleOneLight and its three callees (sampling.glsl:172-230)
LFD void sampleOneLight(const LightRec& light, int numLights, f3 surfacePos, Rng& rng, LightSample& rec) {
    int type = (int)light.type;
    if (type == 0 || type == 1) {
        float r1 = rnd(rng), r2 = rnd(rng);
        f3 lightSurfacePos;
        if (type == 0) lightSurfacePos = light.position + light.u * r1 + light.v * r2;
        else lightSurfacePos = light.position + UniformSampleSphere(r1, r2) * light.radius;
        rec.direction = lightSurfacePos - surfacePos;
        rec.dist = length(rec.direction);
        float distSq = rec.dist * rec.dist;
        rec.direction = rec.direction / rec.dist;
        if (type == 0) rec.normal = normalize(cross(light.u, light.v));
        else rec.normal = normalize(lightSurfacePos - light.position);
        rec.emission = light.emission * (float)numLights;
        rec.pdf = fdiv(distSq, light.area * fabsf(dot(rec.normal, rec.direction)));
    } else {
        rec.direction = normalize(light.position - mk3(0.0f));
        rec.normal = normalize(surfacePos - light.position);
        rec.emission = light.emission * (float)numLights;
        rec.dist = kINF;
        rec.pdf = 1.0f;
    }
}

// ---- environment map: GL sampling rules restated (Renderer.cpp:163-185: hdrTex LINEAR, tables NEAREST, REPEAT)
// GL_REPEAT texel wrap: full modulo for arbitrary texture coordinates (material textures), and a branch-free
// two-select form for indices known to lie in [-n, 2n) (env-map lookups: u, v in [0, 1], bilinear neighbours).
LFD int wrapi(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }
LFD int wrapn(int i, int n) { i += (i < 0) ? n : 0; i -= (i >= n) ? n : 0; return i; }
LFD int nearestIdx(float u, int n) { return wrapn((int)floorf(u * (float)n), n); }
LFN f3 hdrLinear(const DevScene& S, float u, float v) {
    float x = u * (float)S.hdr_w - 0.5f, y = v * (float)S.hdr_h - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float wx = x - fx, wy = y - fy;
    int x0 = wrapn((int)fx, S.hdr_w), x1 = wrapn((int)fx + 1, S.hdr_w), y0 = wrapn((int)fy, S.hdr_h), y1 = wrapn((int)fy + 1, S.hdr_h);
    f3 a = xyz(tex2D<float4>(S.hdr_tex, (float)x0, (float)y0)), b = xyz(tex2D<float4>(S.hdr_tex, (float)x1, (float)y0));
    f3 c = xyz(tex2D<float4>(S.hdr_tex, (float)x0, (float)y1)), e = xyz(tex2D<float4>(S.hdr_tex, (float)x1, (float)y1));
    f3 top = a + (b - a) * wx, bot = c + (e - c) * wx;
    return top + (bot - top) * wy;
}
LFD float2 marginalAt(const DevScene& S, float u) { return __ldg(S.marginal + nearestIdx(u, S.hdr_h)); }
LFD float2 conditionalAt(const DevScene& S, float u, float v) {
    return __ldg(S.conditional + ((size_t)nearestIdx(v, S.hdr_h) * S.hdr_w + nearestIdx(u, S.hdr_w)));
}
LFN float EnvPdf(const DevScene& S, const DevParams& P, f3 dir) {   // sampling.glsl:236-243
    float theta = lf_acos(clampf(dir.y, -1.0f, 1.0f));
    float ux = (kPI + lf_atan2(dir.z, dir.x)) * (1.0f / kTWO_PI), uy = theta * (1.0f / kPI);
    float pdf = conditionalAt(S, ux, uy).y * marginalAt(S, uy).y;
    float st, ct;
    lf_sincos(theta, st, ct);
    return fdiv(pdf * P.hdr_resolution, 2.0f * kPI * kPI * st);
}
// sampling.glsl:246-265; returns direction, pdf in .w
LFN float4 EnvSample(const DevScene& S, const DevParams& P, Rng& rng, f3& color) {
    float r1 = rnd(rng), r2 = rnd(rng);
    float v = marginalAt(S, r1).x;
    float u = conditionalAt(S, r2, v).x;
    color = hdrLinear(S, u, v) * P.hdr_multiplier;
    float pdf = conditionalAt(S, u, v).y * marginalAt(S, v).y;
    float phi = u * kTWO_PI;
    float theta = v * kPI;
    float st, ct, sph, cph;
    lf_sincos(theta, st, ct);
    lf_sincos(phi, sph, cph);
    if (st == 0.0f) pdf = 0.0f;
    return make_float4(-st * cph, ct, -st * sph, fdiv(pdf * P.hdr_resolution, 2.0f * kPI * kPI * st));
}

// material texture array: RGBA8, LINEAR, REPEAT (Renderer.cpp:151-160), fetched unfiltered and filtered the way llvmpipe
// filters 8-bit formats: texel coordinates in 24.8 fixed point, X = round_to_nearest_even(u * W * 256) - 128, texels
// X >> 8 and its +1 neighbour (wrapped), weight X & 255; each lerp is v0 + ((w * (v1 - v0)) >> 8) on the 8-bit channel
// values (x first, then y); the 8-bit result times the float constant 1/255.  (The hardware filter, 9-bit weights on
// normalised floats, would not reproduce it.)
LFD int lerp8(int w, int v0, int v1) { return (v0 + ((w * (v1 - v0)) >> 8)) & 255; }
template <bool COUNT>
LFN float4 texArrayLinear(const DevScene& S, float u, float v, int layer, DevCounters* cnt) {
    layer = max(0, min(layer, S.num_tex - 1));
    bump<COUNT>(cnt, C_TEX);
    int X = __float2int_rn(u * (float)S.tex_w * 256.0f) - 128, Y = __float2int_rn(v * (float)S.tex_h * 256.0f) - 128;
    int wx = X & 255, wy = Y & 255;
    int x0 = wrapi(X >> 8, S.tex_w), x1 = wrapi((X >> 8) + 1, S.tex_w), y0 = wrapi(Y >> 8, S.tex_h), y1 = wrapi((Y >> 8) + 1, S.tex_h);
    uchar4 a = tex2DLayered<uchar4>(S.tex_maps, (float)x0, (float)y0, layer), b = tex2DLayered<uchar4>(S.tex_maps, (float)x1, (float)y0, layer);
    uchar4 c = tex2DLayered<uchar4>(S.tex_maps, (float)x0, (float)y1, layer), e = tex2DLayered<uchar4>(S.tex_maps, (float)x1, (float)y1, layer);
    const float k = (float)(1.0 / 255.0);
    return make_float4((float)lerp8(wy, lerp8(wx, a.x, b.x), lerp8(wx, c.x, e.x)) * k, (float)lerp8(wy, lerp8(wx, a.y, b.y), lerp8(wx, c.y, e.y)) * k,
                       (float)lerp8(wy, lerp8(wx, a.z, b.z), lerp8(wx, c.z, e.z)) * k, (float)lerp8(wy, lerp8(wx, a.w, b.w), lerp8(wx, c.w, e.w)) * k);
}

// ---------------------------------------------------------------------------------------------- disney.glsl
LFN f3 EvalDielectricReflection(const Surf& s, f3 V, f3 N, f3 L, f3 H, float& pdf) {   // disney.glsl:19-33
    pdf = 0.0f;
    if (dot(N, L) <= 0.0f) return mk3(0.0f);
    float F = DielectricFresnel(dot(V, H), s.eta);
    float D = GTR2(dot(N, H), s.mat.roughness);
    pdf = fdiv(D * dot(N, H) * F, 4.0f * fabsf(dot(V, H)));
    float G = SmithG_GGX(fabsf(dot(N, L)), s.mat.roughness) * SmithG_GGX(fabsf(dot(N, V)), s.mat.roughness);
    return s.mat.albedo * F * D * G;
}
LFN f3 EvalDielectricRefraction(const Surf& s, f3 V, f3 N, f3 L, f3 H, float& pdf) {   // disney.glsl:36-52
    pdf = 0.0f;
    if (dot(N, L) >= 0.0f) return mk3(0.0f);
    float F = DielectricFresnel(fabsf(dot(V, H)), s.eta);
    float D = GTR2(dot(N, H), s.mat.roughness);
    float denomSqrt = dot(L, H) + dot(V, H) * s.eta;
    pdf = fdiv(D * dot(N, H) * (1.0f - F) * fabsf(dot(L, H)), denomSqrt * denomSqrt);
    float G = SmithG_GGX(fabsf(dot(N, L)), s.mat.roughness) * SmithG_GGX(fabsf(dot(N, V)), s.mat.roughness);
    return s.mat.albedo * (1.0f - F) * D * G * fabsf(dot(V, H)) * fabsf(dot(L, H)) * 4.0f * s.eta * s.eta / (denomSqrt * denomSqrt);
}
LFN f3 EvalSpecular(const Surf& s, f3 Cspec0, f3 V, f3 N, f3 L, f3 H, float& pdf) {   // disney.glsl:55-69
    pdf = 0.0f;
    if (dot(N, L) <= 0.0f) return mk3(0.0f);
    float D = GTR2(dot(N, H), s.mat.roughness);
    pdf = fdiv(D * dot(N, H), 4.0f * dot(V, H));
    float FH = SchlickFresnel(dot(L, H));
    f3 F = mix3(Cspec0, mk3(1.0f), FH);
    float G = SmithG_GGX(fabsf(dot(N, L)), s.mat.roughness) * SmithG_GGX(fabsf(dot(N, V)), s.mat.roughness);
    return F * D * G;
}
LFN f3 EvalClearcoat(const Surf& s, f3 V, f3 N, f3 L, f3 H, float& pdf) {   // disney.glsl:72-86
    pdf = 0.0f;
    if (dot(N, L) <= 0.0f) return mk3(0.0f);
    float D = GTR1(dot(N, H), mixf(0.1f, 0.001f, s.mat.clearcoatRoughness));
    pdf = fdiv(D * dot(N, H), 4.0f * dot(V, H));
    float FH = SchlickFresnel(dot(L, H));
    float F = mixf(0.04f, 1.0f, FH);
    float G = SmithG_GGX(dot(N, L), 0.25f) * SmithG_GGX(dot(N, V), 0.25f);
    return mk3(0.25f * s.mat.clearcoat * F * D * G);
}
LFN f3 EvalDiffuse(const Surf& s, f3 Csheen, f3 V, f3 N, f3 L, f3 H, float& pdf) {   // disney.glsl:89-113
    pdf = 0.0f;
    if (dot(N, L) <= 0.0f) return mk3(0.0f);
    pdf = dot(N, L) * (1.0f / kPI);
    float FL = SchlickFresnel(dot(N, L));
    float FV = SchlickFresnel(dot(N, V));
    float Fss90 = dot(L, H) * dot(L, H) * s.mat.roughness;
    float Fss = mixf(1.0f, Fss90, FL) * mixf(1.0f, Fss90, FV);
    float ss = 1.f * (Fss * (1.0f / (dot(N, L) + dot(N, V)) - 0.5f) + 0.5f);
    // FH * sheen * Csheen with FH = SchlickFresnel(dot(L, H)) = m2 * m2 * m inlined: Mesa's opt_rebalance_tree turns the scalar product
    // chain ((m2 * m2) * m) * sheen into (m2 * m2) * (m * sheen) (pinned against the reference's own EvalDiffuse on llvmpipe)
    float mH = clampf(1.0f - dot(L, H), 0.0f, 1.0f);
    float mH2 = mH * mH;
    f3 Fsheen = ((mH2 * mH2) * (mH * s.mat.sheen)) * Csheen;
    return ((1.0f / kPI) * (ss + s.mat.subsurface) * s.mat.albedo + Fsheen) * (1.0f - s.mat.metallic);   // :112
}
LFD void disneyTints(const Surf& s, f3& Cspec0, f3& Csheen) {   // disney.glsl:140-145, :266-273
    f3 Cdlin = s.mat.albedo;
    float Cdlum = 0.3f * Cdlin.x + 0.6f * Cdlin.y + 0.1f * Cdlin.z;
    f3 Ctint = Cdlum > 0.0f ? Cdlin / Cdlum : mk3(1.0f);
    Cspec0 = mix3(s.mat.specular * 0.08f * mix3(mk3(1.0f), Ctint, s.mat.specularTint), Cdlin, s.mat.metallic);
    Csheen = mix3(mk3(1.0f), Ctint, s.mat.sheenTint);
}

LFD f3 DisneySample(const Surf& s, f3 V, f3 N, Rng& rng, f3& L, float& pdf) {   // disney.glsl:128-225
    pdf = 0.0f;
    f3 f = mk3(0.0f);
    float r1 = rnd(rng), r2 = rnd(rng);
    float diffuseRatio = 0.5f * (1.0f - s.mat.metallic);
    float transWeight = (1.0f - s.mat.metallic) * s.mat.specTrans;
    f3 Cspec0, Csheen;
    disneyTints(s, Cspec0, Csheen);

    if (rnd(rng) < transWeight) {
        f3 H = ImportanceSampleGTR2(s.mat.roughness, r1, r2);
        H = s.tangent * H.x + s.bitangent * H.y + N * H.z;
        if (dot(V, H) < 0.0f) H = -H;
        f3 R = reflect3(-V, H);
        float F = DielectricFresnel(fabsf(dot(R, H)), s.eta);
        if (rnd(rng) < F) {
            L = normalize(R);
            f = EvalDielectricReflection(s, V, N, L, H, pdf);
        } else {
            L = normalize(refract3(-V, H, s.eta));
            f = EvalDielectricRefraction(s, V, N, L, H, pdf);
        }
        f = f * transWeight;
        pdf *= transWeight;
    } else {
        if (rnd(rng) < diffuseRatio) {
            L = CosineSampleHemisphere(r1, r2);
            L = s.tangent * L.x + s.bitangent * L.y + N * L.z;
            f3 H = normalize(L + V);
            f = EvalDiffuse(s, Csheen, V, N, L, H, pdf);
            pdf *= diffuseRatio;
        } else {
            float primarySpecRatio = 1.0f / (1.0f + s.mat.clearcoat);
            if (rnd(rng) < primarySpecRatio) {
                f3 H = ImportanceSampleGTR2(s.mat.roughness, r1, r2);
                H = s.tangent * H.x + s.bitangent * H.y + N * H.z;
                if (dot(V, H) < 0.0f) H = -H;
                L = normalize(reflect3(-V, H));
                f = EvalSpecular(s, Cspec0, V, N, L, H, pdf);
                pdf *= primarySpecRatio * (1.0f - diffuseRatio);
            } else {
                f3 H = ImportanceSampleGTR1(mixf(0.1f, 0.001f, s.mat.clearcoatRoughness), r1);
                H = s.tangent * H.x + s.bitangent * H.y + N * H.z;
                if (dot(V, H) < 0.0f) H = -H;
                L = normalize(reflect3(-V, H));
                f = EvalClearcoat(s, V, N, L, H, pdf);
                pdf *= (1.0f - primarySpecRatio) * (1.0f - diffuseRatio);
            }
        }
        f = f * (1.0f - transWeight);
        pdf *= (1.0f - transWeight);
    }
    return f;
}

LFN f3 DisneyEval(const Surf& s, f3 V, f3 N, f3 L, float& pdf) {   // disney.glsl:228-291
    f3 H;
    bool refl = dot(N, L) > 0.0f;
    if (refl) H = normalize(L + V);
    else H = normalize(L + V * s.eta);
    if (dot(V, H) < 0.0f) H = -H;

    float diffuseRatio = 0.5f * (1.0f - s.mat.metallic);
    float primarySpecRatio = 1.0f / (1.0f + s.mat.clearcoat);
    float transWeight = (1.0f - s.mat.metallic) * s.mat.specTrans;
    f3 brdf = mk3(0.0f), bsdf = mk3(0.0f);
    float brdfPdf = 0.0f, bsdfPdf = 0.0f;

    if (transWeight > 0.0f) {
        if (refl) bsdf = EvalDielectricReflection(s, V, N, L, H, bsdfPdf);
        else bsdf = EvalDielectricRefraction(s, V, N, L, H, bsdfPdf);
    }
    if (transWeight < 1.0f) {
        float m_pdf;
        f3 Cspec0, Csheen;
        disneyTints(s, Cspec0, Csheen);
        brdf = brdf + EvalDiffuse(s, Csheen, V, N, L, H, m_pdf);
        brdfPdf += m_pdf * diffuseRatio;
        brdf = brdf + EvalSpecular(s, Cspec0, V, N, L, H, m_pdf);
        brdfPdf += m_pdf * primarySpecRatio * (1.0f - diffuseRatio);
        brdf = brdf + EvalClearcoat(s, V, N, L, H, m_pdf);
        brdfPdf += m_pdf * (1.0f - primarySpecRatio) * (1.0f - diffuseRatio);
    }
    pdf = mixf(brdfPdf, bsdfPdf, transWeight);
    return mix3(brdf, bsdf, transWeight);
}

// ---------------------------------------------------------------------------------------------- pathtrace.glsl pieces
LFD void Onb(f3 N, f3& T, f3& B) {   // pathtrace.glsl:7-13
    f3 UpVector = fabsf(N.z) < 0.999f ? mk3(0, 0, 1) : mk3(1, 0, 0);
    T = normalize(cross(UpVector, N));
    B = cross(N, T);
}

// GetNormalsAndTexCoord + GetMaterialsAndTextures (pathtrace.glsl:16-123) for a surface hit.
template <bool COUNT, bool TEX = true>
LFD void load_surface(const DevScene& S, const Hit& hit, f3 rdir, Surf& s, DevCounters* cnt) {
    const float4* np = S.trinrm + (size_t)3 * hit.tri;
    const float4* tp = S.tris + (size_t)kTriStride * hit.tri;
    float4 n1 = ldg4(np), n2 = ldg4(np + 1), n3 = ldg4(np + 2);
    float bw = 1.0f - hit.u - hit.v, bu = hit.u, bv = hit.v;      // state.bary = uvt.wxy
    f3 normal = normalize(xyz(n1) * bw + xyz(n2) * bu + xyz(n3) * bv);
    const float4* ip = S.inst + (size_t)kInstStride * hit.inst;
    float4 m0 = ldg4(ip + 7), m1 = ldg4(ip + 8), m2 = ldg4(ip + 9);   // rows of transpose(inverse(mat3(transform)))
    normal = normalize(mk3((m0.x * normal.x + m0.y * normal.y) + m0.z * normal.z,
                           (m1.x * normal.x + m1.y * normal.y) + m1.z * normal.z,
                           (m2.x * normal.x + m2.y * normal.y) + m2.z * normal.z));
    s.normal = normal;
    s.ffnormal = dot(normal, rdir) <= 0.0f ? normal : normal * -1.0f;
    Onb(s.normal, s.tangent, s.bitangent);

    const float4* mp = S.materials + (size_t)7 * hit.mat;
    float4 p1 = ldg4(mp), p2 = ldg4(mp + 1), p3 = ldg4(mp + 2), p4 = ldg4(mp + 3), p5 = ldg4(mp + 4), p6 = ldg4(mp + 5), p7 = ldg4(mp + 6);
    bump<COUNT>(cnt, C_SHADED);
    Mat& mat = s.mat;
    mat.albedo = mk3(p1.x, p1.y, p1.z); mat.specular = p1.w;
    mat.emission = mk3(p2.x, p2.y, p2.z);
    mat.metallic = p3.x; mat.roughness = gmax(p3.y, 0.001f); mat.subsurface = p3.z; mat.specularTint = p3.w;
    mat.sheen = p4.x; mat.sheenTint = p4.y; mat.clearcoat = p4.z; mat.clearcoatRoughness = p4.w;
    mat.specTrans = p5.x; mat.ior = p5.y; mat.atDistance = p5.z;
    mat.extinction = mk3(p6.x, p6.y, p6.z);
    mat.texA = p7.x; mat.texMR = p7.y; mat.texN = p7.z; mat.texE = p7.w;

    // state.texCoord is read by the four texture lookups only: its u halves sit in the triangle records (48 more bytes from a table the size of
    // the mesh), so they are fetched for hits on materials that have a texture and for no others
    if (TEX && S.num_tex > 0 && ((int)mat.texA >= 0 || (int)mat.texMR >= 0 || (int)mat.texN >= 0 || mat.texE >= 0)) {
        float tu0 = ldg4(tp).w, tu1 = ldg4(tp + 1).w, tu2 = ldg4(tp + 2).w;   // tempTexCoords (closest_hit.glsl:141)
        // state.texCoord = t1 * bary.x + t2 * bary.y + t3 * bary.z
        float tcx = (tu0 * bw + tu1 * bu) + tu2 * bv;
        float tcy = (n1.w * bw + n2.w * bu) + n3.w * bv;
        float tvx = tcx, tvy = 1.0f - tcy;
        if ((int)mat.texA >= 0) {
            float4 c = texArrayLinear<COUNT>(S, tvx, tvy, (int)mat.texA, cnt);
            mat.albedo = mat.albedo * pow3(mk3(c.x, c.y, c.z), 2.2f);
        }
        if ((int)mat.texMR >= 0) {
            float4 c = texArrayLinear<COUNT>(S, tvx, tvy, (int)floorf(mat.texMR + 0.5f), cnt);
            mat.metallic = c.x;
            mat.roughness = gmax(c.y * c.y, 0.001f);
        }
        if ((int)mat.texN >= 0) {
            float4 c = texArrayLinear<COUNT>(S, tvx, tvy, (int)mat.texN, cnt);
            f3 n = normalize(mk3(c.x, c.y, c.z) * 2.0f - mk3(1.0f));
            f3 T, B;
            Onb(s.normal, T, B);
            n = T * n.x + B * n.y + s.normal * n.z;
            s.normal = normalize(n);
            s.ffnormal = dot(s.normal, rdir) <= 0.0f ? s.normal : s.normal * -1.0f;
            Onb(s.normal, s.tangent, s.bitangent);
        }
        if (mat.texE >= 0) {
            float4 c = texArrayLinear<COUNT>(S, tvx, tvy, (int)floorf(mat.texE + 0.5f), cnt);
            mat.emission = pow3(mk3(c.x, c.y, c.z), 2.2f);
        }
    }
    s.eta = dot(rdir, s.normal) < 0.0f ? (1.0f / mat.ior) : mat.ior;   // pathtrace.glsl:122
}

// renderer.glsl:20-62: pixel mapping, RNG seed, tent jitter, thin-lens camera ray for tile-local pixel (lx, ly).
LFD float mapf(float value, float low1, float high1, float low2, float high2) {
    return low2 + fdiv((value - low1) * (high2 - low2), high1 - low1);
}
LFD Ray camera_ray(const DevParams& P, int lx, int ly, int frame, Rng& rng) {
    float resx = (float)P.width, resy = (float)P.height;
    float tcx = ((float)lx + 0.5f) / (float)P.tile_w, tcy = ((float)ly + 0.5f) / (float)P.tile_h;   // TexCoords at the fragment centre
    float xoffset = -1.0f + 2.0f * P.inv_tiles_x * (float)P.tile_x;
    float yoffset = -1.0f + 2.0f * P.inv_tiles_y * (float)P.tile_y;
    float ctx = mapf(tcx, 0.0f, 1.0f, xoffset, xoffset + 2.0f * P.inv_tiles_x);
    float cty = mapf(tcy, 0.0f, 1.0f, yoffset, yoffset + 2.0f * P.inv_tiles_y);
    float fsx = mapf(tcx, 0.0f, 1.0f, P.inv_tiles_x * (float)P.tile_x, P.inv_tiles_x * (float)P.tile_x + P.inv_tiles_x);
    float fsy = mapf(tcy, 0.0f, 1.0f, P.inv_tiles_y * (float)P.tile_y, P.inv_tiles_y * (float)P.tile_y + P.inv_tiles_y);
    float px = fsx * resx, py = fsy * resy;
    rng.x = (unsigned)px; rng.y = (unsigned)py; rng.z = (unsigned)frame; rng.w = (unsigned)px + (unsigned)py;   // InitRNG, globals.glsl:116-120

    float r1 = 2.0f * rnd(rng);
    float r2 = 2.0f * rnd(rng);
    float jx = r1 < 1.0f ? sqrtf(r1) - 1.0f : 1.0f - sqrtf(2.0f - r1);
    float jy = r2 < 1.0f ? sqrtf(r2) - 1.0f : 1.0f - sqrtf(2.0f - r2);
    jx = fdiv(jx, resx * 0.5f);
    jy = fdiv(jy, resy * 0.5f);
    float dx = ctx + jx, dy = cty + jy;
    dy *= fdiv(resy, resx) * P.cam_scale;
    dx *= P.cam_scale;
    f3 right = mk3(P.cam_right[0], P.cam_right[1], P.cam_right[2]), up = mk3(P.cam_up[0], P.cam_up[1], P.cam_up[2]);
    f3 fwd = mk3(P.cam_fwd[0], P.cam_fwd[1], P.cam_fwd[2]), pos = mk3(P.cam_pos[0], P.cam_pos[1], P.cam_pos[2]);
    f3 rayDir = normalize(dx * right + dy * up + fwd);
    f3 focalPoint = P.focal_dist * rayDir;
    float cam_r1 = rnd(rng) * kTWO_PI;
    float cam_r2 = rnd(rng) * P.aperture;
    float sl, cl;
    lf_sincos(cam_r1, sl, cl);
    f3 randomAperturePos = (cl * right + sl * up) * sqrtf(cam_r2);
    f3 finalRayDir = normalize(focalPoint - randomAperturePos);
    Ray r; r.o = pos + randomAperturePos; r.d = finalRayDir;
    return r;
}

// preview_flareon.glsl:22-61: camera ray of fragment (x, y) of the pv_w x pv_h preview viewport.  RNG seeded with
// gl_FragCoord.xy and frame 1 (:24); jitter / screenResolution (:33); d = 2 * TexCoords - 1 (:34); the thin lens only
// under #define USE_DOF (:44-57), its two rand() always drawn (:42-43).
LFD Ray preview_ray(const DevParams& P, int x, int y, Rng& rng) {
    float resx = (float)P.width, resy = (float)P.height;
    float tcx = ((float)x + 0.5f) / (float)P.pv_w, tcy = ((float)y + 0.5f) / (float)P.pv_h;
    rng.x = (unsigned)x; rng.y = (unsigned)y; rng.z = 1u; rng.w = (unsigned)x + (unsigned)y;   // InitRNG(gl_FragCoord.xy, 1)
    float r1 = 2.0f * rnd(rng);
    float r2 = 2.0f * rnd(rng);
    float jx = r1 < 1.0f ? sqrtf(r1) - 1.0f : 1.0f - sqrtf(2.0f - r1);
    float jy = r2 < 1.0f ? sqrtf(r2) - 1.0f : 1.0f - sqrtf(2.0f - r2);
    jx = fdiv(jx, resx);
    jy = fdiv(jy, resy);
    float dx = (2.0f * tcx - 1.0f) + jx, dy = (2.0f * tcy - 1.0f) + jy;
    dy *= fdiv(resy, resx) * P.cam_scale;
    dx *= P.cam_scale;
    f3 right = mk3(P.cam_right[0], P.cam_right[1], P.cam_right[2]), up = mk3(P.cam_up[0], P.cam_up[1], P.cam_up[2]);
    f3 fwd = mk3(P.cam_fwd[0], P.cam_fwd[1], P.cam_fwd[2]), pos = mk3(P.cam_pos[0], P.cam_pos[1], P.cam_pos[2]);
    f3 rayDir = normalize(dx * right + dy * up + fwd);
    f3 focalPoint = P.focal_dist * rayDir;
    float cam_r1 = rnd(rng) * kTWO_PI;
    float cam_r2 = rnd(rng) * P.aperture;
    Ray r;
    if (!P.use_dof) { r.o = pos; r.d = normalize(focalPoint); return r; }
    float sl, cl;
    lf_sincos(cam_r1, sl, cl);
    f3 randomAperturePos = (cl * right + sl * up) * sqrtf(cam_r2);
    r.o = pos + randomAperturePos; r.d = normalize(focalPoint - randomAperturePos);
    return r;
}

}  // namespace lf
