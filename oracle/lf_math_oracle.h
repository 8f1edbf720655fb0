// lf_math_oracle.h — TEST INFRASTRUCTURE.  The oracle's sin/cos/exp/log/pow/acos/atan2.
//
// GLSL leaves the precision of these built-ins to the implementation; the only runnable reference implementation is Mesa
// llvmpipe (tier 1), so this file restates ITS evaluation: plain fp32 (+ - * / sqrt floor and bit operations, every
// operation separately rounded; compiled with -ffp-contract=off), which gives the same bits on any IEEE machine.  Each
// function below is pinned bit for bit against llvmpipe executing the GLSL built-in (oracle/_ref/lp_probe ->
// tests/golden/llvmpipe_builtins.npz, tests/test_oracle_golden.py).  The CUDA side states the same formulas independently
// in lavaframe_b200/csrc/lf_math.cuh, so CUDA, oracle and llvmpipe agree bit for bit on glass/metal chains that amplify
// last-ulp differences into different paths.  Known residue: llvmpipe runs with denormals flushed to zero, this file does not (measured
// to be irrelevant: no denormal is consumed or produced on the path, tests/test_oracle_golden.py::test_no_denormals_on_the_path).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

namespace lfom {

static inline float bits2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline uint32_t f2bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

// sin and cos of one angle: octant reduction with the 3-part pi/4 of Cephes sinf/cosf
static inline void sincos(float x, float& s, float& c) {
    float ax = std::fabs(x);
    int j = (int)(ax * 1.27323954473516f);
    j += (j & 1);
    float y = (float)j;
    float r = ((ax - y * 0.78515625f) - y * 2.4187564849853515625e-4f) - y * 3.77489497744594108e-8f;
    float z = r * r;
    float ps = ((-1.9515295891E-4f * z + 8.3321608736E-3f) * z - 1.6666654611E-1f) * z * r + r;
    float pc = ((2.443315711809948E-005f * z - 1.388731625493765E-003f) * z + 4.166664568298827E-002f) * z * z - 0.5f * z + 1.0f;
    int q = j & 7;
    bool swap = (q == 2) || (q == 6);
    float sv = swap ? pc : ps;
    float cv = swap ? ps : pc;
    if (q == 4 || q == 6) sv = -sv;
    if (q == 2 || q == 4) cv = -cv;
    s = (x < 0.0f) ? -sv : sv;
    c = cv;
}

// tan(x) = sin(x) * (1 / cos(x)) (Mesa lowers tan to sin / cos, and the division to a reciprocal)
static inline float tan(float x) { float s, c; sincos(x, s, c); return s * (1.0f / c); }

// ---- exp / log / pow / acos / atan: the evaluation llvmpipe itself performs, pinned bit for bit with oracle/_ref/lp_probe
// (tests/golden/llvmpipe_builtins.npz).  Mesa's GLSL front end rewrites exp(x) = exp2(x * log2 e), log(x) = log2(x) * ln 2,
// x / y = x * (1 / y); pow stays one instruction that gallivm evaluates as exp2(log2(x) * y); acos / atan are expanded
// into the polynomial expressions of Mesa's builtin_functions.cpp (asin_expr, do_atan, _atan2).  gallivm's exp2 / log2
// (lp_bld_arit.c) are minimax polynomials of degree 5 / 4 evaluated in even / odd halves; llvmpipe's LLVM does not
// contract the multiply-adds.
static inline float mad(float a, float b, float c) { return a * b + c; }

// lp_build_exp2: clamp to [-126.99999, 128], 2^floor(x) by exponent bits, polynomial in fract(x)
static inline float exp2(float x) {
    x = (128.0f < x) ? 128.0f : x;
    x = (-126.99999f > x) ? -126.99999f : x;
    float ip = std::floor(x);
    float fp = x - ip;
    float e = bits2f((uint32_t)((int)ip + 127) << 23);
    float f2 = fp * fp;
    float even = mad(f2, mad(f2, 0.00898934009049466391101f, 0.240153617044375388211f), 1.0f);
    float odd = mad(f2, mad(f2, 0.00187757667519147912699f, 0.0558263180532956664775f), 0.693153073200168932794f);
    return e * mad(odd, fp, even);
}
// lp_build_log2_approx without the edge cases (what pow uses): exponent + y P(y^2), y = (m - 1) / (m + 1); the sign bit is ignored
static inline float log2_raw(float x) {
    uint32_t i = f2bits(x);
    float logexp = (float)((int)((i & 0x7f800000u) >> 23) - 127);
    float mant = bits2f((i & 0x007fffffu) | 0x3f800000u);
    float y = (mant - 1.0f) / (mant + 1.0f);
    float z = y * y;
    float z2 = z * z;
    float even = mad(z2, mad(z2, 0.406718052498846252698f, 0.577440339438736392009f), 2.88539009343309178325f);
    float odd = mad(z2, 0.403343858251329912514f, 0.961791550404184197881f);
    return mad(y, mad(odd, z, even), logexp);
}
// lp_build_log2_safe (the LG2 instruction): + inf, 0 and negative arguments
static inline float log2(float x) {
    float r = log2_raw(x);
    if (x >= bits2f(0x7f800000u)) r = bits2f(0x7f800000u);
    if (x == 0.0f) r = bits2f(0xff800000u);
    if (!(x >= 0.0f)) r = bits2f(0x7fc00000u);   // negative or NaN
    return r;
}
static inline float exp(float x) { return exp2(x * 1.44269504088896340736f); }
static inline float log(float x) { return log2(x) * 0.693147180559945309417f; }
static inline float pow(float x, float y) { return exp2(log2_raw(x) * y); }

static inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
// acos(x) = pi/2 - asin_expr(x, 0.08132463, -0.02363318); NaN outside [-1, 1] (sqrt of a negative number)
static inline float acos(float x) {
    const float PIO2_F = 1.57079632679489661923f;
    float ax = std::fabs(x);
    float as = sign(x) * (PIO2_F - std::sqrt(1.0f - ax) * (PIO2_F + ax * ((0.785398163397448309616f - 1.0f) + ax * (0.08132463f + ax * -0.02363318f))));
    return PIO2_F - as;
}
// do_atan for an argument >= 0
static inline float atan_pos(float a) {
    float mn = a < 1.0f ? a : 1.0f, mx = a > 1.0f ? a : 1.0f;
    float x = mn * (1.0f / mx);
    float t = x * x;
    float r = ((((((((((-0.0121323213173444f * t) + 0.0536813784310406f) * t) - 0.1173503194786851f) * t) + 0.1938924977115610f) * t) - 0.3326756418091246f) * t)
               + 0.9999793128310355f) * x;
    r = r + (a > 1.0f ? 1.0f : 0.0f) * (r * -2.0f + 1.57079632679489661923f);
    return r * sign(a);
}
// atan(y, x) of GLSL as _atan2 expands it: rotate the left half plane by pi/2, tan = |s / t| (1 when |x| == |y|)
static inline float atan2(float y, float x) {
    bool flip = 0.0f >= x;
    float s = flip ? std::fabs(x) : y;
    float t = flip ? y : std::fabs(x);
    float scale = (std::fabs(t) >= 1e18f) ? 0.25f : 1.0f;
    float rcp = 1.0f / (t * scale);
    float sot = (s * scale) * rcp;
    float tn = (std::fabs(x) == std::fabs(y)) ? 1.0f : std::fabs(sot);
    float arc = atan_pos(tn);
    arc = arc + (flip ? 1.0f : 0.0f) * 1.57079632679489661923f;
    float m = y < rcp ? y : rcp;
    return (m < 0.0f) ? -arc : arc;
}

}  // namespace lfom
