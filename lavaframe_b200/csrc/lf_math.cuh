// lf_math.cuh — transcendental functions of the path tracer, written out in plain fp32 arithmetic.
//
// GLSL leaves the precision of sin/cos/tan/pow/exp/log/acos/atan to the implementation, and the only runnable reference
// implementation is Mesa llvmpipe (the tier-1 oracle), so these functions restate ITS evaluation: Cephes-style sincos,
// gallivm's exp2 / log2 polynomials, Mesa's GLSL-level expansions of acos / atan.  They use only + - * / sqrt, floor and
// bit operations, every operation separately rounded (the file is compiled -fmad=false), so:
//   * results are bit-reproducible on any IEEE machine: the CPU oracle states the same formulas (oracle/lf_math_oracle.h)
//     and is itself pinned bit for bit against llvmpipe executing the GLSL built-ins (tests/golden/llvmpipe_builtins.npz),
//     so CUDA, oracle and the reference on llvmpipe agree bit for bit instead of drowning in last-ulp differences that
//     glass and metal chains amplify;
//   * the code is a fraction of the CUDA math library's (no Payne-Hanek paths), which matters because the shade kernel
//     was instruction-cache bound (156 KB of SASS, `no_instruction` the top stall; profiles/README.md).
// Arguments on the path are bounded: angles in [0, 2 pi], cosines in [-1, 1], colours in [0, 1].
#pragma once

#include <cuda_runtime.h>

namespace lf {

// Inlined by default: with llvmpipe's short polynomials the shade kernel is no longer instruction-cache bound (no_instruction
// stall 8.8 -> 1.4 per issue) and inlining measures 3 % faster in k_shade (profiles/r1_experiments/ab_l2_persist_inline_math.txt);
// the Cephes-sized functions of the earlier round were faster out of line.
#ifdef LF_OUTLINE_MATH
#define LFM __device__ __noinline__
#else
#define LFM __device__ __forceinline__
#endif

// sin and cos of the same angle (Cephes sinf/cosf: octant reduction with a 3-part pi/4, degree-7/8 polynomials)
LFM void lf_sincos(float x, float& s, float& c) {
    float ax = fabsf(x);
    int j = (int)(ax * 1.27323954473516f);          // 4/pi
    j += (j & 1);                                    // map zeros to origin
    float y = (float)j;
    float r = ((ax - y * 0.78515625f) - y * 2.4187564849853515625e-4f) - y * 3.77489497744594108e-8f;
    float z = r * r;
    float ps = ((-1.9515295891E-4f * z + 8.3321608736E-3f) * z - 1.6666654611E-1f) * z * r + r;
    float pc = ((2.443315711809948E-005f * z - 1.388731625493765E-003f) * z + 4.166664568298827E-002f) * z * z - 0.5f * z + 1.0f;
    int q = j & 7;                                   // j is even: octant pair 0, 2, 4, 6
    bool swap = (q == 2) || (q == 6);
    float sv = swap ? pc : ps;
    float cv = swap ? ps : pc;
    if (q == 4 || q == 6) sv = -sv;                  // sin negative in the third/fourth quadrant
    if (q == 2 || q == 4) cv = -cv;                  // cos negative in the second/third quadrant
    s = (x < 0.0f) ? -sv : sv;
    c = cv;
}

// tan(x) = sin(x) * (1 / cos(x)): Mesa lowers tan to sin / cos and every division to a multiplication by the reciprocal
LFM float lf_tan(float x) { float s, c; lf_sincos(x, s, c); return s * (1.0f / c); }

// ---- exp / log / pow / acos / atan exactly as llvmpipe evaluates the GLSL built-ins.  Mesa's GLSL front end rewrites
// exp(x) = exp2(x * log2 e) and log(x) = log2(x) * ln 2; pow stays one instruction that gallivm evaluates as
// exp2(log2(x) * y); acos / atan(y, x) are expanded into the polynomial expressions of Mesa's builtin_functions.cpp.
// gallivm's exp2 / log2 are minimax polynomials of degree 5 / 4 evaluated in even / odd halves, no contraction.
__device__ __forceinline__ float lf_mad(float a, float b, float c) { return a * b + c; }   // -fmad=false: two roundings

// exp2: clamp to [-126.99999, 128], 2^floor(x) from exponent bits, polynomial in fract(x)
__device__ __forceinline__ float lf_exp2(float x) {
    x = (128.0f < x) ? 128.0f : x;
    x = (-126.99999f > x) ? -126.99999f : x;
    float ip = floorf(x);
    float fp = x - ip;
    float e = __int_as_float(((int)ip + 127) << 23);
    float f2 = fp * fp;
    float even = lf_mad(f2, lf_mad(f2, 0.00898934009049466391101f, 0.240153617044375388211f), 1.0f);
    float odd = lf_mad(f2, lf_mad(f2, 0.00187757667519147912699f, 0.0558263180532956664775f), 0.693153073200168932794f);
    return e * lf_mad(odd, fp, even);
}
// log2 without edge cases (pow's): exponent + y P(y^2), y = (m - 1) / (m + 1); the sign bit is ignored
__device__ __forceinline__ float lf_log2_raw(float x) {
    int i = __float_as_int(x);
    float logexp = (float)(((i & 0x7f800000) >> 23) - 127);
    float mant = __int_as_float((i & 0x007fffff) | 0x3f800000);
    float y = (mant - 1.0f) / (mant + 1.0f);
    float z = y * y;
    float z2 = z * z;
    float even = lf_mad(z2, lf_mad(z2, 0.406718052498846252698f, 0.577440339438736392009f), 2.88539009343309178325f);
    float odd = lf_mad(z2, 0.403343858251329912514f, 0.961791550404184197881f);
    return lf_mad(y, lf_mad(odd, z, even), logexp);
}
LFM float lf_exp(float x) { return lf_exp2(x * 1.44269504088896340736f); }
// log2 with the LG2 instruction's edge cases: +inf, 0 -> -inf, negative or NaN -> NaN; log(x) = log2(x) * ln 2
LFM float lf_log2(float x) {
    float r = lf_log2_raw(x);
    if (x >= __int_as_float(0x7f800000)) r = __int_as_float(0x7f800000);
    if (x == 0.0f) r = __int_as_float(0xff800000);
    if (!(x >= 0.0f)) r = __int_as_float(0x7fc00000);
    return r;
}
LFM float lf_log(float x) { return lf_log2(x) * 0.693147180559945309417f; }
LFM float lf_pow(float x, float y) { return lf_exp2(lf_log2_raw(x) * y); }

__device__ __forceinline__ float lf_sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
// acos(x) = pi/2 - sign(x) * (pi/2 - sqrt(1 - |x|) * (pi/2 + |x| * (pi/4 - 1 + |x| * (0.08132463 + |x| * -0.02363318))))
LFM float lf_acos(float x) {
    const float PIO2_F = 1.57079632679489661923f;
    float ax = fabsf(x);
    float as = lf_sign(x) * (PIO2_F - sqrtf(1.0f - ax) * (PIO2_F + ax * ((0.785398163397448309616f - 1.0f) + ax * (0.08132463f + ax * -0.02363318f))));
    return PIO2_F - as;
}
// Mesa's do_atan for an argument >= 0
__device__ __forceinline__ float lf_atan_pos(float a) {
    float mn = a < 1.0f ? a : 1.0f, mx = a > 1.0f ? a : 1.0f;
    float x = mn * (1.0f / mx);
    float t = x * x;
    float r = ((((((((((-0.0121323213173444f * t) + 0.0536813784310406f) * t) - 0.1173503194786851f) * t) + 0.1938924977115610f) * t) - 0.3326756418091246f) * t)
               + 0.9999793128310355f) * x;
    r = r + (a > 1.0f ? 1.0f : 0.0f) * (r * -2.0f + 1.57079632679489661923f);
    return r * lf_sign(a);
}
// atan(y, x) of GLSL as Mesa's _atan2 expands it
LFM float lf_atan2(float y, float x) {
    bool flip = 0.0f >= x;
    float s = flip ? fabsf(x) : y;
    float t = flip ? y : fabsf(x);
    float scale = (fabsf(t) >= 1e18f) ? 0.25f : 1.0f;
    float rcp = 1.0f / (t * scale);
    float sot = (s * scale) * rcp;
    float tn = (fabsf(x) == fabsf(y)) ? 1.0f : fabsf(sot);
    float arc = lf_atan_pos(tn);
    arc = arc + (flip ? 1.0f : 0.0f) * 1.57079632679489661923f;
    float m = y < rcp ? y : rcp;
    return (m < 0.0f) ? -arc : arc;
}

}  // namespace lf
