// lf_math.cuh — transcendental functions of the path tracer, written out in plain fp32 arithmetic.
//
// GLSL leaves the precision of sin/cos/pow/exp/log/acos/atan to the implementation (the only runnable reference,
// llvmpipe, evaluates them with its own polynomials), so any accurate implementation is as faithful as another.  These
// use the classic Cephes single-precision kernels (Cody-Waite range reduction + minimax polynomials) restricted to
// + - * / sqrt, floor and bit operations, with every operation separately rounded (the file is compiled -fmad=false).
// Two consequences:
//   * results are bit-reproducible on any IEEE machine: the CPU oracle restates the same formulas and the parity tests
//     compare radiance bit for bit, instead of drowning in last-ulp differences between libm implementations that
//     glass and metal chains amplify;
//   * the code is a fraction of the CUDA math library's (no Payne-Hanek paths), which matters because the shade kernel
//     was instruction-cache bound (156 KB of SASS, `no_instruction` the top stall; profiles/README.md).
// Accuracy: <= 2 ulp for sin/cos on |x| < 8192, <= 1-2 ulp for exp/log, ~1e-6 relative for pow (exp(y log x)),
// <= 2 ulp for acos/atan2.  Arguments on the path are bounded: angles in [0, 2 pi], cosines in [-1, 1], colours in [0, 1].
#pragma once

#include <cuda_runtime.h>

namespace lf {

#ifdef LF_INLINE_MATH
#define LFM __device__ __forceinline__
#else
#define LFM __device__ __noinline__
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

// 2^n as a float for n in [-126, 127]
__device__ __forceinline__ float lf_pow2i(int n) { return __int_as_float((n + 127) << 23); }

// Cephes expf; exact 0 below 2^-126 (no denormal results), +inf above the fp32 range
LFM float lf_exp(float x) {
    if (!(x <= 88.72283905206835f)) return (x != x) ? x : __int_as_float(0x7f800000);
    if (x < -87.33654475055310898657f) return 0.0f;
    float z = floorf(1.44269504088896341f * x + 0.5f);
    int n = (int)z;
    x = (x - z * 0.693359375f) - z * -2.12194440e-4f;
    float xx = x * x;
    float p = (((((1.9875691500E-4f * x + 1.3981999507E-3f) * x + 8.3334519073E-3f) * x + 4.1665795894E-2f) * x + 1.6666665459E-1f) * x
               + 5.0000001201E-1f) * xx + x + 1.0f;
    if (n > 127) return p * lf_pow2i(127) * lf_pow2i(n - 127);
    if (n < -126) return 0.0f;
    return p * lf_pow2i(n);
}

// Cephes logf; log(0) = -inf, log(x < 0) = NaN, denormals are scaled up first
LFM float lf_log(float x) {
    if (!(x > 0.0f)) return (x == 0.0f) ? __int_as_float(0xff800000) : __int_as_float(0x7fc00000);
    if (x == __int_as_float(0x7f800000)) return x;
    int e = 0;
    if (x < 1.17549435e-38f) { x = x * 8388608.0f; e = -23; }
    int bits = __float_as_int(x);
    e += ((bits >> 23) & 0xff) - 126;
    float m = __int_as_float((bits & 0x007fffff) | 0x3f000000);   // mantissa in [0.5, 1)
    if (m < 0.707106781186547524f) { e -= 1; m = m + m - 1.0f; } else { m = m - 1.0f; }
    float z = m * m;
    float y = ((((((((7.0376836292E-2f * m - 1.1514610310E-1f) * m + 1.1676998740E-1f) * m - 1.2420140846E-1f) * m + 1.4249322787E-1f) * m
                  - 1.6668057665E-1f) * m + 2.0000714765E-1f) * m - 2.4999993993E-1f) * m + 3.3333331174E-1f) * m * z;
    float fe = (float)e;
    y = y + -2.12194440e-4f * fe;
    y = y - 0.5f * z;
    return (m + y) + 0.693359375f * fe;
}

// pow for the path's uses (x >= 0): exp(y * log(x)); pow(0, y > 0) = 0
LFM float lf_pow(float x, float y) {
    if (x == 0.0f) return (y > 0.0f) ? 0.0f : ((y == 0.0f) ? 1.0f : __int_as_float(0x7f800000));
    return lf_exp(y * lf_log(x));
}

// Cephes asinf kernel on [0, 0.5]
__device__ __forceinline__ float lf_asin_poly(float a) {
    float z = a * a;
    return ((((4.2163199048E-2f * z + 2.4181311049E-2f) * z + 4.5470025998E-2f) * z + 7.4953002686E-2f) * z + 1.6666752422E-1f) * z * a + a;
}
// acos with the argument clamped to [-1, 1] (the shader's unclamped acos of a unit vector's y can be an ulp outside)
LFM float lf_acos(float x) {
    if (x != x) return x;
    if (x > 1.0f) x = 1.0f;
    if (x < -1.0f) x = -1.0f;
    float a = fabsf(x);
    if (a <= 0.5f) return 1.57079632679489661923f - ((x < 0.0f) ? -lf_asin_poly(a) : lf_asin_poly(a));
    float t = 2.0f * lf_asin_poly(sqrtf(0.5f * (1.0f - a)));
    return (x > 0.0f) ? t : 3.14159265358979323846f - t;
}

// Cephes atanf kernel for t >= 0
__device__ __forceinline__ float lf_atan_pos(float t) {
    float y0;
    if (t > 2.414213562373095f) { y0 = 1.57079632679489661923f; t = -(1.0f / t); }
    else if (t > 0.4142135623730950f) { y0 = 0.785398163397448309616f; t = (t - 1.0f) / (t + 1.0f); }
    else y0 = 0.0f;
    float z = t * t;
    return y0 + ((((8.05374449538e-2f * z - 1.38776856032E-1f) * z + 1.99777106478E-1f) * z - 3.33329491539E-1f) * z * t + t);
}
// atan(y, x) of GLSL: angle of (x, y) in (-pi, pi]
LFM float lf_atan2(float y, float x) {
    if (x != x || y != y) return __int_as_float(0x7fc00000);
    const float PI_F = 3.14159265358979323846f, PIO2_F = 1.57079632679489661923f;
    if (x == 0.0f) return (y > 0.0f) ? PIO2_F : ((y < 0.0f) ? -PIO2_F : 0.0f);
    float a = lf_atan_pos(fabsf(y / x));
    if (x < 0.0f) a = PI_F - a;
    return (y < 0.0f) ? -a : a;
}

}  // namespace lf
