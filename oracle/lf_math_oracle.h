// lf_math_oracle.h — TEST INFRASTRUCTURE.  The oracle's sin/cos/exp/log/pow/acos/atan2.
//
// GLSL leaves the precision of these built-ins to the implementation (llvmpipe, the only runnable reference, uses its
// own polynomials), so the oracle is free to pick any accurate evaluation.  It uses the classic Cephes single-precision
// kernels written in plain fp32 (+ - * / sqrt floor, every operation separately rounded; this file is compiled with
// -ffp-contract=off).  Such formulas give the same bits on any IEEE machine, which is what lets the parity tests compare
// the CUDA kernels with this oracle bit for bit on glass/metal scenes, where libm-vs-libm last-ulp differences would
// otherwise be amplified into different paths.  DESIGN.md lists the formulas; the CUDA side states them independently
// in lavaframe_b200/csrc/lf_math.cuh.
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

static inline float pow2i(int n) { return bits2f((uint32_t)(n + 127) << 23); }

static inline float exp(float x) {
    if (!(x <= 88.72283905206835f)) return (x != x) ? x : bits2f(0x7f800000u);
    if (x < -87.33654475055310898657f) return 0.0f;
    float z = std::floor(1.44269504088896341f * x + 0.5f);
    int n = (int)z;
    x = (x - z * 0.693359375f) - z * -2.12194440e-4f;
    float xx = x * x;
    float p = (((((1.9875691500E-4f * x + 1.3981999507E-3f) * x + 8.3334519073E-3f) * x + 4.1665795894E-2f) * x + 1.6666665459E-1f) * x
               + 5.0000001201E-1f) * xx + x + 1.0f;
    if (n > 127) return p * pow2i(127) * pow2i(n - 127);
    if (n < -126) return 0.0f;
    return p * pow2i(n);
}

static inline float log(float x) {
    if (!(x > 0.0f)) return (x == 0.0f) ? bits2f(0xff800000u) : bits2f(0x7fc00000u);
    if (x == bits2f(0x7f800000u)) return x;
    int e = 0;
    if (x < 1.17549435e-38f) { x = x * 8388608.0f; e = -23; }
    uint32_t b = f2bits(x);
    e += (int)((b >> 23) & 0xff) - 126;
    float m = bits2f((b & 0x007fffffu) | 0x3f000000u);
    if (m < 0.707106781186547524f) { e -= 1; m = m + m - 1.0f; } else { m = m - 1.0f; }
    float z = m * m;
    float y = ((((((((7.0376836292E-2f * m - 1.1514610310E-1f) * m + 1.1676998740E-1f) * m - 1.2420140846E-1f) * m + 1.4249322787E-1f) * m
                  - 1.6668057665E-1f) * m + 2.0000714765E-1f) * m - 2.4999993993E-1f) * m + 3.3333331174E-1f) * m * z;
    float fe = (float)e;
    y = y + -2.12194440e-4f * fe;
    y = y - 0.5f * z;
    return (m + y) + 0.693359375f * fe;
}

static inline float pow(float x, float y) {
    if (x == 0.0f) return (y > 0.0f) ? 0.0f : ((y == 0.0f) ? 1.0f : bits2f(0x7f800000u));
    return exp(y * log(x));
}

static inline float asin_poly(float a) {
    float z = a * a;
    return ((((4.2163199048E-2f * z + 2.4181311049E-2f) * z + 4.5470025998E-2f) * z + 7.4953002686E-2f) * z + 1.6666752422E-1f) * z * a + a;
}
// argument clamped to [-1, 1]
static inline float acos(float x) {
    if (x != x) return x;
    if (x > 1.0f) x = 1.0f;
    if (x < -1.0f) x = -1.0f;
    float a = std::fabs(x);
    if (a <= 0.5f) return 1.57079632679489661923f - ((x < 0.0f) ? -asin_poly(a) : asin_poly(a));
    float t = 2.0f * asin_poly(std::sqrt(0.5f * (1.0f - a)));
    return (x > 0.0f) ? t : 3.14159265358979323846f - t;
}

static inline float atan_pos(float t) {
    float y0;
    if (t > 2.414213562373095f) { y0 = 1.57079632679489661923f; t = -(1.0f / t); }
    else if (t > 0.4142135623730950f) { y0 = 0.785398163397448309616f; t = (t - 1.0f) / (t + 1.0f); }
    else y0 = 0.0f;
    float z = t * t;
    return y0 + ((((8.05374449538e-2f * z - 1.38776856032E-1f) * z + 1.99777106478E-1f) * z - 3.33329491539E-1f) * z * t + t);
}
static inline float atan2(float y, float x) {
    if (x != x || y != y) return bits2f(0x7fc00000u);
    const float PI_F = 3.14159265358979323846f, PIO2_F = 1.57079632679489661923f;
    if (x == 0.0f) return (y > 0.0f) ? PIO2_F : ((y < 0.0f) ? -PIO2_F : 0.0f);
    float a = atan_pos(std::fabs(y / x));
    if (x < 0.0f) a = PI_F - a;
    return (y < 0.0f) ? -a : a;
}

}  // namespace lfom
