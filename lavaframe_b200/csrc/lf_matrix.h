// lf_matrix.h — the two matrix inverses of the instance records, shared by the host re-pack (lf_repack.cpp, g++ -ffp-contract=off)
// and the device-side instance update (lf_tlas.cu, nvcc -fmad=false): one text, the same fp32 operations in the same order on both
// sides, so the records are bit-identical wherever they are made.
#pragma once

#ifdef __CUDACC__
#define LF_HD __host__ __device__
#else
#define LF_HD
#endif

namespace lf {

// inverse(mat4) the way the GLSL built-in is lowered: adjugate from 2x2 sub-factors, times 1/det.
// m[c][r] = column c, row r.
LF_HD inline void inverse4(const float m[4][4], float out[4][4]) {
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
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) out[c][r] = a[c][r] * inv;
}

// inverse(mat3) by cofactors times 1/det.
LF_HD inline void inverse3(const float m[3][3], float out[3][3]) {
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
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) out[c][r] = a[c][r] * inv;
}

}  // namespace lf
