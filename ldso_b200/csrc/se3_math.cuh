// Small double-precision SE(3) helpers usable on host and device (row-major 3x3 + translation).
// Stand in for the Sophus::SE3d calls on the path (thirdparty/sophus/se3.hpp:131-139 Adj, :407-428 exp,
// :560-588 log of the reference): the rotation is kept as a matrix instead of a unit quaternion.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

HD void m3_mul(const double *A, const double *B, double *C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
HD void m3_mulT(const double *A, const double *B, double *C) {  // C = A * B^T
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j * 3] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
}
HD void m3_vec(const double *A, const double *v, double *o) {
    for (int i = 0; i < 3; i++) o[i] = A[i * 3] * v[0] + A[i * 3 + 1] * v[1] + A[i * 3 + 2] * v[2];
}
HD void hat3(const double *w, double *O) {
    O[0] = 0; O[1] = -w[2]; O[2] = w[1];
    O[3] = w[2]; O[4] = 0; O[5] = -w[0];
    O[6] = -w[1]; O[7] = w[0]; O[8] = 0;
}

// SE3::exp(a), a = [upsilon(3), omega(3)]  ->  R (3x3), t = V * upsilon
HD void se3_exp(const double *a, double *R, double *t) {
    const double *om = a + 3;
    const double th2 = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
    const double th = sqrt(th2);
    double O[9], O2[9];
    hat3(om, O);
    m3_mul(O, O, O2);
    double A, B, C;   // R = I + A*O + B*O2 ; V = I + B*O + C*O2
    if (th < 1e-3) {
        A = 1.0 - th2 / 6.0 + th2 * th2 / 120.0;
        B = 0.5 - th2 / 24.0 + th2 * th2 / 720.0;
        C = 1.0 / 6.0 - th2 / 120.0 + th2 * th2 / 5040.0;
    } else {
        A = sin(th) / th;
        B = (1.0 - cos(th)) / th2;
        C = (th - sin(th)) / (th2 * th);
    }
    double V[9];
    for (int i = 0; i < 9; i++) {
        const double I = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        R[i] = I + A * O[i] + B * O2[i];
        V[i] = I + B * O[i] + C * O2[i];
    }
    m3_vec(V, a, t);
}

// T_ab = T_a * T_b
HD void se3_mul(const double *Ra, const double *ta, const double *Rb, const double *tb, double *R, double *t) {
    m3_mul(Ra, Rb, R);
    double r[3];
    m3_vec(Ra, tb, r);
    t[0] = ta[0] + r[0]; t[1] = ta[1] + r[1]; t[2] = ta[2] + r[2];
}
// T_a * T_b^{-1}
HD void se3_mul_inv(const double *Ra, const double *ta, const double *Rb, const double *tb, double *R, double *t) {
    m3_mulT(Ra, Rb, R);
    double r[3];
    m3_vec(R, tb, r);
    t[0] = ta[0] - r[0]; t[1] = ta[1] - r[1]; t[2] = ta[2] - r[2];
}

// float 3x3 inverse the way Eigen does it for Mat33f (cofactors * 1/det), used by FrameFramePrecalc::Set
HD void m33f_inverse(const float *m, float *inv) {
    float c00 = m[4] * m[8] - m[5] * m[7];
    float c01 = m[5] * m[6] - m[3] * m[8];
    float c02 = m[3] * m[7] - m[4] * m[6];
    float det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    float invdet = 1.0f / det;
    inv[0] = c00 * invdet; inv[3] = c01 * invdet; inv[6] = c02 * invdet;
    inv[1] = (m[2] * m[7] - m[1] * m[8]) * invdet;
    inv[4] = (m[0] * m[8] - m[2] * m[6]) * invdet;
    inv[7] = (m[1] * m[6] - m[0] * m[7]) * invdet;
    inv[2] = (m[1] * m[5] - m[2] * m[4]) * invdet;
    inv[5] = (m[2] * m[3] - m[0] * m[5]) * invdet;
    inv[8] = (m[0] * m[4] - m[1] * m[3]) * invdet;
}
HD void m33f_mul(const float *A, const float *B, float *C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float s = A[i * 3 + 0] * B[0 * 3 + j];
            s += A[i * 3 + 1] * B[1 * 3 + j];
            s += A[i * 3 + 2] * B[2 * 3 + j];
            C[i * 3 + j] = s;
        }
}
