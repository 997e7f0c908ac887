// The three dense-algebra hand-offs of oracle/ref_shim/NumTypes.h (TEST INFRASTRUCTURE ONLY): Eigen::LDLT::solve, the PartialPivLU
// inverse and JacobiSVD are not available here (no Eigen on this machine), so the stand-in forwards them to the oracle's restatements
// in oracle/omath.h. Included once by each program that links reference objects (ref_pin/pin_ref.cc, ref_pin/ref_bench.cc).
#pragma once
#include "../omath.h"
// Eigen::LDLT<Mat88 / 77 / 66>::solve as the stand-in forwards it: the oracle's restatement (omath.h) on both sides of the pin
extern "C" void ref_shim_ldlt_solve(int n, const double *A, const double *b, double *x) {
    oracle::MatX M(n, n); oracle::VecXd v(n);
    for (int i = 0; i < n * n; i++) M.d[i] = A[i];
    for (int i = 0; i < n; i++) v[i] = b[i];
    const oracle::VecXd r = oracle::ldlt_solve(M, v);
    for (int i = 0; i < n; i++) x[i] = r[i];
}


// Eigen's PartialPivLU inverse (Mat88::inverse() in marginalizeFrame) and JacobiSVD (orthogonalize), forwarded likewise
extern "C" void ref_shim_inverse_lu(int n, const double *A, double *out) {
    oracle::MatX M(n, n);
    for (int i = 0; i < n * n; i++) M.d[i] = A[i];
    const oracle::MatX I = oracle::inverse_partial_piv_lu(M);
    for (int i = 0; i < n * n; i++) out[i] = I.d[i];
}
extern "C" void ref_shim_jacobi_svd(int m, int n, const double *A, double *U, double *S, double *V) {
    oracle::MatX M(m, n), Uo, Vo; oracle::VecXd So;
    for (int i = 0; i < m * n; i++) M.d[i] = A[i];
    oracle::jacobi_svd(M, Uo, So, Vo);
    memcpy(U, Uo.d.data(), 8 * Uo.d.size()); memcpy(V, Vo.d.data(), 8 * Vo.d.size()); memcpy(S, So.data(), 8 * So.size());
}

