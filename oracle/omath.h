// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, load or execute anything under oracle/.
//
// PARITY PINNED AGAINST THE REFERENCE'S OWN SOURCES, EXCEPT EIGEN / SOPHUS INTERNALS (oracle/ref_pin compiles, unmodified, the reference's Residuals.cc, ImmaturePoint.cc, PointHessian.cc, FrameHessian.cc, FrameFramePrecalc.cc, AccumulatedTopHessian.cc, AccumulatedSCHessian.cc, EnergyFunctional.cc, CoarseTracker.cc, Setting.cc and their headers against stand-in Eigen/Sophus headers and matches the oracle bit for bit on 108 checks; the float SSE / scalar code is pinned completely, the Eigen-expression code structurally - what is added where, in which order - with LDLT, PartialPivLU, JacobiSVD, SE3 exp/log and the product kernels being the oracle's own restatements on both sides; FullSystem.cc's driver loop needs the whole front end and stays restated from the cited lines): the reference (tum-vision/LDSO) ships no golden vectors or tests for
// this path and cannot be compiled here (Eigen3/glog/OpenCV/Pangolin absent). This file is a
// dependency-free CPU restatement of the small dense-math pieces the reference takes from
// Eigen / Sophus:
//   * Sophus::SE3d / SO3d        thirdparty/sophus/se3.hpp:131-139 (Adj), :407-428 (exp),
//                                :560-588 (log); so3.hpp:343-369 (expAndTheta), :491-527 (logAndTheta)
//   * Eigen::LDLT (pivoted)      used at EnergyFunctional.cc:334, CoarseTracker.cc:109
//   * Eigen::JacobiSVD           used at EnergyFunctional.cc:698 (only to build a projector)
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include <limits>

namespace oracle {

// ------------------------------------------------------------------------------------------
// dynamic column-major double matrix (Eigen default layout)
struct MatX {
    int r = 0, c = 0;
    std::vector<double> d;
    MatX() {}
    MatX(int r_, int c_) : r(r_), c(c_), d((size_t) r_ * c_, 0.0) {}
    double &operator()(int i, int j) { return d[(size_t) j * r + i]; }
    double operator()(int i, int j) const { return d[(size_t) j * r + i]; }
    void setZero() { std::fill(d.begin(), d.end(), 0.0); }
    static MatX Zero(int r, int c) { return MatX(r, c); }
};
typedef std::vector<double> VecXd;

inline MatX matmul(const MatX &A, const MatX &B) {
    MatX C(A.r, B.c);
    for (int j = 0; j < B.c; j++)
        for (int k = 0; k < A.c; k++) {
            double b = B(k, j);
            for (int i = 0; i < A.r; i++) C(i, j) += A(i, k) * b;
        }
    return C;
}

inline MatX transpose(const MatX &A) {
    MatX T(A.c, A.r);
    for (int i = 0; i < A.r; i++) for (int j = 0; j < A.c; j++) T(j, i) = A(i, j);
    return T;
}

// ------------------------------------------------------------------------------------------
// Eigen::LDLT<MatXX>::compute + solve restated (lower, unblocked, left-looking, with the
// diagonal pivot search Eigen performs at each step). x = A^{-1} b. A is n x n col-major.
// Inverse of a small dense matrix the way Eigen computes MatrixBase::inverse() for fixed sizes > 4: PartialPivLU
// (unblocked: pivot = largest |entry| of the column at/below the diagonal, rows swapped, multipliers = column / pivot,
// rank-1 update) followed by solve(Identity) (permute, unit-lower forward substitution, upper back substitution).
inline MatX inverse_partial_piv_lu(const MatX &A_in) {
    const int n = A_in.r;
    MatX lu = A_in;
    std::vector<int> perm(n);
    for (int i = 0; i < n; i++) perm[i] = i;
    for (int k = 0; k < n; k++) {
        int piv = k;
        double best = std::fabs(lu(k, k));
        for (int i = k + 1; i < n; i++) if (std::fabs(lu(i, k)) > best) { best = std::fabs(lu(i, k)); piv = i; }
        if (piv != k) {
            for (int j = 0; j < n; j++) std::swap(lu(k, j), lu(piv, j));
            std::swap(perm[k], perm[piv]);
        }
        if (lu(k, k) != 0.0) {
            const double d = lu(k, k);
            for (int i = k + 1; i < n; i++) lu(i, k) /= d;
        }
        for (int j = k + 1; j < n; j++)
            for (int i = k + 1; i < n; i++) lu(i, j) -= lu(i, k) * lu(k, j);
    }
    MatX inv(n, n);
    for (int c = 0; c < n; c++) {
        std::vector<double> x(n);
        for (int i = 0; i < n; i++) x[i] = (perm[i] == c) ? 1.0 : 0.0;       // P * e_c
        for (int i = 0; i < n; i++) for (int j = 0; j < i; j++) x[i] -= lu(i, j) * x[j];
        for (int i = n - 1; i >= 0; i--) {
            for (int j = i + 1; j < n; j++) x[i] -= lu(i, j) * x[j];
            x[i] /= lu(i, i);
        }
        for (int i = 0; i < n; i++) inv(i, c) = x[i];
    }
    return inv;
}

inline VecXd ldlt_solve(const MatX &A_in, const VecXd &b) {
    const int n = A_in.r;
    MatX m = A_in;
    std::vector<int> transp(n);
    std::vector<double> temp(n);
    for (int k = 0; k < n; k++) {
        // largest |diagonal| in the trailing corner
        int big = k;
        double bigv = std::fabs(m(k, k));
        for (int i = k + 1; i < n; i++) {
            double v = std::fabs(m(i, i));
            if (v > bigv) { bigv = v; big = i; }
        }
        transp[k] = big;
        if (k != big) {
            int s = n - big - 1;
            for (int j = 0; j < k; j++) std::swap(m(k, j), m(big, j));
            for (int i = 0; i < s; i++) std::swap(m(big + 1 + i, k), m(big + 1 + i, big));
            std::swap(m(k, k), m(big, big));
            for (int i = k + 1; i < big; i++) {
                double tmp = m(i, k);
                m(i, k) = m(big, i);
                m(big, i) = tmp;
            }
        }
        int rs = n - k - 1;
        if (k > 0) {
            for (int j = 0; j < k; j++) temp[j] = m(j, j) * m(k, j);
            double acc = 0;
            for (int j = 0; j < k; j++) acc += m(k, j) * temp[j];
            m(k, k) -= acc;
            for (int i = 0; i < rs; i++) {
                double a2 = 0;
                for (int j = 0; j < k; j++) a2 += m(k + 1 + i, j) * temp[j];
                m(k + 1 + i, k) -= a2;
            }
        }
        double akk = m(k, k);
        bool valid = std::fabs(akk) > 0.0;
        if (k == 0 && !valid) {
            for (int j = 0; j < n; j++) transp[j] = j;
            break;
        }
        if (rs > 0 && valid) for (int i = 0; i < rs; i++) m(k + 1 + i, k) /= akk;
    }
    // solve: dst = P b ; L^-1 ; D^-1 (pseudo) ; L^-T ; P^T
    VecXd x = b;
    for (int k = 0; k < n; k++) std::swap(x[k], x[transp[k]]);
    for (int i = 0; i < n; i++) {
        double a = x[i];
        for (int j = 0; j < i; j++) a -= m(i, j) * x[j];
        x[i] = a;
    }
    const double tol = std::numeric_limits<double>::min();
    for (int i = 0; i < n; i++) {
        if (std::fabs(m(i, i)) > tol) x[i] /= m(i, i);
        else x[i] = 0;
    }
    for (int i = n - 1; i >= 0; i--) {
        double a = x[i];
        for (int j = i + 1; j < n; j++) a -= m(j, i) * x[j];
        x[i] = a;
    }
    for (int k = n - 1; k >= 0; k--) std::swap(x[k], x[transp[k]]);
    return x;
}

// ------------------------------------------------------------------------------------------
// One-sided (Hestenes) Jacobi SVD of an m x n matrix (m >= n): A = U diag(S) V^T.
// Stands in for Eigen::JacobiSVD<MatXX>(N, ComputeThinU|ComputeThinV) in orthogonalize().
inline void jacobi_svd(const MatX &A, MatX &U, VecXd &S, MatX &V) {
    const int m = A.r, n = A.c;
    U = A;
    V = MatX(n, n);
    for (int i = 0; i < n; i++) V(i, i) = 1;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < m; i++) {
                    alpha += U(i, p) * U(i, p);
                    beta += U(i, q) * U(i, q);
                    gamma += U(i, p) * U(i, q);
                }
                if (gamma == 0) continue;
                off = std::max(off, std::fabs(gamma) / std::sqrt(alpha * beta + 1e-300));
                double zeta = (beta - alpha) / (2.0 * gamma);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < m; i++) {
                    double up = U(i, p), uq = U(i, q);
                    U(i, p) = c * up - s * uq;
                    U(i, q) = s * up + c * uq;
                }
                for (int i = 0; i < n; i++) {
                    double vp = V(i, p), vq = V(i, q);
                    V(i, p) = c * vp - s * vq;
                    V(i, q) = s * vp + c * vq;
                }
            }
        if (off < 1e-15) break;
    }
    S.assign(n, 0.0);
    for (int j = 0; j < n; j++) {
        double nn = 0;
        for (int i = 0; i < m; i++) nn += U(i, j) * U(i, j);
        nn = std::sqrt(nn);
        S[j] = nn;
        if (nn > 0) for (int i = 0; i < m; i++) U(i, j) /= nn;
    }
}

// ------------------------------------------------------------------------------------------
// 3x3 / small fixed helpers (row-major double[9])
struct M3 {
    double m[9];
    double &operator()(int i, int j) { return m[i * 3 + j]; }
    double operator()(int i, int j) const { return m[i * 3 + j]; }
};
struct V3 {
    double v[3];
    double &operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
};
inline M3 m3_identity() { M3 r; memset(r.m, 0, sizeof(r.m)); r(0, 0) = r(1, 1) = r(2, 2) = 1; return r; }
inline M3 m3_mul(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += a(i, k) * b(k, j);
        r(i, j) = s;
    }
    return r;
}
inline V3 m3_mulv(const M3 &a, const V3 &b) {
    V3 r;
    for (int i = 0; i < 3; i++) r[i] = a(i, 0) * b[0] + a(i, 1) * b[1] + a(i, 2) * b[2];
    return r;
}
inline M3 hat(const V3 &w) {  // so3.hpp hat()
    M3 r;
    r(0, 0) = 0; r(0, 1) = -w[2]; r(0, 2) = w[1];
    r(1, 0) = w[2]; r(1, 1) = 0; r(1, 2) = -w[0];
    r(2, 0) = -w[1]; r(2, 1) = w[0]; r(2, 2) = 0;
    return r;
}

// unit quaternion (w,x,y,z) as Sophus::SO3d stores it
struct Quat {
    double w = 1, x = 0, y = 0, z = 0;
};
inline Quat qmul(const Quat &a, const Quat &b) {
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
inline void qnormalize(Quat &q) {  // so3.hpp:196-202
    double l = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    q.w /= l; q.x /= l; q.y /= l; q.z /= l;
}
inline M3 qmat(const Quat &q) {  // Eigen::Quaternion::toRotationMatrix
    M3 r;
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    r(0, 0) = 1 - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
    r(1, 0) = txy + twz; r(1, 1) = 1 - (txx + tzz); r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = 1 - (txx + tyy);
    return r;
}
inline Quat quat_from_mat(const M3 &m) {  // Eigen::Quaternion(Matrix3) (Shepperd)
    Quat q;
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (m(2, 1) - m(1, 2)) * t;
        q.y = (m(0, 2) - m(2, 0)) * t;
        q.z = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (m(k, j) - m(j, k)) * t;
        v[j] = (m(j, i) + m(i, j)) * t;
        v[k] = (m(k, i) + m(i, k)) * t;
        q.x = v[0]; q.y = v[1]; q.z = v[2];
    }
    qnormalize(q);
    return q;
}

// Sophus::SE3d: unit quaternion + translation; tangent = [upsilon(3), omega(3)]
struct SE3 {
    Quat q;
    V3 t{{0, 0, 0}};
    M3 rotationMatrix() const { return qmat(q); }
    const V3 &translation() const { return t; }

    SE3 operator*(const SE3 &o) const {  // se3.hpp operator*= : t += R*o.t ; so3 *= o.so3 (normalised)
        SE3 r;
        V3 rt = m3_mulv(qmat(q), o.t);
        r.t = V3{{t[0] + rt[0], t[1] + rt[1], t[2] + rt[2]}};
        r.q = qmul(q, o.q);
        qnormalize(r.q);
        return r;
    }
    SE3 inverse() const {  // se3.hpp inverse(): invR = so3.inverse(); (invR, invR*(-t))
        SE3 r;
        r.q = Quat{q.w, -q.x, -q.y, -q.z};
        V3 nt{{-t[0], -t[1], -t[2]}};
        r.t = m3_mulv(qmat(r.q), nt);
        return r;
    }
    // se3.hpp:131-139 — Adj = [R, hat(t) R; 0, R]  (row-major 6x6)
    void Adj(double A[36]) const {
        M3 R = rotationMatrix();
        M3 tR = m3_mul(hat(t), R);
        memset(A, 0, sizeof(double) * 36);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                A[i * 6 + j] = R(i, j);
                A[(i + 3) * 6 + (j + 3)] = R(i, j);
                A[i * 6 + (j + 3)] = tR(i, j);
            }
    }
    // so3.hpp:343-369 + se3.hpp:407-428
    static SE3 exp(const double a[6]) {
        V3 omega{{a[3], a[4], a[5]}};
        V3 ups{{a[0], a[1], a[2]}};
        const double eps = 1e-10;  // SophusConstants<double>::epsilon()
        double theta_sq = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
        double theta = std::sqrt(theta_sq);
        double half_theta = 0.5 * theta;
        double imag, real;
        if (theta < eps) {
            double theta_po4 = theta_sq * theta_sq;
            imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
            real = 1.0 - 0.5 * theta_sq + (1.0 / 384.0) * theta_po4;
        } else {
            double s = std::sin(half_theta);
            imag = s / theta;
            real = std::cos(half_theta);
        }
        SE3 r;
        r.q = Quat{real, imag * omega[0], imag * omega[1], imag * omega[2]};
        qnormalize(r.q);  // SO3Group(Quaternion) ctor normalises (so3.hpp:281-283)
        M3 Omega = hat(omega);
        M3 Omega_sq = m3_mul(Omega, Omega);
        M3 V;
        if (theta < eps) {
            V = qmat(r.q);
        } else {
            M3 I = m3_identity();
            double c1 = (1.0 - std::cos(theta)) / theta_sq;
            double c2 = (theta - std::sin(theta)) / (theta_sq * theta);
            for (int i = 0; i < 9; i++) V.m[i] = I.m[i] + c1 * Omega.m[i] + c2 * Omega_sq.m[i];
        }
        r.t = m3_mulv(V, ups);
        return r;
    }
    // so3.hpp:491-527 + se3.hpp:560-588
    void log(double out[6]) const {
        const double eps = 1e-10;
        double sq_n = q.x * q.x + q.y * q.y + q.z * q.z;
        double n = std::sqrt(sq_n);
        double w = q.w;
        double two_atan;
        if (n < eps) {
            double sq_w = w * w;
            two_atan = 2.0 / w - 2.0 * sq_n / (w * sq_w);
        } else {
            if (std::fabs(w) < eps) two_atan = (w > 0 ? M_PI : -M_PI) / n;
            else two_atan = 2.0 * std::atan(n / w) / n;
        }
        double theta = two_atan * n;
        V3 om{{two_atan * q.x, two_atan * q.y, two_atan * q.z}};
        M3 Omega = hat(om);
        M3 O2 = m3_mul(Omega, Omega);
        M3 I = m3_identity();
        M3 Vinv;
        if (std::fabs(theta) < eps) {
            for (int i = 0; i < 9; i++) Vinv.m[i] = I.m[i] - 0.5 * Omega.m[i] + (1. / 12.) * O2.m[i];
        } else {
            double c = (1.0 - theta / (2.0 * std::tan(theta / 2.0))) / (theta * theta);
            for (int i = 0; i < 9; i++) Vinv.m[i] = I.m[i] - 0.5 * Omega.m[i] + c * O2.m[i];
        }
        V3 u = m3_mulv(Vinv, t);
        out[0] = u[0]; out[1] = u[1]; out[2] = u[2];
        out[3] = om[0]; out[4] = om[1]; out[5] = om[2];
    }
    static SE3 fromRt(const double R[9], const double tt[3]) {
        SE3 r;
        M3 m;
        memcpy(m.m, R, sizeof(double) * 9);
        r.q = quat_from_mat(m);
        r.t = V3{{tt[0], tt[1], tt[2]}};
        return r;
    }
};

}  // namespace oracle
