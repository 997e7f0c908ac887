// Host-side double-precision pieces of the drop-in boundary: what EnergyFunctional::setAdjointsF
// (EnergyFunctional.cc:431-489), FrameHessian::setStateZero (FrameHessian.cc:11-42), FullSystem::getNullspaces
// (FullSystem.cc:1711-1760) and EnergyFunctional::orthogonalize (EnergyFunctional.cc:685-717) compute once per
// keyframe window. O(nF^2) work on 68-dimensional objects; the results are uploaded with the frame states.
#pragma once
#include <math.h>
#include <string.h>
#include <vector>
#include "se3_math.cuh"

namespace hostmath {

struct Pose {   // rigid transform as rotation matrix (row-major) + translation
    double R[9], t[3];
};
inline Pose mul(const Pose &a, const Pose &b) { Pose r; se3_mul(a.R, a.t, b.R, b.t, r.R, r.t); return r; }
inline Pose inv(const Pose &a) {
    Pose r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.R[i * 3 + j] = a.R[j * 3 + i];
    double v[3];
    m3_vec(r.R, a.t, v);
    r.t[0] = -v[0]; r.t[1] = -v[1]; r.t[2] = -v[2];
    return r;
}
inline Pose expm(const double a[6]) { Pose r; se3_exp(a, r.R, r.t); return r; }

// SE3 logarithm: rotation vector via the unit quaternion (atan form), translation through V^-1
inline void logm(const Pose &T, double out[6]) {
    const double *m = T.R;
    double qw, qx, qy, qz;
    double tr = m[0] + m[4] + m[8];
    if (tr > 0) {
        double s = sqrt(tr + 1.0) * 2;
        qw = 0.25 * s; qx = (m[7] - m[5]) / s; qy = (m[2] - m[6]) / s; qz = (m[3] - m[1]) / s;
    } else if (m[0] > m[4] && m[0] > m[8]) {
        double s = sqrt(1.0 + m[0] - m[4] - m[8]) * 2;
        qw = (m[7] - m[5]) / s; qx = 0.25 * s; qy = (m[1] + m[3]) / s; qz = (m[2] + m[6]) / s;
    } else if (m[4] > m[8]) {
        double s = sqrt(1.0 + m[4] - m[0] - m[8]) * 2;
        qw = (m[2] - m[6]) / s; qx = (m[1] + m[3]) / s; qy = 0.25 * s; qz = (m[5] + m[7]) / s;
    } else {
        double s = sqrt(1.0 + m[8] - m[0] - m[4]) * 2;
        qw = (m[3] - m[1]) / s; qx = (m[2] + m[6]) / s; qy = (m[5] + m[7]) / s; qz = 0.25 * s;
    }
    double nq = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
    qw /= nq; qx /= nq; qy /= nq; qz /= nq;
    const double n2 = qx * qx + qy * qy + qz * qz, nn = sqrt(n2);
    double k;   // omega = k * q.vec
    if (nn < 1e-10) k = 2.0 / qw - 2.0 * n2 / (qw * qw * qw);
    else if (fabs(qw) < 1e-10) k = (qw > 0 ? M_PI : -M_PI) / nn;
    else k = 2.0 * atan(nn / qw) / nn;
    const double th = k * nn;
    double om[3] = {k * qx, k * qy, k * qz};
    double O[9], O2[9], Vi[9];
    hat3(om, O);
    m3_mul(O, O, O2);
    const double c = (fabs(th) < 1e-10) ? (1.0 / 12.0) : (1.0 - th / (2.0 * tan(th / 2.0))) / (th * th);
    for (int i = 0; i < 9; i++) Vi[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * O[i] + c * O2[i];
    m3_vec(Vi, T.t, out);
    out[3] = om[0]; out[4] = om[1]; out[5] = om[2];
}

// 6x6 adjoint [R, hat(t)R; 0, R], row-major
inline void adjoint(const Pose &T, double A[36]) {
    double tx[9], tR[9];
    hat3(T.t, tx);
    m3_mul(tx, T.R, tR);
    memset(A, 0, sizeof(double) * 36);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            A[i * 6 + j] = T.R[i * 3 + j];
            A[(i + 3) * 6 + j + 3] = T.R[i * 3 + j];
            A[i * 6 + j + 3] = tR[i * 3 + j];
        }
}

// Orthogonal projector onto span(N) (N: dim x k, column-major), dropping directions whose singular value is below
// rel_tol * sigma_max: P = N N^+ symmetrised. One-sided Jacobi rotations on the columns of N.
inline void range_projector(const std::vector<double> &N, int dim, int k, double rel_tol, std::vector<double> &P) {
    std::vector<double> U(N);
    for (int sweep = 0; sweep < 80; sweep++) {
        double worst = 0;
        for (int p = 0; p < k; p++)
            for (int q = p + 1; q < k; q++) {
                double app = 0, aqq = 0, apq = 0;
                for (int i = 0; i < dim; i++) {
                    const double up = U[(size_t) p * dim + i], uq = U[(size_t) q * dim + i];
                    app += up * up; aqq += uq * uq; apq += up * uq;
                }
                if (apq == 0.0 || app == 0.0 || aqq == 0.0) continue;
                const double rel = fabs(apq) / sqrt(app * aqq);
                if (rel > worst) worst = rel;
                const double zeta = (aqq - app) / (2.0 * apq);
                const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
                for (int i = 0; i < dim; i++) {
                    const double up = U[(size_t) p * dim + i], uq = U[(size_t) q * dim + i];
                    U[(size_t) p * dim + i] = cs * up - sn * uq;
                    U[(size_t) q * dim + i] = sn * up + cs * uq;
                }
            }
        if (worst < 1e-15) break;
    }
    std::vector<double> sig(k);
    double smax = 0;
    for (int j = 0; j < k; j++) {
        double s = 0;
        for (int i = 0; i < dim; i++) s += U[(size_t) j * dim + i] * U[(size_t) j * dim + i];
        sig[j] = sqrt(s);
        if (sig[j] > smax) smax = sig[j];
    }
    P.assign((size_t) dim * dim, 0.0);
    for (int j = 0; j < k; j++) {
        if (!(sig[j] > rel_tol * smax)) continue;
        for (int c = 0; c < dim; c++) {
            const double uc = U[(size_t) j * dim + c] / sig[j];
            for (int r = 0; r < dim; r++) P[(size_t) c * dim + r] += (U[(size_t) j * dim + r] / sig[j]) * uc;
        }
    }
    for (int r = 0; r < dim; r++)
        for (int c = r + 1; c < dim; c++) {
            const double v = 0.5 * (P[(size_t) c * dim + r] + P[(size_t) r * dim + c]);
            P[(size_t) c * dim + r] = P[(size_t) r * dim + c] = v;
        }
}

}  // namespace hostmath
