// Sim(3) pose-graph optimisation on the device (BASELINE configs[4]): Map::runPoseGraphOptimization
// (src/Map.cc:75-165) = g2o Gauss-Newton over VertexSim3 / EdgeSim3 (include/internal/PR.h:57-76,151-179) with g2o's numeric
// Jacobians (thirdparty/g2o/g2o/core/base_binary_edge.hpp:131-148, central differences, delta 1e-9, through
// oplus: estimate = Sim3::exp(update) * estimate) and the Sophus Sim3 / RxSO3 / SO3 exp / log of thirdparty/sophus.
//   k_pg_linearize : one WARP per edge; lane = one of the 28 perturbed evaluations (+-delta on each of the 7 + 7 tangent
//                    dimensions) or the nominal one, so the 29 error evaluations of an edge run side by side; Jacobian columns by
//                    shuffle, then the 14x14 block J^T O J, b = -J^T O e and chi2.
//   k_pg_assemble  : one warp per vertex folds its incident edges' diagonal blocks in list order (deterministic) and inverts the
//                    7x7 block (the block-Jacobi preconditioner).
//   k_pg_cg_a / _b : preconditioned conjugate gradients on the block-sparse normal equations, matrix-free over the per-edge
//                    blocks, two launches per iteration, all scalars on the device (partials folded in block order by the last
//                    block to finish). g2o factorises (sparse LDLT); the random long-range loop edges make that fill in heavily,
//                    while the same edges make the graph well connected: CG converges in tens of iterations.
//   k_pg_update    : oplus.
// FP64 throughout (the reference's types); no tensor cores (7x7 blocks, numeric differentiation).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define PG_EPS 1e-10          // SophusConstants<double>::epsilon()
#define PG_DELTA 1e-9         // g2o numeric Jacobian step
#define PG_WARPS 4            // warps (edges / vertices) per CTA

struct PgSim3 { double q[4], t[3]; };      // quaternion (w, x, y, z) with norm = scale; translation

__device__ __forceinline__ void pg_qmul(const double *a, const double *b, double *o) {
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}
__device__ __forceinline__ void pg_qrot(const double *q, const double *v, double *o) {     // s R v, s = |q|
    const double s = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]), is = 1.0 / s;
    const double w = q[0] * is, x = q[1] * is, y = q[2] * is, z = q[3] * is;
    const double tx = 2.0 * (y * v[2] - z * v[1]), ty = 2.0 * (z * v[0] - x * v[2]), tz = 2.0 * (x * v[1] - y * v[0]);
    o[0] = s * (v[0] + w * tx + (y * tz - z * ty));
    o[1] = s * (v[1] + w * ty + (z * tx - x * tz));
    o[2] = s * (v[2] + w * tz + (x * ty - y * tx));
}
__device__ __forceinline__ PgSim3 pg_mul(const PgSim3 &a, const PgSim3 &b) {
    PgSim3 r;
    pg_qmul(a.q, b.q, r.q);
    double rt[3];
    pg_qrot(a.q, b.t, rt);
    for (int i = 0; i < 3; i++) r.t[i] = a.t[i] + rt[i];
    return r;
}
__device__ __forceinline__ PgSim3 pg_inv(const PgSim3 &a) {
    PgSim3 r;
    const double n2 = a.q[0] * a.q[0] + a.q[1] * a.q[1] + a.q[2] * a.q[2] + a.q[3] * a.q[3], in2 = 1.0 / n2;
    r.q[0] = a.q[0] * in2; r.q[1] = -a.q[1] * in2; r.q[2] = -a.q[2] * in2; r.q[3] = -a.q[3] * in2;
    double rt[3];
    pg_qrot(r.q, a.t, rt);
    for (int i = 0; i < 3; i++) r.t[i] = -rt[i];
    return r;
}
// W = A Omega + B Omega^2 + C I (sim3.hpp:609-646)
__device__ __forceinline__ void pg_calc_w(double theta, double sigma, double scale, const double *om, double W[9]) {
    double A, B, C;
    const double th2 = theta * theta;
    if (fabs(sigma) < PG_EPS) {
        C = 1.0;
        if (fabs(theta) < PG_EPS) { A = 0.5; B = 1.0 / 6.0; }
        else { A = (1.0 - cos(theta)) / th2; B = (theta - sin(theta)) / (th2 * theta); }
    } else {
        C = (scale - 1.0) / sigma;
        if (fabs(theta) < PG_EPS) {
            const double s2 = sigma * sigma;
            A = ((sigma - 1.0) * scale + 1.0) / s2;
            B = ((0.5 * sigma * sigma - sigma + 1.0) * scale) / (s2 * sigma);
        } else {
            const double a = scale * sin(theta), b = scale * cos(theta), c = th2 + sigma * sigma;
            A = (a * sigma + (1.0 - b) * theta) / (theta * c);
            B = (C - ((b - 1.0) * sigma + a * theta) / c) * 1.0 / th2;
        }
    }
    const double x = om[0], y = om[1], z = om[2];
    // Omega = hat(om); Omega^2 = om om^T - |om|^2 I
    const double n2 = x * x + y * y + z * z;
    W[0] = B * (x * x - n2) + C;      W[1] = -A * z + B * x * y;       W[2] = A * y + B * x * z;
    W[3] = A * z + B * x * y;         W[4] = B * (y * y - n2) + C;     W[5] = -A * x + B * y * z;
    W[6] = -A * y + B * x * z;        W[7] = A * x + B * y * z;        W[8] = B * (z * z - n2) + C;
}
__device__ __forceinline__ PgSim3 pg_exp(const double a[7]) {
    PgSim3 r;
    const double th2 = a[3] * a[3] + a[4] * a[4] + a[5] * a[5], theta = sqrt(th2), half = 0.5 * theta;
    double imag, real;
    if (theta < PG_EPS) { imag = 0.5 - th2 / 48.0 + th2 * th2 / 3840.0; real = 1.0 - 0.5 * th2 + th2 * th2 / 384.0; }
    else { imag = sin(half) / theta; real = cos(half); }
    const double scale = exp(a[6]);
    r.q[0] = real * scale; r.q[1] = imag * a[3] * scale; r.q[2] = imag * a[4] * scale; r.q[3] = imag * a[5] * scale;
    double W[9];
    pg_calc_w(theta, a[6], scale, a + 3, W);
    for (int i = 0; i < 3; i++) r.t[i] = W[3 * i] * a[0] + W[3 * i + 1] * a[1] + W[3 * i + 2] * a[2];
    return r;
}
__device__ __forceinline__ void pg_log(const PgSim3 &T, double out[7]) {
    const double scale = sqrt(T.q[0] * T.q[0] + T.q[1] * T.q[1] + T.q[2] * T.q[2] + T.q[3] * T.q[3]), is = 1.0 / scale;
    const double sigma = log(scale);
    const double w = T.q[0] * is, x = T.q[1] * is, y = T.q[2] * is, z = T.q[3] * is;
    const double n2 = x * x + y * y + z * z, n = sqrt(n2);
    double f;
    if (n < PG_EPS) f = 2.0 / w - 2.0 * n2 / (w * w * w);
    else if (fabs(w) < PG_EPS) f = (w > 0 ? M_PI : -M_PI) / n;
    else f = 2.0 * atan(n / w) / n;
    const double theta = f * n;
    double om[3] = {f * x, f * y, f * z}, W[9];
    pg_calc_w(theta, sigma, scale, om, W);
    // upsilon = W^-1 t (3x3, cofactors)
    const double c0 = W[4] * W[8] - W[5] * W[7], c1 = W[5] * W[6] - W[3] * W[8], c2 = W[3] * W[7] - W[4] * W[6];
    const double idet = 1.0 / (W[0] * c0 + W[1] * c1 + W[2] * c2);
    const double *t = T.t;
    out[0] = (c0 * t[0] + (W[2] * W[7] - W[1] * W[8]) * t[1] + (W[1] * W[5] - W[2] * W[4]) * t[2]) * idet;
    out[1] = (c1 * t[0] + (W[0] * W[8] - W[2] * W[6]) * t[1] + (W[2] * W[3] - W[0] * W[5]) * t[2]) * idet;
    out[2] = (c2 * t[0] + (W[1] * W[6] - W[0] * W[7]) * t[1] + (W[0] * W[4] - W[1] * W[3]) * t[2]) * idet;
    out[3] = om[0]; out[4] = om[1]; out[5] = om[2]; out[6] = sigma;
}
__device__ __forceinline__ double pg_shfl(double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return __hiloint2double(hi, lo);
}

struct PgGraph {
    int nV, nE, fixed;
    double *q, *t;                      // [nV][4], [nV][3]
    const int *ei, *ej;
    const double *mq, *mt, *info;       // [nE][4], [nE][3], [nE][49] row-major
    double *Hii, *Hij, *Hjj, *bi, *bj, *chi2e;      // per edge: 49, 49, 49, 7, 7, 1
    const int *inc_begin, *inc;         // vertex -> incident (edge * 2 + side) list; side 0: the vertex is the edge's first vertex
    double *D, *Dinv, *b;               // per vertex: 49, 49, 7
    double *x, *r, *z, *p0, *p1, *Ap;   // CG vectors [nV][7]
    double *part;                       // [nblocks] partial sums
    double *scal;                       // [8]: 0 rz, 1 rz_old, 2 pAp, 3 beta, 4 rz0, 5 chi2
    unsigned *counter;
};

__global__ void __launch_bounds__(32 * PG_WARPS) k_pg_linearize(PgGraph g) {
    __shared__ double sJ[PG_WARPS][14][7], sW[PG_WARPS][14][7], sE[PG_WARPS][7];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, e = blockIdx.x * PG_WARPS + wrp;
    if (e >= g.nE) return;
    const int vi = g.ei[e], vj = g.ej[e];
    PgSim3 Vi, Vj, M;
    for (int k = 0; k < 4; k++) { Vi.q[k] = g.q[4 * vi + k]; Vj.q[k] = g.q[4 * vj + k]; M.q[k] = g.mq[4 * e + k]; }
    for (int k = 0; k < 3; k++) { Vi.t[k] = g.t[3 * vi + k]; Vj.t[k] = g.t[3 * vj + k]; M.t[k] = g.mt[3 * e + k]; }
    // lane -> evaluation: 0..6 +delta on i, 7..13 +delta on j, 14..20 -delta on i, 21..27 -delta on j, >= 28 nominal
    if (lane < 28) {
        double up[7] = {0, 0, 0, 0, 0, 0, 0};
        const int d = lane % 7, onj = (lane / 7) & 1;
        up[d] = (lane < 14) ? PG_DELTA : -PG_DELTA;
        const PgSim3 E = pg_exp(up);
        if (onj) Vj = pg_mul(E, Vj); else Vi = pg_mul(E, Vi);
    }
    double err[7];
    pg_log(pg_mul(pg_mul(pg_inv(M), Vi), pg_inv(Vj)), err);
    double J[7], e0[7];
#pragma unroll
    for (int r = 0; r < 7; r++) {
        const double em = pg_shfl(err[r], (lane + 14) & 31);
        J[r] = (err[r] - em) * (1.0 / (2.0 * PG_DELTA));
        e0[r] = pg_shfl(err[r], 28);
    }
    const double *O = g.info + (size_t) 49 * e;
    if (lane < 14) {
        double w[7];
#pragma unroll
        for (int r = 0; r < 7; r++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 7; k++) s = fma(O[r * 7 + k], J[k], s);
            w[r] = s;
        }
        double bb = 0.0;
#pragma unroll
        for (int r = 0; r < 7; r++) { sJ[wrp][lane][r] = J[r]; sW[wrp][lane][r] = w[r]; bb = fma(w[r], e0[r], bb); }
        (lane < 7 ? g.bi : g.bj)[(size_t) 7 * e + (lane % 7)] = -bb;       // b = -J^T O e (base_binary_edge.hpp:63-66)
    }
    if (lane == 28) {
        double c = 0.0;
        for (int r = 0; r < 7; r++) { double s = 0.0; for (int k = 0; k < 7; k++) s = fma(O[r * 7 + k], e0[k], s); c = fma(e0[r], s, c); sE[wrp][r] = e0[r]; }
        g.chi2e[e] = c;
    }
    __syncwarp();
    for (int o = lane; o < 196; o += 32) {       // H[a][c] = J[:,a] . (O J[:,c])
        const int a = o / 14, c = o % 14;
        if (a >= 7 && c < 7) continue;             // H_ji = H_ij^T
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < 7; r++) s = fma(sJ[wrp][a][r], sW[wrp][c][r], s);
        if (a < 7 && c < 7) g.Hii[(size_t) 49 * e + a * 7 + c] = s;
        else if (a < 7) g.Hij[(size_t) 49 * e + a * 7 + (c - 7)] = s;
        else g.Hjj[(size_t) 49 * e + (a - 7) * 7 + (c - 7)] = s;
    }
}

// fold the partial sums of a launch in block order (last block to finish) into scal[slot]
__device__ __forceinline__ void pg_block_sum_to_scalar(double v, const PgGraph &g, int slot, double *s_red, bool also_beta) {
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[wrp] = v;
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < PG_WARPS; w++) s += s_red[w];
        g.part[blockIdx.x] = s;
        __threadfence();
        last = atomicAdd(g.counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        double s = 0.0;
        for (unsigned b = 0; b < gridDim.x; b++) s += ((volatile double *) g.part)[b];
        if (also_beta) { g.scal[1] = g.scal[0]; g.scal[3] = (g.scal[0] != 0.0) ? s / g.scal[0] : 0.0; }
        g.scal[slot] = s;
        *g.counter = 0u;
        __threadfence();
    }
}

// per vertex: D = sum of the incident diagonal blocks, b likewise, Dinv = D^-1 (Gauss-Jordan with partial pivoting); r = b, z = Dinv r,
// x = 0, p = z; partial rz
__global__ void __launch_bounds__(32 * PG_WARPS) k_pg_assemble(PgGraph g) {
    __shared__ double s_red[PG_WARPS];
    __shared__ double sD[PG_WARPS][49], sI[PG_WARPS][49];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, v = blockIdx.x * PG_WARPS + wrp;
    double rz = 0.0;
    if (v < g.nV) {
        const int i0 = g.inc_begin[v], i1 = g.inc_begin[v + 1];
        const bool fix = v == g.fixed;
        for (int o = lane; o < 56; o += 32) {
            double s = 0.0;
            for (int k = i0; k < i1; k++) {
                const int es = g.inc[k], e = es >> 1, side = es & 1;
                s += (o < 49) ? (side ? g.Hjj : g.Hii)[(size_t) 49 * e + o] : (side ? g.bj : g.bi)[(size_t) 7 * e + (o - 49)];
            }
            if (fix) s = (o < 49) ? ((o / 7 == o % 7) ? 1.0 : 0.0) : 0.0;
            if (o < 49) { g.D[(size_t) 49 * v + o] = s; sD[wrp][o] = s; sI[wrp][o] = (o / 7 == o % 7) ? 1.0 : 0.0; }
            else g.b[(size_t) 7 * v + (o - 49)] = s;
        }
        __syncwarp();
        if (lane == 0) {
            double *A = sD[wrp], *I = sI[wrp];
            for (int c = 0; c < 7; c++) {
                int pv = c; double best = fabs(A[c * 7 + c]);
                for (int r = c + 1; r < 7; r++) if (fabs(A[r * 7 + c]) > best) { best = fabs(A[r * 7 + c]); pv = r; }
                if (pv != c) for (int k = 0; k < 7; k++) { double t1 = A[c * 7 + k]; A[c * 7 + k] = A[pv * 7 + k]; A[pv * 7 + k] = t1; t1 = I[c * 7 + k]; I[c * 7 + k] = I[pv * 7 + k]; I[pv * 7 + k] = t1; }
                const double ip = 1.0 / A[c * 7 + c];
                for (int k = 0; k < 7; k++) { A[c * 7 + k] *= ip; I[c * 7 + k] *= ip; }
                for (int r = 0; r < 7; r++) if (r != c) { const double f = A[r * 7 + c]; for (int k = 0; k < 7; k++) { A[r * 7 + k] -= f * A[c * 7 + k]; I[r * 7 + k] -= f * I[c * 7 + k]; } }
            }
        }
        __syncwarp();
        for (int o = lane; o < 49; o += 32) g.Dinv[(size_t) 49 * v + o] = sI[wrp][o];
        __syncwarp();
        if (lane < 7) {
            double zz = 0.0;
            for (int k = 0; k < 7; k++) zz = fma(sI[wrp][lane * 7 + k], g.b[(size_t) 7 * v + k], zz);
            const double rr = g.b[(size_t) 7 * v + lane];
            g.x[(size_t) 7 * v + lane] = 0.0; g.r[(size_t) 7 * v + lane] = rr; g.z[(size_t) 7 * v + lane] = zz; g.p0[(size_t) 7 * v + lane] = zz;
            rz = rr * zz;
        }
    }
    pg_block_sum_to_scalar(rz, g, 0, s_red, false);
}

// CG half-step A: p_new = z + beta p_old (first: p = z), Ap = H p_new (matrix-free over the incident edges), partial p.Ap.
// pin = previous p, pout = this iteration's p; a neighbour's p_new is recomputed from its z and old p (no grid-wide barrier).
__global__ void __launch_bounds__(32 * PG_WARPS) k_pg_cg_a(PgGraph g, const double *pin, double *pout, int first) {
    __shared__ double s_red[PG_WARPS];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, v = blockIdx.x * PG_WARPS + wrp;
    const double beta = first ? 0.0 : g.scal[3];
    double pAp = 0.0;
    if (v < g.nV && v != g.fixed) {
        const int i0 = g.inc_begin[v], i1 = g.inc_begin[v + 1];
        // lane (row r = lane % 7, quarter k = lane / 7 for lanes < 28) -> partial row products; folded by shuffles
        const int r = lane % 7, part = lane / 7;
        double pv[7];
#pragma unroll
        for (int k = 0; k < 7; k++) pv[k] = first ? g.z[(size_t) 7 * v + k] : fma(beta, pin[(size_t) 7 * v + k], g.z[(size_t) 7 * v + k]);
        double acc = 0.0;
        if (lane < 28) {
            if (part == 0) for (int k = 0; k < 7; k++) acc = fma(g.D[(size_t) 49 * v + r * 7 + k], pv[k], acc);
            for (int q = i0 + part; q < i1; q += 4) {
                const int es = g.inc[q], e = es >> 1, side = es & 1;
                const int o = side ? g.ei[e] : g.ej[e];
                if (o == g.fixed) continue;
                const double *Hb = g.Hij + (size_t) 49 * e;
                for (int k = 0; k < 7; k++) {
                    const double po = first ? g.z[(size_t) 7 * o + k] : fma(beta, pin[(size_t) 7 * o + k], g.z[(size_t) 7 * o + k]);
                    acc = fma(side ? Hb[k * 7 + r] : Hb[r * 7 + k], po, acc);          // side 1: H_ji = H_ij^T
                }
            }
        }
        // fold the four quarters: lanes r, r+7, r+14, r+21
        double s = acc;
        s += pg_shfl(acc, (lane + 7) & 31);
        const double s2 = pg_shfl(acc, (lane + 14) & 31) + pg_shfl(acc, (lane + 21) & 31);
        s += s2;
        if (lane < 7) {
            g.Ap[(size_t) 7 * v + lane] = s;
            pout[(size_t) 7 * v + lane] = pv[lane];
            pAp = pv[lane] * s;
        }
    }
    pg_block_sum_to_scalar(pAp, g, 2, s_red, false);
}
// CG half-step B: alpha = rz / pAp; x += alpha p; r -= alpha Ap; z = Dinv r; partial r.z -> rz (and beta = rz_new / rz_old)
__global__ void __launch_bounds__(32 * PG_WARPS) k_pg_cg_b(PgGraph g, const double *p) {
    __shared__ double s_red[PG_WARPS];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, v = blockIdx.x * PG_WARPS + wrp;
    const double pAp = g.scal[2], alpha = (pAp != 0.0) ? g.scal[0] / pAp : 0.0;
    double rz = 0.0;
    if (v < g.nV && v != g.fixed) {
        double rn = 0.0;
        if (lane < 7) {
            g.x[(size_t) 7 * v + lane] = fma(alpha, p[(size_t) 7 * v + lane], g.x[(size_t) 7 * v + lane]);
            rn = fma(-alpha, g.Ap[(size_t) 7 * v + lane], g.r[(size_t) 7 * v + lane]);
            g.r[(size_t) 7 * v + lane] = rn;
        }
        double zz = 0.0;
#pragma unroll
        for (int k = 0; k < 7; k++) { const double rk = pg_shfl(rn, k); if (lane < 7) zz = fma(g.Dinv[(size_t) 49 * v + lane * 7 + k], rk, zz); }
        if (lane < 7) { g.z[(size_t) 7 * v + lane] = zz; rz = rn * zz; }
    }
    pg_block_sum_to_scalar(rz, g, 0, s_red, true);
}
// oplus (PR.h:71-75): estimate = Sim3::exp(dx) * estimate; chi2 of the linearisation folded in edge order by one block
__global__ void k_pg_update(PgGraph g) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= g.nV || v == g.fixed) return;
    double up[7];
    for (int k = 0; k < 7; k++) up[k] = g.x[(size_t) 7 * v + k];
    PgSim3 V;
    for (int k = 0; k < 4; k++) V.q[k] = g.q[4 * v + k];
    for (int k = 0; k < 3; k++) V.t[k] = g.t[3 * v + k];
    V = pg_mul(pg_exp(up), V);
    for (int k = 0; k < 4; k++) g.q[4 * v + k] = V.q[k];
    for (int k = 0; k < 3; k++) g.t[3 * v + k] = V.t[k];
}
__global__ void k_pg_chi2(PgGraph g, double *out) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int e = threadIdx.x; e < g.nE; e += 256) s += g.chi2e[e];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int) threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) *out = sh[0];
}
