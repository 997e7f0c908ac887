// Coarse direct tracker kernels (src/frontend/CoarseTracker.cc of the reference).
//   trk_point   : one reference point through CoarseTracker::calcRes (:475-548) and, for the terms calcRes would
//                 append to buf_warped_*, the 45 weighted products of calcGSSSE / Accumulator9::updateSSE_eighted
//                 (:591-614, MatrixAccumulators.h:1250-1369) — "compute always", no compaction pass.
//   k_trk_eval  : multi-CTA evaluation of one pose (piecewise API, parity tests).
//   k_trk_track : CoarseTracker::trackNewestCoarse (:61-217) with the whole LM loop resident in ONE CTA: block-wide
//                 shuffle reductions, the 8x8 pivoted LDLT and SE3::exp on thread 0, no host round trips.
#pragma once
#include "common.cuh"
#include "se3_math.cuh"

#define TRK_NACC 52      // 45 H entries + E, nE, nSat, nWarped, sumT, sumRT, sumNum
#define TRK_E 45
#define TRK_NE 46
#define TRK_NSAT 47
#define TRK_NW 48
#define TRK_ST 49
#define TRK_SRT 50
#define TRK_SNUM 51

struct TrkLevel {
    const float *pc_u, *pc_v, *pc_idepth, *pc_color;
    int n;
    const float4 *img;     // new frame, this level
    int w, h;
    float fx, fy, cx, cy;
    float Ki[9];
};

struct TrkPose {         // everything calcRes/calcGSSSE derive from (refToNew, aff_g2l)
    float RKi[9], t[3];
    float affLL0, affLL1;  // fromToVecExposure(lastRef, newFrame, lastRef_aff_g2l, aff_g2l)
    float b0;              // lastRef_aff_g2l.b
    float cutoffTH, maxEnergy, huberTH;
};

__device__ __forceinline__ void trk_point(const TrkLevel &L, const TrkPose &P, int i, bool lvl0, float *acc) {
    const float id = L.pc_idepth[i], x = L.pc_u[i], y = L.pc_v[i];
    float pt0 = P.RKi[0] * x; pt0 += P.RKi[1] * y; pt0 += P.RKi[2] * 1.f; pt0 = pt0 + P.t[0] * id;
    float pt1 = P.RKi[3] * x; pt1 += P.RKi[4] * y; pt1 += P.RKi[5] * 1.f; pt1 = pt1 + P.t[1] * id;
    float pt2 = P.RKi[6] * x; pt2 += P.RKi[7] * y; pt2 += P.RKi[8] * 1.f; pt2 = pt2 + P.t[2] * id;
    const float u = pt0 / pt2, v = pt1 / pt2;
    const float Ku = L.fx * u + L.cx, Kv = L.fy * v + L.cy;
    const float new_idepth = id / pt2;

    if (lvl0 && (i % 32 == 0)) {   // flow indicators (:487-517)
        float k0 = L.Ki[0] * x; k0 += L.Ki[1] * y; k0 += L.Ki[2] * 1.f;
        float k1 = L.Ki[3] * x; k1 += L.Ki[4] * y; k1 += L.Ki[5] * 1.f;
        float k2 = L.Ki[6] * x; k2 += L.Ki[7] * y; k2 += L.Ki[8] * 1.f;
        const float a0 = k0 + P.t[0] * id, a1 = k1 + P.t[1] * id, a2 = k2 + P.t[2] * id;
        const float KuT = L.fx * (a0 / a2) + L.cx, KvT = L.fy * (a1 / a2) + L.cy;
        const float b0_ = k0 - P.t[0] * id, b1_ = k1 - P.t[1] * id, b2_ = k2 - P.t[2] * id;
        const float KuT2 = L.fx * (b0_ / b2_) + L.cx, KvT2 = L.fy * (b1_ / b2_) + L.cy;
        float c0 = P.RKi[0] * x; c0 += P.RKi[1] * y; c0 += P.RKi[2] * 1.f; c0 = c0 - P.t[0] * id;
        float c1 = P.RKi[3] * x; c1 += P.RKi[4] * y; c1 += P.RKi[5] * 1.f; c1 = c1 - P.t[1] * id;
        float c2 = P.RKi[6] * x; c2 += P.RKi[7] * y; c2 += P.RKi[8] * 1.f; c2 = c2 - P.t[2] * id;
        const float Ku3 = L.fx * (c0 / c2) + L.cx, Kv3 = L.fy * (c1 / c2) + L.cy;
        acc[TRK_ST] += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
        acc[TRK_ST] += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
        acc[TRK_SRT] += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
        acc[TRK_SRT] += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
        acc[TRK_SNUM] += 2;
    }
    if (!(Ku > 2 && Kv > 2 && Ku < L.w - 3 && Kv < L.h - 3 && new_idepth > 0)) return;

    const float refColor = L.pc_color[i];
    const int ix = (int) Ku, iy = (int) Kv;
    const float dx = Ku - ix, dy = Kv - iy, dxdy = dx * dy;
    const float4 *bp = L.img + ix + iy * L.w;
    const float4 c00 = __ldg(bp), c10 = __ldg(bp + 1), c01 = __ldg(bp + L.w), c11 = __ldg(bp + L.w + 1);
    const float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
    const float hit0 = w11 * c11.x + w01 * c01.x + w10 * c10.x + w00 * c00.x;
    const float hit1 = w11 * c11.y + w01 * c01.y + w10 * c10.y + w00 * c00.y;
    const float hit2 = w11 * c11.z + w01 * c01.z + w10 * c10.z + w00 * c00.z;
    if (!isfinite(hit0)) return;
    const float residual = hit0 - (P.affLL0 * refColor + P.affLL1);
    const float hw = fabsf(residual) < P.huberTH ? 1.f : P.huberTH / fabsf(residual);
    if (fabsf(residual) > P.cutoffTH) {
        acc[TRK_E] += P.maxEnergy;
        acc[TRK_NE] += 1.f;
        acc[TRK_NSAT] += 1.f;
        return;
    }
    acc[TRK_E] += hw * residual * residual * (2 - hw);
    acc[TRK_NE] += 1.f;
    acc[TRK_NW] += 1.f;
    // calcGSSSE row (:591-613)
    const float gx = hit1 * L.fx, gy = hit2 * L.fy;
    float J[9];
    J[0] = new_idepth * gx;
    J[1] = new_idepth * gy;
    J[2] = 0.f - new_idepth * (u * gx + v * gy);
    J[3] = 0.f - ((u * v) * gx + gy * (1.f + v * v));
    J[4] = (u * v) * gy + gx * (1.f + u * u);
    J[5] = u * gy - v * gx;
    J[6] = P.affLL0 * (P.b0 - refColor);
    J[7] = -1.f;
    J[8] = residual;
    int k = 0;
#pragma unroll
    for (int r = 0; r < 9; r++) {
        const float Jw = J[r] * hw;
#pragma unroll
        for (int c = r; c < 9; c++) acc[k++] += Jw * J[c];
    }
}

// block-wide sum of TRK_NACC accumulators -> out[TRK_NACC] (double) in shared memory; blockDim multiple of 32
template<int NT>
__device__ void trk_block_reduce(float *acc, float *s_part /*[NT/32][TRK_NACC]*/, double *out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < TRK_NACC; k++) {
        float v = acc[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[warp * TRK_NACC + k] = v;
    }
    __syncthreads();
    if (threadIdx.x < TRK_NACC) {
        double s = 0.0;
        for (int w = 0; w < NT / 32; w++) s += (double) s_part[w * TRK_NACC + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

// res6 / H / b from the summed accumulators (CoarseTracker.cc:563-571, 616-631)
__device__ void trk_finish(const double *S, double *res6, double *H /*row-major 8x8*/, double *b) {
    res6[0] = (float) S[TRK_E];
    res6[1] = S[TRK_NE];
    res6[2] = (float) S[TRK_ST] / ((float) S[TRK_SNUM] + 0.1);
    res6[3] = 0;
    res6[4] = (float) S[TRK_SRT] / ((float) S[TRK_SNUM] + 0.1);
    res6[5] = (float) ((int) S[TRK_NSAT]) / (float) ((int) S[TRK_NE]);
    if (H == nullptr) return;
    const int nw = (int) (S[TRK_NW] + 0.5);
    const int n = (nw + 3) & ~3;     // buf_warped_n is padded to a multiple of 4 (:550-560)
    const double fac = (double) (1.0f / n);
    const double sc[8] = {SCALE_XI_ROT, SCALE_XI_ROT, SCALE_XI_ROT, SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_A, SCALE_B};
    int k = 0;
    for (int r = 0; r < 9; r++)
        for (int c = r; c < 9; c++, k++) {
            const double v = (double) (float) S[k] * fac;
            if (r < 8 && c < 8) { H[r * 8 + c] = v * sc[c] * sc[r]; H[c * 8 + r] = v * sc[r] * sc[c]; }
            else if (r < 8 && c == 8) b[r] = v * sc[r];
        }
}

#define TRK_EVAL_THREADS 256
// piecewise evaluation: per-CTA partial sums -> global; the last CTA to finish folds them (threadfence reduction)
__global__ void __launch_bounds__(TRK_EVAL_THREADS)
k_trk_eval(TrkLevel L, TrkPose P, int lvl0, float *partials /*[grid][TRK_NACC]*/, unsigned *counter, double *out /*6+64+8*/, int wantH) {
    __shared__ float s_part[(TRK_EVAL_THREADS / 32) * TRK_NACC];
    __shared__ double s_sum[TRK_NACC];
    __shared__ bool last;
    float acc[TRK_NACC];
#pragma unroll
    for (int k = 0; k < TRK_NACC; k++) acc[k] = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L.n; i += gridDim.x * blockDim.x) trk_point(L, P, i, lvl0 != 0, acc);
    trk_block_reduce<TRK_EVAL_THREADS>(acc, s_part, s_sum);
    if (threadIdx.x < TRK_NACC) partials[blockIdx.x * TRK_NACC + threadIdx.x] = (float) s_sum[threadIdx.x];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    if (threadIdx.x < TRK_NACC) {
        double s = 0.0;
        for (unsigned bI = 0; bI < gridDim.x; bI++) s += (double) ((volatile float *) partials)[bI * TRK_NACC + threadIdx.x];
        s_sum[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        trk_finish(s_sum, out, wantH ? out + 6 : nullptr, out + 70);
        *counter = 0;
    }
}

// ---- small dense helpers for the LM loop (thread 0)
// x = A^-1 rhs for the leading m x m block of the row-major 8x8 matrix A, with Eigen::LDLT's algorithm:
// pivot on the largest remaining |diagonal|, lower unit-triangular L, pseudo-inverse of D.
__device__ void ldlt_solve_small(const double *A8, const double *rhs, int m, double *x) {
    double M[64];
    int tr[8];
    for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) M[i * 8 + j] = A8[i * 8 + j];
    for (int k = 0; k < m; k++) {
        int big = k;
        double bv = fabs(M[k * 8 + k]);
        for (int i = k + 1; i < m; i++) if (fabs(M[i * 8 + i]) > bv) { bv = fabs(M[i * 8 + i]); big = i; }
        tr[k] = big;
        if (big != k) {   // symmetric row/column swap of the lower triangle
            for (int j = 0; j < k; j++) { double tmp = M[k * 8 + j]; M[k * 8 + j] = M[big * 8 + j]; M[big * 8 + j] = tmp; }
            for (int i = big + 1; i < m; i++) { double tmp = M[i * 8 + k]; M[i * 8 + k] = M[i * 8 + big]; M[i * 8 + big] = tmp; }
            { double tmp = M[k * 8 + k]; M[k * 8 + k] = M[big * 8 + big]; M[big * 8 + big] = tmp; }
            for (int i = k + 1; i < big; i++) { double tmp = M[i * 8 + k]; M[i * 8 + k] = M[big * 8 + i]; M[big * 8 + i] = tmp; }
        }
        double dk = M[k * 8 + k];
        for (int j = 0; j < k; j++) dk -= M[k * 8 + j] * M[k * 8 + j] * M[j * 8 + j];
        M[k * 8 + k] = dk;
        for (int i = k + 1; i < m; i++) {
            double v = M[i * 8 + k];
            for (int j = 0; j < k; j++) v -= M[i * 8 + j] * M[k * 8 + j] * M[j * 8 + j];
            M[i * 8 + k] = (fabs(dk) > 0.0) ? v / dk : v;
        }
    }
    for (int i = 0; i < m; i++) x[i] = rhs[i];
    for (int k = 0; k < m; k++) { double tmp = x[k]; x[k] = x[tr[k]]; x[tr[k]] = tmp; }
    for (int i = 0; i < m; i++) for (int j = 0; j < i; j++) x[i] -= M[i * 8 + j] * x[j];
    for (int i = 0; i < m; i++) x[i] = (fabs(M[i * 8 + i]) > 2.2250738585072014e-308) ? x[i] / M[i * 8 + i] : 0.0;
    for (int i = m - 1; i >= 0; i--) for (int j = i + 1; j < m; j++) x[i] -= M[j * 8 + i] * x[j];
    for (int k = m - 1; k >= 0; k--) { double tmp = x[k]; x[k] = x[tr[k]]; x[tr[k]] = tmp; }
}

struct TrkShared {   // state of the LM loop, in shared memory, written by thread 0
    TrkPose P;
    double R[9], t[3];          // refToNew_current
    double Rn[9], tn[3];        // refToNew_new
    float aff_a, aff_b, aff_a_new, aff_b_new;
    double resOld[6], resNew[6], H[64], b[8];
    int action;                 // what the CTA does next
    int lvl;
};

struct TrkTrackArgs {
    TrkLevel L[MAXLVL];
    int nLevels;
    float ref_aff_a, ref_aff_b, ref_exposure, new_exposure;
    float huberTH, coarseCutoffTH, affineOptModeA, affineOptModeB;
    double R[9], t[3];
    float aff_a, aff_b;
    int coarsestLvl;
    double minResForAbort[5];
};
struct TrkTrackOut {
    double R[9], t[3];
    float aff_a, aff_b;
    double lastResiduals[5], lastFlowIndicators[3];
    int ok;
    int n_evals;
};

__device__ void trk_make_pose(const TrkTrackArgs &A, const TrkLevel &L, const double *R, const double *t, float aff_a, float aff_b,
                              float cutoffTH, TrkPose &P) {
    float Rf[9];
    for (int i = 0; i < 9; i++) Rf[i] = (float) R[i];
    m33f_mul(Rf, L.Ki, P.RKi);
    for (int i = 0; i < 3; i++) P.t[i] = (float) t[i];
    float eF = A.ref_exposure, eT = A.new_exposure;
    if (eF == 0 || eT == 0) eT = eF = 1;
    const float a = expf(aff_a - A.ref_aff_a) * eT / eF;
    P.affLL0 = a;
    P.affLL1 = aff_b - a * A.ref_aff_b;
    P.b0 = A.ref_aff_b;
    P.cutoffTH = cutoffTH;
    P.huberTH = A.huberTH;
    P.maxEnergy = 2 * A.huberTH * cutoffTH - A.huberTH * A.huberTH;
}

// one starting pose of a batched track (FullSystem::trackNewCoarse tries up to 3 + 27 x 3 of them per frame, FullSystem.cc:290-357)
struct TrkHypothesis {
    double R[9], t[3];
    float aff_a, aff_b;
};
#define TRK_TRACK_THREADS 512
// One CTA runs one whole coarse-to-fine LM loop. hyp == nullptr: the pose in A, grid 1. hyp != nullptr: CTA b starts from hyp[b] and
// writes out[b] -- the hypotheses of a frame tracked side by side, one SM each.
__global__ void __launch_bounds__(TRK_TRACK_THREADS) k_trk_track(TrkTrackArgs A, TrkTrackOut *out, const TrkHypothesis *hyp) {
    if (hyp != nullptr) {
        const TrkHypothesis &h = hyp[blockIdx.x];
        for (int i = 0; i < 9; i++) A.R[i] = h.R[i];
        for (int i = 0; i < 3; i++) A.t[i] = h.t[i];
        A.aff_a = h.aff_a; A.aff_b = h.aff_b;
        out += blockIdx.x;
    }
    __shared__ float s_part[(TRK_TRACK_THREADS / 32) * TRK_NACC];
    __shared__ double s_sum[TRK_NACC];
    __shared__ TrkShared S;
    const int tid = threadIdx.x;

    // evaluate pose currently in S.P at level S.lvl into res6 (+H,b)
    auto eval = [&](double *res6, double *H, double *b) {
        const TrkLevel &L = A.L[S.lvl];
        float acc[TRK_NACC];
#pragma unroll
        for (int k = 0; k < TRK_NACC; k++) acc[k] = 0.f;
        const bool lvl0 = (S.lvl == 0);
        for (int i = tid; i < L.n; i += TRK_TRACK_THREADS) trk_point(L, S.P, i, lvl0, acc);
        trk_block_reduce<TRK_TRACK_THREADS>(acc, s_part, s_sum);
        if (tid == 0) trk_finish(s_sum, res6, H, b);
        __syncthreads();
    };

    __shared__ int s_evals;
    __shared__ double lastRes[5], lastFlow[3];
    if (tid == 0) {
        for (int i = 0; i < 9; i++) S.R[i] = A.R[i];
        for (int i = 0; i < 3; i++) S.t[i] = A.t[i];
        S.aff_a = A.aff_a; S.aff_b = A.aff_b;
        s_evals = 0;
        for (int i = 0; i < 5; i++) lastRes[i] = NAN;
        for (int i = 0; i < 3; i++) lastFlow[i] = 1000;
    }
    __syncthreads();

    const int maxIterations[5] = {10, 20, 50, 50, 50};
    const float lambdaExtrapolationLimit = 0.001f;
    bool haveRepeated = false;
    bool aborted = false;

    for (int lvl = A.coarsestLvl; lvl >= 0; lvl--) {
        float levelCutoffRepeat = 1;
        if (tid == 0) {
            S.lvl = lvl;
            trk_make_pose(A, A.L[lvl], S.R, S.t, S.aff_a, S.aff_b, A.coarseCutoffTH * levelCutoffRepeat, S.P);
            s_evals++;
        }
        __syncthreads();
        eval(S.resOld, S.H, S.b);
        while (S.resOld[5] > 0.6 && levelCutoffRepeat < 50) {    // uniform: S is shared
            levelCutoffRepeat *= 2;
            if (tid == 0) {
                trk_make_pose(A, A.L[lvl], S.R, S.t, S.aff_a, S.aff_b, A.coarseCutoffTH * levelCutoffRepeat, S.P);
                s_evals++;
            }
            __syncthreads();
            eval(S.resOld, S.H, S.b);
        }
        float lambda = 0.01f;
        for (int iteration = 0; iteration < maxIterations[lvl]; iteration++) {
            __shared__ double s_incnorm;
            if (tid == 0) {
                double Hl[64], inc[8], nb[8];
                for (int i = 0; i < 64; i++) Hl[i] = S.H[i];
                for (int i = 0; i < 8; i++) { Hl[i * 8 + i] *= (1 + lambda); nb[i] = -S.b[i]; }
                ldlt_solve_small(Hl, nb, 8, inc);
                const bool fixA = A.affineOptModeA < 0, fixB = A.affineOptModeB < 0;
                if (fixA && fixB) { ldlt_solve_small(Hl, nb, 6, inc); inc[6] = inc[7] = 0; }
                if (!fixA && fixB) { ldlt_solve_small(Hl, nb, 7, inc); inc[7] = 0; }
                if (fixA && !fixB) {
                    double Hs[64], bs[8], is[8];
                    for (int i = 0; i < 64; i++) Hs[i] = Hl[i];
                    for (int i = 0; i < 8; i++) bs[i] = nb[i];
                    for (int r = 0; r < 8; r++) Hs[r * 8 + 6] = Hs[r * 8 + 7];
                    for (int c = 0; c < 8; c++) Hs[6 * 8 + c] = Hs[7 * 8 + c];
                    bs[6] = bs[7];
                    ldlt_solve_small(Hs, bs, 7, is);
                    for (int i = 0; i < 6; i++) inc[i] = is[i];
                    inc[6] = 0;
                    inc[7] = is[6];
                }
                float extrapFac = 1;
                if (lambda < lambdaExtrapolationLimit) extrapFac = sqrtf(sqrt(lambdaExtrapolationLimit / lambda));
                for (int i = 0; i < 8; i++) inc[i] *= extrapFac;
                double incScaled[8];
                for (int i = 0; i < 3; i++) incScaled[i] = inc[i] * SCALE_XI_ROT;
                for (int i = 3; i < 6; i++) incScaled[i] = inc[i] * SCALE_XI_TRANS;
                incScaled[6] = inc[6] * SCALE_A;
                incScaled[7] = inc[7] * SCALE_B;
                double sum = 0;
                for (int i = 0; i < 8; i++) sum += incScaled[i];
                if (!isfinite(sum)) for (int i = 0; i < 8; i++) incScaled[i] = 0;
                double Re[9], te[3];
                se3_exp(incScaled, Re, te);
                se3_mul(Re, te, S.R, S.t, S.Rn, S.tn);
                S.aff_a_new = S.aff_a; S.aff_b_new = S.aff_b;
                S.aff_a_new += incScaled[6];
                S.aff_b_new += incScaled[7];
                trk_make_pose(A, A.L[lvl], S.Rn, S.tn, S.aff_a_new, S.aff_b_new, A.coarseCutoffTH * levelCutoffRepeat, S.P);
                double nrm = 0;
                for (int i = 0; i < 8; i++) nrm += inc[i] * inc[i];
                s_incnorm = sqrt(nrm);
                s_evals++;
            }
            __syncthreads();
            // resNew (and the H,b that calcGSSSE would compute on acceptance) at the trial pose
            __shared__ double Hn[64], bn[8];
            eval(S.resNew, Hn, bn);
            const bool accept = (S.resNew[0] / S.resNew[1]) < (S.resOld[0] / S.resOld[1]);
            __syncthreads();
            if (tid == 0 && accept) {
                for (int i = 0; i < 64; i++) S.H[i] = Hn[i];
                for (int i = 0; i < 8; i++) S.b[i] = bn[i];
                for (int i = 0; i < 6; i++) S.resOld[i] = S.resNew[i];
                S.aff_a = S.aff_a_new; S.aff_b = S.aff_b_new;
                for (int i = 0; i < 9; i++) S.R[i] = S.Rn[i];
                for (int i = 0; i < 3; i++) S.t[i] = S.tn[i];
            }
            if (accept) lambda *= 0.5f;
            else {
                lambda *= 4;
                if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
            }
            const bool stop = !(s_incnorm > 1e-3);
            __syncthreads();
            if (stop) break;
        }
        const float lr = sqrtf((float) (S.resOld[0] / S.resOld[1]));
        if (tid == 0) {
            lastRes[lvl] = lr;
            for (int i = 0; i < 3; i++) lastFlow[i] = S.resOld[2 + i];
        }
        if ((double) lr > 1.5 * A.minResForAbort[lvl]) { aborted = true; break; }
        if (levelCutoffRepeat > 1 && !haveRepeated) { lvl++; haveRepeated = true; }
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) {
        int ok = aborted ? 0 : 1;
        float a_out = S.aff_a, b_out = S.aff_b;
        if (!aborted) {
            for (int i = 0; i < 9; i++) out->R[i] = S.R[i];
            for (int i = 0; i < 3; i++) out->t[i] = S.t[i];
            if ((A.affineOptModeA != 0 && (fabsf(a_out) > 1.2f)) || (A.affineOptModeB != 0 && (fabsf(b_out) > 200))) ok = 0;
            float eF = A.ref_exposure, eT = A.new_exposure;
            if (eF == 0 || eT == 0) eT = eF = 1;
            const float ra = expf(a_out - A.ref_aff_a) * eT / eF;
            const float rb = b_out - ra * A.ref_aff_b;
            if (ok && ((A.affineOptModeA == 0 && (fabsf(logf(ra)) > 1.5f)) || (A.affineOptModeB == 0 && (fabsf(rb) > 200)))) ok = 0;
            if (ok) {
                if (A.affineOptModeA < 0) a_out = 0;
                if (A.affineOptModeB < 0) b_out = 0;
            }
            out->aff_a = a_out;
            out->aff_b = b_out;
        } else {
            for (int i = 0; i < 9; i++) out->R[i] = A.R[i];
            for (int i = 0; i < 3; i++) out->t[i] = A.t[i];
            out->aff_a = A.aff_a;
            out->aff_b = A.aff_b;
        }
        for (int i = 0; i < 5; i++) out->lastResiduals[i] = lastRes[i];
        for (int i = 0; i < 3; i++) out->lastFlowIndicators[i] = lastFlow[i];
        out->ok = ok;
        out->n_evals = s_evals;
    }
}

// =========================================================================================================
// Device-side CoarseTracker::makeCoarseDepthL0 (src/frontend/CoarseTracker.cc:258-438): the reference point cloud
// of the tracker, built from the active points' projections into the newest keyframe.
// (1) scatter idepth*weight and weight at the rounded projection (:262-283)
__global__ void k_cd_scatter(int n, const float *__restrict__ cpt, const float *__restrict__ HdiF, float *idepth0, float *wsum0, int w0, int h0) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int u = (int) (cpt[3 * k + 0] + 0.5f);
    const int v = (int) (cpt[3 * k + 1] + 0.5f);
    if (u < 0 || v < 0 || u >= w0 || v >= h0) return;     // the reference would write out of bounds here
    const float new_idepth = cpt[3 * k + 2];
    const float weight = sqrtf(1e-3 / (HdiF[k] + 1e-12));
    atomicAdd(idepth0 + u + w0 * v, new_idepth * weight);   // >2 points on one pixel: sum order differs by ulps
    atomicAdd(wsum0 + u + w0 * v, weight);
}
// (2) 2x2 sums down the pyramid (:285-310)
__global__ void k_cd_down(const float *__restrict__ idm, const float *__restrict__ wsm, float *idl, float *wsl, int wl, int hl, int wlm1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= wl * hl) return;
    const int x = i % wl, y = i / wl;
    const int b = 2 * x + 2 * y * wlm1;
    idl[i] = idm[b] + idm[b + 1] + idm[b + wlm1] + idm[b + wlm1 + 1];
    wsl[i] = wsm[b] + wsm[b + 1] + wsm[b + wlm1] + wsm[b + wlm1 + 1];
}
// (3) one dilation pass (:312-395): diagonal neighbours on levels 0,1, 4-neighbourhood above. Cells with weight are
// only read, cells without are only written, so the in-place update of idepth is race-free like in the reference.
__global__ void k_cd_dilate(float *idl, float *wsl, const float *__restrict__ wbak, int wl, int hl, int diagonal) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < wl || i >= wl * hl - wl) return;
    if (wbak[i] > 0) return;
    int nb[4];
    if (diagonal) { nb[0] = i + 1 + wl; nb[1] = i - 1 - wl; nb[2] = i + wl - 1; nb[3] = i - wl + 1; }
    else { nb[0] = i + 1; nb[1] = i - 1; nb[2] = i + wl; nb[3] = i - wl; }
    float sum = 0, num = 0, numn = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (wbak[nb[k]] > 0) { sum += idl[nb[k]]; num += wbak[nb[k]]; numn++; }
    if (numn > 0) { idl[i] = sum / numn; wsl[i] = num / numn; }
}
// (4) normalise + ordered compaction (:398-437). One block per image row computes each valid pixel's rank in the row.
__device__ __forceinline__ bool cd_valid(const float *idl, const float *wsl, const float4 *ref, int i, float &id, float &col) {
    if (!(wsl[i] > 0)) return false;
    id = idl[i] / wsl[i];
    col = ref[i].x;
    return isfinite(col) && (id > 0);
}
__global__ void __launch_bounds__(128) k_cd_rowcount(const float *__restrict__ idl, const float *__restrict__ wsl, const float4 *__restrict__ ref,
                                                     int wl, int hl, int *pos, int *rowcount) {
    const int y = blockIdx.x;
    __shared__ int s_warp[4];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const bool rowok = (y >= 2 && y < hl - 2);
    for (int x0 = 0; x0 < wl; x0 += 128) {
        const int x = x0 + threadIdx.x;
        float id, col;
        const bool v = rowok && x >= 2 && x < wl - 2 && cd_valid(idl, wsl, ref, x + y * wl, id, col);
        const unsigned bal = __ballot_sync(0xffffffffu, v);
        const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
        if (lane == 0) s_warp[wp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int k = 0; k < wp; k++) off += s_warp[k];
        if (v) pos[x + y * wl] = off + __popc(bal & ((1u << lane) - 1));
        else if (x < wl) pos[x + y * wl] = -1;
        __syncthreads();
        if (threadIdx.x == 0) s_base += s_warp[0] + s_warp[1] + s_warp[2] + s_warp[3];
        __syncthreads();
    }
    if (threadIdx.x == 0) rowcount[y] = s_base;
}
__global__ void __launch_bounds__(1024) k_cd_rowscan(int *rowcount, int hl, int *total) {
    __shared__ int s[1024];
    const int t = threadIdx.x;
    s[t] = (t < hl) ? rowcount[t] : 0;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = (t >= o) ? s[t - o] : 0;
        __syncthreads();
        s[t] += v;
        __syncthreads();
    }
    if (t < hl) rowcount[t] = s[t] - ((t < hl) ? (s[t] - (t > 0 ? s[t - 1] : 0)) : 0);   // exclusive
    if (t == 0) *total = s[1023];
}
__global__ void k_cd_emit(const float *__restrict__ idl, const float *__restrict__ wsl, const float4 *__restrict__ ref, const int *__restrict__ pos,
                          const int *__restrict__ rowoff, int wl, int hl, float *pc_u, float *pc_v, float *pc_id, float *pc_col) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= wl * hl) return;
    const int p = pos[i];
    if (p < 0) return;
    const int x = i % wl, y = i / wl;
    float id, col;
    cd_valid(idl, wsl, ref, i, id, col);
    const int dst = rowoff[y] + p;
    pc_u[dst] = (float) x; pc_v[dst] = (float) y; pc_id[dst] = id; pc_col[dst] = col;
}
