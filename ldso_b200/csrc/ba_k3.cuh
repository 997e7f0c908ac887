// K3 — one CTA: EnergyFunctional::solveSystemF (EnergyFunctional.cc:240-351, default solver mode), the frame/calib
// part of resubstituteF_MT (:491-507), FullSystem::backupState / doStepFromBackup (FullSystem.cc:1587-1676),
// FrameHessian::setState, FrameFramePrecalc::Set for all nF^2 pairs and EnergyFunctional::setDeltaF.
//
// Everything here is a serial dependency chain on a 68x68 system, so the design goal is latency, not throughput:
//   * the frame/calib records are staged in shared memory once (one global round trip instead of dozens);
//   * the LDL^T is blocked by 4 columns with the diagonal block factored redundantly in every row thread's registers (see the
//     comment in front of k3_rcp for what was measured and why);
//   * after the solve, independent pieces run on different warps at once (frame step + SE3::exp | calibration + canbreak | xAd,
//     adHTdeltaF) instead of one after the other.
#pragma once
#include "common.cuh"
#include "se3_math.cuh"
#include "ba_k2.cuh"

#define K3F_SOLVE 1
#define K3F_STEP 2
#define K3F_BACKUP 4
#define K3F_SELECT 8       // grid 2: CTA 1 runs setNewFrameEnergyTH's order-statistic select beside the solver
#define K3_THREADS 512
#define K3_NP MAXN               // n = 8 nF + 4 is a multiple of the block size 4: no padding
#define K3_LD (K3_NP + 1)        // odd leading dimension: a column of the matrix touches every bank once
#define K3_NB 4
#define K3_A0LD (MAXN + 1)       // staged input: odd column stride (with the natural stride 68 a walk along a row of the column-major matrix
                                 // touches 4 bank groups only: 8-way conflicts on half of the permuted copy's loads)
static_assert((MAXN * K3_A0LD) % 2 == 0, "shared-memory carve-up: 16-byte alignment behind A0");
#define K3_WPLD (K3_NP + 2)      // Wp is [K3_NB][K3_WPLD] (column of the panel major): conflict-free for consecutive rows

struct K3Frames {       // shared-memory staging of the mutable window records
    FrameDev fr[MAXF];
    CalibDev calib;
};

__device__ void calib_set_value(CalibDev &c, const double v[4]) {  // CalibHessian::setValue (CalibHessian.h:71-85)
    for (int i = 0; i < 4; i++) c.value[i] = v[i];
    c.value_scaled[0] = (double) SCALE_F * c.value[0];
    c.value_scaled[1] = (double) SCALE_F * c.value[1];
    c.value_scaled[2] = (double) SCALE_C * c.value[2];
    c.value_scaled[3] = (double) SCALE_C * c.value[3];
    c.fxl = (float) c.value_scaled[0]; c.fyl = (float) c.value_scaled[1];
    c.cxl = (float) c.value_scaled[2]; c.cyl = (float) c.value_scaled[3];
    c.fxli = 1.0f / c.fxl; c.fyli = 1.0f / c.fyl;
    c.cxli = -c.cxl / c.fxl; c.cyli = -c.cyl / c.fyl;
}

__device__ void stage_in(K3Frames *S, const WinState *ws) {
    const int nw = (int) (sizeof(FrameDev) * MAXF / 4);
    const unsigned *src = (const unsigned *) ws->fr;
    unsigned *dst = (unsigned *) S->fr;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = src[i];
    const int nc = (int) (sizeof(CalibDev) / 4);
    const unsigned *srcc = (const unsigned *) &ws->calib;
    unsigned *dstc = (unsigned *) &S->calib;
    for (int i = threadIdx.x; i < nc; i += blockDim.x) dstc[i] = srcc[i];
    __syncthreads();
}
__device__ void stage_out(const K3Frames *S, WinState *ws) {
    __syncthreads();
    const int nw = (int) (sizeof(FrameDev) * MAXF / 4);
    unsigned *dst = (unsigned *) ws->fr;
    const unsigned *src = (const unsigned *) S->fr;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[i] = src[i];
    const int nc = (int) (sizeof(CalibDev) / 4);
    unsigned *dstc = (unsigned *) &ws->calib;
    const unsigned *srcc = (const unsigned *) &S->calib;
    for (int i = threadIdx.x; i < nc; i += blockDim.x) dstc[i] = srcc[i];
}

// part 1 (thread f < nF): FrameHessian::setState for frame f -- PRE_worldToCam = SE3::exp(scaled state) * worldToCam_evalPT, delta
__device__ __forceinline__ void frames_exp_part(K3Frames *S, int f_idx) {
    FrameDev &f = S->fr[f_idx];
    double ss[6];
    for (int i = 0; i < 3; i++) ss[i] = (double) SCALE_XI_TRANS * f.state[i];
    for (int i = 3; i < 6; i++) ss[i] = (double) SCALE_XI_ROT * f.state[i];
    double Re[9], te[3];
    se3_exp(ss, Re, te);
    se3_mul(Re, te, f.evalR, f.evalT, f.preR, f.preT);
    for (int i = 0; i < 8; i++) {
        f.delta[i] = f.state[i] - f.state_zero[i];
        f.delta_prior[i] = f.state[i];
    }
}
// part 2 (all threads of the CTA, after a barrier behind part 1): the nF^2 pair records, three short parallel phases instead of one
// long per-pair chain: F2 (pair,row): one row of R = R_t R_h^T and of t = t_t - R t_h, in double; F3 (pair,row): one row of
// K R K^-1 and K t in float (the reference's Mat33f products, FrameFramePrecalc.cc:21-31); F3' (pair): affine brightness transfer.
__device__ void frames_pairs_part(K3Frames *S, WinState *ws, bool full) {
    const int nF = ws->nF, tid = threadIdx.x;
    __shared__ float sRf[MAXPAIR][9], sTf[MAXPAIR][3];
    __shared__ double sTd[MAXPAIR][3];
    for (int o = tid; o < nF * nF * 3; o += blockDim.x) {
        const int q = o / 3, i = o - 3 * q, h = q % nF, t = q / nF;
        const FrameDev &fh = S->fr[h], &ft = S->fr[t];
        if (full) {     // eval-point (FEJ) part: constant while the window's linearisation point is fixed
            double r0[3], t0 = ft.evalT[i];
            for (int j = 0; j < 3; j++) r0[j] = ft.evalR[i * 3] * fh.evalR[j * 3] + ft.evalR[i * 3 + 1] * fh.evalR[j * 3 + 1] + ft.evalR[i * 3 + 2] * fh.evalR[j * 3 + 2];
            t0 -= r0[0] * fh.evalT[0] + r0[1] * fh.evalT[1] + r0[2] * fh.evalT[2];
            PairRec &pc = ws->pair[q];
            for (int j = 0; j < 3; j++) pc.R0[i * 3 + j] = (float) r0[j];
            pc.t0[i] = (float) t0;
            if (i == 0) { pc.b0 = (float) (fh.state_zero[7] * (double) SCALE_B); pc.pad[0] = pc.pad[1] = pc.pad[2] = pc.pad[3] = 0.f; }
        }
        double r[3], tt = ft.preT[i];
        for (int j = 0; j < 3; j++) r[j] = ft.preR[i * 3] * fh.preR[j * 3] + ft.preR[i * 3 + 1] * fh.preR[j * 3 + 1] + ft.preR[i * 3 + 2] * fh.preR[j * 3 + 2];
        tt -= r[0] * fh.preT[0] + r[1] * fh.preT[1] + r[2] * fh.preT[2];
        for (int j = 0; j < 3; j++) sRf[q][i * 3 + j] = (float) r[j];
        sTf[q][i] = (float) tt;
        sTd[q][i] = tt;
    }
    __syncthreads();
    // outputs [0, 3 nF^2): (pair, row) of K R K^-1 / K t; [3 nF^2, 4 nF^2): the pairs' brightness transfer -- whole warps take one branch
    const int nP3 = nF * nF * 3;
    for (int o = tid; o < nF * nF * 4; o += blockDim.x) {
        if (o < nP3) {
            const int q = o / 3, i = o - 3 * q;
            PairRec &pc = ws->pair[q];
            const CalibDev &c = S->calib;
            const float K[9] = {c.fxl, 0, c.cxl, 0, c.fyl, c.cyl, 0, 0, 1};
            float Ki[9];
            m33f_inverse(K, Ki);
            // row i of K by selects (a runtime index would put K into local memory)
            const float k0 = (i == 0) ? c.fxl : 0.f, k1 = (i == 1) ? c.fyl : 0.f, k2 = (i == 0) ? c.cxl : (i == 1) ? c.cyl : 1.f;
            const float *Rf = sRf[q];
            float tmp[3];
            for (int j = 0; j < 3; j++) {
                float s2 = k0 * Rf[0 * 3 + j];
                s2 += k1 * Rf[1 * 3 + j];
                s2 += k2 * Rf[2 * 3 + j];
                tmp[j] = s2;
            }
            for (int j = 0; j < 3; j++) {
                float s2 = tmp[0] * Ki[0 * 3 + j];
                s2 += tmp[1] * Ki[1 * 3 + j];
                s2 += tmp[2] * Ki[2 * 3 + j];
                pc.KRKi[i * 3 + j] = s2;
            }
            float s2 = k0 * sTf[q][0];
            s2 += k1 * sTf[q][1];
            s2 += k2 * sTf[q][2];
            pc.Kt[i] = s2;
            PairRecFull &pf = ws->pairFull[q];
            for (int j = 0; j < 3; j++) pf.RTll[i * 3 + j] = Rf[i * 3 + j];
            pf.tTll[i] = sTf[q][i];
        } else {
            const int q = o - nP3;
            PairRec &pc = ws->pair[q];
            const int h = q % nF, t = q / nF;
            const FrameDev &fh = S->fr[h], &ft = S->fr[t];
            // AffLight::fromToVecExposure (AffLight.h:27-35) with aff_g2l() = state_scaled[6..7]
            float eF = fh.ab_exposure, eT = ft.ab_exposure;
            if (eF == 0 || eT == 0) eT = eF = 1;
            const float ah = (float) ((double) SCALE_A * fh.state[6]), bh = (float) ((double) SCALE_B * fh.state[7]);
            const float at = (float) ((double) SCALE_A * ft.state[6]), bt = (float) ((double) SCALE_B * ft.state[7]);
            const float aa = expf(at - ah) * eT / eF;
            pc.aff[0] = aa;
            pc.aff[1] = bt - aa * bh;
            pc.distanceLL = (float) sqrt(sTd[q][0] * sTd[q][0] + sTd[q][1] * sTd[q][1] + sTd[q][2] * sTd[q][2]);
        }
    }
}
// adHTdeltaF (EnergyFunctional.cc:406-414), outputs o0, o0 + stride, ... of the nF*nF*8 (pair, column) outputs. With vx != nullptr the
// frame state is formed here as state_backup + (-vx) -- the bits doStepFromBackup stores -- so that the caller can run this beside
// the threads that write state / step (no shared-memory record is read that another warp writes in the same phase).
__device__ __forceinline__ void frames_adHTdelta(const K3Frames *S, WinState *ws, const float *adHF, const float *adTF, int o0, int stride, const double *vx) {
    const int nF = ws->nF;
    for (int o = o0; o < nF * nF * 8; o += stride) {
        const int q = o >> 3, j = o & 7, h = q % nF, t = q / nF;
        const FrameDev &fh = S->fr[h], &ft = S->fr[t];
        const float *AH = adHF + q * 64, *AT = adTF + q * 64;
        float s1 = 0.f, s2 = 0.f;
        if (vx != nullptr) {
            for (int i = 0; i < 8; i++) { const double st = fh.state_backup[i] + (-vx[CPARS + 8 * h + i]); s1 += (float) (st - fh.state_zero[i]) * AH[i * 8 + j]; }
            for (int i = 0; i < 8; i++) { const double st = ft.state_backup[i] + (-vx[CPARS + 8 * t + i]); s2 += (float) (st - ft.state_zero[i]) * AT[i * 8 + j]; }
        } else {
            for (int i = 0; i < 8; i++) s1 += (float) (fh.state[i] - fh.state_zero[i]) * AH[i * 8 + j];
            for (int i = 0; i < 8; i++) s2 += (float) (ft.state[i] - ft.state_zero[i]) * AT[i * 8 + j];
        }
        ws->adHTdeltaF[q][j] = s1 + s2;
    }
}

// FrameHessian::setState (FrameHessian.h:78-91), FrameFramePrecalc::Set for all pairs (FrameFramePrecalc.cc:6-35),
// EnergyFunctional::setDeltaF frame part (EnergyFunctional.cc:403-429). Frame records live in shared memory (S);
// the pair records are written to global. Called by all threads of a CTA with >= 128 threads.
__device__ void frames_refresh(K3Frames *S, WinState *ws, bool full, const float *adHF, const float *adTF) {
    const int nF = ws->nF, tid = threadIdx.x;
    if (tid < nF) frames_exp_part(S, tid);
    if (tid == 64) {
        CalibDev &c = S->calib;
        for (int i = 0; i < 4; i++) c.cDeltaF[i] = (float) (c.value[i] - c.value_zero[i]);
    }
    __syncthreads();
    frames_pairs_part(S, ws, full);
    frames_adHTdelta(S, ws, adHF, adTF, tid, blockDim.x, nullptr);
    __syncthreads();
}

__global__ void __launch_bounds__(128) k_frames_refresh(WinState *ws) {
    __shared__ K3Frames S;
    stage_in(&S, ws);
    frames_refresh(&S, ws, true, &ws->adHostF[0][0], &ws->adTargetF[0][0]);
    stage_out(&S, ws);
}

// ---------------------------------------------------------------------------------------------------------------------
// The 68x68 solve. Everything below is one serial dependency chain on a tiny matrix, so the design goal is the LENGTH OF
// THE CHAIN, not throughput (measured on B200 with tools/panel_bench.cu and ncu's per-instruction stall samples):
//   * LDL^T's inherent chain is one reciprocal + one FMA per pivot (~105 cycles in f64: MUFU seed + 5 dependent DFMAs + the FMA).
//     The factorisation is blocked by 4 columns; inside a block step every row thread factors the 4x4 diagonal block REDUNDANTLY in
//     its own registers and substitutes its own row on the fly, so a block step has no shuffle, no shared-memory exchange and no
//     barrier on the pivot chain. (8-column blocks were measured slower: the redundant O(NB^3) update costs ~2.3 issue cycles per
//     DFMA on the one warp that carries the chain.)
//   * the routine is BRANCH-FREE: every store is unconditional (rows inside the diagonal block write their never-read upper
//     entries, every row thread writes the same reciprocal): ncu attributed 30 % of the routine's time to branch_resolving stalls
//     behind the reconvergence points of its conditional stores.
//   * the next pivot's update is formed BEFORE the reciprocal it is scaled by is known (sq = W^2, then one FMA with 1/d).
//   * the update of the NEXT panel's 4 columns is spread over all threads (one element each) between two barriers; the rest of the
//     trailing update is done by the helper warps while the row threads already factor the next panel.
//   * the right-hand side rides along as matrix row n, so the forward substitution is free; the backward substitution is done by
//     ONE warp with the vector in registers (no barriers), 4 unknowns per step solved redundantly per lane.
// Eigen's LDLT pivots on the largest remaining |diagonal| of the INPUT matrix (its left-looking update never touches later
// diagonal entries before they are chosen), i.e. a descending-|diag| order: computed by a rank sort and applied as a symmetric
// permutation before the (then unpivoted) blocked factorisation; like Eigen, only the lower triangle of the input is referenced.

// ~1 ulp reciprocal without the slow-path branches of __drcp_rn: MUFU.RCP64H seed (a "gross approximation", ~9-10 bits measured:
// with the cubic step alone the solve was only good to 1e-9), one cubically convergent step y1 = y0 (1 + e + e^2), e = 1 - d y0,
// and one Newton step: 5 dependent FMAs
__device__ __forceinline__ double k3_rcp(double d) {
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
    const double e = fma(-d, y0, 1.0);
    const double t = fma(e, e, e);
    const double y1 = fma(y0, t, y0);
    const double e1 = fma(-d, y1, 1.0);
    return fma(y1, e1, y1);
}
__device__ __forceinline__ double shfl_f64(double v, int src) {      // low word first: lands in an aligned register pair
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return __hiloint2double(hi, lo);
}
#define K3_TRI(r, c) ((r) * ((r) + 1) / 2 + (c))      // packed lower triangle of the 4x4 diagonal block

// One block step of the panel for matrix row i (k0 <= i <= n; row n is the right-hand side). On entry the columns k0..k0+3 of all
// rows >= k0 carry every earlier block step's update. Writes, for this row: L (scaled) into A -- the pivot d itself on the diagonal
// -- and the unscaled W into Wp.
__device__ __forceinline__ void k3_panel_row(double *A, double *Wp, int k0, int i) {
    double D[K3_NB * (K3_NB + 1) / 2], a[K3_NB];
#pragma unroll
    for (int r = 0; r < K3_NB; r++)
#pragma unroll
        for (int c = 0; c <= r; c++) D[K3_TRI(r, c)] = A[(k0 + r) * K3_LD + k0 + c];
#pragma unroll
    for (int c = 0; c < K3_NB; c++) a[c] = A[i * K3_LD + k0 + c];      // (rows inside the block: entries right of the diagonal are never read back)
#pragma unroll
    for (int C = 0; C < K3_NB; C++) {
        const double dk = D[K3_TRI(C, C)];
        double sq = 0.0;
        if (C + 1 < K3_NB) sq = D[K3_TRI((C + 1) % K3_NB, C)] * D[K3_TRI((C + 1) % K3_NB, C)];
        const double inv = (fabs(dk) > 0.0) ? k3_rcp(dk) : 1.0;      // "don't scale by an invalid pivot" (Eigen LDLT)
        if (C + 1 < K3_NB) D[K3_TRI((C + 1) % K3_NB, (C + 1) % K3_NB)] = fma(-sq, inv, D[K3_TRI((C + 1) % K3_NB, (C + 1) % K3_NB)]);
        const double w = a[C], l = w * inv;
#pragma unroll
        for (int r = C + 1; r < K3_NB; r++) {
            const double lr = D[K3_TRI(r, C)] * inv;
#pragma unroll
            for (int q = C + 1; q <= r; q++)
                if (!(r == C + 1 && q == C + 1)) D[K3_TRI(r, q)] = fma(-lr, D[K3_TRI(q, C)], D[K3_TRI(r, q)]);
            a[r] = fma(-l, D[K3_TRI(r, C)], a[r]);
        }
        // unconditional stores (a select, no branch): below the pivot the scaled entry, on the diagonal the pivot itself (the
        // pseudo-inverse test of the solve reads it), right of it a value nobody reads
        A[i * K3_LD + k0 + C] = (k0 + C < i) ? l : w;
        Wp[C * K3_WPLD + i] = w;
    }
}

// Clock read that the compiler cannot move across memory operations, and that the hardware cannot execute before a
// preceding barrier has completed: BAR.SYNC.DEFER_BLOCKING lets a warp run ahead until its next memory instruction,
// so the dependent shared-memory load in front pins the read to "after the barrier released this warp".
__device__ __forceinline__ long long clk_fenced() {
    long long t;
    unsigned sink;
    asm volatile("{ .reg .u32 a; mov.u32 a, 0; ld.volatile.shared.u32 %1, [a]; }\n\tmov.u64 %0, %%clock64;" : "=l"(t), "=r"(sink)::"memory");
    if (sink == 0x7f123456u) t ^= 1;      // makes the clock read depend on the load's completion
    return t;
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    const unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned sa = (unsigned) __cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct K3Smem {       // carve-up of the dynamic shared memory block (all 16-byte aligned)
    double *A, *A0, *Wp, *vb, *vS, *vd, *vx, *Pns;
    int *perm;
    K3Frames *S;
    float *adH, *adT;
};
#define K3_A_DOUBLES (((K3_NP + 1) * K3_LD + 1) & ~1)
#define K3_SMEM_DOUBLES (K3_A_DOUBLES + MAXN * K3_A0LD + 2 * K3_NB * K3_WPLD + 4 * K3_NP + MAXN * MAXN + K3_NP / 2 + 4)
__device__ __forceinline__ K3Smem k3_carve(double *base) {
    K3Smem m;
    m.A = base;                                 // [(K3_NP + 1)][K3_LD] permuted, scaled, identity-padded system (+ rhs as row npad), factorised in place
    m.A0 = m.A + K3_A_DOUBLES;                   // [n][K3_A0LD] the assembled system as the stitch kernel left it (column-major, padded columns)
    m.Wp = m.A0 + MAXN * K3_A0LD;                // [2][K3_NB][K3_WPLD] unscaled panel W = L*D of the current / previous block step (68 * 69 is even)
    m.vb = m.Wp + 2 * K3_NB * K3_WPLD; m.vS = m.vb + K3_NP; m.vd = m.vS + K3_NP; m.vx = m.vd + K3_NP;
    m.Pns = m.vx + K3_NP;                        // [n*n] null-space projector
    m.perm = (int *) (m.Pns + MAXN * MAXN);      // [K3_NP]
    m.S = (K3Frames *) (m.perm + K3_NP + 8);
    m.adH = (float *) (m.S + 1);                 // [MAXPAIR][64] adHostF, adTargetF (index h + nF*t)
    m.adT = m.adH + MAXPAIR * 64;
    return m;
}

// Scaled, Eigen-ordered LDL^T solve of the assembled system (EnergyFunctional.cc:326-335): x = S (S A S)^-1 S b.
// In: m.A0 (n x n, column-major with column stride K3_A0LD), m.vb = b, m.vd = diag(A0). Out: m.vx. Called by all K3_THREADS threads.
__device__ void k3_ldlt_solve(const K3Smem &m, int n, long long *prof) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *A = m.A;
    PROF_ONLY(if (tid == 0) prof[9] = clk_fenced();)
    // SVecI = (diag + 10)^-1/2 (:326-327); Eigen's pivot order = descending |diag| of the scaled matrix
    if (tid < n) {
        const double dg = m.vd[tid];
        const double sv = 1.0 / sqrt(dg + 10.0);
        m.vS[tid] = sv;
        m.vd[tid] = fabs(dg * sv * sv);
    }
    __syncthreads();
    PROF_ONLY(if (tid == 0) prof[10] = clk_fenced();)
    for (int i4 = tid; i4 < ((4 * n + 31) & ~31); i4 += K3_THREADS) {      // rank sort, 4 threads per row; whole warps: the shuffles are full-mask
        const int i = i4 >> 2, q = i4 & 3;
        int rank = 0;
        if (i < n) {
            const double di = m.vd[i];
            for (int j = q; j < n; j += 4) {
                const double dj = m.vd[j];
                rank += (dj > di) || (dj == di && j < i);
            }
        }
        rank += __shfl_xor_sync(0xffffffffu, rank, 1);
        rank += __shfl_xor_sync(0xffffffffu, rank, 2);
        if (q == 0 && i < n) m.perm[rank] = i;
    }
    __syncthreads();
    PROF_ONLY(if (tid == 0) prof[11] = clk_fenced();)
    // A = P (S A0 S) P^T (both triangles: the upper one is never read, writing it keeps the copy branch-free); row n = P S b.
    // A lane's columns (lane, lane+32, lane+64) are the same for every row: their permutation / scale / rhs entries are loaded once,
    // and the 5 rows of a warp are independent chains (columns / rows past the end are clamped: duplicate stores of equal values).
    {
        int pcv[3];
        double svc[3], bvc[3];
#pragma unroll
        for (int q = 0; q < 3; q++) {
            pcv[q] = m.perm[min(lane + 32 * q, n - 1)];
            svc[q] = m.vS[pcv[q]];
            bvc[q] = m.vb[pcv[q]];
        }
#pragma unroll
        for (int it = 0; it < (K3_NP + 1 + K3_THREADS / 32 - 1) / (K3_THREADS / 32); it++) {
            const int r = min(warp + it * (K3_THREADS / 32), n);
            const bool rhs = r == n;
            const int pr = m.perm[rhs ? 0 : r];
            const double sr = m.vS[pr];
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const int pc = pcv[q];
                const int hi = max(pr, pc), lo = min(pr, pc);                 // lower triangle of the input (row hi, column lo)
                const double v = (sr * m.A0[lo * K3_A0LD + hi]) * svc[q];
                A[r * K3_LD + min(lane + 32 * q, n - 1)] = rhs ? bvc[q] * svc[q] : v;
            }
        }
    }
    __syncthreads();
    PROF_ONLY(if (tid == 0) prof[0] = clk_fenced();)
    // ---- blocked in-place LDL^T of the matrix augmented with the right-hand side as row n
    const int nblk = n / K3_NB;
    for (int kb = 0; kb < nblk; kb++) {
        const int k0 = kb * K3_NB, m0 = k0 + K3_NB;
        double *Wp = m.Wp + (kb & 1) * K3_NB * K3_WPLD;
        PROF_ONLY(const long long tq0 = clk_fenced();)
        if (tid >= k0 && tid <= n) k3_panel_row(A, Wp, k0, tid);
        PROF_ONLY(if (kb == 4 && tid == 16) prof[4] = clk_fenced() - tq0;)
        // meanwhile the helper warps apply the PREVIOUS block step to the columns right of this panel (far update)
        // (a clamped, branch-free variant of this loop and of the near update below was measured slower: the wasted elements cost
        // more than the branches they remove; the single-warp back-substitution is the other way round)
        if (tid >= 96 && kb > 0) {
            const int pk0 = k0 - K3_NB;
            const double *Wq = m.Wp + ((kb - 1) & 1) * K3_NB * K3_WPLD;
            const int t = tid - 96;
            for (int i = m0 + (t >> 4); i <= n; i += (K3_THREADS - 96) / 16) {
                double li[K3_NB];
#pragma unroll
                for (int c = 0; c < K3_NB; c++) li[c] = A[i * K3_LD + pk0 + c];
                const int jmax = min(i, n - 1);
                for (int j = m0 + (t & 15); j <= jmax; j += 16) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int c = 0; c < K3_NB; c += 2) { s0 = fma(li[c], Wq[c * K3_WPLD + j], s0); s1 = fma(li[c + 1], Wq[(c + 1) * K3_WPLD + j], s1); }
                    A[i * K3_LD + j] -= (s0 + s1);
                }
            }
        }
        PROF_ONLY(if (kb == 4 && tid == 96) prof[5] = clk_fenced() - tq0;)
        PROF_ONLY(const long long tq1 = clk_fenced();)
        __syncthreads();
        PROF_ONLY(const long long tq2 = clk_fenced(); if (kb == 4 && tid == 16) prof[6] = tq2 - tq1;)
        // near update: the next panel's columns, one element per thread: A[i][j] -= sum_c L(i,c) W(j,c), m0 <= j < m0 + NB, j <= i <= n
        if (m0 < n) {
            for (int e = tid; e < (n + 1 - m0) * K3_NB; e += K3_THREADS) {
                const int i = m0 + e / K3_NB, j = m0 + e % K3_NB;
                if (j <= i) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int c = 0; c < K3_NB; c += 2) {
                        s0 = fma(A[i * K3_LD + k0 + c], Wp[c * K3_WPLD + j], s0);
                        s1 = fma(A[i * K3_LD + k0 + c + 1], Wp[(c + 1) * K3_WPLD + j], s1);
                    }
                    A[i * K3_LD + j] -= (s0 + s1);
                }
            }
        }
        PROF_ONLY(const long long tq3 = clk_fenced(); if (kb == 4 && tid == 16) prof[7] = tq3 - tq2;)
        __syncthreads();
        PROF_ONLY(if (kb == 4 && tid == 16) prof[8] = clk_fenced() - tq3;)
    }
    PROF_ONLY(if (tid == 0) prof[1] = clk_fenced();)
    // ---- backward solve L^T x = z by ONE warp, the vector in registers: lane owns rows lane, lane+32, lane+64. Row n holds
    // z = D^-1 L^-1 b (unscaled where the pivot was invalid); Eigen's solve applies the pseudo-inverse of D.
    if (warp == 0) {
        double z[3];
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const int iv = lane + 32 * q, i = min(iv, n - 1);
            const double dk = A[i * K3_LD + i], zi = A[n * K3_LD + i];
            z[q] = (iv < n && fabs(dk) > 2.2250738585072014e-308) ? zi : 0.0;
        }
        // the factor entries a step needs do not depend on the unknowns. The diagonal block's entries, which the step's chain needs
        // first, are loaded one step ahead (software pipeline); the update columns are requested at the top of the step and arrive
        // while the block is solved: a step is the dependency chain shuffle -> 3 FMAs -> update only
        double Lb[K3_NB * (K3_NB - 1) / 2];
#define K3_BS_LOAD_LB(kq_)                                                                                          \
        do {                                                                                                        \
            const int kl_ = (kq_) * K3_NB;                                                                          \
            _Pragma("unroll") for (int j = 1; j < K3_NB; j++)                                                       \
                _Pragma("unroll") for (int c = 0; c < j; c++) Lb[j * (j - 1) / 2 + c] = A[(kl_ + j) * K3_LD + kl_ + c]; \
        } while (0)
        K3_BS_LOAD_LB(nblk - 1);
#pragma unroll 1
        for (int kb = nblk - 1; kb >= 0; kb--) {
            const int k0 = kb * K3_NB, sl = k0 >> 5, l0 = k0 & 31;
            const double zsel = (sl == 0) ? z[0] : (sl == 1) ? z[1] : z[2];
            double x[K3_NB], cLb[K3_NB * (K3_NB - 1) / 2], Lu[3][K3_NB];
#pragma unroll
            for (int e = 0; e < K3_NB * (K3_NB - 1) / 2; e++) cLb[e] = Lb[e];
#pragma unroll
            for (int c = 0; c < K3_NB; c++) x[c] = shfl_f64(zsel, l0 + c);
#pragma unroll
            for (int q = 0; q < 3; q++)
#pragma unroll
                for (int c = 0; c < K3_NB; c++) Lu[q][c] = A[(k0 + c) * K3_LD + min(lane + 32 * q, n - 1)];
            K3_BS_LOAD_LB(max(kb - 1, 0));
            // x_c = z_c - sum_{j > c} L(k0+j, k0+c) x_j, solved redundantly by every lane
#pragma unroll
            for (int j = K3_NB - 1; j >= 1; j--)
#pragma unroll
                for (int c = 0; c < j; c++) x[c] = fma(-cLb[j * (j - 1) / 2 + c], x[j], x[c]);
            // the block's unknowns go back to their owner lanes; earlier rows lose this block's contribution
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const int i = lane + 32 * q;
                const double s0 = fma(Lu[q][1], x[1], Lu[q][0] * x[0]), s1 = fma(Lu[q][3], x[3], Lu[q][2] * x[2]);
                double xo = (i < k0) ? z[q] - (s0 + s1) : z[q];
#pragma unroll
                for (int c = 0; c < K3_NB; c++) xo = (i == k0 + c) ? x[c] : xo;
                z[q] = xo;
            }
        }
#undef K3_BS_LOAD_LB
        // x = S P^T xp
#pragma unroll
        for (int q = 0; q < 3; q++) {
            const int i = lane + 32 * q;
            if (i < n) { const int pi = m.perm[i]; m.vx[pi] = z[q] * m.vS[pi]; }
        }
    }
    __syncthreads();
    PROF_ONLY(if (tid == 0) prof[2] = clk_fenced();)
}

__global__ void __launch_bounds__(K3_THREADS) k3_solve_step(WinState *ws, SolveBufs sb, int flags, int *iteration_dev, const double *sel_red, int sel_n,
                                                            long long *sel_dbg) {
    extern __shared__ double sm3[];
    if (blockIdx.x == 1) {
        // FullSystem::setNewFrameEnergyTH for the linearisation that produced the system being solved: its result is first read by the
        // NEXT linearisation, so it runs beside the solver (another SM) instead of in front of it inside the stitch kernel
        pdl_launch_dependents();
        pdl_wait();
        // ... and it publishes the system being solved as lastHS (EnergyFunctional.cc:285), a 37 KB copy the solver CTA does not need
        if (flags & K3F_SOLVE) {
            const int nn = ws->n * ws->n;
            for (int e = threadIdx.x; e < nn; e += K3_THREADS) sb.lastHS[e] = sb.HSg[e];
        }
        k2_select_body(sel_red, sel_n, ws, sm3, sel_dbg);
        return;
    }
    const K3Smem m = k3_carve(sm3);
    K3Frames *S = m.S;
    const int nF = ws->nF, n = ws->n, tid = threadIdx.x;
    pdl_launch_dependents();
    // ---- before pdl_wait: data that is constant for the whole window (set_frames): the f32 adjoints (xAd, adHTdeltaF) and the
    // null-space projector, as asynchronous 16-byte copies
    for (int e = tid; e < nF * nF * 16; e += K3_THREADS) {
        cp_async16(m.adH + 4 * e, &ws->adHostF[0][0] + 4 * e);
        cp_async16(m.adT + 4 * e, &ws->adTargetF[0][0] + 4 * e);
    }
    for (int e = tid; e < n * n / 2; e += K3_THREADS) cp_async16(m.Pns + 2 * e, sb.Pns + 2 * e);
    pdl_wait();
    // ---- one round trip for everything the previous kernels produced: the assembled system (k2b_stitch, do_assemble;
    // EnergyFunctional.cc:257,283-291: HFinal_top, its diagonal, bFinal_top, HFinal_top - H_sc), the frame / calibration records
    const int iteration = *iteration_dev;
    constexpr int K3_HSCOPY = (MAXN * MAXN + K3_THREADS - 1) / K3_THREADS;
    double hs_pre[K3_HSCOPY], b_pre = 0.0;
    const bool copy_hs = gridDim.x == 1;      // with a second CTA in the grid (the Gauss-Newton loop) that one copies HFinal_top - H_sc to lastHS
    float nid_pre = 0.f, num_pre = 1.f, tho_pre = 0.f;      // doStepFromBackup's canbreak inputs (thread 0 only)
    if (flags & K3F_SOLVE) {
        for (int cc = tid >> 5; cc < n; cc += K3_THREADS / 32)        // a warp per column (no division), 8-byte copies: the padded columns are not 16-byte aligned
            for (int r = tid & 31; r < n; r += 32) cp_async8(m.A0 + cc * K3_A0LD + r, sb.A0g + cc * n + r);
        if (tid < n) { m.vd[tid] = sb.dg[tid]; b_pre = sb.bFg[tid]; m.vb[tid] = b_pre; }
        if (copy_hs) {
#pragma unroll
            for (int k = 0; k < K3_HSCOPY; k++) {
                const int e = tid + k * K3_THREADS;
                hs_pre[k] = (e < n * n) ? sb.HSg[e] : 0.0;
            }
        }
    }
    for (int e = tid; e < (int) (sizeof(K3Frames) / 8); e += K3_THREADS) cp_async8((char *) S + 8 * e, (const char *) ws->fr + 8 * e);
    if (tid == 0) { nid_pre = ws->sumNID; num_pre = ws->numID; tho_pre = ws->S.thOptIterations; }
#ifdef LDSO_B200_PROFILE
    int dbgi = 0;
    long long prof[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define K3_STAMP() do { if (tid == 0) ws->dbg[dbgi] = clk_fenced(); dbgi++; } while (0)
    if (tid == 0) {      // wall-clock timeline of one iteration: K3 span here, K2a/K2b spans by atomics
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        ws->dbg[14] = (long long) gt;
        ws->dbg[16] = 0x7fffffffffffffffLL; ws->dbg[17] = 0; ws->dbg[18] = 0x7fffffffffffffffLL; ws->dbg[19] = 0;
    }
#else
    long long *prof = nullptr;
#define K3_STAMP() do { } while (0)
#endif
    K3_STAMP();   // 0
    cp_async_wait_all();
    __syncthreads();
    K3_STAMP();   // 1: inputs staged

    if (flags & K3F_BACKUP) {
        if (tid < nF) for (int i = 0; i < 10; i++) S->fr[tid].state_backup[i] = S->fr[tid].state[i];
        if (tid == 32) for (int i = 0; i < 4; i++) S->calib.value_backup[i] = S->calib.value[i];
        __syncthreads();
    }
    if (flags & K3F_SOLVE) {
        // The system being solved now becomes the public lastHS / lastbS (EnergyFunctional.cc:285,:335).
        if (copy_hs) {
#pragma unroll
            for (int k = 0; k < K3_HSCOPY; k++) {
                const int e = tid + k * K3_THREADS;
                if (e < n * n) sb.lastHS[e] = hs_pre[k];
            }
        }
        if (tid < n) sb.lastbS[tid] = b_pre;
        k3_ldlt_solve(m, n, prof);
#ifdef LDSO_B200_PROFILE
        if (tid == 0) for (int k = 0; k < 4; k++) ws->dbg[20 + k] = prof[k];
        if (tid == 0) for (int k = 9; k < 12; k++) ws->dbg[31 + k] = prof[k];
        if (tid == 16) for (int k = 4; k < 9; k++) if (k != 5) ws->dbg[20 + k] = prof[k];
        if (tid == 96) ws->dbg[25] = prof[5];
#endif
        K3_STAMP();   // 2: solved
        // orthogonalize(&x, 0) when iteration >= 2 (SOLVER_ORTHOGONALIZE_X_LATER, :339-343): x -= NNpiTS x (NNpiTS is symmetric)
        if (iteration >= 2) {
            for (int i4 = tid; i4 < ((4 * n + 31) & ~31); i4 += K3_THREADS) {
                const int i = i4 >> 2, q = i4 & 3;
                double s0 = 0.0;
                if (i < n) for (int c = q; c < n; c += 4) s0 = fma(m.Pns[i * n + c], m.vx[c], s0);
                s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
                s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
                if (q == 0 && i < n) m.vd[i] = s0;
            }
            __syncthreads();
            if (tid < n) m.vx[tid] -= m.vd[tid];
            __syncthreads();
        }
        K3_STAMP();   // 3: orthogonalised
    }
    if ((flags & K3F_SOLVE) && (flags & K3F_STEP)) {
        // ---- fused tail (the Gauss-Newton loop): resubstituteF_MT's frame part (:495-507), doStepFromBackup's frame / calibration part
        // (FullSystem.cc:1588-1597,1617-1627) and setDeltaF's adHTdeltaF only need x, so they run side by side on different warps:
        //   warp 0     frame steps, states, SE3::exp (the long chain)
        //   warp 1     calibration step + setValue + cDeltaF, canbreak
        //   warps 2..  lastX, xAd, adHTdeltaF
        const int warp = tid >> 5, lane = tid & 31;
        if (warp == 0) {
            if (lane < nF) {
                FrameDev &f = S->fr[lane];
                for (int i = 0; i < 8; i++) f.step[i] = -m.vx[CPARS + 8 * lane + i];
                f.step[8] = f.step[9] = 0.0;
                for (int i = 0; i < 10; i++) f.state[i] = f.state_backup[i] + f.step[i];
                frames_exp_part(S, lane);
            }
        } else if (warp == 1) {
            if (lane < CPARS) {
                S->calib.step[lane] = -m.vx[lane];
                ws->cstep[lane] = (float) m.vx[lane];
            }
            __syncwarp();
            if (lane == 0) {
                CalibDev &c = S->calib;
                double nv[4];
                for (int i = 0; i < 4; i++) nv[i] = c.value_backup[i] + c.step[i];
                calib_set_value(c, nv);
                for (int i = 0; i < 4; i++) c.cDeltaF[i] = (float) (c.value[i] - c.value_zero[i]);
                float sumA = 0, sumB = 0, sumT = 0, sumR = 0;
                for (int h = 0; h < nF; h++) {
                    double st[8];
                    for (int i = 0; i < 8; i++) st[i] = -m.vx[CPARS + 8 * h + i];
                    sumA += st[6] * st[6];
                    sumB += st[7] * st[7];
                    sumT += st[0] * st[0] + st[1] * st[1] + st[2] * st[2];
                    sumR += st[3] * st[3] + st[4] * st[4] + st[5] * st[5];
                }
                sumA /= nF; sumB /= nF; sumR /= nF; sumT /= nF;
                const float sumNID = nid_pre / num_pre;
                const float thO = tho_pre;
                ws->canbreak = (sqrtf(sumA) < 0.0005 * thO && sqrtf(sumB) < 0.00005 * thO && sqrtf(sumR) < 0.00005 * thO &&
                                sqrtf(sumT) * sumNID < 0.00005 * thO) ? 1 : 0;
            }
        } else {
            const int t0 = tid - 64, nt = K3_THREADS - 64;
            for (int e = t0; e < n; e += nt) sb.lastX[e] = m.vx[e];
            for (int o = t0; o < nF * nF * 8; o += nt) {
                const int q = o >> 3, j = o & 7, h = q / nF, t = q % nF;     // xAd[nFrames*h + t]
                const float *AH = m.adH + (h + nF * t) * 64, *AT = m.adT + (h + nF * t) * 64;
                float s1 = 0.f, s2 = 0.f;
                for (int i = 0; i < 8; i++) s1 += (float) m.vx[CPARS + 8 * h + i] * AH[i * 8 + j];
                for (int i = 0; i < 8; i++) s2 += (float) m.vx[CPARS + 8 * t + i] * AT[i * 8 + j];
                ws->xAd[nF * h + t][j] = s1 + s2;
            }
            frames_adHTdelta(S, ws, m.adH, m.adT, t0, nt, m.vx);
        }
        __syncthreads();
        K3_STAMP();   // 4: x distributed, frame poses refreshed
        frames_pairs_part(S, ws, false);      // (stage_out starts with the barrier that closes this phase)
    } else {
    if (flags & K3F_SOLVE) {
        if (tid < n) sb.lastX[tid] = m.vx[tid];
        // resubstituteF_MT frame part (:495-507)
        if (tid < CPARS) {
            S->calib.step[tid] = -m.vx[tid];
            ws->cstep[tid] = (float) m.vx[tid];
        }
        if (tid >= 32 && tid < 32 + nF) {
            const int h = tid - 32;
            for (int i = 0; i < 8; i++) S->fr[h].step[i] = -m.vx[CPARS + 8 * h + i];
            S->fr[h].step[8] = S->fr[h].step[9] = 0.0;
        }
        for (int o = tid; o < nF * nF * 8; o += K3_THREADS) {
            const int q = o >> 3, j = o & 7, h = q / nF, t = q % nF;     // xAd[nFrames*h + t]
            const float *AH = m.adH + (h + nF * t) * 64, *AT = m.adT + (h + nF * t) * 64;
            float s1 = 0.f, s2 = 0.f;
            for (int i = 0; i < 8; i++) s1 += (float) m.vx[CPARS + 8 * h + i] * AH[i * 8 + j];
            for (int i = 0; i < 8; i++) s2 += (float) m.vx[CPARS + 8 * t + i] * AT[i * 8 + j];
            ws->xAd[nF * h + t][j] = s1 + s2;
        }
        __syncthreads();
    }
    K3_STAMP();   // 4: xAd done
    if (flags & K3F_STEP) {
        // doStepFromBackup(1,1,1,1,1), frame/calib part (FullSystem.cc:1588-1597,1617-1627)
        if (tid == 0) {
            double nv[4];
            for (int i = 0; i < 4; i++) nv[i] = S->calib.value_backup[i] + S->calib.step[i];
            calib_set_value(S->calib, nv);
            float sumA = 0, sumB = 0, sumT = 0, sumR = 0;
            for (int h = 0; h < nF; h++) {
                const double *st = S->fr[h].step;
                sumA += st[6] * st[6];
                sumB += st[7] * st[7];
                sumT += st[0] * st[0] + st[1] * st[1] + st[2] * st[2];
                sumR += st[3] * st[3] + st[4] * st[4] + st[5] * st[5];
            }
            sumA /= nF; sumB /= nF; sumR /= nF; sumT /= nF;
            const float sumNID = nid_pre / num_pre;
            const float thO = tho_pre;
            ws->canbreak = (sqrtf(sumA) < 0.0005 * thO && sqrtf(sumB) < 0.00005 * thO && sqrtf(sumR) < 0.00005 * thO &&
                            sqrtf(sumT) * sumNID < 0.00005 * thO) ? 1 : 0;
        }
        if (tid >= 32 && tid < 32 + nF) {
            FrameDev &f = S->fr[tid - 32];
            for (int i = 0; i < 10; i++) f.state[i] = f.state_backup[i] + f.step[i];
        }
        __syncthreads();
        frames_refresh(S, ws, false, m.adH, m.adT);
    }
    }
    K3_STAMP();   // 5: frames refreshed
    stage_out(S, ws);
    K3_STAMP();   // 6
#ifdef LDSO_B200_PROFILE
    if (tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        ws->dbg[15] = (long long) gt;
    }
#endif
    if ((flags & K3F_SOLVE) && tid == 0) *iteration_dev = iteration + 1;
}
#define K3_SMEM_BYTES (K3_SMEM_DOUBLES * sizeof(double) + (K3_NP + 8) * sizeof(int) + sizeof(K3Frames) + 2 * MAXPAIR * 64 * sizeof(float) + 64)
