// K3 — one CTA: EnergyFunctional::solveSystemF (EnergyFunctional.cc:240-351, default solver mode), the frame/calib
// part of resubstituteF_MT (:491-507), FullSystem::backupState / doStepFromBackup (FullSystem.cc:1587-1676),
// FrameHessian::setState, FrameFramePrecalc::Set for all nF^2 pairs and EnergyFunctional::setDeltaF.
//
// Everything here is a serial dependency chain on a 68x68 system, so the design goal is latency, not throughput:
//   * the frame/calib records are staged in shared memory once (one global round trip instead of dozens);
//   * the LDL^T is BLOCKED (8 columns per step): a step is [8x8 diagonal block by one thread, in registers] ->
//     [panel rows, one thread per row] -> [trailing update, all threads], 3 barriers per 8 columns instead of
//     2 barriers per column;
//   * Eigen's LDLT pivots on the largest remaining |diagonal| of the INPUT matrix (its left-looking update never
//     touches later diagonal entries before they are chosen), i.e. a descending-|diag| order: computed by a rank
//     sort and applied as a symmetric permutation before the (then unpivoted) blocked factorisation.
#pragma once
#include "common.cuh"
#include "se3_math.cuh"
#include "ba_k2.cuh"

#define K3F_SOLVE 1
#define K3F_STEP 2
#define K3F_BACKUP 4
#define K3_THREADS 512
#define K3_NP ((MAXN + 7) & ~7)     // system dimension padded to whole 8x8 blocks (identity padding)
#define K3_LD (K3_NP + 1)
#define K3_NB 8
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

// FrameHessian::setState (FrameHessian.h:78-91), FrameFramePrecalc::Set for all pairs (FrameFramePrecalc.cc:6-35),
// EnergyFunctional::setDeltaF frame part (EnergyFunctional.cc:403-429). Frame records live in shared memory (S);
// the pair records are written to global. Called by all threads of a CTA with >= 128 threads.
__device__ void frames_refresh(K3Frames *S, WinState *ws, bool full, const float *adHF, const float *adTF) {
    const int nF = ws->nF, tid = threadIdx.x;
    if (tid < nF) {
        FrameDev &f = S->fr[tid];
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
    if (tid == 64) {
        CalibDev &c = S->calib;
        for (int i = 0; i < 4; i++) c.cDeltaF[i] = (float) (c.value[i] - c.value_zero[i]);
    }
    __syncthreads();
    // pair records, three short parallel phases instead of one long per-pair chain:
    // F2 (pair,row): one row of R = R_t R_h^T and of t = t_t - R t_h, in double; F3 (pair,row): one row of K R K^-1 and K t in
    // float (the reference's Mat33f products, FrameFramePrecalc.cc:21-31); F3' (pair): affine brightness transfer.
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
    for (int o = tid; o < nF * nF * 4; o += blockDim.x) {
        const int q = o >> 2, i = o & 3;
        PairRec &pc = ws->pair[q];
        if (i < 3) {
            const CalibDev &c = S->calib;
            const float K[9] = {c.fxl, 0, c.cxl, 0, c.fyl, c.cyl, 0, 0, 1};
            float Ki[9];
            m33f_inverse(K, Ki);
            const float *Rf = sRf[q];
            float tmp[3];
            for (int j = 0; j < 3; j++) {
                float s2 = K[i * 3 + 0] * Rf[0 * 3 + j];
                s2 += K[i * 3 + 1] * Rf[1 * 3 + j];
                s2 += K[i * 3 + 2] * Rf[2 * 3 + j];
                tmp[j] = s2;
            }
            for (int j = 0; j < 3; j++) {
                float s2 = tmp[0] * Ki[0 * 3 + j];
                s2 += tmp[1] * Ki[1 * 3 + j];
                s2 += tmp[2] * Ki[2 * 3 + j];
                pc.KRKi[i * 3 + j] = s2;
            }
            float s2 = K[i * 3 + 0] * sTf[q][0];
            s2 += K[i * 3 + 1] * sTf[q][1];
            s2 += K[i * 3 + 2] * sTf[q][2];
            pc.Kt[i] = s2;
            PairRecFull &pf = ws->pairFull[q];
            for (int j = 0; j < 3; j++) pf.RTll[i * 3 + j] = Rf[i * 3 + j];
            pf.tTll[i] = sTf[q][i];
        } else {
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
    // adHTdeltaF (EnergyFunctional.cc:406-414): one (pair, column) output per thread pass
    for (int o = tid; o < nF * nF * 8; o += blockDim.x) {
        const int q = o >> 3, j = o & 7, h = q % nF, t = q / nF;
        const FrameDev &fh = S->fr[h], &ft = S->fr[t];
        const float *AH = adHF + q * 64, *AT = adTF + q * 64;
        float s1 = 0.f, s2 = 0.f;
        for (int i = 0; i < 8; i++) s1 += (float) (fh.state[i] - fh.state_zero[i]) * AH[i * 8 + j];
        for (int i = 0; i < 8; i++) s2 += (float) (ft.state[i] - ft.state_zero[i]) * AT[i * 8 + j];
        ws->adHTdeltaF[q][j] = s1 + s2;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(128) k_frames_refresh(WinState *ws) {
    __shared__ K3Frames S;
    stage_in(&S, ws);
    frames_refresh(&S, ws, true, &ws->adHostF[0][0], &ws->adTargetF[0][0]);
    stage_out(&S, ws);
}

// Factor the 8x8 diagonal block at (k0,k0) in place (lower): L below the diagonal (unit), D on it. The matrix is
// padded to whole blocks, so there is no partial-block predicate anywhere on this serial path.
// Executed by ONE WARP: lane j < 8 keeps row j of the block in registers; a step is one broadcast of the pivot,
// one reciprocal, and 7-k shuffles of the unscaled column. (L = W * (1/d): <= 1 ulp from Eigen's W / d.)
__device__ __forceinline__ double shfl_f64(double v, int src) {      // low word first: lands in an aligned register pair
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(0xffffffffu, lo, src);
    hi = __shfl_sync(0xffffffffu, hi, src);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ void ldlt_diag_block_warp(double *A, double *vinv, int k0, int lane) {
    double a[K3_NB];
    const int row = lane & 7;
#pragma unroll
    for (int c = 0; c < K3_NB; c++) a[c] = A[(k0 + row) * K3_LD + k0 + c];
#pragma unroll
    for (int k = 0; k < K3_NB; k++) {
        const double w = a[k];                      // unscaled column entry of this lane's row (rows >= k)
        // all exchanges of this step are issued before the reciprocal: only rcp -> mul -> fma stays on the chain
        const double dk = shfl_f64(w, k);
        double wj[K3_NB];
#pragma unroll
        for (int j = k + 1; j < K3_NB; j++) wj[j] = shfl_f64(w, j);
        const bool valid = fabs(dk) > 0.0;
        const double inv = valid ? __drcp_rn(dk) : 1.0;
        const double l = w * inv;
        // unconditional: for row < j this touches only the (never read, never stored) upper part of the lane's row
#pragma unroll
        for (int j = k + 1; j < K3_NB; j++) a[j] -= l * wj[j];
        if (row > k) a[k] = l;
        if (lane == 0) vinv[k0 + k] = inv;
    }
    if (lane < K3_NB) {
#pragma unroll
        for (int c = 0; c < K3_NB; c++) if (c <= lane) A[(k0 + lane) * K3_LD + k0 + c] = a[c];
    }
}

// z <- L11^{-1} z for the unit-lower block at (k0,k0); one warp, lane c holds z[k0+c]
__device__ __forceinline__ void trsv_lower_warp(const double *A, double *v, int k0, int bs, int lane) {
    const int c = lane & 7;
    double Lr[K3_NB];
#pragma unroll
    for (int j = 0; j < K3_NB; j++) Lr[j] = (c < bs && j < c) ? A[(k0 + c) * K3_LD + k0 + j] : 0.0;
    double z = (c < bs) ? v[k0 + c] : 0.0;
#pragma unroll
    for (int j = 0; j < K3_NB - 1; j++) {
        const double zj = shfl_f64(z, j);
        z -= Lr[j] * zj;          // Lr[j] == 0 for j >= c
    }
    if (lane < bs) v[k0 + lane] = z;
}
// x <- L11^{-T} x
__device__ __forceinline__ void trsv_lower_t_warp(const double *A, double *v, int k0, int bs, int lane) {
    const int c = lane & 7;
    double Lc[K3_NB];
#pragma unroll
    for (int j = 0; j < K3_NB; j++) Lc[j] = (j < bs && c < j) ? A[(k0 + j) * K3_LD + k0 + c] : 0.0;
    double x = (c < bs) ? v[k0 + c] : 0.0;
#pragma unroll
    for (int j = K3_NB - 1; j >= 1; j--) {
        const double xj = shfl_f64(x, j);
        x -= Lc[j] * xj;          // Lc[j] == 0 for j <= c
    }
    if (lane < bs) v[k0 + lane] = x;
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

__global__ void __launch_bounds__(K3_THREADS) k3_solve_step(WinState *ws, SolveBufs sb, int flags, int *iteration_dev) {
    extern __shared__ double sm3[];
    double *A = sm3;                        // [(K3_NP + 1)][K3_LD] permuted, scaled, padded system (+ rhs as row npad), factorised in place
    double *Wp = A + (K3_NP + 1) * K3_LD;    // [K3_NB][K3_WPLD] panel W = L*D of the current block step
    double *vinv = Wp + K3_NB * K3_WPLD;    // [K3_NP] reciprocal pivots
    double *vb = vinv + K3_NP;               // rhs / solution
    double *vS = vb + MAXN;                 // SVecI
    double *vd = vS + MAXN;                 // delta / temp
    double *vx = vd + MAXN;                 // x
    int *perm = (int *) (vx + MAXN);        // [MAXN]
    K3Frames *S = (K3Frames *) (perm + MAXN + 2);
    double *sPns = (double *) (S + 1);      // [n*n] null-space projector
    float *sAdH = (float *) (sPns + MAXN * MAXN);   // [MAXPAIR][64] adHostF, adTargetF (index h + nF*t): staged once per launch
    float *sAdT = sAdH + MAXPAIR * 64;
    const int nF = ws->nF, n = ws->n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();
    constexpr int K3_HSCOPY = (MAXN * MAXN + K3_THREADS - 1) / K3_THREADS;
    double b_pre = 0.0, dg_pre = 0.0, hs_pre[K3_HSCOPY];
    const int npad = (n + K3_NB - 1) & ~(K3_NB - 1);      // identity-padded to whole 8x8 blocks; the rhs is row npad
    // the f32 adjoints (constant for the window) feed xAd (solve) and adHTdeltaF (step) at the very end: asynchronous
    // copies, 16 bytes each, issued before pdl_wait
    for (int e = tid; e < nF * nF * 16; e += K3_THREADS) {
        cp_async16(sAdH + 4 * e, &ws->adHostF[0][0] + 4 * e);
        cp_async16(sAdT + 4 * e, &ws->adTargetF[0][0] + 4 * e);
    }
    pdl_wait();
    const int iteration = *iteration_dev;
    float nid_pre = 0.f, num_pre = 1.f, tho_pre = 0.f;      // doStepFromBackup's canbreak inputs (thread 0 only)
    if (tid == 0) { nid_pre = ws->sumNID; num_pre = ws->numID; tho_pre = ws->S.thOptIterations; }
    if (flags & K3F_SOLVE) {
        // What the stitch kernel left behind (k2b_stitch, do_assemble; EnergyFunctional.cc:257,283-291): HFinal_top and its
        // diagonal, HFinal_top - H_sc (lastHS once solved), bFinal_top (lastbS). The diagonal comes first: it fixes SVecI
        // and the pivot order, and with those the matrix is gathered straight into its permuted place by asynchronous
        // global->shared copies (LDGSTS) that land while the frame state is staged.
        if (tid < n) { dg_pre = sb.dg[tid]; b_pre = sb.bFg[tid]; }
#pragma unroll
        for (int k = 0; k < K3_HSCOPY; k++) {
            const int e = tid + k * K3_THREADS;
            hs_pre[k] = (e < n * n) ? sb.HSg[e] : 0.0;
        }
        if (iteration >= 2)
            for (int e = tid; e < n * n; e += K3_THREADS) cp_async8(sPns + e, sb.Pns + e);
        // SVecI = (diag + 10)^-1/2 (:326-327); Eigen's pivot order = descending |diag| of the scaled matrix
        if (tid < n) {
            const double sv = 1.0 / sqrt(dg_pre + 10.0);
            vS[tid] = sv;
            vd[tid] = fabs(dg_pre * sv * sv);          // |diagonal| of the scaled matrix
            vb[tid] = b_pre;
        }
        __syncthreads();
        {   // rank sort: 4 threads per row fold a quarter of the comparisons each
            const int i = tid >> 2, q = tid & 3;
            int rank = 0;
            if (i < n) {
                const double di = vd[i];
                for (int j = q; j < n; j += 4) {
                    const double dj = vd[j];
                    rank += (dj > di) || (dj == di && j < i);
                }
            }
            rank += __shfl_xor_sync(0xffffffffu, rank, 1);
            rank += __shfl_xor_sync(0xffffffffu, rank, 2);
            if (i < n && q == 0) perm[rank] = i;
        }
        __syncthreads();
        // lower triangle (+ the whole diagonal blocks) of P A0 P^T; padding = identity
        for (int r = warp; r < npad; r += K3_THREADS / 32) {
            const int pr = (r < n) ? perm[r] : 0;
#pragma unroll
            for (int cc = 0; cc < (K3_NP + 31) / 32; cc++) {
                const int c = lane + 32 * cc;
                if (c < npad && (c <= r || (c >> 3) == (r >> 3))) {
                    if (r < n && c < n) cp_async8(A + r * K3_LD + c, sb.A0g + (size_t) perm[c] * n + pr);
                    else A[r * K3_LD + c] = (r == c) ? 1.0 : 0.0;
                }
            }
        }
    }
    int dbgi = 0;
#define K3_STAMP() do { if (tid == 0) ws->dbg[dbgi] = clk_fenced(); dbgi++; } while (0)
    K3_STAMP();
    if (tid == 0) {      // wall-clock timeline of one iteration (development aid): K3 span here, K2a/K2b spans by atomics
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        ws->dbg[14] = (long long) gt;
        ws->dbg[16] = 0x7fffffffffffffffLL; ws->dbg[17] = 0; ws->dbg[18] = 0x7fffffffffffffffLL; ws->dbg[19] = 0;
    }
    stage_in(S, ws);
    K3_STAMP();

    if (flags & K3F_BACKUP) {
        if (tid < nF) for (int i = 0; i < 10; i++) S->fr[tid].state_backup[i] = S->fr[tid].state[i];
        if (tid == 32) for (int i = 0; i < 4; i++) S->calib.value_backup[i] = S->calib.value[i];
        __syncthreads();
    }
    if (flags & K3F_SOLVE) {
        // The system being solved now becomes the public lastHS / lastbS (EnergyFunctional.cc:285,:335).
#pragma unroll
        for (int k = 0; k < K3_HSCOPY; k++) {
            const int e = tid + k * K3_THREADS;
            if (e < n * n) sb.lastHS[e] = hs_pre[k];
        }
        if (tid < n) sb.lastbS[tid] = b_pre;
        K3_STAMP();   // 2
        // A = P (S A0 S) P^T: every thread scales the elements it copied itself (visible to it after the wait); b' = P S b
        cp_async_wait_all();
        for (int r = warp; r < n; r += K3_THREADS / 32) {
            const double sr = vS[perm[r]];
#pragma unroll
            for (int cc = 0; cc < (K3_NP + 31) / 32; cc++) {
                const int c = lane + 32 * cc;
                if (c < n && (c <= r || (c >> 3) == (r >> 3))) A[r * K3_LD + c] = A[r * K3_LD + c] * sr * vS[perm[c]];
            }
        }
        if (tid < npad) A[npad * K3_LD + tid] = (tid < n) ? vb[perm[tid]] * vS[perm[tid]] : 0.0;
        __syncthreads();

        K3_STAMP();   // 3: scaled+permuted
        // ---- blocked in-place LDL^T (lower) of the matrix AUGMENTED with the right-hand side as row n:
        // the panel/trailing steps then leave D^-1 L^-1 b in that row, i.e. the forward solve comes for free.
        // Look-ahead schedule: while warps 1.. apply block step k to the rows below the NEXT diagonal block, warp 0
        // applies it to that block and factorises it right away, so the serial 8-pivot chain of step k+1 is hidden
        // behind the trailing update of step k (two barriers per step).
        long long tp = 0, tw = 0, tb1 = 0, tb2 = 0, tq;     // this thread's clocks in panel / barrier / trailing(+diag) / barrier (development aid)
        // (the loop starts one block early: that pass only factorises diagonal block 0, so the serial pivot code exists
        // once -- a second inlined copy in front of the loop costs ~19 k cycles of cold instruction fetch per launch)
        for (int k0 = -K3_NB; k0 < npad; k0 += K3_NB) {
            const int m0 = k0 + K3_NB;
            tq = clk_fenced();
            if (tid == 0) ws->dbg[33 + (k0 >> 3)] = tq;
            if (k0 >= 0 && tid < npad + 1 - m0) {      // panel row i (incl. the rhs row npad): w = L*D (unscaled), l = L
                const int i = m0 + tid;
                // all operands first (the 28 entries of L11 are warp-uniform broadcasts), then the 8-step substitution
                double a[K3_NB], iv[K3_NB], Lb[K3_NB * (K3_NB - 1) / 2], w[K3_NB];
#pragma unroll
                for (int c = 0; c < K3_NB; c++) { a[c] = A[i * K3_LD + k0 + c]; iv[c] = vinv[k0 + c]; }
#pragma unroll
                for (int c = 1; c < K3_NB; c++)
#pragma unroll
                    for (int j = 0; j < c; j++) Lb[c * (c - 1) / 2 + j] = A[(k0 + c) * K3_LD + k0 + j];
#pragma unroll
                for (int c = 0; c < K3_NB; c++) {
                    double s0 = a[c], s1 = 0.0;
#pragma unroll
                    for (int j = 0; j < c; j += 2) s0 -= w[j] * Lb[c * (c - 1) / 2 + j];
#pragma unroll
                    for (int j = 1; j < c; j += 2) s1 -= w[j] * Lb[c * (c - 1) / 2 + j];
                    w[c] = s0 + s1;
                }
#pragma unroll
                for (int c = 0; c < K3_NB; c++) {
                    A[i * K3_LD + k0 + c] = w[c] * iv[c];
                    Wp[c * K3_WPLD + i] = w[c];
                }
            }
            { const long long t1 = clk_fenced(); tp += t1 - tq; tq = t1; }
            __syncthreads();
            { const long long t1 = clk_fenced(); tb1 += t1 - tq; tq = t1; }
            // trailing update A[i][j] -= sum_c L(i,c) W(j,c), m0 <= j <= min(i, npad-1), i <= npad
            if (warp == 0) {
                if (m0 < npad) {
                    if (k0 >= 0) {
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            const int e = lane + 32 * h, r = e >> 3, cc = e & 7;
                            const int i = m0 + r, j = m0 + cc;
                            double s0 = 0.0, s1 = 0.0;
#pragma unroll
                            for (int c = 0; c < K3_NB; c += 2) {
                                s0 += A[i * K3_LD + k0 + c] * Wp[c * K3_WPLD + j];
                                s1 += A[i * K3_LD + k0 + c + 1] * Wp[(c + 1) * K3_WPLD + j];
                            }
                            if (cc <= r) A[i * K3_LD + j] -= (s0 + s1);
                        }
                    }
                    __syncwarp();
                    ldlt_diag_block_warp(A, vinv, m0, lane);
                }
            } else if (k0 >= 0) {
                const int t = tid - 32;
                for (int i = m0 + K3_NB + (t >> 4); i <= npad; i += (K3_THREADS - 32) / 16) {
                    double li[K3_NB];
#pragma unroll
                    for (int c = 0; c < K3_NB; c++) li[c] = A[i * K3_LD + k0 + c];
                    const int jmax = min(i, npad - 1);
#pragma unroll 2
                    for (int j = m0 + (t & 15); j <= jmax; j += 16) {
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int c = 0; c < K3_NB; c += 2) { s0 += li[c] * Wp[c * K3_WPLD + j]; s1 += li[c + 1] * Wp[(c + 1) * K3_WPLD + j]; }
                        A[i * K3_LD + j] -= (s0 + s1);
                    }
                }
            }
            { const long long t1 = clk_fenced(); tw += t1 - tq; tq = t1; }
            __syncthreads();
            tb2 += clk_fenced() - tq;
        }
        if (tid == 0) { ws->dbg[20] = tp; ws->dbg[21] = tb1; ws->dbg[22] = tw; ws->dbg[23] = tb2; ws->dbg[33 + 9] = clk_fenced(); }
        if (tid == 32) { ws->dbg[24] = tp; ws->dbg[25] = tb1; ws->dbg[26] = tw; ws->dbg[27] = tb2; }
        if (tid == 496) { ws->dbg[28] = tp; ws->dbg[29] = tb1; ws->dbg[30] = tw; ws->dbg[31] = tb2; }
        K3_STAMP();   // 4: factorised
        // row n now holds D^-1 L^-1 b (unscaled where the pivot was invalid): Eigen's solve uses the pseudo-inverse of D
        if (tid < n) {
            const double dk = A[tid * K3_LD + tid];
            vb[tid] = (fabs(dk) > 2.2250738585072014e-308) ? A[npad * K3_LD + tid] : 0.0;
        }
        __syncthreads();
        // ---- backward solve L^T x = z, blocked from the last block up, column oriented: once a block of x is final its
        // contribution is removed from all earlier rows by one thread per row (no reduction on the critical path)
        for (int k0 = ((n - 1) / K3_NB) * K3_NB; k0 >= 0; k0 -= K3_NB) {
            const int bs = min(K3_NB, n - k0);
            if (warp == 0) trsv_lower_t_warp(A, vb, k0, bs, lane);
            __syncthreads();
            if (tid < k0) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int c = 0; c < K3_NB; c += 2) {
                    if (c < bs) s0 += A[(k0 + c) * K3_LD + tid] * vb[k0 + c];
                    if (c + 1 < bs) s1 += A[(k0 + c + 1) * K3_LD + tid] * vb[k0 + c + 1];
                }
                vb[tid] -= (s0 + s1);
            }
            __syncthreads();
        }
        K3_STAMP();   // 5: back-substituted
        if (tid < n) vx[perm[tid]] = vb[tid];
        __syncthreads();
        if (tid < n) vx[tid] *= vS[tid];
        __syncthreads();
        // orthogonalize(&x, 0) when iteration >= 2 (SOLVER_ORTHOGONALIZE_X_LATER, :339-343): x -= NNpiTS x
        if (iteration >= 2) {
            for (int r = warp; r < n; r += K3_THREADS / 32) {
                double s = 0.0;
                for (int c = lane; c < n; c += 32) s += sPns[r * n + c] * vx[c];      // NNpiTS is symmetric
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0) vd[r] = s;
            }
            __syncthreads();
            if (tid < n) vx[tid] -= vd[tid];
            __syncthreads();
        }
        K3_STAMP();   // 6: orthogonalised
        if (tid < n) sb.lastX[tid] = vx[tid];
        // resubstituteF_MT frame part (:495-507)
        if (tid < CPARS) {
            S->calib.step[tid] = -vx[tid];
            ws->cstep[tid] = (float) vx[tid];
        }
        if (tid >= 32 && tid < 32 + nF) {
            const int h = tid - 32;
            for (int i = 0; i < 8; i++) S->fr[h].step[i] = -vx[CPARS + 8 * h + i];
            S->fr[h].step[8] = S->fr[h].step[9] = 0.0;
        }
        for (int o = tid; o < nF * nF * 8; o += K3_THREADS) {
            const int q = o >> 3, j = o & 7, h = q / nF, t = q % nF;     // xAd[nFrames*h + t]
            const float *AH = sAdH + (h + nF * t) * 64, *AT = sAdT + (h + nF * t) * 64;
            float s1 = 0.f, s2 = 0.f;
            for (int i = 0; i < 8; i++) s1 += (float) vx[CPARS + 8 * h + i] * AH[i * 8 + j];
            for (int i = 0; i < 8; i++) s2 += (float) vx[CPARS + 8 * t + i] * AT[i * 8 + j];
            ws->xAd[nF * h + t][j] = s1 + s2;
        }
        __syncthreads();
    }
    K3_STAMP();   // 7: xAd done
    if (flags & K3F_STEP) {
        if (!(flags & K3F_SOLVE)) { cp_async_wait_all(); __syncthreads(); }     // the staged adjoints (the solve path waited already)
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
        frames_refresh(S, ws, false, sAdH, sAdT);
    }
    K3_STAMP();   // 8: frames refreshed
    stage_out(S, ws);
    K3_STAMP();   // 9
    if (tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        ws->dbg[15] = (long long) gt;
    }
    if ((flags & K3F_SOLVE) && tid == 0) *iteration_dev = iteration + 1;
}
#define K3_SMEM_BYTES (((K3_NP + 1) * K3_LD + K3_NB * K3_WPLD + 4 * MAXN + K3_NP) * sizeof(double) + (MAXN + 2) * sizeof(int) + sizeof(K3Frames) + MAXN * MAXN * sizeof(double) + 2 * MAXPAIR * 64 * sizeof(float) + 64)
