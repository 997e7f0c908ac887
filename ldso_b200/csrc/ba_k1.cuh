// K1 — fused [resubstitute + idepth step] + PointFrameResidual::linearize + applyRes + Hessian accumulation.
//
// One CTA per work item = up to `pts_per_item` consecutive points of ONE host keyframe (points are ordered by
// host, EnergyFunctional::makeIDX). Per item:
//   phase R  thread per point : EnergyFunctional::resubstituteFPt (EnergyFunctional.cc:518-547) with the x of the
//                               previous solve + FullSystem::doStepFromBackup's idepth update (FullSystem.cc:1607-1615)
//   phase A  8 lanes/residual : PointFrameResidual::linearize (Residuals.cc:13-214), lane = pattern pixel; the
//                               four bilinear taps are 16-byte texel loads; the 17 per-residual inner products are
//                               folded with 8-lane shuffles; applyRes/takeData (Residuals.h:70-87,123-128) fused;
//                               the compact per-residual record is staged in shared memory at slot [point][target]
//   phase B  warp per target  : AccumulatedTopHessianSSE::addPoint<mode> (AccumulatedTopHessian.cc:9-118): the 91
//                               unique entries of the 13x13 block of (host,target) live in 3 registers per lane
//   phase P  thread per point : Hdd/bd/Hcd sums, HdiF, bdSumF (AccumulatedSCHessian.cc:24-29)
//   phase C  4x4 register tile: AccumulatedSCHessianSSE::addPoint (AccumulatedSCHessian.cc:30-49): the host's
//                               (8nF)x(8nF) matrix D_h = sum_p HdiF v_p v_p^T, plus accE, accEB, accHcc, accbc
// The item's partial sums go to global memory (PART_* layout); K2a folds them over items in double.
// No tensor cores: the work is gather + rank-1/outer-product accumulation.
#pragma once
#include "common.cuh"

__constant__ int c_patx[8] = {0, -1, 1, -2, 0, 2, -1, 0};   // staticPattern[8], src/Setting.cc:221
__constant__ int c_paty[8] = {-2, -1, -1, 0, 0, 0, 1, 2};

__device__ __forceinline__ float group_sum8(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// index tables for the 91 packed entries of AccumulatorApprox (MatrixAccumulators.h:771-801):
// e < 55: upper triangle (r <= c) of the 10x10 [C|xi] block, row-major over r; 55..84: TopRight 10x3; 85..90: BotRight
__device__ __forceinline__ void tri10_rc(int e, int &r, int &c) {
    int rr = 0, rem = e;
    while (rem >= 10 - rr) { rem -= 10 - rr; rr++; }
    r = rr;
    c = rr + rem;
}

#define recs_dot(base, ri, k) (base)[2 * (ri) + (k)]

struct K1Shared {   // offsets (in floats) inside the dynamic shared memory block
    int recs, pair, ptin, ptout, misc;
};
__host__ __device__ inline K1Shared k1_layout(int pts_per_item) {
    K1Shared L;
    L.recs = 0;
    L.pair = L.recs + pts_per_item * MAXF * REC;
    L.ptin = L.pair + MAXF * 32;
    L.ptout = L.ptin + pts_per_item * 8;      // u v idepth idepth_zero priorF deltaF sel pad
    L.misc = L.ptout + pts_per_item * 2 * MAXF;   // phase P: HdiF bdSumF Hcd[4] w ngood (8/pt); phase R: 2 floats per residual
    return L;
}
#define K1_MISC_FLOATS (MAXF /*frameEnergyTH*/ + 8 /*calib*/ + MAXF * 8 /*xAd rows of host*/ + 4 /*cstep*/ + MAXF * 8 /*adHTdeltaF*/ + 4 /*cDeltaF*/ + 4)
__host__ __device__ inline size_t k1_smem_bytes(int pts_per_item) {
    return (size_t) (k1_layout(pts_per_item).misc + K1_MISC_FLOATS + 64) * sizeof(float);
}

__global__ void __launch_bounds__(K1_THREADS)
k1_linearize_accumulate(DevWindow d, const WinState *__restrict__ ws, int flags, const uint8_t *__restrict__ pt_sel) {
    extern __shared__ float smem[];
    const K1Shared L = k1_layout(d.pts_per_item);
    float *recs = smem + L.recs;
    float *s_pair = smem + L.pair;
    float *s_ptin = smem + L.ptin;
    float *s_ptout = smem + L.ptout;
    float *s_thr = smem + L.misc;              // [MAXF] frameEnergyTH
    float *s_cal = s_thr + MAXF;               // fxl fyl cxl cyl fxli fyli wM3G hM3G
    float *s_xAd = s_cal + 8;                  // [MAXF][8] rows (host*nF + t)
    float *s_cstep = s_xAd + MAXF * 8;         // [4]
    float *s_dHT = s_cstep + 4;                // [MAXF][8] adHTdeltaF[host + nF*t]
    float *s_cD = s_dHT + MAXF * 8;            // [4] cDeltaF
    float *s_set = s_cD + 4;                   // huberTH, outlierTHSumComponent, affA, affB
    __shared__ double s_red[K1_THREADS / 32][4];

    const int item = blockIdx.x;
    const int4 it = d.items[item];
    const int host = it.x, p0 = it.y, p1 = it.z;
    const int npts = p1 - p0;
    const int nF = ws->nF;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mode = (flags >> K1F_MODE_SHIFT) & 3;

#ifdef LDSO_B200_PROFILE
    int dbgi = 0;
    long long k1_t0 = clock64();
    unsigned long long k1_gt0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(k1_gt0));
#define K1_STAMP() do { if (tid == 0 && blockIdx.x == 0) d.dbg[dbgi] = clock64(); dbgi++; } while (0)
#else
#define K1_STAMP() do { } while (0)
#endif
    // ---------------- phase 0a (before pdl_wait: iteration-constant data only): clear records, request the point scalars
    // and the first round of residual records
    pdl_launch_dependents();
    for (int i = tid; i < npts * MAXF * 9; i += K1_THREADS) {     // only the active flag and JpJdF must start at zero
        const int q = i / 9, k = i - q * 9;
        recs[q * REC + REC_ACTIVE + k] = 0.f;
    }
    const int rbeg = d.pt_res_begin[p0], rend = d.pt_res_begin[p1];
    const int nres = rend - rbeg;
    float *s_rdot = s_ptout;        // [nres <= pts_per_item*MAXF] reuse: s_ptout is not live before phase P
    // software prefetch: everything phases R2 and A(round 0) read from global is requested here, before the first
    // barrier, so that its L2/HBM round trip overlaps the staging above instead of serialising behind each barrier
    float pf_idepth = 0.f, pf_idz = 0.f, pf_bd = 0.f, pf_step = 0.f, pf_hdi = 0.f, pf_u = 0.f, pf_v = 0.f, pf_prior = 0.f;
    float4 pf_hcd = make_float4(0.f, 0.f, 0.f, 0.f);
    int pf_r0 = 0, pf_r1 = 0;
    if (tid < npts) {        // window inputs that no kernel writes
        const int p = p0 + tid;
        pf_u = d.pt_u[p]; pf_v = d.pt_v[p]; pf_prior = d.pt_priorF[p];
        if (flags & K1F_APPLY_STEP) { pf_r0 = d.pt_res_begin[p] - rbeg; pf_r1 = d.pt_res_begin[p + 1] - rbeg; }
    }
    const int grp = tid >> 3, idx = tid & 7;
    int nx_p, nx_t; uint8_t nx_state, nx_lin; float nx_energy; int nx_slot;
    // topology (constant) and state (written by the previous iteration's K1) of the residual a lane group handles next
#define K1_PREFETCH_RES_TOPO(base_)                                                                     \
    do {                                                                                                \
        const int ri_ = (base_) + grp;                                                                  \
        const int r_ = rbeg + ((ri_ < nres) ? ri_ : 0);                                                 \
        nx_p = d.res_point[r_]; nx_t = d.res_target[r_];                                                \
        nx_slot = (d.res_newest_slot != nullptr) ? d.res_newest_slot[r_] : -1;                          \
    } while (0)
#define K1_PREFETCH_RES_STATE(base_)                                                                    \
    do {                                                                                                \
        const int ri_ = (base_) + grp;                                                                  \
        const int r_ = rbeg + ((ri_ < nres) ? ri_ : 0);                                                 \
        nx_state = d.res_state[r_]; nx_energy = d.res_energy[r_]; nx_lin = d.res_lin[r_];               \
    } while (0)
#define K1_PREFETCH_RES(base_) do { K1_PREFETCH_RES_TOPO(base_); K1_PREFETCH_RES_STATE(base_); } while (0)
    nx_p = p0; nx_t = 0; nx_state = 0; nx_lin = 0; nx_energy = 0.f; nx_slot = -1;
    if (nres > 0) K1_PREFETCH_RES_TOPO(0);
    // ---------------- phase 0b: everything the solver kernel produced (pair records, xAd, calibration, thresholds)
    pdl_wait();
#ifdef LDSO_B200_PROFILE
    if (tid == 0) {          // per-CTA wall-clock span (3 stores per CTA; after pdl_wait: no global writes before it)
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        d.dbg[32 + 3 * blockIdx.x] = (long long) k1_gt0;
        d.dbg[32 + 3 * blockIdx.x + 2] = smid;
        if (blockIdx.x == 0) d.dbg[dbgi] = k1_t0;
    }
    dbgi++;
#endif
    if (tid < npts) {        // point state written by the previous iteration's K1
        const int p = p0 + tid;
        pf_idepth = d.pt_idepth[p]; pf_idz = d.pt_idepth_zero[p];
        if (flags & K1F_APPLY_STEP) { pf_bd = d.pt_bdSumF[p]; pf_step = d.pt_step[p]; pf_hdi = d.pt_HdiF[p]; pf_hcd = *(const float4 *) (d.pt_Hcd + 4 * p); }
    }
    if (nres > 0) K1_PREFETCH_RES_STATE(0);
    for (int i = tid; i < MAXF * 32; i += K1_THREADS) {
        int t = i >> 5, k = i & 31;
        s_pair[i] = (t < nF) ? ((const float *) &ws->pair[host + nF * t])[k] : 0.f;
    }
    if (tid < MAXF) s_thr[tid] = (tid < nF) ? ws->frameEnergyTH[tid] : 0.f;
    if (tid == 0) {
        s_cal[0] = ws->calib.fxl; s_cal[1] = ws->calib.fyl; s_cal[2] = ws->calib.cxl; s_cal[3] = ws->calib.cyl;
        s_cal[4] = ws->calib.fxli; s_cal[5] = ws->calib.fyli; s_cal[6] = ws->wM3G; s_cal[7] = ws->hM3G;
        s_set[0] = ws->S.huberTH; s_set[1] = ws->S.outlierTHSumComponent;
        s_set[2] = ws->S.affineOptModeA; s_set[3] = ws->S.affineOptModeB;
    }
    if (tid < MAXF * 8) {
        int t = tid >> 3, k = tid & 7;
        s_xAd[tid] = (t < nF) ? ws->xAd[host * nF + t][k] : 0.f;
        s_dHT[tid] = (t < nF) ? ws->adHTdeltaF[host + nF * t][k] : 0.f;
    }
    if (tid < 4) { s_cstep[tid] = ws->cstep[tid]; s_cD[tid] = ws->calib.cDeltaF[tid]; }
    if (flags & K1F_APPLY_STEP) {   // R1 (no dependence on the staged constants: overlaps their load latency)
        for (int ri = tid; ri < nres; ri += K1_THREADS) {
            const int r = rbeg + ri;
            float s = 0.f;
            float act = 0.f;
            if (d.res_active[r]) {
                const float4 j0 = *(const float4 *) (d.res_JpJdF + 8 * r), j1 = *(const float4 *) (d.res_JpJdF + 8 * r + 4);
                const float4 *xa = (const float4 *) ws->xAd[host * nF + d.res_target[r]];
                const float4 x0 = xa[0], x1 = xa[1];
                s += x0.x * j0.x; s += x0.y * j0.y; s += x0.z * j0.z; s += x0.w * j0.w;
                s += x1.x * j1.x; s += x1.y * j1.y; s += x1.z * j1.z; s += x1.w * j1.w;
                act = 1.f;
            }
            recs_dot(s_rdot, ri, 0) = s;
            recs_dot(s_rdot, ri, 1) = act;
        }
    }
    __syncthreads();

    K1_STAMP();   // 1: phase 0 done
    // ---------------- phase R: resubstitute + step, stage point inputs.
    // R1: one thread per residual computes xAd[h,t] . JpJdF_r (EnergyFunctional.cc:539-542) into shared memory;
    // R2: one thread per point folds them in residual-list order and applies the idepth step.
    double my_sumNID = 0.0, my_numID = 0.0;
    if (tid < npts) {
        const int p = p0 + tid;
        float idepth = pf_idepth;
        float idepth_zero = pf_idz;
        if (flags & K1F_APPLY_STEP) {
            const int r0 = pf_r0, r1 = pf_r1;
            int ngood = 0;
            float b = pf_bd;
            {
                const float4 hc = pf_hcd;
                float s = 0.f;
                s += s_cstep[0] * hc.x; s += s_cstep[1] * hc.y; s += s_cstep[2] * hc.z; s += s_cstep[3] * hc.w;
                b -= s;
            }
            for (int ri = r0; ri < r1; ri++) {
                if (recs_dot(s_rdot, ri, 1) == 0.f) continue;
                ngood++;
                b -= recs_dot(s_rdot, ri, 0);
            }
            float step = pf_step;
            if (ngood == 0) step = 0.f;
            else if (isfinite(b)) step = -b * pf_hdi;          // non-finite b: this point keeps its step (the reference leaves its whole thread range, :544)
            d.pt_step[p] = step;
            const float backup = idepth;                       // FullSystem::backupState
            d.pt_idepth_backup[p] = backup;
            idepth = backup + step;                            // setIdepth(idepth_backup + stepfacD*step)
            idepth_zero = idepth;                              // setIdepthZero(...)
            d.pt_idepth[p] = idepth;
            d.pt_idepth_zero[p] = idepth_zero;
        }
        my_sumNID = fabsf(idepth);
        my_numID = 1.0;
        float *pi = s_ptin + tid * 8;
        pi[0] = pf_u;
        pi[1] = pf_v;
        pi[2] = idepth;
        pi[3] = idepth_zero;
        pi[4] = pf_prior;
        pi[5] = idepth - idepth_zero;                          // deltaF (EnergyFunctional.cc:424)
        pi[6] = (pt_sel == nullptr || pt_sel[p]) ? 1.f : 0.f;
    }
    __syncthreads();

    K1_STAMP();   // 2: phase R done
    // ---------------- phase A: one 8-lane group per residual
    double my_energy = 0.0, my_nres = 0.0;
    const float fxl = s_cal[0], fyl = s_cal[1], cxl = s_cal[2], cyl = s_cal[3], fxli = s_cal[4], fyli = s_cal[5];
    const float wM3G = s_cal[6], hM3G = s_cal[7];
    const float huberTH = s_set[0], outTH = s_set[1];
    const bool zeroA = s_set[2] < 0.f, zeroB = s_set[3] < 0.f;
    const int imgw = ws->w;

    for (int base = 0; base < nres; base += K1_GROUPS) {
        const int ri = base + grp;
        const bool valid = ri < nres;
        const int r = rbeg + (valid ? ri : 0);
        const int p = nx_p;
        const int t = nx_t;
        const int pl = p - p0;
        const float *pin = s_ptin + pl * 8;
        const float *pr = s_pair + t * 32;
        uint8_t old_state = (flags & K1F_RESET_OOB) ? (uint8_t) LDSO_B200_RES_IN : nx_state;
        float old_energy = (flags & K1F_RESET_OOB) ? 0.f : nx_energy;
        const bool lin = nx_lin != 0;
        const int my_slot = nx_slot;
        if (base + K1_GROUPS < nres) K1_PREFETCH_RES(base + K1_GROUPS);    // next round's indices, in flight during this round
        // FullSystem::optimize only puts non-linearized residuals into activeResiduals (:744-750): a linearized
        // residual is neither reset nor re-linearized, and mode-0 accumulation skips it.
        const bool touch = valid && !((flags & K1F_LINEARIZE) && lin) && (pin[6] != 0.f);   // pt_sel restricts the pass

        float x10[10], y10[10], Jpdd0, Jpdd1;
        float v_res, v_ji0, v_ji1, v_jab0, v_jab1;       // this lane's pixel row: resF, JIdx[0..1], JabF[0..1]
        float sA, sB, sC, sJabJI00, sJabJI01, sJabJI10, sJabJI11, sJab2_00, sJab2_01, sJab2_11;
        float energyLeft = 0.f, wJI2 = 0.f;
        bool oob = false;
        float Ku = 0.f, Kv = 0.f, cKu = 0.f, cKv = 0.f, cId = 0.f;

        if (flags & K1F_LINEARIZE) {
            const float u = pin[0], v = pin[1], idepth = pin[2], idepth_zero = pin[3];
            // ---- centre projection at the evaluation point (ResidualProjections.h:57-84)
            const float KliP0 = (u + 0 - cxl) * fxli;
            const float KliP1 = (v + 0 - cyl) * fyli;
            float ptp0 = pr[0] * KliP0; ptp0 += pr[1] * KliP1; ptp0 += pr[2] * 1.f; ptp0 = ptp0 + pr[9] * idepth_zero;
            float ptp1 = pr[3] * KliP0; ptp1 += pr[4] * KliP1; ptp1 += pr[5] * 1.f; ptp1 = ptp1 + pr[10] * idepth_zero;
            float ptp2 = pr[6] * KliP0; ptp2 += pr[7] * KliP1; ptp2 += pr[8] * 1.f; ptp2 = ptp2 + pr[11] * idepth_zero;
            const float drescale = 1.0f / ptp2;
            const float new_idepth = idepth_zero * drescale;
            bool ok_c = drescale > 0;
            const float uu = ptp0 * drescale, vv = ptp1 * drescale;
            cKu = uu * fxl + cxl;
            cKv = vv * fyl + cyl;
            cId = new_idepth;
            ok_c = ok_c && cKu > 1.1f && cKv > 1.1f && cKu < wM3G && cKv < hM3G;
            // ---- geometric Jacobians (Residuals.cc:67-104)
            const float t0x = pr[9], t0y = pr[10], t0z = pr[11];
            Jpdd0 = drescale * (t0x - t0z * uu) * SCALE_IDEPTH * fxl;
            Jpdd1 = drescale * (t0y - t0z * vv) * SCALE_IDEPTH * fyl;
            float dCx2 = drescale * (pr[6] * uu - pr[0]);
            float dCx3 = fxl * drescale * (pr[7] * uu - pr[1]) * fyli;
            float dCx0 = KliP0 * dCx2;
            float dCx1 = KliP1 * dCx3;
            float dCy2 = fyl * drescale * (pr[6] * vv - pr[3]) * fxli;
            float dCy3 = drescale * (pr[7] * vv - pr[4]);
            float dCy0 = KliP0 * dCy2;
            float dCy1 = KliP1 * dCy3;
            x10[0] = (dCx0 + uu) * SCALE_F; x10[1] = dCx1 * SCALE_F; x10[2] = (dCx2 + 1) * SCALE_C; x10[3] = dCx3 * SCALE_C;
            y10[0] = dCy0 * SCALE_F; y10[1] = (dCy1 + vv) * SCALE_F; y10[2] = dCy2 * SCALE_C; y10[3] = (dCy3 + 1) * SCALE_C;
            x10[4] = new_idepth * fxl; x10[5] = 0; x10[6] = -new_idepth * uu * fxl;
            x10[7] = -uu * vv * fxl; x10[8] = (1 + uu * uu) * fxl; x10[9] = -vv * fxl;
            y10[4] = 0; y10[5] = new_idepth * fyl; y10[6] = -new_idepth * vv * fyl;
            y10[7] = -(1 + vv * vv) * fyl; y10[8] = uu * vv * fyl; y10[9] = uu * fyl;

            // ---- this lane's pattern pixel (ResidualProjections.h:24-33, Residuals.cc:126-188)
            const float pu = u + c_patx[idx], pv = v + c_paty[idx];
            float q0 = pr[12] * pu; q0 += pr[13] * pv; q0 += pr[14] * 1.f; q0 = q0 + pr[21] * idepth;
            float q1 = pr[15] * pu; q1 += pr[16] * pv; q1 += pr[17] * 1.f; q1 = q1 + pr[22] * idepth;
            float q2 = pr[18] * pu; q2 += pr[19] * pv; q2 += pr[20] * 1.f; q2 = q2 + pr[23] * idepth;
            Ku = q0 / q2;
            Kv = q1 / q2;
            const bool ok_p = Ku > 1.1f && Kv > 1.1f && Ku < wM3G && Kv < hM3G;
            const float sx = ok_p ? Ku : 1.5f, sy = ok_p ? Kv : 1.5f;   // keep the gather in-bounds for dead lanes
            const int ix = (int) sx, iy = (int) sy;
            const float dx = sx - ix, dy = sy - iy, dxdy = dx * dy;
            const float4 *bp = ws->img0[t] + ix + iy * imgw;
            const float4 c00 = __ldg(bp), c10 = __ldg(bp + 1), c01 = __ldg(bp + imgw), c11 = __ldg(bp + imgw + 1);
            const float w11 = dxdy, w01 = dy - dxdy, w10 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
            float hit0 = w11 * c11.x + w01 * c01.x + w10 * c10.x + w00 * c00.x;
            float hit1 = w11 * c11.y + w01 * c01.y + w10 * c10.y + w00 * c00.y;
            float hit2 = w11 * c11.z + w01 * c01.z + w10 * c10.z + w00 * c00.z;
            const float color = __ldg(d.pt_color + 8 * p + idx), weight = __ldg(d.pt_weights + 8 * p + idx);
            const float residual = hit0 - (pr[24] * color + pr[25]);
            const float drdA = color - pr[26];
            const bool lane_bad = !ok_c || !ok_p || !isfinite(hit0);
            const unsigned bal = __ballot_sync(0xffffffffu, lane_bad);
            oob = ((bal >> (lane & 24)) & 0xffu) != 0;

            float w = sqrtf(outTH / (outTH + (hit1 * hit1 + hit2 * hit2)));
            w = 0.5f * (w + weight);
            float hw = fabsf(residual) < huberTH ? 1.f : huberTH / fabsf(residual);
            float e_px = w * w * hw * residual * residual * (2 - hw);
            if (hw < 1) hw = sqrtf(hw);
            hw = hw * w;
            hit1 *= hw;
            hit2 *= hw;
            v_res = residual * hw;
            v_ji0 = hit1;
            v_ji1 = hit2;
            v_jab0 = drdA * hw;
            v_jab1 = hw;
            if (lane_bad) { e_px = 0.f; v_res = v_ji0 = v_ji1 = v_jab0 = v_jab1 = 0.f; }   // result is discarded (OOB)
            sA = group_sum8(v_ji0 * v_ji0);
            sC = group_sum8(v_ji1 * v_ji1);
            sB = group_sum8(v_ji0 * v_ji1);
            sJabJI00 = group_sum8(v_jab0 * v_ji0);
            sJabJI01 = group_sum8(v_jab0 * v_ji1);
            sJabJI10 = group_sum8(v_jab1 * v_ji0);
            sJabJI11 = group_sum8(v_jab1 * v_ji1);
            sJab2_00 = group_sum8(v_jab0 * v_jab0);
            sJab2_01 = group_sum8(v_jab0 * v_jab1);
            sJab2_11 = group_sum8(v_jab1 * v_jab1);
            wJI2 = group_sum8(v_jab1 * v_jab1 * (v_ji0 * v_ji0 + v_ji1 * v_ji1));
            energyLeft = group_sum8(e_px);
            if (zeroA) v_jab0 = 0.f;     // Residuals.cc:185-186 (after the inner products)
            if (zeroB) v_jab1 = 0.f;
        } else {
            // ---- rebuild from the stored RawResidualJacobian (piecewise API / marginalisation modes)
            const float *J = d.res_J + (size_t) 74 * r;
            v_res = J[0 + idx];
            for (int i = 0; i < 6; i++) { x10[4 + i] = J[8 + i]; y10[4 + i] = J[14 + i]; }
            for (int i = 0; i < 4; i++) { x10[i] = J[20 + i]; y10[i] = J[24 + i]; }
            Jpdd0 = J[28]; Jpdd1 = J[29];
            v_ji0 = J[30 + idx]; v_ji1 = J[38 + idx]; v_jab0 = J[46 + idx]; v_jab1 = J[54 + idx];
            sA = J[62]; sB = J[63]; sC = J[65];
            sJabJI00 = J[66]; sJabJI01 = J[67]; sJabJI10 = J[68]; sJabJI11 = J[69];
            sJab2_00 = J[70]; sJab2_01 = J[71]; sJab2_11 = J[73];
        }

        // ---- state decision (Residuals.cc:203-213) and applyRes (Residuals.h:70-87)
        uint8_t new_state;
        float new_energy, energy_wo, ret_energy;
        bool active;
        if (flags & K1F_LINEARIZE) {
            if (old_state == LDSO_B200_RES_OOB) oob = true;
            if (oob) {
                new_state = LDSO_B200_RES_OOB;
                energy_wo = -1.f;
                ret_energy = old_energy;
                new_energy = old_energy;     // state_NewEnergy is left untouched by the early returns
            } else {
                energy_wo = energyLeft;
                const float th = fmaxf(s_thr[host], s_thr[t]);
                if (energyLeft > th || wJI2 < 2) { energyLeft = th; new_state = LDSO_B200_RES_OUTLIER; }
                else new_state = LDSO_B200_RES_IN;
                new_energy = energyLeft;
                ret_energy = energyLeft;
            }
            active = (old_state != LDSO_B200_RES_OOB) && (new_state == LDSO_B200_RES_IN);
            if (!(flags & K1F_APPLY_RES)) active = false;
        } else {
            new_state = old_state; new_energy = old_energy; energy_wo = -1.f; ret_energy = 0.f;
            const bool act = d.res_active[r] != 0;
            // mode 0: addPoint<0> (active, not linearized); 1: addPoint<1> (active, linearized); 2: addPoint<2> (all active,
            // res_toZeroF); 3: modes 0 and 1 in ONE pass -- what solveSystemF sums anyway (HA + HL, Hdd_accAF + Hdd_accLF, ...)
            active = (mode == 0) ? (act && !lin) : (mode == 1) ? (act && lin) : act;
            if (pin[6] == 0.f) active = false;
        }
        if (!touch) active = false;

        // ---- resApprox per accumulate mode (AccumulatedTopHessian.cc:40-64) and JI_r, Jab_r, rr (:66-76)
        float resApprox = v_res;
        const int rmode = (mode == 3) ? (lin ? 1 : 0) : mode;
        if (!(flags & K1F_LINEARIZE) && rmode != 0) {
            const float rtz = d.res_toZero[8 * r + idx];
            if (rmode == 2) resApprox = rtz;
            else {
                const float *dp = s_dHT + t * 8;
                float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
                for (int i = 0; i < 6; i++) { a0 += x10[4 + i] * dp[i]; a1 += y10[4 + i] * dp[i]; }
                for (int i = 0; i < 4; i++) { b0 += x10[i] * s_cD[i]; b1 += y10[i] * s_cD[i]; }
                const float Jpdx = a0 + b0 + Jpdd0 * pin[5];
                const float Jpdy = a1 + b1 + Jpdd1 * pin[5];
                float q = rtz;
                q = q + v_ji0 * Jpdx;
                q = q + v_ji1 * Jpdy;
                q = q + v_jab0 * dp[6];
                q = q + v_jab1 * dp[7];
                resApprox = q;
            }
        }
        const float JI_r0 = group_sum8(resApprox * v_ji0);
        const float JI_r1 = group_sum8(resApprox * v_ji1);
        const float Jab_r0 = group_sum8(resApprox * v_jab0);
        const float Jab_r1 = group_sum8(resApprox * v_jab1);
        const float rr = group_sum8(resApprox * resApprox);

        // takeData (Residuals.h:123-128)
        const float JIJd0 = sA * Jpdd0 + sB * Jpdd1;
        const float JIJd1 = sB * Jpdd0 + sC * Jpdd1;
        float JpJdF[8];
#pragma unroll
        for (int i = 0; i < 6; i++) JpJdF[i] = x10[4 + i] * JIJd0 + y10[4 + i] * JIJd1;
        JpJdF[6] = sJabJI00 * Jpdd0 + sJabJI01 * Jpdd1;
        JpJdF[7] = sJabJI10 * Jpdd0 + sJabJI11 * Jpdd1;

        if (touch && idx == 0) {
            if (flags & K1F_LINEARIZE) {
                d.res_new_state[r] = new_state;
                d.res_new_energy[r] = new_energy;
                d.res_new_energy_wo[r] = energy_wo;
                const bool skip_apply = (old_state == LDSO_B200_RES_OOB);
                if (flags & K1F_APPLY_RES) {
                    if (!skip_apply) {
                        d.res_state[r] = new_state;
                        d.res_energy[r] = new_energy;
                        d.res_active[r] = active ? 1 : 0;
                        if (active) {
#pragma unroll
                            for (int i = 0; i < 8; i++) d.res_JpJdF[8 * r + i] = JpJdF[i];
                        }
                    } else if (flags & K1F_RESET_OOB) {
                        d.res_state[r] = old_state;
                        d.res_energy[r] = old_energy;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) d.res_JpJdF_new[8 * r + i] = JpJdF[i];
                    if (flags & K1F_RESET_OOB) { d.res_state[r] = old_state; d.res_energy[r] = old_energy; }
                }
                if (my_slot >= 0) d.red[RED_SELECT + my_slot] = (double) energy_wo;
                my_energy += (double) ret_energy;
            }
            if (active) my_nres += 1.0;
            // stage the record for the accumulation phases
            float *rec = recs + (pl * MAXF + t) * REC;
            if (active) {
#pragma unroll
                for (int i = 0; i < 10; i++) { rec[REC_X + i] = x10[i]; rec[REC_Y + i] = y10[i]; }
                rec[REC_A] = sA; rec[REC_B] = sB; rec[REC_C] = sC;
                rec[REC_JABJI + 0] = sJabJI00; rec[REC_JABJI + 1] = sJabJI01;
                rec[REC_JABJI + 2] = sJabJI10; rec[REC_JABJI + 3] = sJabJI11;
                rec[REC_JIR] = JI_r0; rec[REC_JIR + 1] = JI_r1;
                rec[REC_JAB2] = sJab2_00; rec[REC_JAB2 + 1] = sJab2_01; rec[REC_JAB2 + 2] = sJab2_11;
                rec[REC_JABR] = Jab_r0; rec[REC_JABR + 1] = Jab_r1;
                rec[REC_RR] = rr;
                rec[REC_ACTIVE] = 1.f;
                if (flags & K1F_LINEARIZE) {
#pragma unroll
                    for (int i = 0; i < 8; i++) rec[REC_JPJD + i] = JpJdF[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) rec[REC_JPJD + i] = d.res_JpJdF[8 * r + i];
                }
                rec[REC_HDD] = JIJd0 * Jpdd0 + JIJd1 * Jpdd1;                 // AccumulatedTopHessian.cc:96
                rec[REC_BD] = JI_r0 * Jpdd0 + JI_r1 * Jpdd1;                  // :95
#pragma unroll
                for (int i = 0; i < 4; i++) rec[REC_HCD + i] = x10[i] * JIJd0 + y10[i] * JIJd1;   // :97
                rec[REC_JPDD] = Jpdd0; rec[REC_JPDD + 1] = Jpdd1;
            }
        }
        if ((flags & K1F_STORE_J) && (flags & K1F_LINEARIZE) && touch) {
            float *J = d.res_J + (size_t) 74 * r;
            J[0 + idx] = v_res;
            J[30 + idx] = v_ji0; J[38 + idx] = v_ji1; J[46 + idx] = v_jab0; J[54 + idx] = v_jab1;
            d.res_proj[16 * r + 2 * idx] = Ku;
            d.res_proj[16 * r + 2 * idx + 1] = Kv;
            if (idx == 0) {
                for (int i = 0; i < 6; i++) { J[8 + i] = x10[4 + i]; J[14 + i] = y10[4 + i]; }
                for (int i = 0; i < 4; i++) { J[20 + i] = x10[i]; J[24 + i] = y10[i]; }
                J[28] = Jpdd0; J[29] = Jpdd1;
                J[62] = sA; J[63] = sB; J[64] = sB; J[65] = sC;
                J[66] = sJabJI00; J[67] = sJabJI01; J[68] = sJabJI10; J[69] = sJabJI11;
                J[70] = sJab2_00; J[71] = sJab2_01; J[72] = sJab2_01; J[73] = sJab2_11;
                d.res_cpt[3 * r] = cKu; d.res_cpt[3 * r + 1] = cKv; d.res_cpt[3 * r + 2] = cId;
            }
        }
    }

    K1_STAMP();   // 3: phase A done
    // ---------------- item statistics: energy, active residual count, sum |idepth|, point count
    {
        double v0 = my_energy, v1 = my_nres, v2 = my_sumNID, v3 = my_numID;
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
            v2 += __shfl_xor_sync(0xffffffffu, v2, o);
            v3 += __shfl_xor_sync(0xffffffffu, v3, o);
        }
        if (lane == 0) { s_red[warp][0] = v0; s_red[warp][1] = v1; s_red[warp][2] = v2; s_red[warp][3] = v3; }
    }
    __syncthreads();
    if (tid < 4) {
        double s = 0.0;
        for (int w2 = 0; w2 < K1_THREADS / 32; w2++) s += s_red[w2][tid];
        d.item_stats[4 * item + tid] = s;
    }
    if (!(flags & K1F_ACCUMULATE)) return;

    float *part = d.partials + (size_t) item * PART_STRIDE;
    K1_STAMP();   // 4: stats done

    // ---------------- phase B: top Hessian blocks, warp <-> target (AccumulatedTopHessian.cc:78-92)
    // 4 lane-slots of uniform entry type (no divergence): [tri 0..31] [tri 32..54] [TopRight 0..29] [BotRight 0..5]
    {
        const int t = warp;     // K1_THREADS/32 == MAXF
        int r0, c0, r1, c1;
        tri10_rc(lane, r0, c0);
        tri10_rc(min(32 + lane, 54), r1, c1);
        const bool on1 = lane < 23, on2 = lane < 30, on3 = lane < 6;
        const int tri = min(lane, 29) / 3, trk = min(lane, 29) % 3;
        const int troff0 = (trk == 0) ? REC_JABJI + 0 : (trk == 1) ? REC_JABJI + 2 : REC_JIR + 0;
        const int troff1 = (trk == 0) ? REC_JABJI + 1 : (trk == 1) ? REC_JABJI + 3 : REC_JIR + 1;
        const int b3 = min(lane, 5);   // Jab2(0,0) Jab2(0,1) Jab_r[0] Jab2(1,1) Jab_r[1] rr
        const int broff = (b3 == 0) ? REC_JAB2 : (b3 == 1) ? REC_JAB2 + 1 : (b3 == 2) ? REC_JABR : (b3 == 3) ? REC_JAB2 + 2
                                                                          : (b3 == 4) ? REC_JABR + 1 : REC_RR;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        if (t < nF && t != host) {
            for (int pl = 0; pl < npts; pl++) {
                const float *rec = recs + (pl * MAXF + t) * REC;
                if (rec[REC_ACTIVE] == 0.f) continue;
                const float a = rec[REC_A], b = rec[REC_B], c = rec[REC_C];
                {
                    const float xr = rec[REC_X + r0], xc = rec[REC_X + c0], yr = rec[REC_Y + r0], yc = rec[REC_Y + c0];
                    acc0 += a * xc * xr + c * yc * yr + b * (xc * yr + yc * xr);
                }
                {
                    const float xr = rec[REC_X + r1], xc = rec[REC_X + c1], yr = rec[REC_Y + r1], yc = rec[REC_Y + c1];
                    acc1 += a * xc * xr + c * yc * yr + b * (xc * yr + yc * xr);
                }
                acc2 += rec[REC_X + tri] * rec[troff0] + rec[REC_Y + tri] * rec[troff1];
                acc3 += rec[broff];
            }
        }
        float *pt = part + PART_TOP + t * 96;
        pt[lane] = acc0;
        if (on1) pt[32 + lane] = acc1;
        if (on2) pt[55 + lane] = acc2;
        if (on3) pt[85 + lane] = acc3;
        if (lane < 5) pt[91 + lane] = 0.f;
    }

    K1_STAMP();   // 5: phase B done
    // ---------------- phase P: per-point sums (AccumulatedTopHessian.cc:94-116, AccumulatedSCHessian.cc:11-29)
    if (tid < npts) {
        const int pl = tid, p = p0 + pl;
        float Hdd = 0.f, bd = 0.f, Hcd[4] = {0.f, 0.f, 0.f, 0.f};
        int ngood = 0;
        for (int t = 0; t < nF; t++) {
            const float *rec = recs + (pl * MAXF + t) * REC;
            if (rec[REC_ACTIVE] == 0.f) continue;
            ngood++;
            bd += rec[REC_BD];
            Hdd += rec[REC_HDD];
            for (int i = 0; i < 4; i++) Hcd[i] += rec[REC_HCD + i];
        }
        const float *pin = s_ptin + pl * 8;
        float HdiF = 0.f, bdSumF = 0.f;
        if (ngood > 0) {
            float H = Hdd + pin[4];
            if (H < 1e-10f) H = 1e-10f;
            HdiF = 1.0f / H;
            bdSumF = bd;
            if (!(flags & K1F_NO_SHIFT_PRIOR)) bdSumF += pin[4] * pin[5];
        } else {
            Hcd[0] = Hcd[1] = Hcd[2] = Hcd[3] = 0.f;
        }
        float *po = s_ptout + pl * 8;
        po[0] = HdiF; po[1] = bdSumF; po[2] = Hcd[0]; po[3] = Hcd[1]; po[4] = Hcd[2]; po[5] = Hcd[3];
        po[6] = HdiF * bdSumF;
        po[7] = (float) ngood;
        if (pin[6] != 0.f) {
            d.pt_HdiF[p] = HdiF; d.pt_bdSumF[p] = bdSumF; d.pt_Hdd[p] = Hdd; d.pt_bd[p] = bd;
            for (int i = 0; i < 4; i++) d.pt_Hcd[4 * p + i] = Hcd[i];
        }
    }
    __syncthreads();

    K1_STAMP();   // 6: phase P done
    // ---------------- phase C: Schur complement accumulators (AccumulatedSCHessian.cc:30-49)
    {
        const int ty = tid >> 4, tx = tid & 15;      // 16x16 threads x (4x4) tile of the 64x64 matrix D_host
        const int rt = ty >> 1, ro = (ty & 1) * 4;   // row block (target t1) and offset inside JpJdF
        const int ct = tx >> 1, co = (tx & 1) * 4;
        float D[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) D[i][j] = 0.f;
        // accE: thread -> (t = tid>>5, a = (tid>>2)&7, c = tid&3)
        const int et = tid >> 5, ea = (tid >> 2) & 7, ec = tid & 3;
        float accE = 0.f, accX = 0.f;   // accX: EB (tid<64), Hcc (64..79), bc (80..83)
        for (int pl = 0; pl < npts; pl++) {
            const float *po = s_ptout + pl * 8;
            const float HdiF = po[0];
            if (po[7] == 0.f) continue;
            const float4 vr = *(const float4 *) (recs + (pl * MAXF + rt) * REC + REC_JPJD + ro);
            const float4 vc = *(const float4 *) (recs + (pl * MAXF + ct) * REC + REC_JPJD + co);
            const float wr[4] = {HdiF * vr.x, HdiF * vr.y, HdiF * vr.z, HdiF * vr.w};
            const float cc[4] = {vc.x, vc.y, vc.z, vc.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) D[i][j] += wr[i] * cc[j];
            const float jp = recs[(pl * MAXF + et) * REC + REC_JPJD + ea];
            accE += (HdiF * jp) * po[2 + ec];
            if (tid < 64) accX += (po[6]) * recs[(pl * MAXF + (tid >> 3)) * REC + REC_JPJD + (tid & 7)];
            else if (tid < 80) accX += (HdiF * po[2 + ((tid - 64) >> 2)]) * po[2 + ((tid - 64) & 3)];
            else if (tid < 84) accX += (po[1] * HdiF) * po[2 + (tid - 80)];
        }
        float *pd = part + PART_D + (rt * MAXF + ct) * 64;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) pd[(ro + i) * 8 + (co + j)] = D[i][j];
        part[PART_E + et * 32 + ea * 4 + ec] = accE;
        if (tid < 64) part[PART_EB + tid] = accX;
        else if (tid < 80) part[PART_HCC + (tid - 64)] = accX;
        else if (tid < 84) part[PART_BC + (tid - 80)] = accX;
    }
    K1_STAMP();   // 7: phase C done
#ifdef LDSO_B200_PROFILE
    if (tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        d.dbg[32 + 3 * blockIdx.x + 1] = (long long) gt;
    }
#endif
}
