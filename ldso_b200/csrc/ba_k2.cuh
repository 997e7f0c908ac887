// K2a (fold per-item partials in double), K2b (stitch through the adjoints + newest-frame energy threshold),
// K3 (assemble, scaled LDLT solve, orthogonalize, frame/calib step, frame-pair precalc) and the small
// piecewise kernels (applyRes, resubstitute/step on points).
#pragma once
#include "common.cuh"
#include "se3_math.cuh"

// ---------------------------------------------------------------------------------------------------------
// K2a: red[h][e] = sum over the work items of host h of partial[item][e], in double
// (the reference casts its per-thread float accumulators to double before summing them,
//  AccumulatedTopHessian.cc:215-219, AccumulatedSCHessian.cc:78-83,101-105).
// The last block folds the per-item scalar statistics.
__global__ void __launch_bounds__(256) k2a_reduce(DevWindow d, WinState *ws, int full, int multi) {
    const int nF = ws->nF;
    if (blockIdx.x == gridDim.x - 1) {
        // stats: warp w (<4) reduces stat w over all items
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (w < 4) {
            double s = 0.0;
            for (int i = lane; i < d.nItems; i += 32) s += d.item_stats[4 * i + w];
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) {
                d.red[RED_STATS + w] = s;
                if (!multi) {     // single GPU: publish now; multi GPU: K2b publishes after the all-reduce
                    if (w == 0) ws->energy = s;
                    if (w == 1) ws->resInA = (int) (s + 0.5);
                    if (w == 2) ws->sumNID = (float) s;
                    if (w == 3) ws->numID = (float) s;
                }
            }
        }
        return;
    }
    if (!full) return;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= MAXF * PART_USED) return;
    const int h = g / PART_USED, e = g - h * PART_USED;
    double s = 0.0;
    if (h < nF) {
        const int i0 = d.host_item_begin[h], i1 = d.host_item_begin[h + 1];
        const float *p = d.partials + (size_t) i0 * PART_STRIDE + e;
        for (int i = i0; i < i1; i++, p += PART_STRIDE) s += (double) *p;
    }
    d.red[g] = s;
}

// ---------------------------------------------------------------------------------------------------------
// packed index of element (r,c) of the symmetric 13x13 AccumulatorApprox block [C(4)|xi(6)|ab(2)|r(1)]
__device__ __forceinline__ int packed13(int r, int c) {
    if (r > c) { int t = r; r = c; c = t; }
    if (c < 10) return r * 10 - (r * (r - 1)) / 2 + (c - r);
    if (r < 10) return 55 + 3 * r + (c - 10);
    return 85 + ((r == 10) ? (c - 10) : (r == 11) ? 3 + (c - 11) : 5);
}

struct SolveBufs {
    double *H_A, *b_A, *H_sc, *b_sc;      // stitched pieces, column-major n x n / n
    double *HM, *bM;                      // marginalisation prior
    double *Pns;                          // null-space projector (n x n, col-major)
    double *lastHS, *lastbS, *lastX;
};

// K2b: one CTA per 8x8 output block (a,b) of H_A and H_sc, nF CTAs for the calibration rows + b, one CTA for
// the calibration corner, and a last CTA for FullSystem::setNewFrameEnergyTH (FullSystem.cc:1762-1793).
// Formulas: AccumulatedTopHessian.cc:221-240 + .h:95-104 and AccumulatedSCHessian.cc:85-118 + .h:93-97,
// regrouped by output block so that no two CTAs write the same element (deterministic, no atomics).
#define K2B_THREADS 256
__device__ __forceinline__ double top_elem(const double *red, int h, int t, int r13, int c13) {
    return red[h * PART_USED + PART_TOP + t * 96 + packed13(r13, c13)];
}

__global__ void __launch_bounds__(K2B_THREADS) k2b_stitch(DevWindow d, WinState *ws, SolveBufs sb, int do_stitch, int do_select) {
    const int nF = ws->nF, n = ws->n;
    const int tid = threadIdx.x;
    const double *red = d.red;
    const int nBlocks = nF * nF;
    __shared__ double Ts[4][64];
    __shared__ double outp[2][4][64];
    __shared__ unsigned hist[256];
    __shared__ unsigned sel_prefix, sel_k, sel_count;

    if ((int) blockIdx.x < nBlocks) {
        if (!do_stitch) return;
        const int a = blockIdx.x % nF, b = blockIdx.x / nF;
        const int slot = tid >> 6, e = tid & 63, r = e >> 3, c = e & 7;
        double accA = 0.0, accS = 0.0;
        // ---- top (active) Hessian terms
        const int nqA = (a == b) ? 2 * nF : 2;
        for (int q0 = 0; q0 < nqA; q0 += 4) {
            const int q = q0 + slot;
            bool on = q < nqA;
            const double *Lm = nullptr, *Rm = nullptr;
            int mh = 0, mt = 0;
            bool tr = false;
            if (on) {
                if (a == b) {
                    if (q < nF) { mh = a; mt = q; Lm = Rm = ws->adHost[a + nF * q]; }
                    else { mh = q - nF; mt = a; Lm = Rm = ws->adTarget[mh + nF * a]; }
                    if (mh == mt) on = false;
                } else {
                    if (q == 0) { mh = a; mt = b; Lm = ws->adHost[a + nF * b]; Rm = ws->adTarget[a + nF * b]; }
                    else { mh = b; mt = a; Lm = ws->adHost[b + nF * a]; Rm = ws->adTarget[b + nF * a]; tr = true; }
                }
            }
            if (on) {
                double s = 0.0;
                for (int i = 0; i < 8; i++) s += Lm[r * 8 + i] * top_elem(red, mh, mt, 4 + i, 4 + c);
                Ts[slot][e] = s;
            }
            __syncthreads();
            if (on) {
                // out[r][c] += sum_j T[r][j] R[c][j]; transposed term contributes out[c][r]
                const int rr = tr ? c : r, cc = tr ? r : c;
                double s = 0.0;
                for (int j = 0; j < 8; j++) s += Ts[slot][rr * 8 + j] * Rm[cc * 8 + j];
                accA += s;
            }
            __syncthreads();
        }
        // ---- Schur complement terms
        // q in [0,nF): i=q      : AdT_ia D_i[a,b] AdT_ib^T
        // q in [nF,2nF): k=q-nF : AdT_ba D_b[a,k] AdH_bk^T
        // q in [2nF,3nF): j     : AdH_aj D_a[j,b] AdT_ab^T
        // a==b: q in [3nF, 3nF+nF*nF): (j,k): AdH_aj D_a[j,k] AdH_ak^T
        const int nqS = 3 * nF + ((a == b) ? nF * nF : 0);
        for (int q0 = 0; q0 < nqS; q0 += 4) {
            const int q = q0 + slot;
            const bool on = q < nqS;
            const double *Lm = nullptr, *Rm = nullptr, *Dm = nullptr;
            if (on) {
                if (q < nF) {
                    const int i = q;
                    Lm = ws->adTarget[i + nF * a]; Rm = ws->adTarget[i + nF * b];
                    Dm = red + i * PART_USED + PART_D + (a * MAXF + b) * 64;
                } else if (q < 2 * nF) {
                    const int k = q - nF;
                    Lm = ws->adTarget[b + nF * a]; Rm = ws->adHost[b + nF * k];
                    Dm = red + b * PART_USED + PART_D + (a * MAXF + k) * 64;
                } else if (q < 3 * nF) {
                    const int j = q - 2 * nF;
                    Lm = ws->adHost[a + nF * j]; Rm = ws->adTarget[a + nF * b];
                    Dm = red + a * PART_USED + PART_D + (j * MAXF + b) * 64;
                } else {
                    const int jk = q - 3 * nF, j = jk % nF, k = jk / nF;
                    Lm = ws->adHost[a + nF * j]; Rm = ws->adHost[a + nF * k];
                    Dm = red + a * PART_USED + PART_D + (j * MAXF + k) * 64;
                }
                double s = 0.0;
                for (int i = 0; i < 8; i++) s += Lm[r * 8 + i] * Dm[i * 8 + c];
                Ts[slot][e] = s;
            }
            __syncthreads();
            if (on) {
                double s = 0.0;
                for (int j = 0; j < 8; j++) s += Ts[slot][r * 8 + j] * Rm[c * 8 + j];
                accS += s;
            }
            __syncthreads();
        }
        outp[0][slot][e] = accA;
        outp[1][slot][e] = accS;
        __syncthreads();
        if (tid < 64) {
            const double vA = ((outp[0][0][e] + outp[0][1][e]) + outp[0][2][e]) + outp[0][3][e];
            const double vS = ((outp[1][0][e] + outp[1][1][e]) + outp[1][2][e]) + outp[1][3][e];
            const int row = CPARS + 8 * a + r, col = CPARS + 8 * b + c;
            sb.H_A[(size_t) col * n + row] = vA;
            sb.H_sc[(size_t) col * n + row] = vS;
        }
        return;
    }
    if ((int) blockIdx.x < nBlocks + nF) {
        if (!do_stitch) return;
        // ---- calibration rows of frame a and b segments
        const int a = blockIdx.x - nBlocks;
        if (tid < 32) {                 // H_A[a, c]  (8x4)
            const int r = tid >> 2, c = tid & 3;
            double s = 0.0;
            for (int t = 0; t < nF; t++) {
                if (t == a) continue;
                const double *AH = ws->adHost[a + nF * t], *AT = ws->adTarget[t + nF * a];
                for (int i = 0; i < 8; i++) {
                    s += AH[r * 8 + i] * top_elem(red, a, t, 4 + i, c);
                    s += AT[r * 8 + i] * top_elem(red, t, a, 4 + i, c);
                }
            }
            sb.H_A[(size_t) c * n + (CPARS + 8 * a + r)] = s;
            sb.H_A[(size_t) (CPARS + 8 * a + r) * n + c] = s;
        } else if (tid < 40) {          // b_A[a]
            const int r = tid - 32;
            double s = 0.0;
            for (int t = 0; t < nF; t++) {
                if (t == a) continue;
                const double *AH = ws->adHost[a + nF * t], *AT = ws->adTarget[t + nF * a];
                for (int i = 0; i < 8; i++) {
                    s += AH[r * 8 + i] * top_elem(red, a, t, 4 + i, 12);
                    s += AT[r * 8 + i] * top_elem(red, t, a, 4 + i, 12);
                }
            }
            sb.b_A[CPARS + 8 * a + r] = s;
        } else if (tid >= 64 && tid < 96) {   // H_sc[a, c]
            const int r = (tid - 64) >> 2, c = (tid - 64) & 3;
            double s = 0.0;
            for (int j = 0; j < nF; j++) {
                const double *AH = ws->adHost[a + nF * j], *AT = ws->adTarget[j + nF * a];
                const double *Eaj = red + a * PART_USED + PART_E + j * 32;
                const double *Eja = red + j * PART_USED + PART_E + a * 32;
                for (int i = 0; i < 8; i++) {
                    s += AH[r * 8 + i] * Eaj[i * 4 + c];
                    s += AT[r * 8 + i] * Eja[i * 4 + c];
                }
            }
            sb.H_sc[(size_t) c * n + (CPARS + 8 * a + r)] = s;
            sb.H_sc[(size_t) (CPARS + 8 * a + r) * n + c] = s;
        } else if (tid >= 96 && tid < 104) {  // b_sc[a]
            const int r = tid - 96;
            double s = 0.0;
            for (int j = 0; j < nF; j++) {
                const double *AH = ws->adHost[a + nF * j], *AT = ws->adTarget[j + nF * a];
                const double *Baj = red + a * PART_USED + PART_EB + j * 8;
                const double *Bja = red + j * PART_USED + PART_EB + a * 8;
                for (int i = 0; i < 8; i++) {
                    s += AH[r * 8 + i] * Baj[i];
                    s += AT[r * 8 + i] * Bja[i];
                }
            }
            sb.b_sc[CPARS + 8 * a + r] = s;
        }
        return;
    }
    if ((int) blockIdx.x == nBlocks + nF) {
        if (!do_stitch) return;
        // ---- calibration corner
        if (tid < 16) {
            const int r = tid >> 2, c = tid & 3;
            double sA = 0.0, sS = 0.0;
            for (int h = 0; h < nF; h++) {
                for (int t = 0; t < nF; t++) if (t != h) sA += top_elem(red, h, t, r, c);
                sS += red[h * PART_USED + PART_HCC + r * 4 + c];
            }
            sb.H_A[(size_t) c * n + r] = sA;
            sb.H_sc[(size_t) c * n + r] = sS;
        } else if (tid < 20) {
            const int r = tid - 16;
            double sA = 0.0, sS = 0.0;
            for (int h = 0; h < nF; h++) {
                for (int t = 0; t < nF; t++) if (t != h) sA += top_elem(red, h, t, r, 12);
                sS += red[h * PART_USED + PART_BC + r];
            }
            sb.b_A[r] = sA;
            sb.b_sc[r] = sS;
        }
        return;
    }
    // ---- last CTA: exact k-th order statistic of the newest frame's residual energies (radix select)
    if (tid == 0) {   // publish the (all-reduced) scalar statistics
        ws->energy = red[RED_STATS + 0];
        ws->resInA = (int) (red[RED_STATS + 1] + 0.5);
        ws->sumNID = (float) red[RED_STATS + 2];
        ws->numID = (float) red[RED_STATS + 3];
    }
    if (!do_select) return;
    {
        const int N = d.newest_total;
        const double *vals = red + RED_SELECT;
        // count valid (energy >= 0) values
        if (tid == 0) { sel_count = 0; sel_prefix = 0; }
        __syncthreads();
        unsigned cnt = 0;
        for (int i = tid; i < N; i += K2B_THREADS) if (vals[i] >= 0.0) cnt++;
        atomicAdd(&sel_count, cnt);
        __syncthreads();
        const unsigned m = sel_count;
        float th;
        if (m == 0) {
            th = 12 * 12 * LDSO_B200_PATTERN;
        } else {
            if (tid == 0) sel_k = (unsigned) (int) (ws->S.frameEnergyTHN * (float) m);
            __syncthreads();
            for (int pass = 3; pass >= 0; pass--) {
                for (int i = tid; i < 256; i += K2B_THREADS) hist[i] = 0;
                __syncthreads();
                const unsigned pref = sel_prefix;
                const unsigned himask = (pass == 3) ? 0u : (0xffffffffu << (8 * (pass + 1)));
                for (int i = tid; i < N; i += K2B_THREADS) {
                    const double v = vals[i];
                    if (v < 0.0) continue;
                    const unsigned key = __float_as_uint((float) v);
                    if ((key & himask) == pref) atomicAdd(&hist[(key >> (8 * pass)) & 0xffu], 1u);
                }
                __syncthreads();
                if (tid < 32) {
                    // warp 0: lane owns 8 consecutive bins; find the bin holding rank sel_k
                    unsigned loc = 0;
                    for (int q = 0; q < 8; q++) loc += hist[8 * tid + q];
                    unsigned incl = loc;
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (tid >= o) incl += v;
                    }
                    const unsigned excl = incl - loc;
                    const unsigned k0 = sel_k;
                    const bool mine = (k0 >= excl) && (k0 < incl);
                    __syncwarp();
                    if (mine) {
                        unsigned k = k0 - excl, bin = 8 * tid;
                        for (int q = 0; q < 8; q++, bin++) {
                            if (k < hist[bin]) break;
                            k -= hist[bin];
                        }
                        sel_k = k;
                        sel_prefix = pref | (bin << (8 * pass));
                    }
                }
                __syncthreads();
            }
            const float nthElement = sqrtf(__uint_as_float(sel_prefix));
            th = nthElement * ws->S.frameEnergyTHFacMedian;
            th = 26.0f * ws->S.frameEnergyTHConstWeight + th * (1 - ws->S.frameEnergyTHConstWeight);
            th = th * th;
            th *= ws->S.overallEnergyTHWeight * ws->S.overallEnergyTHWeight;
        }
        if (tid == 0) ws->fr[nF - 1].frameEnergyTH = th;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Frame-state update: FrameHessian::setState (FrameHessian.h:78-91), FrameFramePrecalc::Set for all nF^2
// pairs (FrameFramePrecalc.cc:6-35), EnergyFunctional::setDeltaF frame part (EnergyFunctional.cc:403-429).
// Called by all threads of one CTA (>= 64 threads); contains __syncthreads.
__device__ void frames_refresh(WinState *ws) {
    const int nF = ws->nF, tid = threadIdx.x;
    if (tid < nF) {
        FrameDev &f = ws->fr[tid];
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
        CalibDev &c = ws->calib;
        for (int i = 0; i < 4; i++) c.cDeltaF[i] = (float) (c.value[i] - c.value_zero[i]);
    }
    __syncthreads();
    if (tid < nF * nF) {
        const int h = tid % nF, t = tid / nF;
        const FrameDev &fh = ws->fr[h], &ft = ws->fr[t];
        PairRec &pc = ws->pair[h + nF * t];
        PairRecFull &pf = ws->pairFull[h + nF * t];
        double R0[9], t0[3], R[9], tt[3];
        se3_mul_inv(ft.evalR, ft.evalT, fh.evalR, fh.evalT, R0, t0);
        se3_mul_inv(ft.preR, ft.preT, fh.preR, fh.preT, R, tt);
        float Rf[9], tf[3];
        for (int i = 0; i < 9; i++) { pc.R0[i] = (float) R0[i]; Rf[i] = (float) R[i]; pf.RTll[i] = Rf[i]; }
        for (int i = 0; i < 3; i++) { pc.t0[i] = (float) t0[i]; tf[i] = (float) tt[i]; pf.tTll[i] = tf[i]; }
        pc.distanceLL = (float) sqrt(tt[0] * tt[0] + tt[1] * tt[1] + tt[2] * tt[2]);
        const CalibDev &c = ws->calib;
        float K[9] = {c.fxl, 0, c.cxl, 0, c.fyl, c.cyl, 0, 0, 1};
        float Ki[9], tmp[9];
        m33f_inverse(K, Ki);
        m33f_mul(K, Rf, tmp);
        m33f_mul(tmp, Ki, pc.KRKi);
        for (int i = 0; i < 3; i++) {
            float s = K[i * 3 + 0] * tf[0];
            s += K[i * 3 + 1] * tf[1];
            s += K[i * 3 + 2] * tf[2];
            pc.Kt[i] = s;
        }
        // AffLight::fromToVecExposure (AffLight.h:27-35) with aff_g2l() = state_scaled[6..7]
        float eF = fh.ab_exposure, eT = ft.ab_exposure;
        if (eF == 0 || eT == 0) eT = eF = 1;
        const float ah = (float) ((double) SCALE_A * fh.state[6]), bh = (float) ((double) SCALE_B * fh.state[7]);
        const float at = (float) ((double) SCALE_A * ft.state[6]), bt = (float) ((double) SCALE_B * ft.state[7]);
        const float aa = expf(at - ah) * eT / eF;
        pc.aff[0] = aa;
        pc.aff[1] = bt - aa * bh;
        pc.b0 = (float) (fh.state_zero[7] * (double) SCALE_B);
        // adHTdeltaF (EnergyFunctional.cc:406-414)
        float dh[8], dt[8];
        for (int i = 0; i < 8; i++) {
            dh[i] = (float) (fh.state[i] - fh.state_zero[i]);
            dt[i] = (float) (ft.state[i] - ft.state_zero[i]);
        }
        const float *AH = ws->adHostF[h + nF * t], *AT = ws->adTargetF[h + nF * t];
        for (int j = 0; j < 8; j++) {
            float s1 = 0.f, s2 = 0.f;
            for (int i = 0; i < 8; i++) s1 += dh[i] * AH[i * 8 + j];
            for (int i = 0; i < 8; i++) s2 += dt[i] * AT[i * 8 + j];
            ws->adHTdeltaF[h + nF * t][j] = s1 + s2;
        }
    }
    __syncthreads();
}

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

__global__ void __launch_bounds__(128) k_frames_refresh(WinState *ws) { frames_refresh(ws); }

// ---------------------------------------------------------------------------------------------------------
// K3: EnergyFunctional::solveSystemF (EnergyFunctional.cc:240-351, default solver mode) + the frame/calib part of
// resubstituteF_MT (:491-507), FullSystem::backupState and doStepFromBackup (FullSystem.cc:1587-1676). One CTA.
#define K3F_SOLVE 1
#define K3F_STEP 2
#define K3F_BACKUP 4
#define K3_THREADS 256
#define K3_LD (MAXN + 1)

__global__ void __launch_bounds__(K3_THREADS) k3_solve_step(WinState *ws, SolveBufs sb, int flags, int *iteration_dev) {
    extern __shared__ double sm3[];
    double *A0 = sm3;                      // [n][K3_LD] row-major assembled matrix
    double *A = A0 + MAXN * K3_LD;         // permuted copy, factorised in place
    double *vb = A + MAXN * K3_LD;         // rhs / solution
    double *vS = vb + MAXN;                // SVecI
    double *vd = vS + MAXN;                // delta / temp
    double *vx = vd + MAXN;                // x
    int *perm = (int *) (vx + MAXN);       // [n]
    const int nF = ws->nF, n = ws->n, tid = threadIdx.x;
    const int iteration = *iteration_dev;

    if (flags & K3F_BACKUP) {
        if (tid < nF) for (int i = 0; i < 10; i++) ws->fr[tid].state_backup[i] = ws->fr[tid].state[i];
        if (tid == 32) for (int i = 0; i < 4; i++) ws->calib.value_backup[i] = ws->calib.value[i];
        __syncthreads();
    }
    if (flags & K3F_SOLVE) {
        const double lambda = 1e-5;        // SOLVER_FIX_LAMBDA (EnergyFunctional.cc:243)
        // delta = getStitchedDeltaF (EnergyFunctional.h:178-184)
        if (tid < n) {
            double dv;
            if (tid < CPARS) dv = (double) ws->calib.cDeltaF[tid];
            else dv = ws->fr[(tid - CPARS) >> 3].delta[(tid - CPARS) & 7];
            vd[tid] = dv;
        }
        __syncthreads();
        // HFinal_top = HL + HM + HA ; bFinal_top = bL + bM_top + bA - b_sc  (:283-284)
        for (int e = tid; e < n * n; e += K3_THREADS) {
            const int r = e % n, c = e / n;
            double v = sb.H_A[e] + sb.HM[e];
            if (r == c) v += (r < CPARS) ? ws->cPrior[r] : ws->fr[(r - CPARS) >> 3].prior[(r - CPARS) & 7];
            A0[r * K3_LD + c] = v;
        }
        if (tid < n) {
            double bm = sb.bM[tid];
            for (int c = 0; c < n; c++) bm += sb.HM[(size_t) c * n + tid] * vd[c];
            double bl;
            if (tid < CPARS) bl = ws->cPrior[tid] * (double) ws->calib.cDeltaF[tid];
            else {
                const FrameDev &f = ws->fr[(tid - CPARS) >> 3];
                bl = f.prior[(tid - CPARS) & 7] * f.delta_prior[(tid - CPARS) & 7];
            }
            const double bf = bl + bm + sb.b_A[tid] - sb.b_sc[tid];
            vb[tid] = bf;
            sb.lastbS[tid] = bf;
        }
        __syncthreads();
        // lastHS = HFinal_top - H_sc ; diag *= (1+lambda) ; HFinal_top -= H_sc / (1+lambda)   (:286-291)
        const double inv1l = 1.0 / (1.0 + lambda);
        for (int e = tid; e < n * n; e += K3_THREADS) {
            const int r = e % n, c = e / n;
            const double hsc = sb.H_sc[e];
            double v = A0[r * K3_LD + c];
            sb.lastHS[e] = v - hsc;
            if (r == c) v *= (1 + lambda);
            A0[r * K3_LD + c] = v - hsc * inv1l;
        }
        __syncthreads();
        // SVecI = (diag + 10)^-1/2 ; scale (:326-327)
        if (tid < n) vS[tid] = 1.0 / sqrt(A0[tid * K3_LD + tid] + 10.0);
        __syncthreads();
        for (int e = tid; e < n * n; e += K3_THREADS) {
            const int r = e / n, c = e % n;
            A0[r * K3_LD + c] *= vS[r] * vS[c];
        }
        if (tid < n) vb[tid] *= vS[tid];
        __syncthreads();
        // Eigen::LDLT pivots on the largest remaining |diagonal| of the *input* matrix (its left-looking update
        // never touches later diagonal entries), i.e. a descending-|diag| ordering: rank sort.
        if (tid < n) {
            const double di = fabs(A0[tid * K3_LD + tid]);
            int rank = 0;
            for (int j = 0; j < n; j++) {
                const double dj = fabs(A0[j * K3_LD + j]);
                rank += (dj > di) || (dj == di && j < tid);
            }
            perm[rank] = tid;
        }
        __syncthreads();
        for (int e = tid; e < n * n; e += K3_THREADS) {
            const int r = e / n, c = e % n;
            A[r * K3_LD + c] = A0[perm[r] * K3_LD + perm[c]];
        }
        if (tid < n) vd[tid] = vb[perm[tid]];
        __syncthreads();
        if (tid < n) vb[tid] = vd[tid];
        __syncthreads();
        // in-place LDL^T (lower): A[i][k] <- L(i,k), A[k][k] <- D(k)
        for (int k = 0; k < n; k++) {
            const double dk = A[k * K3_LD + k];
            const bool valid = fabs(dk) > 0.0;
            const int rs = n - k - 1;
            if (tid < rs) {
                const int i = k + 1 + tid;
                vd[i] = A[i * K3_LD + k];                      // unscaled column (l_i * d_k)
                if (valid) A[i * K3_LD + k] = vd[i] / dk;
            }
            __syncthreads();
            // trailing update (lower triangle): A[i][j] -= L(i,k) * (d_k L(j,k)) = A[i][k]*vd[j]
            for (int i = k + 1 + (tid >> 4); i < n; i += 16) {
                const double lik = A[i * K3_LD + k];
                for (int j = k + 1 + (tid & 15); j <= i; j += 16) A[i * K3_LD + j] -= lik * vd[j];
            }
            __syncthreads();
        }
        // solve: L z = b
        for (int k = 0; k < n; k++) {
            const double zk = vb[k];
            const int i = k + 1 + tid;
            if (i < n) vb[i] -= A[i * K3_LD + k] * zk;
            __syncthreads();
        }
        if (tid < n) {
            const double dk = A[tid * K3_LD + tid];
            vb[tid] = (fabs(dk) > 2.2250738585072014e-308) ? vb[tid] / dk : 0.0;
        }
        __syncthreads();
        for (int k = n - 1; k >= 0; k--) {
            const double xk = vb[k];
            if (tid < k) vb[tid] -= A[k * K3_LD + tid] * xk;
            __syncthreads();
        }
        if (tid < n) vx[perm[tid]] = vb[tid];
        __syncthreads();
        if (tid < n) vx[tid] *= vS[tid];
        __syncthreads();
        // orthogonalize(&x, 0) when iteration >= 2 (SOLVER_ORTHOGONALIZE_X_LATER, :339-343): x -= NNpiTS * x
        if (iteration >= 2) {
            double px = 0.0;
            if (tid < n) for (int c = 0; c < n; c++) px += sb.Pns[(size_t) c * n + tid] * vx[c];
            __syncthreads();
            if (tid < n) vx[tid] -= px;
            __syncthreads();
        }
        if (tid < n) sb.lastX[tid] = vx[tid];
        // resubstituteF_MT frame part (:495-507)
        if (tid < CPARS) {
            ws->calib.step[tid] = -vx[tid];
            ws->cstep[tid] = (float) vx[tid];
        }
        if (tid >= 32 && tid < 32 + nF) {
            const int h = tid - 32;
            for (int i = 0; i < 8; i++) ws->fr[h].step[i] = -vx[CPARS + 8 * h + i];
            ws->fr[h].step[8] = ws->fr[h].step[9] = 0.0;
        }
        if (tid >= 64 && tid < 64 + nF * nF) {
            const int q = tid - 64, h = q / nF, t = q % nF;     // xAd[nFrames*h + t]
            const float *AH = ws->adHostF[h + nF * t], *AT = ws->adTargetF[h + nF * t];
            for (int j = 0; j < 8; j++) {
                float s1 = 0.f, s2 = 0.f;
                for (int i = 0; i < 8; i++) s1 += (float) vx[CPARS + 8 * h + i] * AH[i * 8 + j];
                for (int i = 0; i < 8; i++) s2 += (float) vx[CPARS + 8 * t + i] * AT[i * 8 + j];
                ws->xAd[nF * h + t][j] = s1 + s2;
            }
        }
        __syncthreads();
    }
    if (flags & K3F_STEP) {
        // doStepFromBackup(1,1,1,1,1), frame/calib part (FullSystem.cc:1588-1597,1617-1627)
        if (tid == 0) {
            double nv[4];
            for (int i = 0; i < 4; i++) nv[i] = ws->calib.value_backup[i] + ws->calib.step[i];
            calib_set_value(ws->calib, nv);
            float sumA = 0, sumB = 0, sumT = 0, sumR = 0;
            for (int h = 0; h < nF; h++) {
                const double *st = ws->fr[h].step;
                sumA += st[6] * st[6];
                sumB += st[7] * st[7];
                sumT += st[0] * st[0] + st[1] * st[1] + st[2] * st[2];
                sumR += st[3] * st[3] + st[4] * st[4] + st[5] * st[5];
            }
            sumA /= nF; sumB /= nF; sumR /= nF; sumT /= nF;
            const float sumNID = ws->sumNID / ws->numID;
            const float thO = ws->S.thOptIterations;
            ws->canbreak = (sqrtf(sumA) < 0.0005 * thO && sqrtf(sumB) < 0.00005 * thO && sqrtf(sumR) < 0.00005 * thO &&
                            sqrtf(sumT) * sumNID < 0.00005 * thO) ? 1 : 0;
        }
        if (tid >= 32 && tid < 32 + nF) {
            FrameDev &f = ws->fr[tid - 32];
            for (int i = 0; i < 10; i++) f.state[i] = f.state_backup[i] + f.step[i];
        }
        __syncthreads();
        frames_refresh(ws);
    }
    if ((flags & K3F_SOLVE) && tid == 0) *iteration_dev = iteration + 1;
}
#define K3_SMEM_BYTES ((2 * MAXN * K3_LD + 4 * MAXN) * sizeof(double) + MAXN * sizeof(int) + 64)

// ---------------------------------------------------------------------------------------------------------
// piecewise helpers
// PointFrameResidual::applyRes(true) on every active residual (Residuals.h:70-87)
__global__ void k_apply_res(DevWindow d) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= d.nR) return;
    if (d.res_lin[r]) return;
    if (d.res_state[r] == LDSO_B200_RES_OOB) return;
    const uint8_t ns = d.res_new_state[r];
    if (ns == LDSO_B200_RES_IN) {
        d.res_active[r] = 1;
        for (int i = 0; i < 8; i++) d.res_JpJdF[8 * r + i] = d.res_JpJdF_new[8 * r + i];
    } else d.res_active[r] = 0;
    d.res_state[r] = ns;
    d.res_energy[r] = d.res_new_energy[r];
}

// what: 1 = backupState (idepth_backup = idepth), 2 = resubstituteFPt (step only), 4 = apply step
__global__ void k_points(DevWindow d, const WinState *__restrict__ ws, int what) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.nP) return;
    const int nF = ws->nF;
    if (what & 1) d.pt_idepth_backup[p] = d.pt_idepth[p];
    if (what & 2) {
        const int host = d.pt_host[p];
        const int r0 = d.pt_res_begin[p], r1 = d.pt_res_begin[p + 1];
        int ngood = 0;
        float b = d.pt_bdSumF[p];
        {
            float s = 0.f;
            for (int i = 0; i < 4; i++) s += ws->cstep[i] * d.pt_Hcd[4 * p + i];
            b -= s;
        }
        for (int r = r0; r < r1; r++) {
            if (!d.res_active[r]) continue;
            ngood++;
            const float *xa = ws->xAd[host * nF + d.res_target[r]];
            float s = 0.f;
            for (int i = 0; i < 8; i++) s += xa[i] * d.res_JpJdF[8 * r + i];
            b -= s;
        }
        if (ngood == 0) d.pt_step[p] = 0.f;
        else if (isfinite(b)) d.pt_step[p] = -b * d.pt_HdiF[p];
    }
    if (what & 4) {
        const float nid = d.pt_idepth_backup[p] + d.pt_step[p];
        d.pt_idepth[p] = nid;
        d.pt_idepth_zero[p] = nid;
    }
}

// sum |idepth_backup| for canbreak in the piecewise do_step
__global__ void k_sum_nid(DevWindow d, WinState *ws) {
    __shared__ float sh[256];
    float s = 0.f;
    for (int p = threadIdx.x; p < d.nP; p += 256) s += fabsf(d.pt_idepth_backup[p]);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) { ws->sumNID = sh[0]; ws->numID = (float) d.nP; }
}
