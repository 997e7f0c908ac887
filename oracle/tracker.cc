// ORACLE — TEST INFRASTRUCTURE ONLY (see omath.h header). PARITY PINNED AGAINST THE REFERENCE'S OWN SOURCES, EXCEPT EIGEN / SOPHUS INTERNALS (oracle/ref_pin compiles, unmodified, the reference's Residuals.cc, ImmaturePoint.cc, PointHessian.cc, FrameHessian.cc, FrameFramePrecalc.cc, AccumulatedTopHessian.cc, AccumulatedSCHessian.cc, EnergyFunctional.cc, CoarseTracker.cc, Setting.cc and their headers against stand-in Eigen/Sophus headers and matches the oracle bit for bit on 108 checks; the float SSE / scalar code is pinned completely, the Eigen-expression code structurally - what is added where, in which order - with LDLT, PartialPivLU, JacobiSVD, SE3 exp/log and the product kernels being the oracle's own restatements on both sides; FullSystem.cc's driver loop needs the whole front end and stays restated from the cited lines).
// CPU restatement of src/frontend/CoarseTracker.cc of the reference (file:line cited per function).
#include "tracker.h"

namespace oracle {

static void m33f_inverse_t(const float *m, float *inv) {  // Eigen 3x3 inverse: cofactors * (1/det)
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

// CoarseTracker::CoarseTracker — CoarseTracker.cc:30-59
CoarseTracker::CoarseTracker(int ww, int hh, int levels) {
    pyrLevelsUsed = levels;
    for (int lvl = 0; lvl < levels; lvl++) {
        int wl = ww >> lvl, hl = hh >> lvl;
        idepth[lvl].assign((size_t) wl * hl, 0.f);
        weightSums[lvl].assign((size_t) wl * hl, 0.f);
        weightSums_bak[lvl].assign((size_t) wl * hl, 0.f);
        pc_u[lvl].assign((size_t) wl * hl, 0.f);
        pc_v[lvl].assign((size_t) wl * hl, 0.f);
        pc_idepth[lvl].assign((size_t) wl * hl, 0.f);
        pc_color[lvl].assign((size_t) wl * hl, 0.f);
        pc_n[lvl] = 0;
        refDIp[lvl] = newDIp[lvl] = nullptr;
        w[lvl] = wl;
        h[lvl] = hl;
    }
    size_t n = (size_t) ww * hh;
    buf_warped_idepth.assign(n, 0.f); buf_warped_u.assign(n, 0.f); buf_warped_v.assign(n, 0.f);
    buf_warped_dx.assign(n, 0.f); buf_warped_dy.assign(n, 0.f); buf_warped_residual.assign(n, 0.f);
    buf_warped_weight.assign(n, 0.f); buf_warped_refColor.assign(n, 0.f);
    for (int i = 0; i < 5; i++) lastResiduals[i] = NAN;
    for (int i = 0; i < 3; i++) lastFlowIndicators[i] = 1000;
}

// CoarseTracker::makeK — :219-246
void CoarseTracker::makeK(float fxl, float fyl, float cxl, float cyl) {
    fx[0] = fxl; fy[0] = fyl; cx[0] = cxl; cy[0] = cyl;
    for (int level = 1; level < pyrLevelsUsed; ++level) {
        fx[level] = fx[level - 1] * 0.5;
        fy[level] = fy[level - 1] * 0.5;
        cx[level] = (cx[0] + 0.5) / ((int) 1 << level) - 0.5;
        cy[level] = (cy[0] + 0.5) / ((int) 1 << level) - 0.5;
    }
    for (int level = 0; level < pyrLevelsUsed; ++level) {
        float *Kl = K[level];
        Kl[0] = fx[level]; Kl[1] = 0; Kl[2] = cx[level];
        Kl[3] = 0; Kl[4] = fy[level]; Kl[5] = cy[level];
        Kl[6] = 0; Kl[7] = 0; Kl[8] = 1;
        m33f_inverse_t(Kl, Ki[level]);
        fxi[level] = Ki[level][0];
        fyi[level] = Ki[level][4];
        cxi[level] = Ki[level][2];
        cyi[level] = Ki[level][5];
    }
}

// CoarseTracker::makeCoarseDepthL0 — :258-438
void CoarseTracker::makeCoarseDepthL0(int n, const float *cpt, const float *HdiF) {
    std::fill(idepth[0].begin(), idepth[0].end(), 0.f);
    std::fill(weightSums[0].begin(), weightSums[0].end(), 0.f);
    for (int k = 0; k < n; k++) {
        int u = cpt[3 * k + 0] + 0.5f;
        int v = cpt[3 * k + 1] + 0.5f;
        float new_idepth = cpt[3 * k + 2];
        float weight = sqrtf(1e-3 / (HdiF[k] + 1e-12));
        idepth[0][u + w[0] * v] += new_idepth * weight;
        weightSums[0][u + w[0] * v] += weight;
    }
    for (int lvl = 1; lvl < pyrLevelsUsed; lvl++) {
        int lvlm1 = lvl - 1;
        int wl = w[lvl], hl = h[lvl], wlm1 = w[lvlm1];
        float *idepth_l = idepth[lvl].data(), *weightSums_l = weightSums[lvl].data();
        float *idepth_lm = idepth[lvlm1].data(), *weightSums_lm = weightSums[lvlm1].data();
        for (int y = 0; y < hl; y++)
            for (int x = 0; x < wl; x++) {
                int bidx = 2 * x + 2 * y * wlm1;
                idepth_l[x + y * wl] = idepth_lm[bidx] + idepth_lm[bidx + 1] + idepth_lm[bidx + wlm1] + idepth_lm[bidx + wlm1 + 1];
                weightSums_l[x + y * wl] = weightSums_lm[bidx] + weightSums_lm[bidx + 1] + weightSums_lm[bidx + wlm1] +
                                           weightSums_lm[bidx + wlm1 + 1];
            }
    }
    // dilate idepth by 1 (diagonal neighbours) on levels 0,1 — :312-355
    for (int lvl = 0; lvl < 2 && lvl < pyrLevelsUsed; lvl++) {
        int wh = w[lvl] * h[lvl] - w[lvl];
        int wl = w[lvl];
        float *weightSumsl = weightSums[lvl].data();
        float *weightSumsl_bak = weightSums_bak[lvl].data();
        memcpy(weightSumsl_bak, weightSumsl, (size_t) w[lvl] * h[lvl] * sizeof(float));
        float *idepthl = idepth[lvl].data();
        for (int i = w[lvl]; i < wh; i++) {
            if (weightSumsl_bak[i] <= 0) {
                float sum = 0, num = 0, numn = 0;
                if (weightSumsl_bak[i + 1 + wl] > 0) { sum += idepthl[i + 1 + wl]; num += weightSumsl_bak[i + 1 + wl]; numn++; }
                if (weightSumsl_bak[i - 1 - wl] > 0) { sum += idepthl[i - 1 - wl]; num += weightSumsl_bak[i - 1 - wl]; numn++; }
                if (weightSumsl_bak[i + wl - 1] > 0) { sum += idepthl[i + wl - 1]; num += weightSumsl_bak[i + wl - 1]; numn++; }
                if (weightSumsl_bak[i - wl + 1] > 0) { sum += idepthl[i - wl + 1]; num += weightSumsl_bak[i - wl + 1]; numn++; }
                if (numn > 0) { idepthl[i] = sum / numn; weightSumsl[i] = num / numn; }
            }
        }
    }
    // dilate idepth by 1 (4-neighbourhood) on levels >= 2 — :358-395
    for (int lvl = 2; lvl < pyrLevelsUsed; lvl++) {
        int wh = w[lvl] * h[lvl] - w[lvl];
        int wl = w[lvl];
        float *weightSumsl = weightSums[lvl].data();
        float *weightSumsl_bak = weightSums_bak[lvl].data();
        memcpy(weightSumsl_bak, weightSumsl, (size_t) w[lvl] * h[lvl] * sizeof(float));
        float *idepthl = idepth[lvl].data();
        for (int i = w[lvl]; i < wh; i++) {
            if (weightSumsl_bak[i] <= 0) {
                float sum = 0, num = 0, numn = 0;
                if (weightSumsl_bak[i + 1] > 0) { sum += idepthl[i + 1]; num += weightSumsl_bak[i + 1]; numn++; }
                if (weightSumsl_bak[i - 1] > 0) { sum += idepthl[i - 1]; num += weightSumsl_bak[i - 1]; numn++; }
                if (weightSumsl_bak[i + wl] > 0) { sum += idepthl[i + wl]; num += weightSumsl_bak[i + wl]; numn++; }
                if (weightSumsl_bak[i - wl] > 0) { sum += idepthl[i - wl]; num += weightSumsl_bak[i - wl]; numn++; }
                if (numn > 0) { idepthl[i] = sum / numn; weightSumsl[i] = num / numn; }
            }
        }
    }
    // normalize idepths and weights — :398-437
    for (int lvl = 0; lvl < pyrLevelsUsed; lvl++) {
        float *weightSumsl = weightSums[lvl].data();
        float *idepthl = idepth[lvl].data();
        const float *dIRefl = refDIp[lvl];
        int wl = w[lvl], hl = h[lvl];
        int lpc_n = 0;
        float *lpc_u = pc_u[lvl].data(), *lpc_v = pc_v[lvl].data();
        float *lpc_idepth = pc_idepth[lvl].data(), *lpc_color = pc_color[lvl].data();
        for (int y = 2; y < hl - 2; y++)
            for (int x = 2; x < wl - 2; x++) {
                int i = x + y * wl;
                if (weightSumsl[i] > 0) {
                    idepthl[i] /= weightSumsl[i];
                    lpc_u[lpc_n] = x;
                    lpc_v[lpc_n] = y;
                    lpc_idepth[lpc_n] = idepthl[i];
                    lpc_color[lpc_n] = dIRefl[3 * i];
                    if (!std::isfinite(lpc_color[lpc_n]) || !(idepthl[i] > 0)) {
                        idepthl[i] = -1;
                        continue;
                    }
                    lpc_n++;
                } else
                    idepthl[i] = -1;
                weightSumsl[i] = 1;
            }
        pc_n[lvl] = lpc_n;
    }
}

// CoarseTracker::calcRes — :440-572
void CoarseTracker::calcRes(int lvl, const SE3 &refToNew, float aff_a, float aff_b, float cutoffTH, double rs[6]) {
    float E = 0;
    int numTermsInE = 0, numTermsInWarped = 0, numSaturated = 0;
    int wl = w[lvl], hl = h[lvl];
    const float *dINewl = newDIp[lvl];
    float fxl = fx[lvl], fyl = fy[lvl], cxl = cx[lvl], cyl = cy[lvl];

    M3 Rd = refToNew.rotationMatrix();
    float Rf[9], RKi[9];
    for (int i = 0; i < 9; i++) Rf[i] = (float) Rd.m[i];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float s = Rf[i * 3 + 0] * Ki[lvl][0 * 3 + j];
            s += Rf[i * 3 + 1] * Ki[lvl][1 * 3 + j];
            s += Rf[i * 3 + 2] * Ki[lvl][2 * 3 + j];
            RKi[i * 3 + j] = s;
        }
    float t[3] = {(float) refToNew.t[0], (float) refToNew.t[1], (float) refToNew.t[2]};
    double ab[2];
    fromToVecExposure(lastRef_ab_exposure, newFrame_ab_exposure, lastRef_aff_a, lastRef_aff_b, aff_a, aff_b, ab);
    float affLL[2] = {(float) ab[0], (float) ab[1]};

    float sumSquaredShiftT = 0, sumSquaredShiftRT = 0, sumSquaredShiftNum = 0;
    float maxEnergy = 2 * S.huberTH * cutoffTH - S.huberTH * S.huberTH;

    int nl = pc_n[lvl];
    const float *lpc_u = pc_u[lvl].data(), *lpc_v = pc_v[lvl].data();
    const float *lpc_idepth = pc_idepth[lvl].data(), *lpc_color = pc_color[lvl].data();
    const float *Kil = Ki[lvl];

    for (int i = 0; i < nl; i++) {
        float id = lpc_idepth[i];
        float x = lpc_u[i];
        float y = lpc_v[i];
        float pt[3];
        for (int r = 0; r < 3; r++) {
            float s = RKi[r * 3 + 0] * x;
            s += RKi[r * 3 + 1] * y;
            s += RKi[r * 3 + 2] * 1.0f;
            pt[r] = s + t[r] * id;
        }
        float u = pt[0] / pt[2];
        float v = pt[1] / pt[2];
        float Ku = fxl * u + cxl;
        float Kv = fyl * v + cyl;
        float new_idepth = id / pt[2];

        if (lvl == 0 && i % 32 == 0) {
            float kp[3];
            for (int r = 0; r < 3; r++) {
                float s = Kil[r * 3 + 0] * x;
                s += Kil[r * 3 + 1] * y;
                s += Kil[r * 3 + 2] * 1.0f;
                kp[r] = s;
            }
            float ptT[3] = {kp[0] + t[0] * id, kp[1] + t[1] * id, kp[2] + t[2] * id};
            float uT = ptT[0] / ptT[2], vT = ptT[1] / ptT[2];
            float KuT = fxl * uT + cxl, KvT = fyl * vT + cyl;
            float ptT2[3] = {kp[0] - t[0] * id, kp[1] - t[1] * id, kp[2] - t[2] * id};
            float uT2 = ptT2[0] / ptT2[2], vT2 = ptT2[1] / ptT2[2];
            float KuT2 = fxl * uT2 + cxl, KvT2 = fyl * vT2 + cyl;
            float pt3[3];
            for (int r = 0; r < 3; r++) {
                float s = RKi[r * 3 + 0] * x;
                s += RKi[r * 3 + 1] * y;
                s += RKi[r * 3 + 2] * 1.0f;
                pt3[r] = s - t[r] * id;
            }
            float u3 = pt3[0] / pt3[2], v3 = pt3[1] / pt3[2];
            float Ku3 = fxl * u3 + cxl, Kv3 = fyl * v3 + cyl;
            sumSquaredShiftT += (KuT - x) * (KuT - x) + (KvT - y) * (KvT - y);
            sumSquaredShiftT += (KuT2 - x) * (KuT2 - x) + (KvT2 - y) * (KvT2 - y);
            sumSquaredShiftRT += (Ku - x) * (Ku - x) + (Kv - y) * (Kv - y);
            sumSquaredShiftRT += (Ku3 - x) * (Ku3 - x) + (Kv3 - y) * (Kv3 - y);
            sumSquaredShiftNum += 2;
        }

        if (!(Ku > 2 && Kv > 2 && Ku < wl - 3 && Kv < hl - 3 && new_idepth > 0)) continue;

        float refColor = lpc_color[i];
        float hitColor[3];
        getInterpolatedElement33(dINewl, Ku, Kv, wl, hitColor);
        if (!std::isfinite((float) hitColor[0])) continue;
        float residual = hitColor[0] - (float) (affLL[0] * refColor + affLL[1]);
        float hw = fabsf(residual) < S.huberTH ? 1 : S.huberTH / fabsf(residual);

        if (fabsf(residual) > cutoffTH) {
            E += maxEnergy;
            numTermsInE++;
            numSaturated++;
        } else {
            E += hw * residual * residual * (2 - hw);
            numTermsInE++;
            buf_warped_idepth[numTermsInWarped] = new_idepth;
            buf_warped_u[numTermsInWarped] = u;
            buf_warped_v[numTermsInWarped] = v;
            buf_warped_dx[numTermsInWarped] = hitColor[1];
            buf_warped_dy[numTermsInWarped] = hitColor[2];
            buf_warped_residual[numTermsInWarped] = residual;
            buf_warped_weight[numTermsInWarped] = hw;
            buf_warped_refColor[numTermsInWarped] = lpc_color[i];
            numTermsInWarped++;
        }
    }
    while (numTermsInWarped % 4 != 0) {
        buf_warped_idepth[numTermsInWarped] = 0;
        buf_warped_u[numTermsInWarped] = 0;
        buf_warped_v[numTermsInWarped] = 0;
        buf_warped_dx[numTermsInWarped] = 0;
        buf_warped_dy[numTermsInWarped] = 0;
        buf_warped_residual[numTermsInWarped] = 0;
        buf_warped_weight[numTermsInWarped] = 0;
        buf_warped_refColor[numTermsInWarped] = 0;
        numTermsInWarped++;
    }
    buf_warped_n = numTermsInWarped;
    rs[0] = E;
    rs[1] = numTermsInE;
    rs[2] = sumSquaredShiftT / (sumSquaredShiftNum + 0.1);
    rs[3] = 0;
    rs[4] = sumSquaredShiftRT / (sumSquaredShiftNum + 0.1);
    rs[5] = numSaturated / (float) numTermsInE;
}

// CoarseTracker::calcGSSSE — :574-632
void CoarseTracker::calcGSSSE(int lvl, double H_out[64], double b_out[8], const SE3 &refToNew, float aff_a, float aff_b) {
    (void) refToNew;
    acc.initialize();
    float fxl = fx[lvl], fyl = fy[lvl];
    float b0 = lastRef_aff_b;
    double ab[2];
    fromToVecExposure(lastRef_ab_exposure, newFrame_ab_exposure, lastRef_aff_a, lastRef_aff_b, aff_a, aff_b, ab);
    float a = (float) ab[0];
    int n = buf_warped_n;
    for (int i = 0; i < n; i += 4) {
        float J[9][4], wv[4];
        for (int l = 0; l < 4; l++) {
            float dx = buf_warped_dx[i + l] * fxl;
            float dy = buf_warped_dy[i + l] * fyl;
            float u = buf_warped_u[i + l];
            float v = buf_warped_v[i + l];
            float id = buf_warped_idepth[i + l];
            J[0][l] = id * dx;
            J[1][l] = id * dy;
            J[2][l] = 0.0f - id * (u * dx + v * dy);
            J[3][l] = 0.0f - ((u * v) * dx + dy * (1.0f + v * v));
            J[4][l] = (u * v) * dy + dx * (1.0f + u * u);
            J[5][l] = u * dy - v * dx;
            J[6][l] = a * (b0 - buf_warped_refColor[i + l]);
            J[7][l] = -1.0f;
            J[8][l] = buf_warped_residual[i + l];
            wv[l] = buf_warped_weight[i + l];
        }
        acc.updateSSE_eighted(J, wv);
    }
    acc.finish();
    double fac = (double) (1.0f / n);
    for (int r = 0; r < 8; r++) {
        for (int c = 0; c < 8; c++) H_out[r * 8 + c] = (double) acc.H[r * 9 + c] * fac;
        b_out[r] = (double) acc.H[r * 9 + 8] * fac;
    }
    const double sc[8] = {SCALE_XI_ROT, SCALE_XI_ROT, SCALE_XI_ROT, SCALE_XI_TRANS, SCALE_XI_TRANS, SCALE_XI_TRANS,
                          SCALE_A, SCALE_B};
    for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) H_out[r * 8 + c] *= sc[c];
    for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) H_out[r * 8 + c] *= sc[r];
    for (int r = 0; r < 8; r++) b_out[r] *= sc[r];
}

static void solve_sub(const double *Hl, const double *b, int n, double *inc) {  // Hl.topLeftCorner<n,n>().ldlt().solve(-b.head<n>())
    MatX A(n, n);
    VecXd rhs(n);
    for (int r = 0; r < n; r++) {
        for (int c = 0; c < n; c++) A(r, c) = Hl[r * 8 + c];
        rhs[r] = -b[r];
    }
    VecXd x = ldlt_solve(A, rhs);
    for (int r = 0; r < n; r++) inc[r] = x[r];
}

// CoarseTracker::trackNewestCoarse — :61-217
bool CoarseTracker::trackNewestCoarse(SE3 &lastToNew_out, float &aff_a_out, float &aff_b_out, int coarsestLvl,
                                      const double minResForAbort[5]) {
    for (int i = 0; i < 5; i++) lastResiduals[i] = NAN;
    for (int i = 0; i < 3; i++) lastFlowIndicators[i] = 1000;
    lm_iterations_total = 0;
    int maxIterations[] = {10, 20, 50, 50, 50};
    float lambdaExtrapolationLimit = 0.001;
    SE3 refToNew_current = lastToNew_out;
    float aff_a_cur = aff_a_out, aff_b_cur = aff_b_out;
    bool haveRepeated = false;

    for (int lvl = coarsestLvl; lvl >= 0; lvl--) {
        double H[64], b[8];
        float levelCutoffRepeat = 1;
        double resOld[6];
        calcRes(lvl, refToNew_current, aff_a_cur, aff_b_cur, S.coarseCutoffTH * levelCutoffRepeat, resOld);
        lm_iterations_total++;
        while (resOld[5] > 0.6 && levelCutoffRepeat < 50) {
            levelCutoffRepeat *= 2;
            calcRes(lvl, refToNew_current, aff_a_cur, aff_b_cur, S.coarseCutoffTH * levelCutoffRepeat, resOld);
            lm_iterations_total++;
        }
        calcGSSSE(lvl, H, b, refToNew_current, aff_a_cur, aff_b_cur);
        float lambda = 0.01;

        for (int iteration = 0; iteration < maxIterations[lvl]; iteration++) {
            double Hl[64];
            memcpy(Hl, H, sizeof(Hl));
            for (int i = 0; i < 8; i++) Hl[i * 8 + i] *= (1 + lambda);
            double inc[8];
            solve_sub(Hl, b, 8, inc);
            if (S.affineOptModeA < 0 && S.affineOptModeB < 0) {
                solve_sub(Hl, b, 6, inc);
                inc[6] = inc[7] = 0;
            }
            if (!(S.affineOptModeA < 0) && S.affineOptModeB < 0) {
                solve_sub(Hl, b, 7, inc);
                inc[7] = 0;
            }
            if (S.affineOptModeA < 0 && !(S.affineOptModeB < 0)) {
                double HlStitch[64], bStitch[8];
                memcpy(HlStitch, Hl, sizeof(Hl));
                memcpy(bStitch, b, sizeof(bStitch));
                for (int r = 0; r < 8; r++) HlStitch[r * 8 + 6] = HlStitch[r * 8 + 7];
                for (int c = 0; c < 8; c++) HlStitch[6 * 8 + c] = HlStitch[7 * 8 + c];
                bStitch[6] = bStitch[7];
                double incStitch[8];
                solve_sub(HlStitch, bStitch, 7, incStitch);
                for (int i = 0; i < 8; i++) inc[i] = 0;
                for (int i = 0; i < 6; i++) inc[i] = incStitch[i];
                inc[6] = 0;
                inc[7] = incStitch[6];
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
            if (!std::isfinite(sum)) for (int i = 0; i < 8; i++) incScaled[i] = 0;

            SE3 refToNew_new = SE3::exp(incScaled) * refToNew_current;
            float aff_a_new = aff_a_cur, aff_b_new = aff_b_cur;
            aff_a_new += incScaled[6];
            aff_b_new += incScaled[7];

            double resNew[6];
            calcRes(lvl, refToNew_new, aff_a_new, aff_b_new, S.coarseCutoffTH * levelCutoffRepeat, resNew);
            lm_iterations_total++;
            bool accept = (resNew[0] / resNew[1]) < (resOld[0] / resOld[1]);
            if (accept) {
                calcGSSSE(lvl, H, b, refToNew_new, aff_a_new, aff_b_new);
                memcpy(resOld, resNew, sizeof(resOld));
                aff_a_cur = aff_a_new;
                aff_b_cur = aff_b_new;
                refToNew_current = refToNew_new;
                lambda *= 0.5;
            } else {
                lambda *= 4;
                if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
            }
            double nrm = 0;
            for (int i = 0; i < 8; i++) nrm += inc[i] * inc[i];
            if (!(std::sqrt(nrm) > 1e-3)) break;
        }
        lastResiduals[lvl] = sqrtf((float) (resOld[0] / resOld[1]));
        for (int i = 0; i < 3; i++) lastFlowIndicators[i] = resOld[2 + i];
        if (lastResiduals[lvl] > 1.5 * minResForAbort[lvl]) return false;
        if (levelCutoffRepeat > 1 && !haveRepeated) {
            lvl++;
            haveRepeated = true;
        }
    }
    lastToNew_out = refToNew_current;
    aff_a_out = aff_a_cur;
    aff_b_out = aff_b_cur;
    if ((S.affineOptModeA != 0 && (fabsf(aff_a_out) > 1.2)) || (S.affineOptModeB != 0 && (fabsf(aff_b_out) > 200)))
        return false;
    double rel[2];
    fromToVecExposure(lastRef_ab_exposure, newFrame_ab_exposure, lastRef_aff_a, lastRef_aff_b, aff_a_out, aff_b_out, rel);
    float relAff[2] = {(float) rel[0], (float) rel[1]};
    if ((S.affineOptModeA == 0 && (fabsf(logf((float) relAff[0])) > 1.5)) ||
        (S.affineOptModeB == 0 && (fabsf((float) relAff[1]) > 200)))
        return false;
    if (S.affineOptModeA < 0) aff_a_out = 0;
    if (S.affineOptModeB < 0) aff_b_out = 0;
    return true;
}

// ------------------------------------------------------------------------------------------
// CoarseDistanceMap — CoarseTracker.cc:634-870
CoarseDistanceMap::CoarseDistanceMap(int ww, int hh, int levels) {
    pyrLevelsUsed = levels;
    fwdWarpedIDDistFinal.assign((size_t) ww * hh / 4, 0.f);
    bfsList1.assign((size_t) ww * hh / 4 * 2, 0);
    bfsList2.assign((size_t) ww * hh / 4 * 2, 0);
    for (int l = 0; l < PYR_LEVELS; l++) w[l] = h[l] = 0;
    w[0] = ww; h[0] = hh;       // the reference takes them from wG / hG in makeK
}

void CoarseDistanceMap::makeK(float fxl, float fyl, float cxl, float cyl) {   // :657-685 (same arithmetic as CoarseTracker::makeK)
    fx[0] = fxl; fy[0] = fyl; cx[0] = cxl; cy[0] = cyl;
    for (int level = 1; level < pyrLevelsUsed; ++level) {
        w[level] = w[0] >> level;
        h[level] = h[0] >> level;
        fx[level] = fx[level - 1] * 0.5;
        fy[level] = fy[level - 1] * 0.5;
        cx[level] = (cx[0] + 0.5) / ((int) 1 << level) - 0.5;
        cy[level] = (cy[0] + 0.5) / ((int) 1 << level) - 0.5;
    }
    for (int level = 0; level < pyrLevelsUsed; ++level) {
        float *Kl = K[level];
        Kl[0] = fx[level]; Kl[1] = 0; Kl[2] = cx[level];
        Kl[3] = 0; Kl[4] = fy[level]; Kl[5] = cy[level];
        Kl[6] = 0; Kl[7] = 0; Kl[8] = 1;
        m33f_inverse_t(Kl, Ki[level]);
    }
}

static void m33f_mul_t(const float *a, const float *b, float *c) {   // 3x3 float product, row . column accumulated left to right
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float s = a[i * 3 + 0] * b[0 * 3 + j];
            s += a[i * 3 + 1] * b[1 * 3 + j];
            s += a[i * 3 + 2] * b[2 * 3 + j];
            c[i * 3 + j] = s;
        }
}

// KRKi = K[1] * R * Ki[0], Kt = K[1] * t (CoarseTracker.cc:705-706, FullSystem.cc:1093-1094)
void CoarseDistanceMap::hostToNewest(const float R[9], const float t[3], float KRKi[9], float Kt[3]) const {
    float KR[9];
    m33f_mul_t(K[1], R, KR);
    m33f_mul_t(KR, Ki[0], KRKi);
    for (int i = 0; i < 3; i++) {
        float s = K[1][i * 3 + 0] * t[0];
        s += K[1][i * 3 + 1] * t[1];
        s += K[1][i * 3 + 2] * t[2];
        Kt[i] = s;
    }
}

void CoarseDistanceMap::beginDistanceMap() {   // :690-698
    int wh1 = w[1] * h[1];
    for (int i = 0; i < wh1; i++) fwdWarpedIDDistFinal[i] = 1000;
    numItems = 0;
}

void CoarseDistanceMap::addFramePoints(const float R[9], const float t[3], int n, const float *u_, const float *v_, const float *idepth_scaled) {   // :701-722
    float KRKi[9], Kt[3];
    hostToNewest(R, t, KRKi, Kt);
    int w1 = w[1];
    for (int i = 0; i < n; i++) {
        float ptp[3];
        for (int r = 0; r < 3; r++) {
            float s = KRKi[r * 3 + 0] * u_[i];
            s += KRKi[r * 3 + 1] * v_[i];
            s += KRKi[r * 3 + 2] * 1.0f;
            ptp[r] = s + Kt[r] * idepth_scaled[i];
        }
        int u = ptp[0] / ptp[2] + 0.5f;
        int v = ptp[1] / ptp[2] + 0.5f;
        if (!(u > 0 && v > 0 && u < w[1] && v < h[1])) continue;
        fwdWarpedIDDistFinal[u + w1 * v] = 0;
        bfsList1[2 * numItems] = u; bfsList1[2 * numItems + 1] = v;
        numItems++;
    }
}

void CoarseDistanceMap::growDistBFS(int bfsNum) {   // :728-812
    int w1 = w[1], h1 = h[1];
    float *D = fwdWarpedIDDistFinal.data();
    for (int k = 1; k < 40; k++) {
        int bfsNum2 = bfsNum;
        std::swap(bfsList1, bfsList2);
        bfsNum = 0;
        const int n4[4] = {1, -1, w1, -w1}, dx4[4] = {1, -1, 0, 0}, dy4[4] = {0, 0, 1, -1};
        const int n8[4] = {1 + w1, -1 + w1, -1 - w1, 1 - w1}, dx8[4] = {1, -1, -1, 1}, dy8[4] = {1, 1, -1, -1};
        for (int i = 0; i < bfsNum2; i++) {
            int x = bfsList2[2 * i], y = bfsList2[2 * i + 1];
            if (x == 0 || y == 0 || x == w1 - 1 || y == h1 - 1) continue;
            int idx = x + y * w1;
            for (int q = 0; q < 4; q++)
                if (D[idx + n4[q]] > k) {
                    D[idx + n4[q]] = k;
                    bfsList1[2 * bfsNum] = x + dx4[q]; bfsList1[2 * bfsNum + 1] = y + dy4[q];
                    bfsNum++;
                }
            if (k % 2 != 0)      // odd steps also grow diagonally
                for (int q = 0; q < 4; q++)
                    if (D[idx + n8[q]] > k) {
                        D[idx + n8[q]] = k;
                        bfsList1[2 * bfsNum] = x + dx8[q]; bfsList1[2 * bfsNum + 1] = y + dy8[q];
                        bfsNum++;
                    }
        }
    }
}

void CoarseDistanceMap::addIntoDistFinal(int u, int v) {   // :814-819
    if (w[0] == 0) return;
    bfsList1[0] = u; bfsList1[1] = v;
    fwdWarpedIDDistFinal[u + w[1] * v] = 0;
    growDistBFS(1);
}

// FullSystem::activatePointsMT — FullSystem.cc:1097-1150 (the body of the loop over one host keyframe's immature features)
void selectActivation(CoarseDistanceMap &M, const float R[9], const float t[3], bool hostFlaggedForMarginalization, float currentMinActDist,
                      float minTraceQuality, int n, const ActivationCand *c, unsigned char *action) {
    float KRKi[9], Kt[3];
    M.hostToNewest(R, t, KRKi, Kt);
    const int w1 = M.w[1], h1 = M.h[1];
    for (int i = 0; i < n; i++) {
        const ActivationCand &ph = c[i];
        if (!std::isfinite(ph.idepth_max) || ph.lastTraceStatus == 2 /*IPS_OUTLIER*/) { action[i] = 2; continue; }          // :1103-1107
        bool canActivate = (ph.lastTraceStatus == 0 /*GOOD*/ || ph.lastTraceStatus == 3 /*SKIPPED*/ || ph.lastTraceStatus == 4 /*BADCONDITION*/ ||
                            ph.lastTraceStatus == 1 /*OOB*/)
                           && ph.lastTracePixelInterval < 8 && ph.quality > minTraceQuality && (ph.idepth_max + ph.idepth_min) > 0;   // :1109-1115
        if (!canActivate) {                                                                                                 // :1117-1125
            action[i] = (hostFlaggedForMarginalization || ph.lastTraceStatus == 1) ? 2 : 0;
            continue;
        }
        float ptp[3];                                                                                                       // :1128-1131
        const float idm = 0.5f * (ph.idepth_max + ph.idepth_min);
        for (int r = 0; r < 3; r++) {
            float s = KRKi[r * 3 + 0] * ph.u;
            s += KRKi[r * 3 + 1] * ph.v;
            s += KRKi[r * 3 + 2] * 1.0f;
            ptp[r] = s + Kt[r] * idm;
        }
        int u = ptp[0] / ptp[2] + 0.5f;
        int v = ptp[1] / ptp[2] + 0.5f;
        if (u > 0 && v > 0 && u < w1 && v < h1) {                                                                           // :1133-1143
            float dist = M.fwdWarpedIDDistFinal[u + w1 * v] + (ptp[0] - floorf((float) (ptp[0])));
            if (dist >= currentMinActDist * ph.my_type) {
                M.addIntoDistFinal(u, v);
                action[i] = 1;
            } else action[i] = 0;
        } else action[i] = 2;                                                                                               // :1144-1148
    }
}

}  // namespace oracle
