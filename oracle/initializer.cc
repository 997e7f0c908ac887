// ORACLE — TEST INFRASTRUCTURE ONLY (see initializer.h). CPU restatement of CoarseInitializer::makeK and calcResAndGS
// (src/frontend/CoarseInitializer.cc:689-715, :181-405), keeping the reference's float arithmetic and evaluation order.
#include "initializer.h"

namespace oracle {

CoarseInitializer::CoarseInitializer(int ww, int hh, int levels) {
    pyrLevelsUsed = levels;
    for (int l = 0; l < PYR_LEVELS; l++) { w[l] = h[l] = 0; firstDIp[l] = newDIp[l] = nullptr; }
    w[0] = ww; h[0] = hh;
    JbBuffer_new.assign((size_t) ww * hh, std::array<float, 10>{});
}

static void m33d_inverse(const double *m, double *inv) {     // Eigen's 3x3 inverse (cofactors * 1/det), double
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    const double invdet = 1.0 / det;
    inv[0] = c00 * invdet; inv[3] = c01 * invdet; inv[6] = c02 * invdet;
    inv[1] = (m[2] * m[7] - m[1] * m[8]) * invdet; inv[4] = (m[0] * m[8] - m[2] * m[6]) * invdet; inv[7] = (m[1] * m[6] - m[0] * m[7]) * invdet;
    inv[2] = (m[1] * m[5] - m[2] * m[4]) * invdet; inv[5] = (m[2] * m[3] - m[0] * m[5]) * invdet; inv[8] = (m[0] * m[4] - m[1] * m[3]) * invdet;
}

void CoarseInitializer::makeK(float fxl, float fyl, float cxl, float cyl) {   // :689-715 (fx.. are doubles here)
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
        double *Kl = K[level];
        Kl[0] = fx[level]; Kl[1] = 0; Kl[2] = cx[level];
        Kl[3] = 0; Kl[4] = fy[level]; Kl[5] = cy[level];
        Kl[6] = 0; Kl[7] = 0; Kl[8] = 1;
        m33d_inverse(Kl, Ki[level]);
    }
}

static inline float interp31(const float *mat, float x, float y, int width) {     // getInterpolatedElement31, GlobalFuncs.h:145-159
    int ix = (int) x, iy = (int) y;
    float dx = x - ix, dy = y - iy, dxdy = dx * dy;
    const float *bp = mat + 3 * (ix + iy * width);
    return dxdy * bp[3 * (1 + width)] + (dy - dxdy) * bp[3 * width] + (dx - dxdy) * bp[3] + (1 - dx - dy + dxdy) * bp[0];
}

void CoarseInitializer::calcResAndGS(int lvl, float H_out[64], float b_out[8], float H_out_sc[64], float b_out_sc[8], const SE3 &refToNew, float aff_a,
                                     float aff_b, float res3[3]) {
    int wl = w[lvl], hl = h[lvl];
    const float *colorRef = firstDIp[lvl], *colorNew = newDIp[lvl];
    // RKi = (R * Ki).cast<float>(): the product in double, then rounded (:190)
    M3 Rd = refToNew.rotationMatrix();
    float RKi[9], t[3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = Rd(i, 0) * Ki[lvl][0 * 3 + j];
            s += Rd(i, 1) * Ki[lvl][1 * 3 + j];
            s += Rd(i, 2) * Ki[lvl][2 * 3 + j];
            RKi[i * 3 + j] = (float) s;
        }
    for (int i = 0; i < 3; i++) t[i] = (float) refToNew.t[i];
    const float r2new_aff[2] = {std::exp(aff_a), aff_b};
    float fxl = fx[lvl], fyl = fy[lvl], cxl = cx[lvl], cyl = cy[lvl];

    Accumulator11 E;
    acc9.initialize();
    E.initialize();
    int npts = (int) points[lvl].size();
    InitPnt *ptsl = points[lvl].data();
    for (int i = 0; i < npts; i++) {
        InitPnt *point = ptsl + i;
        point->maxstep = 1e10;
        if (!point->isGood) {
            E.updateSingle((float) (point->energy[0]));
            point->energy_new[0] = point->energy[0]; point->energy_new[1] = point->energy[1];
            point->isGood_new = false;
            continue;
        }
        alignas(16) float dp[9][8];        // dp0..dp7 and r: [k][idx]
        float dd[8];
        std::array<float, 10> &Jb = JbBuffer_new[i];
        for (int k = 0; k < 10; k++) Jb[k] = 0;
        bool isGood = true;
        float energy = 0;
        for (int idx = 0; idx < patternNum; idx++) {
            int dx = patternP[idx][0], dy = patternP[idx][1];
            float pt[3];
            const float px = point->u + dx, py = point->v + dy;
            for (int r = 0; r < 3; r++) {
                float s = RKi[r * 3 + 0] * px;
                s += RKi[r * 3 + 1] * py;
                s += RKi[r * 3 + 2] * 1.0f;
                pt[r] = s + t[r] * point->idepth_new;
            }
            float u = pt[0] / pt[2], v = pt[1] / pt[2];
            float Ku = fxl * u + cxl, Kv = fyl * v + cyl;
            float new_idepth = point->idepth_new / pt[2];
            if (!(Ku > 1 && Kv > 1 && Ku < wl - 2 && Kv < hl - 2 && new_idepth > 0)) { isGood = false; break; }
            float hitColor[3];
            getInterpolatedElement33(colorNew, Ku, Kv, wl, hitColor);
            float rlR = interp31(colorRef, point->u + dx, point->v + dy, wl);
            if (!std::isfinite(rlR) || !std::isfinite((float) hitColor[0])) { isGood = false; break; }
            float residual = hitColor[0] - r2new_aff[0] * rlR - r2new_aff[1];
            float hw = fabs(residual) < S.huberTH ? 1 : S.huberTH / fabs(residual);
            energy += hw * residual * residual * (2 - hw);
            float dxdd = (t[0] - t[2] * u) / pt[2];
            float dydd = (t[1] - t[2] * v) / pt[2];
            if (hw < 1) hw = sqrtf(hw);
            float dxInterp = hw * hitColor[1] * fxl;
            float dyInterp = hw * hitColor[2] * fyl;
            dp[0][idx] = new_idepth * dxInterp;
            dp[1][idx] = new_idepth * dyInterp;
            dp[2][idx] = -new_idepth * (u * dxInterp + v * dyInterp);
            dp[3][idx] = -u * v * dxInterp - (1 + v * v) * dyInterp;
            dp[4][idx] = (1 + u * u) * dxInterp + u * v * dyInterp;
            dp[5][idx] = -v * dxInterp + u * dyInterp;
            dp[6][idx] = -hw * r2new_aff[0] * rlR;
            dp[7][idx] = -hw * 1;
            dd[idx] = dxInterp * dxdd + dyInterp * dydd;
            dp[8][idx] = hw * residual;
            {   // 1 / Vec2f(dxdd*fxl, dydd*fyl).norm()
                const float a = dxdd * fxl, b = dydd * fyl;
                float maxstep = 1.0f / std::sqrt(a * a + b * b);
                if (maxstep < point->maxstep) point->maxstep = maxstep;
            }
            for (int k = 0; k < 8; k++) Jb[k] += dp[k][idx] * dd[idx];
            Jb[8] += dp[8][idx] * dd[idx];
            Jb[9] += dd[idx] * dd[idx];
        }
        if (!isGood || energy > point->outlierTH * 20) {
            E.updateSingle((float) (point->energy[0]));
            point->isGood_new = false;
            point->energy_new[0] = point->energy[0]; point->energy_new[1] = point->energy[1];
            continue;
        }
        E.updateSingle(energy);
        point->isGood_new = true;
        point->energy_new[0] = energy;
        for (int i4 = 0; i4 + 3 < patternNum; i4 += 4) {
            float J[9][4];
            for (int k = 0; k < 9; k++) for (int l = 0; l < 4; l++) J[k][l] = dp[k][i4 + l];
            acc9.updateSSE(J);
        }
        for (int i1 = ((patternNum >> 2) << 2); i1 < patternNum; i1++) {
            float J[9];
            for (int k = 0; k < 9; k++) J[k] = dp[k][i1];
            acc9.updateSingle(J);
        }
    }
    E.finish();
    acc9.finish();

    // alpha energy (:336-356). The reference adds the regulariser terms to E (already finished), not to EAlpha: kept.
    Accumulator11 EAlpha;
    EAlpha.initialize();
    for (int i = 0; i < npts; i++) {
        InitPnt *point = ptsl + i;
        if (!point->isGood_new) E.updateSingle((float) (point->energy[1]));
        else {
            point->energy_new[1] = (point->idepth_new - 1) * (point->idepth_new - 1);
            E.updateSingle((float) (point->energy_new[1]));
        }
    }
    EAlpha.finish();
    const double tsq = refToNew.t[0] * refToNew.t[0] + refToNew.t[1] * refToNew.t[1] + refToNew.t[2] * refToNew.t[2];
    float alphaEnergy = alphaW * (EAlpha.A + tsq * npts);
    float alphaOpt;
    if (alphaEnergy > alphaK * npts) { alphaOpt = 0; alphaEnergy = alphaK * npts; }
    else alphaOpt = alphaW;

    acc9SC.initialize();
    for (int i = 0; i < npts; i++) {
        InitPnt *point = ptsl + i;
        if (!point->isGood_new) continue;
        std::array<float, 10> &Jb = JbBuffer_new[i];
        point->lastHessian_new = Jb[9];
        Jb[8] += alphaOpt * (point->idepth_new - 1);
        Jb[9] += alphaOpt;
        if (alphaOpt == 0) {
            Jb[8] += couplingWeight * (point->idepth_new - point->iR);
            Jb[9] += couplingWeight;
        }
        Jb[9] = 1 / (1 + Jb[9]);
        acc9SC.updateSingleWeighted(Jb.data(), Jb[9]);
    }
    acc9SC.finish();
    for (int r = 0; r < 8; r++) {
        for (int c = 0; c < 8; c++) { H_out[r * 8 + c] = acc9.H[r * 9 + c]; H_out_sc[r * 8 + c] = acc9SC.H[r * 9 + c]; }
        b_out[r] = acc9.H[r * 9 + 8]; b_out_sc[r] = acc9SC.H[r * 9 + 8];
    }
    H_out[0 * 8 + 0] += alphaOpt * npts;
    H_out[1 * 8 + 1] += alphaOpt * npts;
    H_out[2 * 8 + 2] += alphaOpt * npts;
    double lg[6];
    refToNew.log(lg);
    const float tlog[3] = {(float) lg[0], (float) lg[1], (float) lg[2]};
    b_out[0] += tlog[0] * alphaOpt * npts;
    b_out[1] += tlog[1] * alphaOpt * npts;
    b_out[2] += tlog[2] * alphaOpt * npts;
    res3[0] = E.A; res3[1] = alphaEnergy; res3[2] = (float) E.num;
}

}  // namespace oracle
