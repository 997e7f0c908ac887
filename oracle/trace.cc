// ORACLE — TEST INFRASTRUCTURE ONLY. See trace.h for the reference lines each function follows.
#include "trace.h"
#include "ba.h"

namespace oracle {

static const int kPattern[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};   // Setting.cc:221 (pattern 8)

// GlobalFuncs.h:145-159
static inline float getInterpolatedElement31(const float *mat, float x, float y, int width) {
    int ix = (int) x, iy = (int) y;
    float dx = x - ix, dy = y - iy, dxdy = dx * dy;
    const float *bp = mat + 3 * (ix + iy * width);
    return dxdy * bp[3 * (1 + width)] + (dy - dxdy) * bp[3 * width] + (dx - dxdy) * bp[3] + (1 - dx - dy + dxdy) * bp[0];
}

// ImmaturePoint.cc:14-38
void immature_init(ImmaturePt &p, const float *dI_host, int w, float u, float v, const TraceSettings &S) {
    p = ImmaturePt();
    p.u = u; p.v = v;
    for (int idx = 0; idx < 8; idx++) {
        int dx = kPattern[idx][0], dy = kPattern[idx][1];
        float ptc[3];
        getInterpolatedElement33BiLin(dI_host, u + dx, v + dy, w, ptc);
        p.color[idx] = ptc[0];
        if (!std::isfinite(p.color[idx])) { p.energyTH = NAN; return; }
        p.gradH[0] += ptc[1] * ptc[1]; p.gradH[1] += ptc[1] * ptc[2];
        p.gradH[2] += ptc[2] * ptc[1]; p.gradH[3] += ptc[2] * ptc[2];
        p.weights[idx] = sqrtf(S.outlierTHSumComponent / (S.outlierTHSumComponent + (ptc[1] * ptc[1] + ptc[2] * ptc[2])));
    }
    p.energyTH = 8 * S.outlierTH;
    p.energyTH *= S.overallEnergyTHWeight * S.overallEnergyTHWeight;
}

// ImmaturePoint.cc:46-314
int trace_on(ImmaturePt &p, const float *dI, int w, int h, const float KRKi[9], const float Kt[3], const float aff[2], const TraceSettings &S) {
    if (p.lastTraceStatus == IPS_OOB) return p.lastTraceStatus;
    float maxPixSearch = (w + h) * S.maxPixSearch;
    auto oob = [&]() { p.lastTraceUV[0] = p.lastTraceUV[1] = -1; p.lastTracePixelInterval = 0; return p.lastTraceStatus = IPS_OOB; };

    // project min and max (:57-68)
    float pr[3];
    for (int i = 0; i < 3; i++) pr[i] = KRKi[i * 3 + 0] * p.u + KRKi[i * 3 + 1] * p.v + KRKi[i * 3 + 2] * 1.0f;
    float ptpMin[3];
    for (int i = 0; i < 3; i++) ptpMin[i] = pr[i] + Kt[i] * p.idepth_min;
    float uMin = ptpMin[0] / ptpMin[2], vMin = ptpMin[1] / ptpMin[2];
    if (!(uMin > 4 && vMin > 4 && uMin < w - 5 && vMin < h - 5)) return oob();

    float dist, uMax, vMax, ptpMax[3];
    if (std::isfinite(p.idepth_max)) {                 // :77-98
        for (int i = 0; i < 3; i++) ptpMax[i] = pr[i] + Kt[i] * p.idepth_max;
        uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
        if (!(uMax > 4 && vMax > 4 && uMax < w - 5 && vMax < h - 5)) return oob();
        dist = (uMin - uMax) * (uMin - uMax) + (vMin - vMax) * (vMin - vMax);
        dist = sqrtf(dist);
        if (dist < S.trace_slackInterval) {
            p.lastTraceUV[0] = (uMax + uMin) * 0.5f; p.lastTraceUV[1] = (vMax + vMin) * 0.5f;
            p.lastTracePixelInterval = dist;
            return p.lastTraceStatus = IPS_SKIPPED;
        }
    } else {                                           // :99-124
        dist = maxPixSearch;
        for (int i = 0; i < 3; i++) ptpMax[i] = pr[i] + Kt[i] * 0.01f;
        uMax = ptpMax[0] / ptpMax[2]; vMax = ptpMax[1] / ptpMax[2];
        float dx = uMax - uMin, dy = vMax - vMin;
        float d = 1.0f / sqrtf(dx * dx + dy * dy);
        uMax = uMin + dist * dx * d;
        vMax = vMin + dist * dy * d;
        if (!(uMax > 4 && vMax > 4 && uMax < w - 5 && vMax < h - 5)) return oob();
    }
    // scale change too big (:127-131)
    if (!(p.idepth_min < 0 || (ptpMin[2] > 0.75f && ptpMin[2] < 1.5f))) return oob();

    // error bounds (:134-148)
    float dx = S.trace_stepsize * (uMax - uMin);
    float dy = S.trace_stepsize * (vMax - vMin);
    // (v^T * gradH) * v, left to right like Eigen evaluates `v.transpose() * gradH * v`
    float a = (dx * p.gradH[0] + dy * p.gradH[2]) * dx + (dx * p.gradH[1] + dy * p.gradH[3]) * dy;
    float b = (dy * p.gradH[0] + (-dx) * p.gradH[2]) * dy + (dy * p.gradH[1] + (-dx) * p.gradH[3]) * (-dx);
    float errorInPixel = 0.2f + 0.2f * (a + b) / a;
    if (errorInPixel * S.trace_minImprovementFactor > dist && std::isfinite(p.idepth_max)) {
        p.lastTraceUV[0] = (uMax + uMin) * 0.5f; p.lastTraceUV[1] = (vMax + vMin) * 0.5f;
        p.lastTracePixelInterval = dist;
        return p.lastTraceStatus = IPS_BADCONDITION;
    }
    if (errorInPixel > 10) errorInPixel = 10;

    // discrete search (:151-217)
    dx /= dist;
    dy /= dist;
    if (dist > maxPixSearch) {
        uMax = uMin + maxPixSearch * dx;
        vMax = vMin + maxPixSearch * dy;
        dist = maxPixSearch;
    }
    int numSteps = 1.9999f + dist / S.trace_stepsize;
    const float Rp[4] = {KRKi[0], KRKi[1], KRKi[3], KRKi[4]};
    float randShift = uMin * 1000 - floorf(uMin * 1000);
    float ptx = uMin - randShift * dx;
    float pty = vMin - randShift * dy;
    float rot[8][2];
    for (int idx = 0; idx < 8; idx++) {
        rot[idx][0] = Rp[0] * kPattern[idx][0] + Rp[1] * kPattern[idx][1];
        rot[idx][1] = Rp[2] * kPattern[idx][0] + Rp[3] * kPattern[idx][1];
    }
    if (!std::isfinite(dx) || !std::isfinite(dy)) { p.lastTracePixelInterval = 0; p.lastTraceUV[0] = p.lastTraceUV[1] = -1; return p.lastTraceStatus = IPS_OOB; }

    float errors[100];
    float bestU = 0, bestV = 0, bestEnergy = 1e10;
    int bestIdx = -1;
    if (numSteps >= 100) numSteps = 99;
    for (int i = 0; i < numSteps; i++) {
        float energy = 0;
        for (int idx = 0; idx < 8; idx++) {
            float hitColor = getInterpolatedElement31(dI, (float) (ptx + rot[idx][0]), (float) (pty + rot[idx][1]), w);
            if (!std::isfinite(hitColor)) { energy += 1e5; continue; }
            float residual = hitColor - (float) (aff[0] * p.color[idx] + aff[1]);
            float hw = fabs(residual) < S.huberTH ? 1 : S.huberTH / fabs(residual);
            energy += hw * residual * residual * (2 - hw);
        }
        errors[i] = energy;
        if (energy < bestEnergy) { bestU = ptx; bestV = pty; bestEnergy = energy; bestIdx = i; }
        ptx += dx;
        pty += dy;
    }
    // best score outside a +-2px radius (:220-227)
    float secondBest = 1e10;
    for (int i = 0; i < numSteps; i++)
        if ((i < bestIdx - S.minTraceTestRadius || i > bestIdx + S.minTraceTestRadius) && errors[i] < secondBest) secondBest = errors[i];
    float newQuality = secondBest / bestEnergy;
    if (newQuality < p.quality || numSteps > 10) p.quality = newQuality;

    // GN optimisation (:231-278)
    float uBak = bestU, vBak = bestV, gnstepsize = 1, stepBack = 0;
    if (S.trace_GNIterations > 0) bestEnergy = 1e5;
    for (int it = 0; it < S.trace_GNIterations; it++) {
        float H = 1, bb = 0, energy = 0;
        for (int idx = 0; idx < 8; idx++) {
            float hitColor[3];
            getInterpolatedElement33(dI, (float) (bestU + rot[idx][0]), (float) (bestV + rot[idx][1]), w, hitColor);
            if (!std::isfinite((float) hitColor[0])) { energy += 1e5; continue; }
            float residual = hitColor[0] - (aff[0] * p.color[idx] + aff[1]);
            float dResdDist = dx * hitColor[1] + dy * hitColor[2];
            float hw = fabs(residual) < S.huberTH ? 1 : S.huberTH / fabs(residual);
            H += hw * dResdDist * dResdDist;
            bb += hw * residual * dResdDist;
            energy += p.weights[idx] * p.weights[idx] * hw * residual * residual * (2 - hw);
        }
        if (energy > bestEnergy) {
            stepBack *= 0.5;
            bestU = uBak + stepBack * dx;
            bestV = vBak + stepBack * dy;
        } else {
            float step = -gnstepsize * bb / H;
            if (step < -0.5) step = -0.5;
            else if (step > 0.5) step = 0.5;
            if (!std::isfinite(step)) step = 0;
            uBak = bestU; vBak = bestV; stepBack = step;
            bestU += step * dx;
            bestV += step * dy;
            bestEnergy = energy;
        }
        if (fabsf(stepBack) < S.trace_GNThreshold) break;
    }
    // energy-based outlier (:281-288)
    if (!(bestEnergy < p.energyTH * S.trace_extraSlackOnTH)) {
        p.lastTracePixelInterval = 0;
        p.lastTraceUV[0] = p.lastTraceUV[1] = -1;
        if (p.lastTraceStatus == IPS_OUTLIER) return p.lastTraceStatus = IPS_OOB;
        return p.lastTraceStatus = IPS_OUTLIER;
    }
    // new interval (:291-310)
    if (dx * dx > dy * dy) {
        p.idepth_min = (pr[2] * (bestU - errorInPixel * dx) - pr[0]) / (Kt[0] - Kt[2] * (bestU - errorInPixel * dx));
        p.idepth_max = (pr[2] * (bestU + errorInPixel * dx) - pr[0]) / (Kt[0] - Kt[2] * (bestU + errorInPixel * dx));
    } else {
        p.idepth_min = (pr[2] * (bestV - errorInPixel * dy) - pr[1]) / (Kt[1] - Kt[2] * (bestV - errorInPixel * dy));
        p.idepth_max = (pr[2] * (bestV + errorInPixel * dy) - pr[1]) / (Kt[1] - Kt[2] * (bestV + errorInPixel * dy));
    }
    if (p.idepth_min > p.idepth_max) { float t = p.idepth_min; p.idepth_min = p.idepth_max; p.idepth_max = t; }
    if (!std::isfinite(p.idepth_min) || !std::isfinite(p.idepth_max) || (p.idepth_max < 0)) {
        p.lastTracePixelInterval = 0;
        p.lastTraceUV[0] = p.lastTraceUV[1] = -1;
        return p.lastTraceStatus = IPS_OUTLIER;
    }
    p.lastTracePixelInterval = 2 * errorInPixel;
    p.lastTraceUV[0] = bestU; p.lastTraceUV[1] = bestV;
    return p.lastTraceStatus = IPS_GOOD;
}

}  // namespace oracle
