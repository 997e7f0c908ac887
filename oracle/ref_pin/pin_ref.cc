// oracle/ref_pin — pins the oracle restatement against the REFERENCE'S OWN CODE for the pieces of the hot path that compile
// from the reference's sources where they lie (TEST INFRASTRUCTURE ONLY; built by `make -C oracle ref_pin` into oracle/_ref/,
// only when /root/reference is present):
//   include/internal/OptimizationBackend/MatrixAccumulators.h   AccumulatorApprox, Accumulator9, Accumulator11, AccumulatorXX, AccumulatorX
//   include/internal/GlobalFuncs.h                              getInterpolatedElement33 / 31 / 33BiLin
//   include/AffLight.h                                          AffLight::fromToVecExposure
//   include/internal/ResidualProjections.h                      projectPoint (both overloads), derive_idepth
//   src/Setting.cc (+ include/Settings.h)                       every setting_* constant and the residual pattern the path reads
// compiled UNMODIFIED against oracle/ref_shim/NumTypes.h (a stand-in for the Eigen types those headers use; Eigen3, Sophus, glog,
// DBoW3 are not in this image). Every comparison is bit-exact (memcmp). Exit code 0 and "PIN OK" on success.
#include <cstdio>
#include <cstdlib>
#include <xmmintrin.h>
// the oracle side first: the reference's Settings.h defines `patternP` as a macro
#include "../accumulators.h"
#include "../ba.h"
#include "../trace.h"
static inline int oracle_pattern(int i, int k) { return oracle::patternP[i][k]; }
// the reference's own sources, unmodified, from /root/reference (their `#include "NumTypes.h"` is satisfied by the stand-in)
#include "../ref_shim/NumTypes.h"
#include "Settings.h"
#include "AffLight.h"
#include "internal/GlobalFuncs.h"
#include "internal/OptimizationBackend/MatrixAccumulators.h"
// ResidualProjections.h expects CalibHessian to be complete (the reference's includers pull in internal/CalibHessian.h, which drags
// in the camera/frame classes): the accessors it calls are all it needs. The globals of GlobalCalib.h are defined below.
namespace ldso { namespace internal {
struct CalibHessian {
    float fx, fy, cx, cy, fxi, fyi;
    float fxl() const { return fx; } float fyl() const { return fy; } float cxl() const { return cx; } float cyl() const { return cy; }
    float fxli() const { return fxi; } float fyli() const { return fyi; }
};
} }
#include "internal/ResidualProjections.h"
namespace ldso { namespace internal { float wM3G, hM3G; } }

static unsigned long long rng_state = 88172645463325252ull;
static inline float frand(float lo, float hi) {
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return lo + (hi - lo) * (float) ((rng_state >> 11) * (1.0 / 9007199254740992.0));
}
static int fails = 0, checks = 0;
#define CHECK(cond, what) do { checks++; if (!(cond)) { fails++; printf("PIN MISMATCH: %s\n", what); } } while (0)

static void pin_accumulators() {
    using namespace ldso::internal;
    // AccumulatorApprox: > 1e6 updates would take long; 2500 updates cross the 1000-entry tier twice
    {
        AccumulatorApprox R; oracle::AccumulatorApprox O;
        R.initialize(); O.initialize();
        for (int k = 0; k < 2500; k++) {
            float x4[4], x6[6], y4[4], y6[6];
            for (int i = 0; i < 4; i++) { x4[i] = frand(-3, 3); y4[i] = frand(-3, 3); }
            for (int i = 0; i < 6; i++) { x6[i] = frand(-300, 300); y6[i] = frand(-300, 300); }
            const float a = frand(0, 50), b = frand(-20, 20), c = frand(0, 50);
#ifdef PIN_SELFTEST_BREAK      // negative control: a one-ulp-scale change on the oracle side must be detected
            R.update(x4, x6, y4, y6, a, b, c); O.update(x4, x6, y4, y6, a * 1.0000002f, b, c);
#else
            R.update(x4, x6, y4, y6, a, b, c); O.update(x4, x6, y4, y6, a, b, c);
#endif
            float t[6]; for (int i = 0; i < 6; i++) t[i] = frand(-40, 40);
            R.updateTopRight(x4, x6, y4, y6, t[0], t[1], t[2], t[3], t[4], t[5]); O.updateTopRight(x4, x6, y4, y6, t[0], t[1], t[2], t[3], t[4], t[5]);
            R.updateBotRight(t[0] * t[0], t[0] * t[1], t[1] * t[2], t[3] * t[3], t[4], t[5] * t[5]);
            O.updateBotRight(t[0] * t[0], t[0] * t[1], t[1] * t[2], t[3] * t[3], t[4], t[5] * t[5]);
        }
        R.finish(); O.finish();
        bool same = true;
        for (int r = 0; r < 13; r++) for (int c = 0; c < 13; c++) same &= memcmp(&R.H(r, c), &O.H[r * 13 + c], 4) == 0;
        CHECK(same, "AccumulatorApprox H after update/updateTopRight/updateBotRight/finish");
        CHECK(R.num == O.num, "AccumulatorApprox num");
    }
    {   // Accumulator9::updateSSE_eighted (CoarseTracker::calcGSSSE)
        Accumulator9 R; oracle::Accumulator9 O;
        R.initialize(); O.initialize();
        for (int k = 0; k < 2600; k++) {
            alignas(16) float J[9][4], w[4];
            for (int i = 0; i < 9; i++) for (int l = 0; l < 4; l++) J[i][l] = frand(-50, 50);
            for (int l = 0; l < 4; l++) w[l] = frand(0, 1);
            R.updateSSE_eighted(_mm_load_ps(J[0]), _mm_load_ps(J[1]), _mm_load_ps(J[2]), _mm_load_ps(J[3]), _mm_load_ps(J[4]), _mm_load_ps(J[5]),
                                _mm_load_ps(J[6]), _mm_load_ps(J[7]), _mm_load_ps(J[8]), _mm_load_ps(w));
            O.updateSSE_eighted(J, w);
        }
        R.finish(); O.finish();
        bool same = true;
        for (int r = 0; r < 9; r++) for (int c = 0; c < 9; c++) same &= memcmp(&R.H(r, c), &O.H[r * 9 + c], 4) == 0;
        CHECK(same, "Accumulator9 H after updateSSE_eighted/finish");
    }
    {   // Accumulator11 (energy sums)
        Accumulator11 R; oracle::Accumulator11 O;
        R.initialize(); O.initialize();
        for (int k = 0; k < 3000; k++) {
            const float v = frand(0, 400);
            R.updateSingle(v); O.updateSingle(v);
            alignas(16) float q[4] = {frand(0, 9), frand(0, 9), frand(0, 9), frand(0, 9)};
            R.updateSSENoShift(_mm_load_ps(q)); O.updateSSENoShift(q);
        }
        R.finish(); O.finish();
        CHECK(memcmp(&R.A, &O.A, 4) == 0, "Accumulator11 A");
    }
    {   // AccumulatorXX<8,4>, <8,8>, AccumulatorX<8> (Schur complement accumulators)
        AccumulatorXX<8, CPARS> R84; oracle::AccumulatorXX<8, 4> O84;
        AccumulatorXX<8, 8> R88; oracle::AccumulatorXX<8, 8> O88;
        AccumulatorX<8> R8; oracle::AccumulatorX<8> O8;
        R84.initialize(); O84.initialize(); R88.initialize(); O88.initialize(); R8.initialize(); O8.initialize();
        for (int k = 0; k < 2300; k++) {
            Eigen::Matrix<float, 8, 1> L, L2; Eigen::Matrix<float, CPARS, 1> Rc;
            for (int i = 0; i < 8; i++) { L[i] = frand(-100, 100); L2[i] = frand(-100, 100); }
            for (int i = 0; i < 4; i++) Rc[i] = frand(-10, 10);
            const float w = frand(0, 2);
            R84.update(L, Rc, w); O84.update(L.d, Rc.d, w);
            R88.update(L, L2, w); O88.update(L.d, L2.d, w);
            R8.update(L, w); O8.update(L.d, w);
        }
        R84.finish(); O84.finish(); R88.finish(); O88.finish(); R8.finish(); O8.finish();
        CHECK(memcmp(R84.A1m.d, O84.A1m, sizeof(O84.A1m)) == 0, "AccumulatorXX<8,4> A1m");
        CHECK(memcmp(R88.A1m.d, O88.A1m, sizeof(O88.A1m)) == 0, "AccumulatorXX<8,8> A1m");
        CHECK(memcmp(R8.A1m.d, O8.A1m, sizeof(O8.A1m)) == 0, "AccumulatorX<8> A1m");
    }
}

static void pin_samplers() {
    using namespace ldso::internal;
    const int w = 64, h = 48;
    std::vector<Eigen::Vector3f> img(w * h);
    std::vector<float> flat(3 * w * h);
    for (int i = 0; i < w * h; i++) for (int k = 0; k < 3; k++) { const float v = frand(-255, 255); img[i][k] = v; flat[3 * i + k] = v; }
    bool s33 = true, s31 = true, sbl = true;
    for (int k = 0; k < 20000; k++) {
        const float x = frand(1.0f, w - 2.5f), y = frand(1.0f, h - 2.5f);
        const Eigen::Vector3f a = getInterpolatedElement33(img.data(), x, y, w);
        float b[3]; oracle::getInterpolatedElement33(flat.data(), x, y, w, b);
        s33 &= memcmp(a.d, b, 12) == 0;
        const float c = getInterpolatedElement31(img.data(), x, y, w);
        oracle::ImmaturePt dummy; (void) dummy;
        // oracle::trace.cc's 31-sampler is file-local; the 33 sampler's first component is the same expression (GlobalFuncs.h:145-159)
        s31 &= memcmp(&c, &b[0], 4) == 0;
        const Eigen::Vector3f e = getInterpolatedElement33BiLin(img.data(), x, y, w);
        float f[3]; oracle::getInterpolatedElement33BiLin(flat.data(), x, y, w, f);
        sbl &= memcmp(e.d, f, 12) == 0;
    }
    CHECK(s33, "getInterpolatedElement33"); CHECK(s31, "getInterpolatedElement31 == first component of 33"); CHECK(sbl, "getInterpolatedElement33BiLin");
}

static void pin_afflight() {
    bool same = true;
    for (int k = 0; k < 5000; k++) {
        const float eF = (k % 7 == 0) ? 0.f : frand(0.001f, 0.05f), eT = (k % 11 == 0) ? 0.f : frand(0.001f, 0.05f);
        const float aF = frand(-0.3f, 0.3f), bF = frand(-20, 20), aT = frand(-0.3f, 0.3f), bT = frand(-20, 20);
        const Vec2 r = ldso::AffLight::fromToVecExposure(eF, eT, ldso::AffLight(aF, bF), ldso::AffLight(aT, bT));
        double o[2]; oracle::fromToVecExposure(eF, eT, aF, bF, aT, bT, o);
        same &= memcmp(r.d, o, 16) == 0;
    }
    CHECK(same, "AffLight::fromToVecExposure");
}

static void pin_projections() {
    using namespace ldso::internal;
    oracle::Calib OC;
    OC.value_scaledf[0] = 400.25f; OC.value_scaledf[1] = 401.5f; OC.value_scaledf[2] = 319.5f; OC.value_scaledf[3] = 239.5f;
    OC.value_scaledi[0] = 1.0f / OC.value_scaledf[0]; OC.value_scaledi[1] = 1.0f / OC.value_scaledf[1];
    auto HC = std::make_shared<CalibHessian>();
    HC->fx = OC.fxl(); HC->fy = OC.fyl(); HC->cx = OC.cxl(); HC->cy = OC.cyl(); HC->fxi = OC.fxli(); HC->fyi = OC.fyli();
    wM3G = 640 - 3; hM3G = 480 - 3;
    bool okA = true, okB = true, okD = true;
    for (int k = 0; k < 20000; k++) {
        Mat33f KRKi, R; Vec3f Kt, t;
        float kr[9], rr[9], kt[3], tt[3];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
            kr[i * 3 + j] = (i == j ? 1.f : 0.f) + frand(-0.05f, 0.05f); rr[i * 3 + j] = (i == j ? 1.f : 0.f) + frand(-0.05f, 0.05f);
            KRKi(i, j) = kr[i * 3 + j]; R(i, j) = rr[i * 3 + j];
        }
        for (int i = 0; i < 3; i++) { kt[i] = frand(-30, 30); tt[i] = frand(-0.2f, 0.2f); Kt[i] = kt[i]; t[i] = tt[i]; }
        kr[2] = frand(-20, 20); kr[5] = frand(-20, 20); KRKi(0, 2) = kr[2]; KRKi(1, 2) = kr[5];
        const float u = frand(-50, 700), v = frand(-50, 530), id = frand(-0.2f, 2.f);
        float Ku, Kv, oKu, oKv;
        const bool a = projectPoint(u, v, id, KRKi, Kt, Ku, Kv);
        const bool b = oracle::projectPointA(u, v, id, kr, kt, wM3G, hM3G, oKu, oKv);
        okA &= a == b && memcmp(&Ku, &oKu, 4) == 0 && memcmp(&Kv, &oKv, 4) == 0;
        const int dx = (int) frand(-2.99f, 2.99f), dy = (int) frand(-2.99f, 2.99f);
        float drescale, uu, vv, Ku2, Kv2, nid, odr, ouu, ovv, oKu2, oKv2, onid, oKliP[3];
        Vec3f KliP;
        shared_ptr<CalibHessian> HCc = HC;
        const bool c = projectPoint(u, v, id, dx, dy, HCc, R, t, drescale, uu, vv, Ku2, Kv2, KliP, nid);
        const bool d = oracle::projectPointB(u, v, id, dx, dy, OC, rr, tt, wM3G, hM3G, odr, ouu, ovv, oKu2, oKv2, oKliP, onid);
        okB &= c == d && memcmp(&drescale, &odr, 4) == 0 && memcmp(&nid, &onid, 4) == 0 && memcmp(KliP.d, oKliP, 12) == 0;
        if (c && d) okB &= memcmp(&uu, &ouu, 4) == 0 && memcmp(&vv, &ovv, 4) == 0 && memcmp(&Ku2, &oKu2, 4) == 0 && memcmp(&Kv2, &oKv2, 4) == 0;
        const float dxi = frand(-50, 50), dyi = frand(-50, 50), dr = frand(0.5f, 2.f);
        const float e = derive_idepth(t, uu, vv, dx, dy, dxi, dyi, dr), f = oracle::derive_idepth(tt, uu, vv, dx, dy, dxi, dyi, dr);
        okD &= memcmp(&e, &f, 4) == 0 || (e != e && f != f);
    }
    CHECK(okA, "projectPoint (pattern, KRKi/Kt form)"); CHECK(okB, "projectPoint (centre, R/t/K^-1 form)"); CHECK(okD, "derive_idepth");
}

static void pin_settings() {
    using namespace ldso;
    oracle::Settings S; oracle::TraceSettings T;
    CHECK(setting_huberTH == S.huberTH && setting_huberTH == T.huberTH, "setting_huberTH");
    CHECK(setting_outlierTHSumComponent == S.outlierTHSumComponent && setting_outlierTHSumComponent == T.outlierTHSumComponent, "setting_outlierTHSumComponent");
    CHECK(setting_outlierTH == T.outlierTH, "setting_outlierTH");
    CHECK(setting_overallEnergyTHWeight == S.overallEnergyTHWeight && setting_overallEnergyTHWeight == T.overallEnergyTHWeight, "setting_overallEnergyTHWeight");
    CHECK(setting_initialRotPrior == S.initialRotPrior && setting_initialTransPrior == S.initialTransPrior, "initial pose priors");
    CHECK(setting_initialAffAPrior == S.initialAffAPrior && setting_initialAffBPrior == S.initialAffBPrior, "initial affine priors");
    CHECK(setting_initialCalibHessian == S.initialCalibHessian, "setting_initialCalibHessian");
    CHECK(setting_solverModeDelta == S.solverModeDelta, "setting_solverModeDelta");
    CHECK(setting_idepthFixPrior == S.idepthFixPrior, "setting_idepthFixPrior");
    CHECK(setting_affineOptModeA == S.affineOptModeA && setting_affineOptModeB == S.affineOptModeB, "setting_affineOptModeA/B");
    CHECK(setting_frameEnergyTHConstWeight == S.frameEnergyTHConstWeight && setting_frameEnergyTHN == S.frameEnergyTHN &&
          setting_frameEnergyTHFacMedian == S.frameEnergyTHFacMedian, "frameEnergyTH settings");
    CHECK(setting_coarseCutoffTH == S.coarseCutoffTH, "setting_coarseCutoffTH");
    CHECK(setting_thOptIterations == S.thOptIterations, "setting_thOptIterations");
    CHECK(setting_solverMode == (SOLVER_FIX_LAMBDA | SOLVER_ORTHOGONALIZE_X_LATER), "setting_solverMode = FIX_LAMBDA | ORTHOGONALIZE_X_LATER");
    CHECK(setting_forceAceptStep == true, "setting_forceAceptStep");
    CHECK(setting_maxPixSearch == T.maxPixSearch && setting_minTraceTestRadius == T.minTraceTestRadius, "trace search settings");
    CHECK(setting_trace_stepsize == T.trace_stepsize && setting_trace_GNIterations == T.trace_GNIterations && setting_trace_GNThreshold == T.trace_GNThreshold &&
          setting_trace_extraSlackOnTH == T.trace_extraSlackOnTH && setting_trace_slackInterval == T.trace_slackInterval &&
          setting_trace_minImprovementFactor == T.trace_minImprovementFactor, "setting_trace_*");
    CHECK(setting_minIdepthH_act == 100 && setting_GNItsOnPointActivation == 3, "point-activation settings");
    CHECK(SCALE_IDEPTH == oracle::SCALE_IDEPTH && SCALE_XI_ROT == oracle::SCALE_XI_ROT && SCALE_XI_TRANS == oracle::SCALE_XI_TRANS && SCALE_F == oracle::SCALE_F &&
          SCALE_C == oracle::SCALE_C && SCALE_A == oracle::SCALE_A && SCALE_B == oracle::SCALE_B, "SCALE_* constants");
    CHECK(patternNum == oracle::patternNum, "patternNum");
    bool pat = true;
    for (int i = 0; i < 8; i++) pat &= patternP[i][0] == oracle_pattern(i, 0) && patternP[i][1] == oracle_pattern(i, 1);
    CHECK(pat, "patternP (residual pattern 8)");
    CHECK(NUM_THREADS == 6, "NUM_THREADS");
}

int main() {
    pin_accumulators();
    pin_samplers();
    pin_afflight();
    pin_projections();
    pin_settings();
    if (fails) { printf("PIN FAILED: %d of %d checks\n", fails, checks); return 1; }
    printf("PIN OK: %d checks against the reference's own MatrixAccumulators.h, GlobalFuncs.h, ResidualProjections.h, AffLight.h and Setting.cc\n", checks);
    return 0;
}
