// oracle/ref_pin — pins the oracle restatement against the REFERENCE'S OWN CODE for the pieces of the hot path that compile
// from the reference's sources where they lie (TEST INFRASTRUCTURE ONLY; built by `make -C oracle ref_pin` into oracle/_ref/,
// only when /root/reference is present):
//   include/internal/OptimizationBackend/MatrixAccumulators.h   AccumulatorApprox, Accumulator9, Accumulator11, AccumulatorXX, AccumulatorX
//   include/internal/GlobalFuncs.h                              getInterpolatedElement33 / 31 / 33BiLin
//   include/AffLight.h                                          AffLight::fromToVecExposure
//   src/internal/Residuals.cc (+ Residuals.h, RawResidualJacobian.h, FrameFramePrecalc.h)   PointFrameResidual::linearize, fixLinearizationF, applyRes/takeData
//   src/internal/ImmaturePoint.cc (+ ImmaturePoint.h, Feature.h)   ImmaturePoint::ImmaturePoint, traceOn, linearizeResidual
//   include/internal/ResidualProjections.h                      projectPoint (both overloads), derive_idepth
//   src/internal/OptimizationBackend/AccumulatedTopHessian.cc, AccumulatedSCHessian.cc (+ their headers)   addPoint<0,1,2>, SC addPoint, stitchDouble, stitchDoubleMT, stitchDoubleInternal
//   src/frontend/CoarseTracker.cc (+ CoarseTracker.h, Feature.h, Point.h)   CoarseTracker::makeK, setCoarseTrackingRef / makeCoarseDepthL0, calcRes, calcGSSSE, trackNewestCoarse
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
#include "../tracker.h"
#include "../initializer.h"
static inline int oracle_pattern(int i, int k) { return oracle::patternP[i][k]; }
// the reference's own sources, unmodified, from /root/reference (their `#include "NumTypes.h"` is satisfied by the stand-in)
#include "../ref_shim/NumTypes.h"
#include "Settings.h"
#include "AffLight.h"
#include "internal/GlobalFuncs.h"
#include "internal/OptimizationBackend/MatrixAccumulators.h"
#include "../ref_shim/ref_classes.h"      // Frame stand-in + the reference's own FrameHessian / PointHessian / CalibHessian headers
#include "internal/OptimizationBackend/EnergyFunctional.h"
#include "internal/Residuals.h"           // the reference's own PointFrameResidual (its linearize lives in src/internal/Residuals.cc)
#include "internal/ImmaturePoint.h"       // the reference's own ImmaturePoint (+ Feature.h); bodies in src/internal/ImmaturePoint.cc
#include "internal/ResidualProjections.h"
#define private public                    // CoarseTracker.cc is compiled with -Dprivate=public too (Makefile): calcRes / calcGSSSE / buffers
#include "frontend/CoarseTracker.h"       // the reference's own CoarseTracker; bodies in src/frontend/CoarseTracker.cc
#include "internal/OptimizationBackend/AccumulatedTopHessian.h"     // the reference's own accumulators; bodies in AccumulatedTopHessian.cc / AccumulatedSCHessian.cc
#include "internal/OptimizationBackend/AccumulatedSCHessian.h"
#include "frontend/CoarseInitializer.h"    // the reference's own initializer; calcResAndGS lives in src/frontend/CoarseInitializer.cc
#undef private
namespace ldso { namespace internal { float wM3G, hM3G; int wG[PYR_LEVELS], hG[PYR_LEVELS]; } }

// src/Camera.cc and src/Point.cc are not compiled (they reach into the front end); these are their plain constructors
ldso::Camera::Camera(double fx_, double fy_, double cx_, double cy_) { fx = fx_; fy = fy_; cx = cx_; cy = cy_; }
ldso::Point::Point() {}
static shared_ptr<ldso::internal::CalibHessian> make_calib(double fx, double fy, double cx, double cy) {
    return std::make_shared<ldso::internal::CalibHessian>(std::make_shared<ldso::Camera>(fx, fy, cx, cy));
}
// FrameHessian's destructor frees the pyramids makeImages() allocated; frames whose pyramids point into harness buffers drop them first
static shared_ptr<ldso::internal::FrameHessian> make_fh(shared_ptr<ldso::Frame> fr, bool ownsPyramid = false) {
    auto *p = new ldso::internal::FrameHessian(fr);
    for (int i = 0; i < PYR_LEVELS; i++) { p->dIp[i] = nullptr; p->absSquaredGrad[i] = nullptr; }
    if (ownsPyramid) return shared_ptr<ldso::internal::FrameHessian>(p);
    return shared_ptr<ldso::internal::FrameHessian>(p, [](ldso::internal::FrameHessian *q) {
        for (int i = 0; i < PYR_LEVELS; i++) { q->dIp[i] = nullptr; q->absSquaredGrad[i] = nullptr; }
        delete q;
    });
}

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
    auto HC = make_calib(400.25, 401.5, 319.5, 239.5);
    if (memcmp(HC->value_scaledf.d, OC.value_scaledf, 16) != 0 || HC->fxli() != OC.fxli() || HC->fyli() != OC.fyli()) { printf("PIN MISMATCH: CalibHessian constructor\n"); fails++; }
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

// ---- PointFrameResidual::linearize / applyRes+takeData / fixLinearizationF: the reference's src/internal/Residuals.cc itself
extern "C" {
void *oracle_ba_create(int w, int h, int threads_mode);
void oracle_ba_destroy(void *o);
void oracle_ba_set_calib(void *o, const double value_scaled[4]);
void oracle_ba_set_calib_delta(void *o, const double delta[4]);
int oracle_ba_add_frame(void *o, const double R[9], const double t[3], const double state_zero[10], const double state[10], float ab_exposure,
                        int frame_id, const float *dI);
int oracle_ba_add_point(void *o, int host, float u, float v, float idepth_zero, float idepth, int hasDepthPrior, const float color[8],
                        const float weights[8]);
int oracle_ba_add_residual(void *o, int point, int target);
void oracle_ba_set_frame_energy_th(void *o, int frame, float th);
void oracle_ba_finalize(void *o);
void oracle_ba_set_marg_prior(void *o, const double *HM_colmajor, const double *bM);
}
// a small window (oracle side) and its reference-side mirror, shared by pin_linearize and pin_hessians
struct Scene {
    int w, h, nF;
    std::vector<std::vector<float>> imgs;
    void *o; oracle::Window *W;
    shared_ptr<ldso::internal::CalibHessian> HC;
    std::vector<shared_ptr<ldso::internal::FrameHessian>> FH;
    shared_ptr<ldso::internal::EnergyFunctional> EF;
};
static Scene *make_scene(int nPper, int threads_mode = 0, bool depthPriors = false) {
    using namespace ldso::internal;
    Scene *S = new Scene();
    const int w = 160, h = 120, nF = 4;
    S->w = w; S->h = h; S->nF = nF;
    // images: smooth texture + central-difference gradients, (I, dx, dy) AoS like FrameHessian::dI
    std::vector<std::vector<float>> &imgs = S->imgs;
    imgs.assign(nF, std::vector<float>(3 * w * h));
    for (int f = 0; f < nF; f++) {
        std::vector<float> I(w * h);
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++)
            I[y * w + x] = 128.f + 60.f * sinf(0.11f * x) * cosf(0.07f * y) + 30.f * sinf(0.05f * (x + 2 * y)) + 0.3f * f;   // same scene, tiny brightness offset
        for (int i = 0; i < w * h; i++) {
            imgs[f][3 * i] = I[i];
            const int x = i % w, y = i / w;
            imgs[f][3 * i + 1] = (x > 0 && x < w - 1) ? 0.5f * (I[i + 1] - I[i - 1]) : 0.f;
            imgs[f][3 * i + 2] = (y > 0 && y < h - 1) ? 0.5f * (I[i + w] - I[i - w]) : 0.f;
        }
    }
    void *o = oracle_ba_create(w, h, threads_mode);
    oracle::Window *W = (oracle::Window *) o;
    S->o = o; S->W = W;
    const double K[4] = {110.0, 112.0, 79.5, 59.5};
    oracle_ba_set_calib(o, K);
    const double cd[4] = {1e-4, -2e-4, 3e-4, 1e-4};
    oracle_ba_set_calib_delta(o, cd);
    for (int f = 0; f < nF; f++) {
        double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0.004 * f, 0.001 * f, 0.0005 * f}, sz[10] = {0}, st[10] = {0};
        const double a = 0.002 * f;      // small rotation about y
        R[0] = cos(a); R[2] = sin(a); R[6] = -sin(a); R[8] = cos(a);
        sz[6] = 0.002 * f; sz[7] = 0.001 * f;
        for (int i = 0; i < 10; i++) st[i] = sz[i];
        for (int i = 0; i < 6; i++) st[i] += (f ? 1e-3 : 0.0) * (i + 1) * (f % 2 ? 1 : -1);
        st[6] += f ? 1e-4 : 0; st[7] -= f ? 2e-4 : 0;
        oracle_ba_add_frame(o, R, t, sz, st, 1.0f, f, imgs[f].data());
        oracle_ba_set_frame_energy_th(o, f, (f == 2) ? 2000.f : 8 * 8 * 8);
    }
    for (int f = 0; f < nF; f++)
        for (int k = 0; k < nPper; k++) {
            float col[8], wt[8];
            const float u = (float) (int) frand(3, w - 3), v = (float) (int) frand(3, h - 3);
            for (int i = 0; i < 8; i++) {       // colour / weight of the pattern pixel on the host image (ImmaturePoint.cc:21-35), a few perturbed
                float c3[3];
                oracle::getInterpolatedElement33BiLin(imgs[f].data(), u + oracle_pattern(i, 0), v + oracle_pattern(i, 1), w, c3);
                col[i] = c3[0] + ((k % 9 == 0) ? frand(-40, 40) : 0.f);
                wt[i] = sqrtf(2500.f / (2500.f + c3[1] * c3[1] + c3[2] * c3[2]));
            }
            const float idz = frand(0.2f, 1.5f);
            const int p = oracle_ba_add_point(o, f, u, v, idz, idz + frand(-0.02f, 0.02f), (depthPriors && k % 5 == 0) ? 1 : 0, col, wt);
            for (int t = 0; t < nF; t++) if (t != f) oracle_ba_add_residual(o, p, t);
        }
    oracle_ba_finalize(o);
    // the reference-side mirror of the window
    wG[0] = w; hG[0] = h; wM3G = w - 3; hM3G = h - 3;
    S->HC = make_calib(K[0], K[1], K[2], K[3]);
    auto HC = S->HC;
    for (int i = 0; i < 4; i++) {       // mirror the oracle's calibration state (value, zero point, float copies)
        HC->value_zero[i] = W->HCalib.value_zero[i]; HC->value[i] = W->HCalib.value[i]; HC->value_scaled[i] = W->HCalib.value_scaled[i];
        HC->value_minus_value_zero[i] = W->HCalib.value_minus_value_zero[i]; HC->value_scaledf[i] = W->HCalib.value_scaledf[i]; HC->value_scaledi[i] = W->HCalib.value_scaledi[i];
    }
    std::vector<shared_ptr<FrameHessian>> &FH = S->FH;
    FH.resize(nF);
    for (int f = 0; f < nF; f++) {
        FH[f] = make_fh(nullptr);
        FH[f]->idx = f; FH[f]->dI = (Eigen::Vector3f *) imgs[f].data(); FH[f]->frameEnergyTH = W->frames[f].frameEnergyTH;
        FH[f]->targetPrecalc.resize(nF);
        for (int t = 0; t < nF; t++) {
            const oracle::FramePrecalc &s = W->frames[f].targetPrecalc[t];
            FrameFramePrecalc &d = FH[f]->targetPrecalc[t];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
                d.PRE_RTll(i, j) = s.PRE_RTll[i * 3 + j]; d.PRE_RTll_0(i, j) = s.PRE_RTll_0[i * 3 + j];
                d.PRE_KRKiTll(i, j) = s.PRE_KRKiTll[i * 3 + j]; d.PRE_RKiTll(i, j) = s.PRE_RKiTll[i * 3 + j];
            }
            for (int i = 0; i < 3; i++) { d.PRE_tTll[i] = s.PRE_tTll[i]; d.PRE_tTll_0[i] = s.PRE_tTll_0[i]; d.PRE_KtTll[i] = s.PRE_KtTll[i]; }
            d.PRE_aff_mode[0] = s.PRE_aff_mode[0]; d.PRE_aff_mode[1] = s.PRE_aff_mode[1]; d.PRE_b0_mode = s.PRE_b0_mode; d.distanceLL = s.distanceLL;
        }
    }
    S->EF = std::make_shared<EnergyFunctional>();
    auto EF = S->EF;
    EF->nFrames = nF;
    EF->adHTdeltaF = new Mat18f[nF * nF];        // owned (delete[]d) by the reference's EnergyFunctional
    for (int q = 0; q < nF * nF; q++) for (int i = 0; i < 8; i++) EF->adHTdeltaF[q][i] = W->adHTdeltaF[8 * q + i];
    for (int i = 0; i < 4; i++) EF->cDeltaF[i] = W->cDeltaF[i];
    // adjoints, calibration prior and frame priors for the stitching (EnergyFunctional::setAdjointsF, FrameHessian::takeData)
    EF->adHost = new Mat88[nF * nF]; EF->adTarget = new Mat88[nF * nF];
    for (int q = 0; q < nF * nF; q++) for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) {
        EF->adHost[q](i, j) = W->adHost[64 * q + 8 * i + j]; EF->adTarget[q](i, j) = W->adTarget[64 * q + 8 * i + j];
    }
    for (int i = 0; i < 4; i++) EF->cPrior[i] = W->cPrior[i];
    EF->frames = FH;
    for (int f = 0; f < nF; f++) for (int i = 0; i < 8; i++) { FH[f]->prior[i] = W->frames[f].prior[i]; FH[f]->delta_prior[i] = W->frames[f].delta_prior[i]; }
    return S;
}

static void pin_linearize() {
    using namespace ldso::internal;
    Scene *S = make_scene(150);
    oracle::Window *W = S->W;
    auto HC = S->HC; auto EF = S->EF;
    std::vector<shared_ptr<FrameHessian>> &FH = S->FH;
    std::vector<std::vector<float>> &imgs = S->imgs;
    const int nF = S->nF, w = S->w, h = S->h; (void) w; (void) h;
    bool okRet = true, okState = true, okJ = true, okProj = true, okTake = true, okFix = true;
    int nIn = 0, nOob = 0, nOut = 0;
    for (size_t ri = 0; ri < W->residuals.size(); ri++) {
        oracle::Residual &orr = W->residuals[ri];
        const oracle::Point &op = W->points[orr.point];
        auto PH = std::make_shared<PointHessian>();
        PH->u = op.u; PH->v = op.v; PH->idepth_scaled = op.idepth_scaled; PH->idepth_zero_scaled = op.idepth_zero_scaled; PH->deltaF = op.deltaF;
        memcpy(PH->color, op.color, 32); memcpy(PH->weights, op.weights, 32);
        PointFrameResidual R(PH, FH[orr.host], FH[orr.target]);
        R.hostIDX = orr.host; R.targetIDX = orr.target;
        orr.hostIDX = orr.host; orr.targetIDX = orr.target;
        const double er = R.linearize(HC);            // the reference's own code
        const double eo = W->linearize(orr);          // the restatement
        okRet &= memcmp(&er, &eo, 8) == 0;
        okState &= (int) R.state_NewState == (int) orr.state_NewState && memcmp(&R.state_NewEnergy, &orr.state_NewEnergy, 8) == 0 &&
                   memcmp(&R.state_NewEnergyWithOutlier, &orr.state_NewEnergyWithOutlier, 8) == 0;
        nIn += R.state_NewState == ResState::IN; nOob += R.state_NewState == ResState::OOB; nOut += R.state_NewState == ResState::OUTLIER;
        if (R.state_NewState != ResState::OOB) {
            const RawResidualJacobian &a = *R.J; const oracle::RawResidualJacobian &b = orr.J;
            bool j = memcmp(a.resF.d, b.resF, 32) == 0;
            for (int k = 0; k < 2; k++) j &= memcmp(a.Jpdxi[k].d, b.Jpdxi[k], 24) == 0 && memcmp(a.Jpdc[k].d, b.Jpdc[k], 16) == 0 &&
                                             memcmp(a.JIdx[k].d, b.JIdx[k], 32) == 0 && memcmp(a.JabF[k].d, b.JabF[k], 32) == 0;
            j &= memcmp(a.Jpdd.d, b.Jpdd, 8) == 0;
            const float i2[4] = {a.JIdx2(0, 0), a.JIdx2(0, 1), a.JIdx2(1, 0), a.JIdx2(1, 1)}, ji[4] = {a.JabJIdx(0, 0), a.JabJIdx(0, 1), a.JabJIdx(1, 0), a.JabJIdx(1, 1)},
                        a2[4] = {a.Jab2(0, 0), a.Jab2(0, 1), a.Jab2(1, 0), a.Jab2(1, 1)};
            j &= memcmp(i2, b.JIdx2, 16) == 0 && memcmp(ji, b.JabJIdx, 16) == 0 && memcmp(a2, b.Jab2, 16) == 0;
            okJ &= j;
            bool pj = memcmp(R.centerProjectedTo.d, orr.centerProjectedTo, 12) == 0;
            for (int k = 0; k < 8; k++) pj &= memcmp(R.projectedTo[k].d, orr.projectedTo[k], 8) == 0;
            okProj &= pj;
            // applyRes(true) -> takeData, then fixLinearizationF
            R.applyRes(true); W->applyRes(orr, true);
            okTake &= R.isActive() == orr.isActive() && (int) R.state_state == (int) orr.state_state;
            if (R.isActive()) {
                okTake &= memcmp(R.JpJdF.d, orr.JpJdF, 32) == 0;
                R.fixLinearizationF(EF); W->fixLinearizationF(orr);
                okFix &= memcmp(R.res_toZeroF.d, orr.res_toZeroF, 32) == 0 && R.isLinearized == orr.isLinearized;
            }
        }
    }
    printf("  linearize pin: %zu residuals (%d IN, %d OUTLIER, %d OOB)\n", W->residuals.size(), nIn, nOut, nOob);
    CHECK(nIn > 300 && nOob > 0 && nOut > 50, "linearize scenario exercises the IN, OUTLIER and OOB branches");
    CHECK(okRet, "PointFrameResidual::linearize return value"); CHECK(okState, "linearize state_NewState / state_NewEnergy / state_NewEnergyWithOutlier");
    CHECK(okJ, "linearize RawResidualJacobian (all 74 floats)"); CHECK(okProj, "linearize projectedTo / centerProjectedTo");
    CHECK(okTake, "applyRes(true) + takeData (JpJdF, isActive)"); CHECK(okFix, "fixLinearizationF (res_toZeroF)");

    // ---- ImmaturePoint: constructor, traceOn (twice, on two frames), linearizeResidual — the reference's src/internal/ImmaturePoint.cc
    {
        oracle::TraceSettings TS;
        std::vector<shared_ptr<ldso::Frame>> FR(nF);
        for (int f = 0; f < nF; f++) { FR[f] = std::make_shared<ldso::Frame>(); FR[f]->frameHessian = FH[f]; }
        bool okCtor = true, okTrace = true, okLin = true;
        int hist[6] = {0, 0, 0, 0, 0, 0}, nLin = 0, nLinOob = 0;
        for (int k = 0; k < 1200; k++) {
            const int hst = k % (nF - 1);
            const float u = (float) (int) frand(6, w - 6), v = (float) (int) frand(6, h - 6);
            auto feat = std::make_shared<ldso::Feature>(u, v, FR[hst]);
            shared_ptr<CalibHessian> HCc = HC;
            ImmaturePoint ip(FR[hst], feat, 1, HCc);                          // the reference's constructor
            oracle::ImmaturePt op;
            oracle::immature_init(op, imgs[hst].data(), w, u, v, TS);
            const float g[4] = {ip.gradH(0, 0), ip.gradH(0, 1), ip.gradH(1, 0), ip.gradH(1, 1)};
            okCtor &= memcmp(ip.color, op.color, 32) == 0 && memcmp(ip.weights, op.weights, 32) == 0 && memcmp(g, op.gradH, 16) == 0 &&
                      memcmp(&ip.energyTH, &op.energyTH, 4) == 0;
            if (k % 5 == 0) { ip.idepth_min = op.idepth_min = frand(0.1f, 0.6f); ip.idepth_max = op.idepth_max = ip.idepth_min + frand(0.05f, 1.2f); }
            for (int pass = 0; pass < 2; pass++) {
                const int nw = (pass == 0) ? nF - 1 : (hst + 1) % nF;
                if (nw == hst) continue;
                const oracle::FramePrecalc &pc = W->frames[hst].targetPrecalc[nw];     // host -> traced frame, as FullSystem.cc:1027-1032 builds it
                Mat33f KRKi; Vec3f Kt; Vec2f aff(pc.PRE_aff_mode[0], pc.PRE_aff_mode[1]);
                for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) KRKi(i, j) = pc.PRE_KRKiTll[i * 3 + j]; Kt[i] = pc.PRE_KtTll[i] * ((k % 3) ? 1.f : 40.f); }
                float kt[3] = {Kt[0], Kt[1], Kt[2]};
                const int sr = (int) ip.traceOn(FH[nw], KRKi, Kt, aff, HC);
                const int so = oracle::trace_on(op, imgs[nw].data(), w, h, pc.PRE_KRKiTll, kt, pc.PRE_aff_mode, TS);
                hist[sr]++;
                okTrace &= sr == so && (int) ip.lastTraceStatus == op.lastTraceStatus && memcmp(&ip.idepth_min, &op.idepth_min, 4) == 0 &&
                           memcmp(&ip.idepth_max, &op.idepth_max, 4) == 0 && memcmp(&ip.quality, &op.quality, 4) == 0;
                if (sr != IPS_OOB || pass == 0)
                    okTrace &= memcmp(ip.lastTraceUV.d, op.lastTraceUV, 8) == 0 && memcmp(&ip.lastTracePixelInterval, &op.lastTracePixelInterval, 4) == 0;
            }
            // linearizeResidual against every other frame at a few depths
            oracle::Window::ImmatureCand c;
            c.u = u; c.v = v; c.host = hst; c.energyTH = op.energyTH; c.idepth_min = 0; c.idepth_max = 0;
            memcpy(c.color, op.color, 32); memcpy(c.weights, op.weights, 32);
            for (int t = 0; t < nF; t++) {
                if (t == hst) continue;
                auto tr = std::make_shared<ImmaturePointTemporaryResidual>();
                tr->state_state = ResState::IN; tr->state_energy = 0; tr->state_NewState = ResState::OUTLIER; tr->state_NewEnergy = 0; tr->target = FH[t];
                oracle::TmpRes ot = {oracle::RS_IN, oracle::RS_OUTLIER, 0.f, 0.f};
                const float idp = (k % 4 == 0) ? frand(-0.5f, 6.f) : frand(0.2f, 1.5f), slack = (k % 2) ? 1.f : 1000.f;
                float Hr = 0.5f, br = -0.25f, Ho = 0.5f, bo = -0.25f;
                const double er = ip.linearizeResidual(HC, slack, tr, Hr, br, idp);
                const double eo = oracle::immatureLinearizeResidual(*W, c, t, slack, ot, Ho, bo, idp);
                nLin++; nLinOob += tr->state_NewState == ResState::OOB;
                okLin &= memcmp(&er, &eo, 8) == 0 && memcmp(&Hr, &Ho, 4) == 0 && memcmp(&br, &bo, 4) == 0 && (int) tr->state_NewState == ot.state_NewState &&
                         (float) tr->state_NewEnergy == ot.state_NewEnergy;
            }
        }
        printf("  immature pin: traceOn statuses GOOD %d OOB %d OUTLIER %d SKIPPED %d BADCONDITION %d; linearizeResidual %d calls (%d OOB)\n",
               hist[0], hist[1], hist[2], hist[3], hist[4], nLin, nLinOob);
        CHECK(hist[0] > 100 && hist[1] > 10 && (hist[3] + hist[4]) > 10 && nLinOob > 5, "immature scenario exercises GOOD, OOB, SKIPPED/BADCONDITION and the OOB early return");
        CHECK(okCtor, "ImmaturePoint constructor (color, weights, gradH, energyTH)");
        CHECK(okTrace, "ImmaturePoint::traceOn (status, idepth interval, quality, lastTraceUV, lastTracePixelInterval)");
        CHECK(okLin, "ImmaturePoint::linearizeResidual (energy, Hdd, bd, state)");
    }
    oracle_ba_destroy(S->o); delete S;
}



// ---- AccumulatedTopHessianSSE / AccumulatedSCHessianSSE: the reference's own addPoint<mode> and stitching against oracle/ba.cc.
// addPoint is float SSE / scalar code with an explicit order: bit-exact pin. The stitching is written in Eigen expressions; compiled
// against the stand-in their 8x8 products are row-times-column sums accumulated left to right and A*B*C^T is (A*B)*C^T, which is what the
// restatement assumes too, so equality here pins the block / index / ordering structure of the stitch, not Eigen's
// own product kernels.
static bool same_dyn(const MatXX &A, const oracle::MatX &B) { return A.r == B.r && A.c == B.c && memcmp(A.d.data(), B.d.data(), 8 * A.d.size()) == 0; }
static bool same_dyn(const VecX &a, const oracle::VecXd &b) { return a.d.size() == b.size() && memcmp(a.d.data(), b.data(), 8 * b.size()) == 0; }
static void pin_hessians() {
    using namespace ldso::internal;
    Scene *S = make_scene(150);
    oracle::Window *W = S->W;
    auto HC = S->HC; auto EF = S->EF;
    const int nF = S->nF, nP = (int) W->points.size();
    std::vector<shared_ptr<PointHessian>> PH(nP);
    int nLin = 0, nAct = 0;
    for (int pi = 0; pi < nP; pi++) {
        oracle::Point &op = W->points[pi];
        if (pi % 5 == 0) op.priorF = frand(1.f, 2000.f);
        PH[pi] = std::make_shared<PointHessian>();
        auto ph = PH[pi];
        ph->u = op.u; ph->v = op.v; ph->idepth_scaled = op.idepth_scaled; ph->idepth_zero_scaled = op.idepth_zero_scaled; ph->deltaF = op.deltaF; ph->priorF = op.priorF;
        memcpy(ph->color, op.color, 32); memcpy(ph->weights, op.weights, 32);
        int k = 0;
        for (int ri : op.residuals) {
            oracle::Residual &orr = W->residuals[ri];
            auto r = std::make_shared<PointFrameResidual>(ph, S->FH[orr.host], S->FH[orr.target]);
            r->hostIDX = orr.hostIDX = orr.host; r->targetIDX = orr.targetIDX = orr.target;
            r->linearize(HC); W->linearize(orr);
            r->applyRes(true); W->applyRes(orr, true);
            // every 4th point has all its active residuals linearised (a marginalisation candidate), the others two out of three
            if (r->isActive() && (pi % 4 == 0 || k % 3 != 0)) { r->fixLinearizationF(EF); W->fixLinearizationF(orr); nLin++; }
            nAct += r->isActive();
            ph->residuals.push_back(r); k++;
        }
    }
    printf("  hessian pin: %d points, %d active residuals, %d of them linearised\n", nP, nAct, nLin);
    CHECK(nAct > 800 && nLin > 300 && nLin < nAct, "hessian scenario has active, linearised and inactive residuals");
    auto same_top_acc = [&](AccumulatedTopHessianSSE &R, oracle::AccumulatedTopHessianSSE &O, int tid) {
        bool ok = R.nres[tid] == O.nres[tid];
        for (int q = 0; q < nF * nF; q++) {
            ok &= R.acc[tid][q].num == O.acc[tid][q].num;
            for (int r = 0; r < 13; r++) for (int c = 0; c < 13; c++) ok &= memcmp(&R.acc[tid][q].H(r, c), &O.acc[tid][q].H[r * 13 + c], 4) == 0;
        }
        return ok;
    };
    auto same_point_acc = [&](bool L) {
        bool ok = true;
        for (int pi = 0; pi < nP; pi++) {
            const oracle::Point &op = W->points[pi]; const PointHessian &ph = *PH[pi];
            if (L) ok &= memcmp(&ph.Hdd_accLF, &op.Hdd_accLF, 4) == 0 && memcmp(&ph.bd_accLF, &op.bd_accLF, 4) == 0 && memcmp(ph.Hcd_accLF.d, op.Hcd_accLF, 16) == 0;
            else ok &= memcmp(&ph.Hdd_accAF, &op.Hdd_accAF, 4) == 0 && memcmp(&ph.bd_accAF, &op.bd_accAF, 4) == 0 && memcmp(ph.Hcd_accAF.d, op.Hcd_accAF, 16) == 0;
        }
        return ok;
    };
    MatXX Hr; VecX br; oracle::MatX Ho; oracle::VecXd bo;
    // active residuals: addPoint<0>, stitchDouble and stitchDoubleMT(MT = false), no prior
    AccumulatedTopHessianSSE RA; oracle::AccumulatedTopHessianSSE OA;
    RA.setZero(nF); OA.setZero(nF, 0);
    for (int pi = 0; pi < nP; pi++) { RA.addPoint<0>(PH[pi], EF.get()); W->topAddPoint<0>(OA, W->points[pi], 0); }
    CHECK(same_point_acc(false), "AccumulatedTopHessianSSE::addPoint<0>: Hdd_accAF, bd_accAF, Hcd_accAF of every point");
    RA.stitchDouble(Hr, br, EF.get(), false, false); W->topStitchDouble(OA, Ho, bo, false);
    CHECK(same_top_acc(RA, OA, 0), "addPoint<0>: all nF*nF 13x13 AccumulatorApprox blocks after finish, nres");
    if (!same_dyn(Hr, Ho)) { int nb = 0; double mx = 0; for (size_t i = 0; i < Hr.d.size(); i++) if (Hr.d[i] != Ho.d[i]) { if (nb < 5) printf("   H[%zu,%zu] ref %.17g oracle %.17g\n", i % Hr.r, i / Hr.r, Hr.d[i], Ho.d[i]); nb++; mx = std::max(mx, fabs(Hr.d[i] - Ho.d[i])); } printf("   %d entries differ, max abs %.3g\n", nb, mx); }
    CHECK(same_dyn(Hr, Ho) && same_dyn(br, bo), "AccumulatedTopHessianSSE::stitchDouble (active, no prior): H, b");
    RA.stitchDoubleMT(nullptr, Hr, br, EF.get(), false, false); W->topStitchDoubleMT(OA, Ho, bo, false, false);
    CHECK(same_dyn(Hr, Ho) && same_dyn(br, bo), "AccumulatedTopHessianSSE::stitchDoubleMT(MT = false) (active): H, b");
    // linearised residuals: addPoint<1>, with the frame / calibration priors
    AccumulatedTopHessianSSE RL; oracle::AccumulatedTopHessianSSE OL;
    RL.setZero(nF); OL.setZero(nF, 0);
    for (int pi = 0; pi < nP; pi++) { RL.addPoint<1>(PH[pi], EF.get()); W->topAddPoint<1>(OL, W->points[pi], 0); }
    CHECK(same_point_acc(true), "AccumulatedTopHessianSSE::addPoint<1>: Hdd_accLF, bd_accLF, Hcd_accLF of every point");
    RL.stitchDoubleMT(nullptr, Hr, br, EF.get(), true, false); W->topStitchDoubleMT(OL, Ho, bo, true, false);
    CHECK(same_top_acc(RL, OL, 0), "addPoint<1>: all AccumulatorApprox blocks after finish, nres");
    CHECK(same_dyn(Hr, Ho) && same_dyn(br, bo), "stitchDoubleMT(MT = false) (linearised, usePrior): H, b");
    RL.stitchDouble(Hr, br, EF.get(), true, true); W->topStitchDouble(OL, Ho, bo, true);
    CHECK(same_dyn(Hr, Ho) && same_dyn(br, bo), "stitchDouble (linearised, usePrior): H, b");
    // Schur complement accumulator: addPoint(shiftPriorToZero = true) over all points
    AccumulatedSCHessianSSE RS; oracle::AccumulatedSCHessianSSE &OS = W->accSSE_bot;
    RS.setZero(nF); OS.setZero(nF, 0);
    for (int pi = 0; pi < nP; pi++) { RS.addPoint(PH[pi], true); W->scAddPoint(W->points[pi], true, 0); }
    {
        bool ok = true;
        for (int pi = 0; pi < nP; pi++) {
            const oracle::Point &op = W->points[pi]; const PointHessian &ph = *PH[pi];
            ok &= memcmp(&ph.HdiF, &op.HdiF, 4) == 0 && memcmp(&ph.bdSumF, &op.bdSumF, 4) == 0 && memcmp(&ph.idepth_hessian, &op.idepth_hessian, 4) == 0;
        }
        CHECK(ok, "AccumulatedSCHessianSSE::addPoint: HdiF, bdSumF, idepth_hessian of every point");
    }
    RS.stitchDouble(Hr, br, EF.get()); W->scStitchDouble(Ho, bo);
    {
        bool ok = true;
        for (int q = 0; q < nF * nF; q++) {
            ok &= RS.accE[0][q].num == OS.accE[0][q].num && RS.accEB[0][q].num == OS.accEB[0][q].num;
            ok &= memcmp(RS.accEB[0][q].A1m.d, OS.accEB[0][q].A1m, 32) == 0 && memcmp(RS.accE[0][q].A1m.d, OS.accE[0][q].A1m, 128) == 0;      // both column-major
        }
        for (int q = 0; q < nF * nF * nF; q++) {
            ok &= RS.accD[0][q].num == OS.accD[0][q].num;
            ok &= memcmp(RS.accD[0][q].A1m.d, OS.accD[0][q].A1m, 256) == 0;
        }
        ok &= memcmp(RS.accbc[0].A1m.d, OS.accbc[0].A1m, 16) == 0 && memcmp(RS.accHcc[0].A1m.d, OS.accHcc[0].A1m, 64) == 0;
        CHECK(ok, "SC addPoint: accE, accEB, accD, accHcc, accbc after finish");
    }
    CHECK(same_dyn(Hr, Ho) && same_dyn(br, bo), "AccumulatedSCHessianSSE::stitchDouble: H_sc, b_sc");
    RS.stitchDoubleMT(nullptr, Hr, br, EF.get(), false); W->scStitchDoubleMT(Ho, bo, false);
    CHECK(same_dyn(Hr, Ho) && same_dyn(br, bo), "AccumulatedSCHessianSSE::stitchDoubleMT(MT = false): H_sc, b_sc");
    // marginalisation: addPoint<2> + SC addPoint(false) on the points whose active residuals are all linearised (marginalizePointsF)
    {
        AccumulatedTopHessianSSE RM; oracle::AccumulatedTopHessianSSE OM;
        RM.setZero(nF); OM.setZero(nF, 0); RS.setZero(nF); OS.setZero(nF, 0);
        int nM = 0;
        for (int pi = 0; pi < nP; pi += 4) { RM.addPoint<2>(PH[pi], EF.get()); W->topAddPoint<2>(OM, W->points[pi], 0); RS.addPoint(PH[pi], false); W->scAddPoint(W->points[pi], false, 0); nM++; }
        bool ok = same_point_acc(true) && same_point_acc(false);
        RM.stitchDouble(Hr, br, EF.get(), false, false); W->topStitchDouble(OM, Ho, bo, false);
        ok &= same_top_acc(RM, OM, 0) && same_dyn(Hr, Ho) && same_dyn(br, bo);
        RS.stitchDouble(Hr, br, EF.get()); W->scStitchDouble(Ho, bo);
        ok &= same_dyn(Hr, Ho) && same_dyn(br, bo);
        CHECK(ok && nM > 100, "addPoint<2> + SC addPoint(shiftPriorToZero = false) + both stitchDouble on the marginalisation subset");
    }
    // the multi-threaded layout, driven by hand with a fixed chunk -> thread assignment (the reference's worker threads grab chunks
    // dynamically, so its own MT result depends on scheduling): addPoint into acc[tid], stitchDoubleInternal(min, max, tid) aggregating
    // over the NUM_THREADS accumulators
    {
        AccumulatedTopHessianSSE RT; oracle::AccumulatedTopHessianSSE OT;
        const int per = (nP + NUM_THREADS - 1) / NUM_THREADS;
        for (int tid = 0; tid < NUM_THREADS; tid++) {
            RT.setZero(nF, 0, 0, 0, tid); OT.setZero(nF, tid); RS.setZero(nF, 0, 0, 0, tid); OS.setZero(nF, tid);
            for (int pi = tid * per; pi < std::min(nP, (tid + 1) * per); pi++) {
                RT.addPoint<0>(PH[pi], EF.get(), tid); W->topAddPoint<0>(OT, W->points[pi], tid);
                RS.addPoint(PH[pi], true, tid); W->scAddPoint(W->points[pi], true, tid);
            }
        }
        const int n = nF * 8 + CPARS, nk = nF * nF, kper = (nk + NUM_THREADS - 1) / NUM_THREADS;
        MatXX Hs[NUM_THREADS], Hc[NUM_THREADS]; VecX bs[NUM_THREADS], bc[NUM_THREADS];
        oracle::MatX Hso[NUM_THREADS], Hco[NUM_THREADS]; oracle::VecXd bso[NUM_THREADS], bco[NUM_THREADS];
        for (int i = 0; i < NUM_THREADS; i++) {
            Hs[i] = MatXX::Zero(n, n); bs[i] = VecX::Zero(n); Hc[i] = MatXX::Zero(n, n); bc[i] = VecX::Zero(n);
            Hso[i] = oracle::MatX(n, n); bso[i].assign(n, 0.0); Hco[i] = oracle::MatX(n, n); bco[i].assign(n, 0.0);
        }
        bool ok = true;
        for (int tid = 0; tid < NUM_THREADS; tid++) {
            const int mn = std::min(nk, tid * kper), mx = std::min(nk, (tid + 1) * kper);
            RT.stitchDoubleInternal(Hs, bs, EF.get(), true, mn, mx, nullptr, tid); W->topStitchDoubleInternal(OT, Hso, bso, true, mn, mx, tid);
            RS.stitchDoubleInternal(Hc, bc, EF.get(), mn, mx, nullptr, tid); W->scStitchDoubleInternal(Hco, bco, mn, mx, tid);
        }
        for (int tid = 0; tid < NUM_THREADS; tid++) ok &= same_dyn(Hs[tid], Hso[tid]) && same_dyn(bs[tid], bso[tid]) && same_dyn(Hc[tid], Hco[tid]) && same_dyn(bc[tid], bco[tid]);
        CHECK(ok, "stitchDoubleInternal (top with prior, and SC) aggregating NUM_THREADS accumulators, per-thread partial H / b");
    }
    oracle_ba_destroy(S->o); delete S;
}

#include "ref_hooks.h"      // LDLT / PartialPivLU / JacobiSVD hand-offs of the stand-in, implemented with the oracle's omath.h

// ---- the back end as a whole: the reference's own FrameHessian.cc, FrameFramePrecalc.cc, PointHessian.h and EnergyFunctional.cc driven
// through the calls FullSystem makes (FullSystem.cc itself needs the whole front end and is not compiled), against oracle/ba.cc.
// Single-threaded mode (multiThreading = false): the reference's worker threads pick chunks dynamically, so its 6-thread sums are not
// reproducible run to run. Dynamic-size Eigen expressions (HM * delta, the diagonal scalings, the Schur complement of marginalizeFrame,
// orthogonalize) are evaluated by the stand-in as plain left-to-right sums, LDLT / PartialPivLU / JacobiSVD / SE3 by the oracle's own
// restatements: equality pins the STRUCTURE of these functions (what is added where, in which order, with which scaling, permutation and
// sign), not Eigen's kernels.
static bool same_se3(const SE3 &a, const oracle::SE3 &b) { return memcmp(&a.q, &b.q, sizeof(b.q)) == 0 && memcmp(a.t.d, &b.t, 24) == 0; }
static void pin_backend() {
    using namespace ldso; using namespace ldso::internal;
    Scene *S = make_scene(150, 1, true);
    oracle::Window *W = S->W;
    const int nF = S->nF, nP = (int) W->points.size();
    multiThreading = false;
    // calibration: the reference's own CalibHessian constructor + setValue
    const double K[4] = {110.0, 112.0, 79.5, 59.5}, cd[4] = {1e-4, -2e-4, 3e-4, 1e-4};
    auto HC = make_calib(K[0], K[1], K[2], K[3]);
    { VecC v; for (int i = 0; i < 4; i++) v[i] = HC->value_zero[i] + cd[i]; HC->setValue(v); }
    {
        bool ok = memcmp(HC->value.d, W->HCalib.value, 32) == 0 && memcmp(HC->value_zero.d, W->HCalib.value_zero, 32) == 0 && memcmp(HC->value_scaled.d, W->HCalib.value_scaled, 32) == 0 &&
                  memcmp(HC->value_minus_value_zero.d, W->HCalib.value_minus_value_zero, 32) == 0 && memcmp(HC->value_scaledf.d, W->HCalib.value_scaledf, 16) == 0 &&
                  memcmp(HC->value_scaledi.d, W->HCalib.value_scaledi, 16) == 0;
        CHECK(ok, "CalibHessian constructor / setValue: value, value_zero, value_scaled, value_minus_value_zero, float copies");
    }
    // frames: setState / setStateZero (nullspaces) / setState as FrameHessian::setEvalPT + setState do
    std::vector<shared_ptr<Frame>> FR(nF); std::vector<shared_ptr<FrameHessian>> FH(nF);
    bool okState = true, okNull = true;
    for (int f = 0; f < nF; f++) {
        const oracle::Frame &of = W->frames[f];
        FR[f] = std::make_shared<Frame>(); FR[f]->id = of.id;
        FH[f] = make_fh(FR[f]); FR[f]->frameHessian = FH[f];
        FH[f]->frameID = of.frameID; FH[f]->ab_exposure = of.ab_exposure; FH[f]->frameEnergyTH = of.frameEnergyTH;
        FH[f]->dI = (Vec3f *) S->imgs[f].data();
        Vec10 sz, st; for (int i = 0; i < 10; i++) { sz[i] = of.state_zero[i]; st[i] = of.state[i]; }
        FH[f]->setEvalPT(SE3(of.worldToCam_evalPT), sz);
        FH[f]->setState(st);
        okState &= memcmp(FH[f]->state.d, of.state, 80) == 0 && memcmp(FH[f]->state_scaled.d, of.state_scaled, 80) == 0 && memcmp(FH[f]->state_zero.d, of.state_zero, 80) == 0 &&
                   same_se3(FH[f]->PRE_worldToCam, of.PRE_worldToCam) && same_se3(FH[f]->PRE_camToWorld, of.PRE_camToWorld);
        for (int r = 0; r < 6; r++) { okNull &= memcmp(&FH[f]->nullspaces_scale[r], &of.nullspaces_scale[r], 8) == 0; for (int c = 0; c < 6; c++) okNull &= memcmp(&FH[f]->nullspaces_pose(r, c), &of.nullspaces_pose[r][c], 8) == 0; }
        for (int r = 0; r < 4; r++) for (int c = 0; c < 2; c++) okNull &= memcmp(&FH[f]->nullspaces_affine(r, c), &of.nullspaces_affine[r][c], 8) == 0;
    }
    CHECK(okState, "FrameHessian::setEvalPT / setState: state, state_scaled, state_zero, PRE_worldToCam, PRE_camToWorld");
    CHECK(okNull, "FrameHessian::setStateZero: nullspaces_pose, nullspaces_scale, nullspaces_affine");
    // points and residuals behind the frames' features, the way makeIDX / setDeltaF / accumulate*_MT reach them
    std::vector<shared_ptr<PointHessian>> PH(nP); std::vector<shared_ptr<Point>> PT(nP);
    bool okPt = true;
    for (int pi = 0; pi < nP; pi++) {
        const oracle::Point &op = W->points[pi];
        auto feat = std::make_shared<Feature>(op.u, op.v, FR[op.host]); feat->status = Feature::FeatureStatus::VALID;
        PT[pi] = std::make_shared<Point>(); PT[pi]->status = Point::PointStatus::ACTIVE; PT[pi]->mHostFeature = feat; feat->point = PT[pi];
        PH[pi] = std::make_shared<PointHessian>(); PT[pi]->mpPH = PH[pi]; PH[pi]->point = PT[pi];
        auto ph = PH[pi];
        ph->u = op.u; ph->v = op.v; ph->hasDepthPrior = op.hasDepthPrior;
        ph->setIdepthZero(op.idepth_zero); ph->setIdepth(op.idepth);
        memcpy(ph->color, op.color, 32); memcpy(ph->weights, op.weights, 32);
        ph->takeData();
        okPt &= memcmp(&ph->idepth_scaled, &op.idepth_scaled, 4) == 0 && memcmp(&ph->idepth_zero_scaled, &op.idepth_zero_scaled, 4) == 0 && memcmp(&ph->nullspaces_scale, &op.nullspaces_scale, 4) == 0 &&
                memcmp(&ph->priorF, &op.priorF, 4) == 0 && memcmp(&ph->deltaF, &op.deltaF, 4) == 0;
        FR[op.host]->features.push_back(feat);
        for (int ri : op.residuals) {
            const oracle::Residual &orr = W->residuals[ri];
            ph->residuals.push_back(std::make_shared<PointFrameResidual>(ph, FH[orr.host], FH[orr.target]));
        }
    }
    CHECK(okPt, "PointHessian::setIdepthZero / setIdepth / takeData: scaled idepths, nullspaces_scale, priorF, deltaF");
    // EnergyFunctional::insertFrame (takeData, setAdjointsF, makeIDX)
    auto EF = std::make_shared<EnergyFunctional>();
    EF->red = new IndexThreadReduce<Vec10>();
    for (int f = 0; f < nF; f++) EF->insertFrame(FH[f], HC);
    {
        bool ok = EF->nFrames == nF && (int) EF->allPoints.size() == nP;
        for (int f = 0; f < nF; f++) ok &= memcmp(FH[f]->prior.d, W->frames[f].prior, 64) == 0 && memcmp(FH[f]->delta.d, W->frames[f].delta, 64) == 0 && memcmp(FH[f]->delta_prior.d, W->frames[f].delta_prior, 64) == 0 && FH[f]->idx == f;
        CHECK(ok, "insertFrame: FrameHessian::takeData / getPrior (prior, delta, delta_prior), makeIDX");
        bool ad = true;
        for (int q = 0; q < nF * nF; q++) for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) {
            ad &= memcmp(&EF->adHost[q](i, j), &W->adHost[64 * q + 8 * i + j], 8) == 0 && memcmp(&EF->adTarget[q](i, j), &W->adTarget[64 * q + 8 * i + j], 8) == 0;
            ad &= memcmp(&EF->adHostF[q](i, j), &W->adHostF[64 * q + 8 * i + j], 4) == 0 && memcmp(&EF->adTargetF[q](i, j), &W->adTargetF[64 * q + 8 * i + j], 4) == 0;
        }
        ad &= memcmp(EF->cPrior.d, W->cPrior, 32) == 0 && memcmp(EF->cPriorF.d, W->cPriorF, 16) == 0;
        CHECK(ad, "EnergyFunctional::setAdjointsF: adHost, adTarget (double and float), cPrior");
    }
    // FullSystem::setPrecalcValues: FrameFramePrecalc::Set for every pair, then setDeltaF
    {
        bool ok = true;
        for (int f = 0; f < nF; f++) {
            FH[f]->targetPrecalc.resize(nF);
            for (int t = 0; t < nF; t++) {
                FH[f]->targetPrecalc[t].Set(FH[f], FH[t], HC);
                const FrameFramePrecalc &d = FH[f]->targetPrecalc[t]; const oracle::FramePrecalc &o = W->frames[f].targetPrecalc[t];
                for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
                    ok &= memcmp(&d.PRE_RTll(i, j), &o.PRE_RTll[i * 3 + j], 4) == 0 && memcmp(&d.PRE_RTll_0(i, j), &o.PRE_RTll_0[i * 3 + j], 4) == 0 &&
                          memcmp(&d.PRE_KRKiTll(i, j), &o.PRE_KRKiTll[i * 3 + j], 4) == 0 && memcmp(&d.PRE_RKiTll(i, j), &o.PRE_RKiTll[i * 3 + j], 4) == 0;
                ok &= memcmp(d.PRE_tTll.d, o.PRE_tTll, 12) == 0 && memcmp(d.PRE_tTll_0.d, o.PRE_tTll_0, 12) == 0 && memcmp(d.PRE_KtTll.d, o.PRE_KtTll, 12) == 0 &&
                      memcmp(d.PRE_aff_mode.d, o.PRE_aff_mode, 8) == 0 && memcmp(&d.PRE_b0_mode, &o.PRE_b0_mode, 4) == 0 && memcmp(&d.distanceLL, &o.distanceLL, 4) == 0;
            }
        }
        CHECK(ok, "FrameFramePrecalc::Set for all frame pairs (R, t, K R K^-1, R K^-1, K t, affine mode, distance)");
        EF->setDeltaF(HC);
        bool dl = memcmp(EF->cDeltaF.d, W->cDeltaF, 16) == 0;
        for (int q = 0; q < nF * nF; q++) dl &= memcmp(EF->adHTdeltaF[q].d, &W->adHTdeltaF[8 * q], 32) == 0;
        CHECK(dl, "EnergyFunctional::setDeltaF: adHTdeltaF, cDeltaF");
    }
    // linearise everything (pinned separately), fix a part
    int nAct = 0;
    for (int pi = 0; pi < nP; pi++) {
        int k = 0;
        for (int ri : W->points[pi].residuals) {
            oracle::Residual &orr = W->residuals[ri]; auto r = PH[pi]->residuals[k];
            r->linearize(HC); W->linearize(orr);
            r->applyRes(true); W->applyRes(orr, true);
            if (r->isActive() && (pi % 4 == 0 || k % 3 != 0)) { r->fixLinearizationF(EF); W->fixLinearizationF(orr); }
            nAct += r->isActive(); k++;
        }
    }
    // a marginalisation prior HM, bM (symmetric, diagonally dominant) and the gauge nullspaces FullSystem::getNullspaces builds
    const int n = 8 * nF + CPARS;
    {
        std::vector<double> HMc((size_t) n * n), bMc(n);
        for (int j = 0; j < n; j++) { bMc[j] = frand(-50, 50); for (int i = 0; i <= j; i++) { const double v = (i == j) ? frand(2000, 9000) : frand(-30, 30); HMc[(size_t) j * n + i] = HMc[(size_t) i * n + j] = v; } }
        oracle_ba_set_marg_prior(S->o, HMc.data(), bMc.data());
        EF->HM = MatXX::Zero(n, n); EF->bM = VecX::Zero(n);
        for (int i = 0; i < n * n; i++) EF->HM.d[i] = HMc[i];
        for (int i = 0; i < n; i++) EF->bM[i] = bMc[i];
        W->getNullspaces();
        auto cp = [&](const std::vector<oracle::VecXd> &src, std::vector<VecX> &dst) { dst.clear(); for (auto &v : src) { VecX x = VecX::Zero((int) v.size()); for (size_t i = 0; i < v.size(); i++) x[i] = v[i]; dst.push_back(x); } };
        cp(W->lastNullspaces_pose, EF->lastNullspaces_pose); cp(W->lastNullspaces_scale, EF->lastNullspaces_scale);
        cp(W->lastNullspaces_affA, EF->lastNullspaces_affA); cp(W->lastNullspaces_affB, EF->lastNullspaces_affB);
    }
    // solveSystemF: iteration 0 (no orthogonalisation of x) and iteration 2 (SOLVER_ORTHOGONALIZE_X_LATER)
    for (int it = 0; it <= 2; it += 2) {
        EF->solveSystemF(it, 1e-4, HC); W->solveSystemF(it, 1e-4);
        bool ok = same_dyn(EF->lastHS, W->lastHS) && same_dyn(EF->lastbS, W->lastbS);
        CHECK(ok, it == 0 ? "solveSystemF(0): lastHS, lastbS (accumulate*_MT single-threaded, prior shift HM * delta, assembly)" : "solveSystemF(2): lastHS, lastbS");
        if (!same_dyn(EF->lastX, W->lastX)) { double mx = 0; for (int i = 0; i < n; i++) mx = std::max(mx, fabs(EF->lastX[i] - W->lastX[i])); printf("   lastX max abs diff %.3g\n", mx); }
        CHECK(same_dyn(EF->lastX, W->lastX), it == 0 ? "solveSystemF(0): lastX (scaling, LDLT hand-off, unscaling)" : "solveSystemF(2): lastX after orthogonalize(&x, 0)");
        bool st = memcmp(HC->step.d, W->HCalib.step, 32) == 0;
        for (int f = 0; f < nF; f++) st &= memcmp(FH[f]->step.d, W->frames[f].step, 80) == 0;
        for (int pi = 0; pi < nP; pi++) st &= memcmp(&PH[pi]->step, &W->points[pi].step, 4) == 0;
        CHECK(st, it == 0 ? "resubstituteF_MT / resubstituteFPt (0): calibration, frame and point steps" : "resubstituteF_MT / resubstituteFPt (2): steps");
        CHECK(EF->resInA == W->resInA && EF->resInL == W->resInL, "solveSystemF: resInA, resInL");
    }
    // orthogonalize on a vector and a matrix (SOLVER_ORTHOGONALIZE_SYSTEM / POINTMARG code path), energies
    {
        VecX b = EF->lastbS; MatXX H = EF->lastHS; oracle::VecXd bo = W->lastbS; oracle::MatX Ho = W->lastHS;
        EF->orthogonalize(&b, &H); W->orthogonalize(&bo, &Ho);
        CHECK(same_dyn(b, bo) && same_dyn(H, Ho), "EnergyFunctional::orthogonalize(b, H)");
        const double mr = EF->calcMEnergyF(), mo = W->calcMEnergyF();
        CHECK(memcmp(&mr, &mo, 8) == 0, "calcMEnergyF");
        Vec10 sr; sr.setZero(); double so[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        EF->calcLEnergyPt(0, nP, &sr, 0); W->calcLEnergyPt(0, nP, so, 0);
        CHECK(memcmp(&sr[0], &so[0], 8) == 0, "calcLEnergyPt over all points");
        const double lr = EF->calcLEnergyF_MT(), lo = W->calcLEnergyF_MT();       // the reference adds the per-chunk results in thread completion order
        CHECK(fabs(lr - lo) <= 1e-9 * fabs(lo), "calcLEnergyF_MT (to 1e-9: chunk sums arrive in thread order)");
    }
    // marginalizePointsF on the points whose active residuals are all linearised, then marginalizeFrame of frame 1
    {
        std::vector<int> idx;
        for (int pi = 0; pi < nP; pi += 4) { idx.push_back(pi); PT[pi]->status = Point::PointStatus::MARGINALIZED; W->points[pi].priorF *= setting_idepthFixPriorMargFac; }
        EF->marginalizePointsF(); W->marginalizePointsF(idx);
        CHECK(same_dyn(EF->HM, W->HM) && same_dyn(EF->bM, W->bM), "EnergyFunctional::marginalizePointsF: HM, bM");
        CHECK(EF->resInM == W->resInM && (int) EF->allPoints.size() == nP - (int) idx.size(), "marginalizePointsF: resInM, remaining points");
#ifdef PIN_SELFTEST_BREAK
        W->HM(9, 20) = std::nextafter(W->HM(9, 20), 1e300);
#endif
        EF->marginalizeFrame(FH[1]); W->marginalizeFramePrior(1);
        if (!same_dyn(EF->HM, W->HM)) { double mx = 0; for (size_t i = 0; i < EF->HM.d.size() && i < W->HM.d.size(); i++) mx = std::max(mx, fabs(EF->HM.d[i] - W->HM.d[i])); printf("   marginalizeFrame HM %dx%d vs %dx%d max abs diff %.3g\n", EF->HM.r, EF->HM.c, W->HM.r, W->HM.c, mx); }
        CHECK(same_dyn(EF->HM, W->HM) && same_dyn(EF->bM, W->bM), "EnergyFunctional::marginalizeFrame (a middle frame): HM, bM");
        CHECK(EF->nFrames == nF - 1 && FH[2]->idx == 1 && FH[3]->idx == 2, "marginalizeFrame bookkeeping");
    }
    { double nx = 0, nh = 0; for (int i = 0; i < n; i++) nx += EF->lastX[i] * EF->lastX[i]; for (double v : EF->HM.d) nh += v * v;
      printf("  backend pin: %d frames, %d points, %d active residuals, system size %d; |lastX| = %.6g, |HM after marginalizeFrame| = %.6g (%dx%d)\n", nF, nP, nAct, n, sqrt(nx), sqrt(nh), EF->HM.r, EF->HM.c); }
    delete EF->red; EF->red = nullptr;
    oracle_ba_destroy(S->o); delete S;
}

// ---- CoarseTracker: the reference's src/frontend/CoarseTracker.cc against oracle/tracker.cc
static void pin_tracker() {
    using namespace ldso; using namespace ldso::internal;
    const int w = 640, h = 480, L = 4;
    pyrLevelsUsed = L;
    for (int l = 0; l < L; l++) { wG[l] = w >> l; hG[l] = h >> l; }
    wM3G = w - 3; hM3G = h - 3;
    // a fronto-parallel textured plane at inverse depth id0 seen from two cameras a small x/y translation apart: the new image is the
    // reference image shifted by (fx*tx*id0, fy*ty*id0) pixels, plus an affine brightness change
    const float fxl = 520.f, fyl = 522.f, cxl = 318.3f, cyl = 241.1f, id0 = 0.8f;
    const float shx = 2.6f, shy = -1.3f;
    auto tex = [](float x, float y) {
        return 120.f + 40.f * sinf(0.013f * x + 0.3f) * cosf(0.017f * y) + 25.f * sinf(0.045f * (x + 0.6f * y)) + 14.f * cosf(0.11f * x - 0.07f * y) +
               9.f * sinf(0.31f * x) * sinf(0.27f * y);
    };
    std::vector<float> colRef(w * h), colNew(w * h);
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
        colRef[y * w + x] = tex(x, y);
        colNew[y * w + x] = 1.04f * tex(x - shx, y - shy) + 3.f;
    }
    std::vector<std::vector<float>> pyrRef(L), pyrNew(L);
    float *pr[PYR_LEVELS] = {}, *pn[PYR_LEVELS] = {};
    for (int l = 0; l < L; l++) { pyrRef[l].assign(3 * (w >> l) * (h >> l), 0.f); pyrNew[l].assign(3 * (w >> l) * (h >> l), 0.f); pr[l] = pyrRef[l].data(); pn[l] = pyrNew[l].data(); }
    oracle::makeImages(colRef.data(), w, h, L, pr);
    oracle::makeImages(colNew.data(), w, h, L, pn);
    {   // FrameHessian::makeImages, the reference's own (src/internal/FrameHessian.cc:44-100), against the oracle's pyramid
        auto fhI = make_fh(nullptr, true);
        fhI->makeImages(colRef.data(), make_calib(520, 522, 318.3, 241.1));
        bool ok = fhI->dI == fhI->dIp[0];
        // the reference leaves the first and last image row's gradients uninitialised beyond its byte-count memset; compare the rows it writes
        for (int l = 0; l < L; l++) { const int wl = w >> l, hl = h >> l; for (int i = wl; i < wl * (hl - 1); i++) ok &= memcmp(fhI->dIp[l][i].d, pr[l] + 3 * i, 12) == 0;
                                      for (int i = 0; i < wl * hl; i++) ok &= memcmp(&fhI->dIp[l][i][0], pr[l] + 3 * i, 4) == 0; }
        CHECK(ok, "FrameHessian::makeImages: intensity pyramid and gradients on every level");
    }

    // reference side: frames, features, points, their newest residual
    auto HC = make_calib(fxl, fyl, cxl, cyl);
    const int nKF = 3, nPer = 900;
    std::vector<shared_ptr<FrameHessian>> FH(nKF);
    std::vector<shared_ptr<Frame>> FR(nKF);
    for (int f = 0; f < nKF; f++) {
        FR[f] = std::make_shared<Frame>(); FH[f] = make_fh(FR[f]);
        FR[f]->frameHessian = FH[f]; FR[f]->id = 10 + f; FH[f]->idx = f;
    }
    auto lastRef = FH[nKF - 1];
    for (int l = 0; l < L; l++) lastRef->dIp[l] = (Vec3f *) pr[l];
    lastRef->dI = lastRef->dIp[0]; lastRef->ab_exposure = 1.0f; lastRef->setEvalPT_scaled(SE3(), AffLight(0.02f, -1.5f));
    auto newFH = make_fh(nullptr);
    for (int l = 0; l < L; l++) newFH->dIp[l] = (Vec3f *) pn[l];
    newFH->dI = newFH->dIp[0]; newFH->ab_exposure = 1.061f;     // so that the affine brightness (0, 0) is close to the truth (gain 1.04)
    std::vector<float> cpt, hdi;            // the oracle's input: the contributions in the order the reference visits them
    std::vector<shared_ptr<PointFrameResidual>> keep;
    int nSkipped = 0;
    for (int f = 0; f < nKF; f++)
        for (int k = 0; k < nPer; k++) {
            auto feat = std::make_shared<Feature>(0.f, 0.f, FR[f]);
            auto pt = std::make_shared<Point>();
            auto ph = std::make_shared<PointHessian>();
            feat->point = pt; pt->mpPH = ph; feat->status = Feature::FeatureStatus::VALID; pt->status = Point::PointStatus::ACTIVE;
            auto r = std::make_shared<PointFrameResidual>(ph, FH[f], lastRef);
            r->isActiveAndIsGoodNEW = true;
            // projected position in lastRef; every 11th point lands on the pixel of the previous one (accumulation), a band is left empty (dilation)
            float u = frand(1.f, w - 2.f), v = frand(1.f, h - 2.f);
            if (v > 200 && v < 260 && u > 100 && u < 400) v += 70;
            if (k % 11 == 0 && !cpt.empty()) { u = cpt[cpt.size() - 3]; v = cpt[cpt.size() - 2]; }
            r->centerProjectedTo = Vec3f(u, v, id0 * (1.f + frand(-0.02f, 0.02f)));
            ph->HdiF = frand(0.5f, 400.f);
            ph->lastResiduals[0] = std::make_pair(r, ResState::IN);
            int skip = 0;
            if (k % 17 == 3) { feat->status = Feature::FeatureStatus::OUTLIER; skip = 1; }
            if (k % 19 == 4) { pt->status = Point::PointStatus::MARGINALIZED; skip = 1; }
            if (k % 23 == 5) { ph->lastResiduals[0].second = ResState::OOB; skip = 1; }
            if (k % 29 == 6) { ph->lastResiduals[0].first = nullptr; skip = 1; }
            FR[f]->features.push_back(feat); keep.push_back(r);
            if (skip) { nSkipped++; continue; }
            cpt.push_back(r->centerProjectedTo[0]); cpt.push_back(r->centerProjectedTo[1]); cpt.push_back(r->centerProjectedTo[2]);
            hdi.push_back(ph->HdiF);
        }
    CoarseTracker R(w, h);
    R.makeK(HC);
    R.setCoarseTrackingRef(FH);
    oracle::CoarseTracker O(w, h, L);
    O.makeK(fxl, fyl, cxl, cyl);
    for (int l = 0; l < L; l++) { O.refDIp[l] = pr[l]; O.newDIp[l] = pn[l]; }
    O.lastRef_aff_a = lastRef->aff_g2l().a; O.lastRef_aff_b = lastRef->aff_g2l().b; O.lastRef_ab_exposure = lastRef->ab_exposure; O.newFrame_ab_exposure = newFH->ab_exposure;
    O.makeCoarseDepthL0((int) hdi.size(), cpt.data(), hdi.data());

    bool okK = true, okPc = true, okMap = true;
    for (int l = 0; l < L; l++) {
        okK &= R.w[l] == O.w[l] && R.h[l] == O.h[l] && memcmp(&R.fx[l], &O.fx[l], 4) == 0 && memcmp(&R.fy[l], &O.fy[l], 4) == 0 && memcmp(&R.cx[l], &O.cx[l], 4) == 0 &&
               memcmp(&R.cy[l], &O.cy[l], 4) == 0 && memcmp(&R.fxi[l], &O.fxi[l], 4) == 0 && memcmp(&R.fyi[l], &O.fyi[l], 4) == 0 && memcmp(&R.cxi[l], &O.cxi[l], 4) == 0 &&
               memcmp(&R.cyi[l], &O.cyi[l], 4) == 0;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) okK &= memcmp(&R.K[l](i, j), &O.K[l][i * 3 + j], 4) == 0 && memcmp(&R.Ki[l](i, j), &O.Ki[l][i * 3 + j], 4) == 0;
        okPc &= R.pc_n[l] == O.pc_n[l];
        if (R.pc_n[l] == O.pc_n[l]) {
            const size_t nb = 4 * (size_t) R.pc_n[l];
            okPc &= memcmp(R.pc_u[l], O.pc_u[l].data(), nb) == 0 && memcmp(R.pc_v[l], O.pc_v[l].data(), nb) == 0 && memcmp(R.pc_idepth[l], O.pc_idepth[l].data(), nb) == 0 &&
                    memcmp(R.pc_color[l], O.pc_color[l].data(), nb) == 0;
        }
        const size_t nm = 4 * (size_t) R.w[l] * R.h[l];
        okMap &= memcmp(R.idepth[l], O.idepth[l].data(), nm) == 0 && memcmp(R.weightSums[l], O.weightSums[l].data(), nm) == 0;
    }
    printf("  tracker pin: %zu contributing points (%d filtered), pc_n = %d %d %d %d\n", hdi.size(), nSkipped, R.pc_n[0], R.pc_n[1], R.pc_n[2], R.pc_n[3]);
    CHECK(okK, "CoarseTracker::makeK (w, h, fx.., K, Ki per level)");
    CHECK(R.pc_n[0] > 1500 && R.pc_n[0] < R.pc_n[1] + 100000 && nSkipped > 100, "tracker scenario: enough points, filters exercised");
    CHECK(okPc, "setCoarseTrackingRef / makeCoarseDepthL0: pc_n, pc_u, pc_v, pc_idepth, pc_color on every level");
    CHECK(okMap, "makeCoarseDepthL0: idepth and weightSums maps on every level");
    CHECK(R.refFrameID == 10 + nKF - 1 && R.lastRef_aff_g2l.a == lastRef->aff_g2l().a, "setCoarseTrackingRef bookkeeping");


    // ---- CoarseDistanceMap (same file): makeK, makeDistanceMap (projection of the ACTIVE points of the other keyframes + BFS), addIntoDistFinal
    {
        const int nH = 3, nPts = 700;
        std::vector<shared_ptr<FrameHessian>> DF(nH + 1); std::vector<shared_ptr<Frame>> DR(nH + 1);
        CoarseDistanceMap RD(w, h); oracle::CoarseDistanceMap OD(w, h, L);
        RD.makeK(HC); OD.makeK(fxl, fyl, cxl, cyl);
        for (int f = 0; f <= nH; f++) {
            DR[f] = std::make_shared<Frame>(); DF[f] = make_fh(DR[f]); DR[f]->frameHessian = DF[f];
            Vec6 xi; for (int i = 0; i < 6; i++) xi[i] = frand(-1.f, 1.f) * (i < 3 ? 0.05 : 0.02);
            DF[f]->PRE_worldToCam = SE3::exp(xi); DF[f]->PRE_camToWorld = DF[f]->PRE_worldToCam.inverse();
        }
        auto newest = DF[nH];
        OD.beginDistanceMap();
        for (int f = 0; f < nH; f++) {
            std::vector<float> pu, pv, pid;
            for (int k = 0; k < nPts; k++) {
                auto feat = std::make_shared<Feature>(0.f, 0.f, DR[f]); auto pt = std::make_shared<Point>(); auto ph = std::make_shared<PointHessian>();
                feat->point = pt; pt->mpPH = ph; pt->status = (k % 13 == 5) ? Point::PointStatus::OUTLIER : Point::PointStatus::ACTIVE;
                ph->u = frand(-20.f, w + 20.f); ph->v = frand(-20.f, h + 20.f); ph->idepth_scaled = frand(0.05f, 2.5f);
                if (k < 40) { ph->u = frand(0.f, 3.f); ph->v = frand(0.f, h - 1.f); }       // some land on the left border column of level 1
                DR[f]->features.push_back(feat);
                if (k % 17 == 3) feat->point = nullptr;
                else if (pt->status == Point::PointStatus::ACTIVE) { pu.push_back(ph->u); pv.push_back(ph->v); pid.push_back(ph->idepth_scaled); }
            }
            const SE3 fhToNew = newest->PRE_worldToCam * DF[f]->PRE_camToWorld;
            const Mat33f Rf = fhToNew.rotationMatrix().cast<float>(); const Vec3f tf = fhToNew.translation().cast<float>();
            float Rr[9], tr[3]; for (int i = 0; i < 3; i++) { tr[i] = tf[i]; for (int j = 0; j < 3; j++) Rr[i * 3 + j] = Rf(i, j); }
            OD.addFramePoints(Rr, tr, (int) pu.size(), pu.data(), pv.data(), pid.data());
        }
        OD.finishDistanceMap();
        RD.makeDistanceMap(DF, newest);
        const size_t nb = 4 * (size_t) RD.w[1] * RD.h[1];
        int nz = 0, nfar = 0; for (int i = 0; i < RD.w[1] * RD.h[1]; i++) { nz += RD.fwdWarpedIDDistFinal[i] == 0; nfar += RD.fwdWarpedIDDistFinal[i] > 5; }
        bool okK2 = true; for (int l = 0; l < L; l++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) okK2 &= memcmp(&RD.K[l](i, j), &OD.K[l][i * 3 + j], 4) == 0 && memcmp(&RD.Ki[l](i, j), &OD.Ki[l][i * 3 + j], 4) == 0;
        CHECK(okK2 && RD.w[1] == OD.w[1] && RD.h[1] == OD.h[1], "CoarseDistanceMap::makeK");
        CHECK(nz > 800 && nfar > 500, "distance map scenario: many seeds, regions farther than 5");
        CHECK(memcmp(RD.fwdWarpedIDDistFinal, OD.fwdWarpedIDDistFinal.data(), nb) == 0, "CoarseDistanceMap::makeDistanceMap: the level-1 distance map");
        bool okAdd = true;
        for (int k = 0; k < 400; k++) {
            const int u = 1 + (int) frand(0.f, RD.w[1] - 2.f), v = 1 + (int) frand(0.f, RD.h[1] - 2.f);
            RD.addIntoDistFinal(u, v); OD.addIntoDistFinal(u, v);
            if (k % 50 == 49) okAdd &= memcmp(RD.fwdWarpedIDDistFinal, OD.fwdWarpedIDDistFinal.data(), nb) == 0;
        }
        CHECK(okAdd, "CoarseDistanceMap::addIntoDistFinal x400 (incremental BFS)");
        printf("  distance map pin: %d seed cells, %d cells farther than 5 before the incremental adds\n", nz, nfar);
    }

    // calcRes / calcGSSSE at several poses, brightness parameters, cutoffs and levels
    R.newFrame = newFH;
    bool okRes = true, okBuf = true, okH = true; int nSat = 0, nEval = 0;
    for (int trial = 0; trial < 12; trial++) {
        Vec6 xi; double xia[6];
        for (int i = 0; i < 6; i++) { xia[i] = (trial == 0) ? 0.0 : frand(-1.f, 1.f) * (i < 3 ? 0.01 : 0.004) * (1 + trial % 3); xi[i] = xia[i]; }
        const SE3 Tr = SE3::exp(xi); const oracle::SE3 To = oracle::SE3::exp(xia);
        const AffLight aff(frand(-0.03f, 0.03f), frand(-4.f, 4.f));
        const float cutoff = (trial % 4 == 1) ? 6.f : (trial % 4 == 2 ? 40.f : 20.f);
        for (int l = L - 1; l >= 0; l--) {
            const Vec6 rr = R.calcRes(l, Tr, aff, cutoff);
            double ro[6]; O.calcRes(l, To, aff.a, aff.b, cutoff, ro);
#ifdef PIN_SELFTEST_BREAK
            if (trial == 3 && l == 1) ro[0] = std::nextafter(ro[0], 1e30);
#endif
            okRes &= memcmp(rr.d, ro, 48) == 0;
            okBuf &= R.buf_warped_n == O.buf_warped_n;
            if (R.buf_warped_n == O.buf_warped_n) {
                const size_t nb = 4 * (size_t) R.buf_warped_n;
                okBuf &= memcmp(R.buf_warped_idepth, O.buf_warped_idepth.data(), nb) == 0 && memcmp(R.buf_warped_u, O.buf_warped_u.data(), nb) == 0 &&
                         memcmp(R.buf_warped_v, O.buf_warped_v.data(), nb) == 0 && memcmp(R.buf_warped_dx, O.buf_warped_dx.data(), nb) == 0 &&
                         memcmp(R.buf_warped_dy, O.buf_warped_dy.data(), nb) == 0 && memcmp(R.buf_warped_residual, O.buf_warped_residual.data(), nb) == 0 &&
                         memcmp(R.buf_warped_weight, O.buf_warped_weight.data(), nb) == 0 && memcmp(R.buf_warped_refColor, O.buf_warped_refColor.data(), nb) == 0;
            }
            nSat += rr[5] > 0; nEval++;
            Mat88 Hr; Vec8 br; R.calcGSSSE(l, Hr, br, Tr, aff);
            double Ho[64], bo[8]; O.calcGSSSE(l, Ho, bo, To, aff.a, aff.b);
            for (int i = 0; i < 8; i++) { okH &= memcmp(&br[i], &bo[i], 8) == 0; for (int j = 0; j < 8; j++) okH &= memcmp(&Hr(i, j), &Ho[i * 8 + j], 8) == 0; }
        }
    }
    printf("  tracker pin: %d of %d calcRes evaluations had saturated residuals\n", nSat, nEval);
    CHECK(nSat > 0 && nSat < nEval, "calcRes scenario exercises the saturated (cutoff) branch on some evaluations");
    CHECK(okRes, "CoarseTracker::calcRes return vector (E, count, flow indicators, saturated ratio)");
    CHECK(okBuf, "calcRes warped buffers (idepth, u, v, dx, dy, residual, weight, refColor, padded count)");
    CHECK(okH, "CoarseTracker::calcGSSSE H (8x8) and b after the SCALE_* rescale");

    // the whole coarse-to-fine LM loop, from three starts (identity; a start that is too far and fails the residual check; wrong brightness)
    bool okTrack = true; int nTrue = 0, nFalse = 0, its = 0;
    for (int run = 0; run < 4; run++) {
        Vec6 xi; double xia[6] = {0, 0, 0, 0, 0, 0};
        if (run == 1) { xia[0] = 0.004; xia[4] = -0.003; }
        if (run == 2) { xia[0] = 0.25; xia[1] = -0.2; xia[5] = 0.3; }
        for (int i = 0; i < 6; i++) xi[i] = xia[i];
        SE3 Tr = SE3::exp(xi); oracle::SE3 To = oracle::SE3::exp(xia);
        AffLight aff(run == 3 ? 0.3f : 0.f, run == 3 ? 20.f : 0.f); float oa = aff.a, ob = aff.b;
        Vec5 minRes; double minResO[5];
        for (int i = 0; i < 5; i++) { minRes[i] = minResO[i] = (run == 2) ? 1.0 : NAN; }
        const bool gr = R.trackNewestCoarse(newFH, Tr, aff, L - 1, minRes);
        const bool go = O.trackNewestCoarse(To, oa, ob, L - 1, minResO);
        nTrue += gr; nFalse += !gr; its += O.lm_iterations_total;
        bool same = gr == go && memcmp(&aff.a, &oa, 4) == 0 && memcmp(&aff.b, &ob, 4) == 0 && memcmp(&Tr.q, &To.q, sizeof(To.q)) == 0 && memcmp(Tr.t.d, &To.t, 24) == 0;
        same &= memcmp(R.lastResiduals.d, O.lastResiduals, 40) == 0 && memcmp(R.lastFlowIndicators.d, O.lastFlowIndicators, 24) == 0;
        okTrack &= same;
        printf("  tracker pin: track run %d -> %s, t = (%.5f %.5f %.5f) (true %.5f %.5f 0), a = %.4f b = %.3f, %d calcRes evaluations\n", run, gr ? "good" : "lost",
                             Tr.t[0], Tr.t[1], Tr.t[2], shx / (fxl * id0), shy / (fyl * id0), aff.a, aff.b, O.lm_iterations_total);
    }
    CHECK(nTrue >= 2 && nFalse >= 1, "trackNewestCoarse scenario has converging and aborting runs");
    CHECK(okTrack, "CoarseTracker::trackNewestCoarse: return value, pose, affine brightness, lastResiduals, lastFlowIndicators");
}

// ---- CoarseInitializer::makeK / calcResAndGS: the reference's src/frontend/CoarseInitializer.cc against oracle/initializer.cc
static void pin_initializer() {
    using namespace ldso; using namespace ldso::internal;
    const int w = 640, h = 480, L = 4;
    pyrLevelsUsed = L;
    for (int l = 0; l < L; l++) { wG[l] = w >> l; hG[l] = h >> l; }
    auto tex = [](float x, float y) { return 120.f + 40.f * sinf(0.013f * x + 0.3f) * cosf(0.017f * y) + 25.f * sinf(0.045f * (x + 0.6f * y)) + 14.f * cosf(0.11f * x - 0.07f * y); };
    std::vector<float> c0(w * h), c1(w * h);
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) { c0[y * w + x] = tex(x, y); c1[y * w + x] = 1.03f * tex(x - 1.7f, y + 0.9f) + 2.f; }
    std::vector<std::vector<float>> p0(L), p1(L); float *q0[PYR_LEVELS] = {}, *q1[PYR_LEVELS] = {};
    for (int l = 0; l < L; l++) { p0[l].assign(3 * (w >> l) * (h >> l), 0.f); p1[l].assign(3 * (w >> l) * (h >> l), 0.f); q0[l] = p0[l].data(); q1[l] = p1[l].data(); }
    oracle::makeImages(c0.data(), w, h, L, q0); oracle::makeImages(c1.data(), w, h, L, q1);
    auto HC = make_calib(520.f, 522.f, 318.3f, 241.1f);
    auto f0 = make_fh(nullptr), f1 = make_fh(nullptr);
    for (int l = 0; l < L; l++) { f0->dIp[l] = (Vec3f *) q0[l]; f1->dIp[l] = (Vec3f *) q1[l]; }
    CoarseInitializer RI(w, h); oracle::CoarseInitializer OI(w, h, L);
    RI.makeK(HC); OI.makeK(520.f, 522.f, 318.3f, 241.1f);
    RI.firstFrame = f0; RI.newFrame = f1;
    RI.alphaK = OI.alphaK; RI.alphaW = OI.alphaW; RI.couplingWeight = OI.couplingWeight; RI.regWeight = 0.8f;
    bool okK = true;
    for (int l = 0; l < L; l++) { okK &= RI.w[l] == OI.w[l] && RI.h[l] == OI.h[l] && memcmp(&RI.fx[l], &OI.fx[l], 8) == 0 && memcmp(&RI.cy[l], &OI.cy[l], 8) == 0;
                                  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) okK &= memcmp(&RI.K[l](i, j), &OI.K[l][i * 3 + j], 8) == 0 && memcmp(&RI.Ki[l](i, j), &OI.Ki[l][i * 3 + j], 8) == 0; }
    CHECK(okK, "CoarseInitializer::makeK (double K, Ki per level)");
    bool okRes = true, okH = true, okPts = true, okJb = true; int nGood = 0, nBad = 0, nCapped = 0;
    for (int l = 0; l < L; l++) {
        OI.firstDIp[l] = q0[l]; OI.newDIp[l] = q1[l];
        const int wl = w >> l, hl = h >> l, n = (l == 0) ? 3000 : 1200 >> l;
        RI.points[l] = new Pnt[n]; RI.numPoints[l] = n; OI.points[l].assign(n, oracle::InitPnt());
        for (int i = 0; i < n; i++) {
            Pnt &a = RI.points[l][i]; oracle::InitPnt &b = OI.points[l][i];
            a.u = b.u = (float) (int) frand(3.f, wl - 4.f); a.v = b.v = (float) (int) frand(3.f, hl - 4.f);
            if (i % 37 == 5) { a.u = b.u = 2.f; }                                   // pattern touches the border test of the warped position
            a.idepth = b.idepth = 1; a.idepth_new = b.idepth_new = frand(0.4f, 2.2f); a.iR = b.iR = frand(0.8f, 1.2f);
            a.isGood = b.isGood = (i % 11 != 3);
            a.energy = Vec2f(frand(0.f, 300.f), frand(0.f, 1.f)); b.energy[0] = a.energy[0]; b.energy[1] = a.energy[1];
            a.energy_new = Vec2f(0, 0); a.isGood_new = false; a.lastHessian = a.lastHessian_new = 0; a.maxstep = 0;
            a.outlierTH = b.outlierTH = (i % 13 == 7) ? 0.5f : 8 * 12 * 12.f; a.my_type = 1;
        }
    }
    for (int trial = 0; trial < 6; trial++) {
        Vec6 xi; double xia[6];
        for (int i = 0; i < 6; i++) { xia[i] = (trial == 0) ? 0.0 : frand(-1.f, 1.f) * (i < 3 ? (trial >= 4 ? 0.3 : 0.01) : 0.003); xi[i] = xia[i]; }
        const SE3 Tr = SE3::exp(xi); const oracle::SE3 To = oracle::SE3::exp(xia);
        const AffLight aff(frand(-0.05f, 0.05f), frand(-3.f, 3.f));
        for (int l = L - 1; l >= 0; l--) {
            Mat88f H, Hsc; Vec8f b, bsc; float Ho[64], bo[8], Hsco[64], bsco[8], ro[3];
            const Vec3f rr = RI.calcResAndGS(l, H, b, Hsc, bsc, Tr, aff, false);
            OI.calcResAndGS(l, Ho, bo, Hsco, bsco, To, aff.a, aff.b, ro);
            okRes &= memcmp(rr.d, ro, 12) == 0;
            if (ro[1] == OI.alphaK * OI.points[l].size()) nCapped++;
            for (int i = 0; i < 8; i++) { okH &= memcmp(&b[i], &bo[i], 4) == 0 && memcmp(&bsc[i], &bsco[i], 4) == 0;
                                          for (int j = 0; j < 8; j++) okH &= memcmp(&H(i, j), &Ho[i * 8 + j], 4) == 0 && memcmp(&Hsc(i, j), &Hsco[i * 8 + j], 4) == 0; }
            for (size_t i = 0; i < OI.points[l].size(); i++) {
                const Pnt &a = RI.points[l][i]; const oracle::InitPnt &c = OI.points[l][i];
                okPts &= a.isGood_new == c.isGood_new && memcmp(a.energy_new.d, c.energy_new, 8) == 0 && memcmp(&a.maxstep, &c.maxstep, 4) == 0 &&
                         memcmp(&a.lastHessian_new, &c.lastHessian_new, 4) == 0;
                if (a.isGood) okJb &= memcmp(RI.JbBuffer_new[i].d, OI.JbBuffer_new[i].data(), 40) == 0;
                nGood += a.isGood_new; nBad += !a.isGood_new;
            }
        }
    }
    printf("  initializer pin: %d good / %d rejected point evaluations, alpha energy capped in %d of 24 calls\n", nGood, nBad, nCapped);
    CHECK(nGood > 10000 && nBad > 2000 && nCapped > 0 && nCapped < 24, "initializer scenario: good and rejected points, both alpha branches");
    CHECK(okRes, "CoarseInitializer::calcResAndGS return vector (E.A, alphaEnergy, E.num)");
    CHECK(okH, "calcResAndGS H, b, H_sc, b_sc");
    CHECK(okPts, "calcResAndGS per-point isGood_new, energy_new, maxstep, lastHessian_new");
    CHECK(okJb, "calcResAndGS JbBuffer_new (10 floats per point)");
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
    pin_linearize();
    pin_hessians();
    pin_backend();
    pin_tracker();
    pin_initializer();
    pin_settings();
    if (fails) { printf("PIN FAILED: %d of %d checks\n", fails, checks); return 1; }
    printf("PIN OK: %d checks against the reference's own MatrixAccumulators.h, GlobalFuncs.h, ResidualProjections.h, AffLight.h, Setting.cc, Residuals.cc, ImmaturePoint.cc, PointHessian.cc, FrameHessian.cc, FrameFramePrecalc.cc, AccumulatedTopHessian.cc, AccumulatedSCHessian.cc, EnergyFunctional.cc, CoarseTracker.cc and CoarseInitializer.cc\n", checks);
    return 0;
}
