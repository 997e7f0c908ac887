// ORACLE — TEST INFRASTRUCTURE ONLY (see omath.h header). PARITY PINNED AGAINST THE REFERENCE'S OWN SOURCES, EXCEPT EIGEN / SOPHUS INTERNALS (oracle/ref_pin compiles, unmodified, the reference's Residuals.cc, ImmaturePoint.cc, PointHessian.cc, FrameHessian.cc, FrameFramePrecalc.cc, AccumulatedTopHessian.cc, AccumulatedSCHessian.cc, EnergyFunctional.cc, CoarseTracker.cc, Setting.cc and their headers against stand-in Eigen/Sophus headers and matches the oracle bit for bit on 108 checks; the float SSE / scalar code is pinned completely, the Eigen-expression code structurally - what is added where, in which order - with LDLT, PartialPivLU, JacobiSVD, SE3 exp/log and the product kernels being the oracle's own restatements on both sides; FullSystem.cc's driver loop needs the whole front end and stays restated from the cited lines).
// Data records of the windowed photometric BA, restated from the reference without the
// shared_ptr graph (indices instead):
//   RawResidualJacobian   include/internal/RawResidualJacobian.h:13-39
//   PointFrameResidual    include/internal/Residuals.h:40-130
//   PointHessian          include/internal/PointHessian.h:19-132
//   FrameHessian          include/internal/FrameHessian.h:27-214
//   CalibHessian          include/internal/CalibHessian.h:16-140
//   FrameFramePrecalc     include/internal/FrameFramePrecalc.h:22-45
//   EnergyFunctional      include/internal/OptimizationBackend/EnergyFunctional.h:54-231
#pragma once
#include <vector>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <atomic>
#include "omath.h"
#include "accumulators.h"

namespace oracle {

// ---- compile-time constants (include/Settings.h:8-43,163 ; include/NumTypes.h:26,28)
static const int NUM_THREADS = 6;
static const int CPARS = 4;
static const int patternNum = 8;
static const float SCALE_IDEPTH = 1.0f;
static const float SCALE_XI_ROT = 1.0f;
static const float SCALE_XI_TRANS = 0.5f;
static const float SCALE_F = 50.0f;
static const float SCALE_C = 50.0f;
static const float SCALE_A = 10.0f;
static const float SCALE_B = 1000.0f;
static const float SCALE_XI_ROT_INVERSE = 1.0f / SCALE_XI_ROT;
static const float SCALE_XI_TRANS_INVERSE = 1.0f / SCALE_XI_TRANS;
static const float SCALE_F_INVERSE = 1.0f / SCALE_F;
static const float SCALE_C_INVERSE = 1.0f / SCALE_C;
static const float SCALE_A_INVERSE = 1.0f / SCALE_A;
static const float SCALE_B_INVERSE = 1.0f / SCALE_B;
// staticPattern[8] (src/Setting.cc:221)
static const int patternP[8][2] = {{0, -2}, {-1, -1}, {1, -1}, {-2, 0}, {0, 0}, {2, 0}, {-1, 1}, {0, 2}};

// ---- run-time settings with the reference defaults (src/Setting.cc:18-23,41,65-66,73,76-81)
struct Settings {
    float initialRotPrior = 1e11f, initialTransPrior = 1e10f;
    float initialAffBPrior = 1e14f, initialAffAPrior = 1e14f;
    float initialCalibHessian = 5e9f;
    double solverModeDelta = 0.00001;
    float idepthFixPrior = 50 * 50;
    float outlierTHSumComponent = 50 * 50;
    float affineOptModeA = 1e12f, affineOptModeB = 1e8f;
    float huberTH = 9;
    float frameEnergyTHConstWeight = 0.5f, frameEnergyTHN = 0.7f;
    float frameEnergyTHFacMedian = 1.5f, overallEnergyTHWeight = 1;
    float coarseCutoffTH = 20;
    float thOptIterations = 1.2f;
    bool multiThreading = true;
};

enum ResState { RS_IN = 0, RS_OOB = 1, RS_OUTLIER = 2 };

struct RawResidualJacobian {
    float resF[8];
    float Jpdxi[2][6];
    float Jpdc[2][4];
    float Jpdd[2];
    float JIdx[2][8];
    float JabF[2][8];
    float JIdx2[4];    // (0,0) (0,1) (1,0) (1,1)
    float JabJIdx[4];
    float Jab2[4];
};

struct Residual {
    int point = -1, host = -1, target = -1;  // indices into Window::points / frames
    ResState state_state = RS_OUTLIER;
    double state_energy = 0;
    ResState state_NewState = RS_OUTLIER;
    double state_NewEnergy = 0;
    double state_NewEnergyWithOutlier = 0;
    RawResidualJacobian J;
    bool isNew = true;
    float projectedTo[8][2];
    float centerProjectedTo[3];
    int hostIDX = 0, targetIDX = 0;
    float res_toZeroF[8];
    float JpJdF[8];
    bool isLinearized = false;
    bool isActiveAndIsGoodNEW = false;
    bool isActive() const { return isActiveAndIsGoodNEW; }
    Residual() {
        memset(&J, 0, sizeof(J)); memset(projectedTo, 0, sizeof(projectedTo));
        memset(centerProjectedTo, 0, sizeof(centerProjectedTo));
        memset(res_toZeroF, 0, sizeof(res_toZeroF)); memset(JpJdF, 0, sizeof(JpJdF));
    }
    void resetOOB() {  // Residuals.h:63-67
        state_NewEnergy = state_energy = 0;
        state_NewState = RS_OUTLIER;
        state_state = RS_IN;
    }
};

struct Point {
    int host = -1;
    float u = 0, v = 0;
    bool hasDepthPrior = false;
    float idepth_scaled = 0, idepth_zero_scaled = 0, idepth_zero = 0, idepth = 0;
    float step = 0, step_backup = 0, idepth_backup = 0;
    float nullspaces_scale = 0, idepth_hessian = 0, maxRelBaseline = 0;
    std::vector<int> residuals;  // indices into Window::residuals
    float color[8], weights[8];
    float priorF = 0, deltaF = 0;
    float bdSumF = 0, HdiF = 0;
    float Hdd_accLF = 0, Hcd_accLF[4] = {0, 0, 0, 0}, bd_accLF = 0;
    float Hdd_accAF = 0, Hcd_accAF[4] = {0, 0, 0, 0}, bd_accAF = 0;
    void setIdepth(float id) { idepth = id; idepth_scaled = SCALE_IDEPTH * id; }          // PointHessian.h:29-36
    void setIdepthZero(float id) {                                                          // :47-51
        idepth_zero = id; idepth_zero_scaled = SCALE_IDEPTH * id;
        nullspaces_scale = -(id * 1.001 - id / 1.001) * 500;
    }
};

struct FramePrecalc {  // row-major 3x3 floats
    float PRE_RTll[9], PRE_RTll_0[9], PRE_tTll[3], PRE_tTll_0[3];
    float PRE_KRKiTll[9], PRE_RKiTll[9], PRE_aff_mode[2], PRE_b0_mode, PRE_KtTll[3], distanceLL;
};

struct Calib {
    double value_zero[4] = {0, 0, 0, 0}, value_scaled[4], value[4], step[4] = {0, 0, 0, 0};
    double value_backup[4], value_minus_value_zero[4];
    float value_scaledf[4], value_scaledi[4];
    float fxl() const { return value_scaledf[0]; }
    float fyl() const { return value_scaledf[1]; }
    float cxl() const { return value_scaledf[2]; }
    float cyl() const { return value_scaledf[3]; }
    float fxli() const { return value_scaledi[0]; }
    float fyli() const { return value_scaledi[1]; }
    void setValue(const double v[4]);
    void setValueScaled(const double vs[4]);
};

struct Frame {
    int frameID = 0;   // key-frame id in window order (FrameHessian::frameID)
    int id = 0;        // Frame::id (0 => carries the gauge prior, FrameHessian.h:129)
    int w = 0, h = 0;
    const float *dI = nullptr;  // level-0 (I,dx,dy) AoS, not owned
    float frameEnergyTH = 8 * 8 * patternNum;
    float ab_exposure = 1;
    SE3 worldToCam_evalPT;
    double state[10], state_zero[10], state_scaled[10], step[10], step_backup[10], state_backup[10];
    SE3 PRE_worldToCam, PRE_camToWorld;
    double nullspaces_pose[6][6];   // [row][col]
    double nullspaces_affine[4][2];
    double nullspaces_scale[6];
    std::vector<FramePrecalc> targetPrecalc;
    double prior[8], delta_prior[8], delta[8];
    int idx = 0;

    Frame() {
        memset(state, 0, sizeof(state)); memset(state_zero, 0, sizeof(state_zero));
        memset(state_scaled, 0, sizeof(state_scaled)); memset(step, 0, sizeof(step));
        memset(step_backup, 0, sizeof(step_backup)); memset(state_backup, 0, sizeof(state_backup));
    }
    void aff_g2l(float &a, float &b) const { a = (float) state_scaled[6]; b = (float) state_scaled[7]; }
    void aff_g2l_0(float &a, float &b) const { a = (float) (state_zero[6] * SCALE_A); b = (float) (state_zero[7] * SCALE_B); }
    void setState(const double s[10]);
    void setStateScaled(const double ss[10]);
    void setStateZero(const double sz[10]);
    void getPrior(const Settings &S, double p[10]) const;
    void takeData(const Settings &S);
};

// ---- IndexThreadReduce restated (include/internal/IndexThreadReduce.h:26-170).
// The reference hands [min,max) chunks to whichever of its 6 workers asks first, so its float
// sums are run-to-run non-deterministic. Here chunk c is always executed by worker c % 6, and
// per-chunk stats are added in chunk order, which is one of the orders the reference can produce.
struct ThreadReduce {
    typedef std::function<void(int, int, double *, int)> Fn;  // (min, max, stats[10], tid)
    double stats[10];
    ThreadReduce(bool spawn);
    ~ThreadReduce();
    void reduce(const Fn &fn, int first, int end, int stepSize);
private:
    bool threaded;
    std::thread workers[NUM_THREADS];
    std::mutex mtx;
    std::condition_variable cv_go, cv_done;
    unsigned long generation = 0;
    int n_done = 0;
    bool running = true;
    const Fn *cur = nullptr;
    int cur_first = 0, cur_end = 0, cur_step = 1;
    std::vector<double> chunk_stats;
    void run_tid(int tid);
    void loop(int tid);
};

struct AccumulatedTopHessianSSE {
    int nframes[NUM_THREADS];
    std::vector<AccumulatorApprox> acc[NUM_THREADS];
    int nres[NUM_THREADS];
    AccumulatedTopHessianSSE() { for (int i = 0; i < NUM_THREADS; i++) { nframes[i] = 0; nres[i] = 0; } }
    void setZero(int nFrames, int tid) {
        acc[tid].resize((size_t) nFrames * nFrames);
        for (auto &a : acc[tid]) a.initialize();
        nframes[tid] = nFrames;
        nres[tid] = 0;
    }
};

struct AccumulatedSCHessianSSE {
    int nframes[NUM_THREADS];
    std::vector<AccumulatorXX<8, 4>> accE[NUM_THREADS];
    std::vector<AccumulatorX<8>> accEB[NUM_THREADS];
    std::vector<AccumulatorXX<8, 8>> accD[NUM_THREADS];
    AccumulatorXX<4, 4> accHcc[NUM_THREADS];
    AccumulatorX<4> accbc[NUM_THREADS];
    AccumulatedSCHessianSSE() { for (int i = 0; i < NUM_THREADS; i++) nframes[i] = 0; }
    void setZero(int n, int tid) {
        accE[tid].resize((size_t) n * n); accEB[tid].resize((size_t) n * n); accD[tid].resize((size_t) n * n * n);
        accbc[tid].initialize(); accHcc[tid].initialize();
        for (auto &a : accE[tid]) a.initialize();
        for (auto &a : accEB[tid]) a.initialize();
        for (auto &a : accD[tid]) a.initialize();
        nframes[tid] = n;
    }
};

// The window: FullSystem's frames/points/activeResiduals + EnergyFunctional in one object.
struct Window {
    Settings S;
    int wG0 = 0, hG0 = 0;
    float wM3G = 0, hM3G = 0;
    Calib HCalib;
    std::vector<Frame> frames;
    std::vector<Point> points;          // == ef->allPoints order (sorted by host, makeIDX)
    std::vector<Residual> residuals;
    std::vector<int> activeResiduals;   // indices
    ThreadReduce *red = nullptr;

    // EnergyFunctional state
    int nFrames = 0;
    MatX HM; VecXd bM;
    int resInA = 0, resInL = 0, resInM = 0;
    MatX lastHS; VecXd lastbS, lastX;
    MatX last_HA, last_Hsc, last_HL; VecXd last_bA, last_bsc, last_bL;  // extra taps for parity tests
    std::vector<VecXd> lastNullspaces_pose, lastNullspaces_scale, lastNullspaces_affA, lastNullspaces_affB;
    std::vector<float> adHTdeltaF;        // [nF*nF][8]
    std::vector<double> adHost, adTarget; // [nF*nF][64] row-major 8x8
    std::vector<float> adHostF, adTargetF;
    double cPrior[4]; float cDeltaF[4]; float cPriorF[4];
    AccumulatedTopHessianSSE accSSE_top_L, accSSE_top_A;
    AccumulatedSCHessianSSE accSSE_bot;
    double lastEnergyP = 0;

    Window(int w, int h, int nthreads_mode);
    ~Window();

    // FullSystem pieces restricted to the path
    void setPrecalcValues();                         // FullSystem.cc:1423-1431
    double linearizeAll(bool fixLinearization);      // :1442-1492 (returns lastEnergyP)
    void applyResAll();                              // :1706-1709
    void setNewFrameEnergyTH();                      // :1762-1793
    void backupState();                              // :1662-1676 (non-momentum branch)
    bool doStepFromBackup(float stepfacC, float stepfacT, float stepfacR, float stepfacA, float stepfacD);  // :1587-1622
    void getNullspaces();                            // :1711-1760
    void optimizeBegin();                            // FullSystem.cc:734-771
    bool gnIteration(int iteration);                 // body of the loop :777-831 (forceAcceptStep)

    // PointFrameResidual
    double linearize(Residual &r);                   // Residuals.cc:13-214
    void applyRes(Residual &r, bool copyJacobians);  // Residuals.h:70-87
    void takeData(Residual &r);                      // Residuals.h:123-128
    void fixLinearizationF(Residual &r);             // Residuals.cc:216-242

    // FrameFramePrecalc::Set                         FrameFramePrecalc.cc:6-35
    void precalcSet(FramePrecalc &pc, const Frame &host, const Frame &target);

    // EnergyFunctional
    void insertFrames();                             // insertFrame x nF: takeData + setAdjointsF + makeIDX
    void setAdjointsF();                             // EnergyFunctional.cc:431-489
    void setDeltaF();                                // :403-429
    void makeIDX();                                  // :385-401
    void solveSystemF(int iteration, double lambda); // :240-351
    void accumulateAF_MT(MatX &H, VecXd &b, bool MT);
    void accumulateLF_MT(MatX &H, VecXd &b, bool MT);
    void accumulateSCF_MT(MatX &H, VecXd &b, bool MT);
    void resubstituteF_MT(const VecXd &x, bool MT);  // :491-516
    void resubstituteFPt(const float xc[4], const float *xAd, int min, int max);  // :518-547
    void orthogonalize(VecXd *b, MatX *H);           // :685-717
    VecXd getStitchedDeltaF() const;
    double calcMEnergyF();                           // :353-359
    double calcLEnergyF_MT();                        // :361-378
    void calcLEnergyPt(int min, int max, double *stats, int tid);  // :627-682

    // accumulators
    template<int mode> void topAddPoint(AccumulatedTopHessianSSE &A, Point &p, int tid);  // AccumulatedTopHessian.cc:9-118
    void topStitchDoubleInternal(AccumulatedTopHessianSSE &A, MatX *H, VecXd *b, bool usePrior, int min, int max, int tid, bool hostOuter = false);  // :193-255
    void topStitchDoubleMT(AccumulatedTopHessianSSE &A, MatX &H, VecXd &b, bool usePrior, bool MT);  // .h:64-105
    void topStitchDouble(AccumulatedTopHessianSSE &A, MatX &H, VecXd &b, bool usePrior, int tid = 0);  // .cc:129-191
    void scAddPoint(Point &p, bool shiftPriorToZero, int tid);  // AccumulatedSCHessian.cc:9-51
    void scStitchDoubleInternal(MatX *H, VecXd *b, int min, int max, int tid, bool hostOuter = false);  // :53-119
    void scStitchDoubleMT(MatX &H, VecXd &b, bool MT);  // .h:64-98
    void scStitchDouble(MatX &H, VecXd &b, int tid = 0);  // .cc:121-177

    // marginalisation algebra
    void marginalizePointsF(const std::vector<int> &pointIdx);   // EnergyFunctional.cc:165-222
    void marginalizeFramePrior(int idx);                         // EnergyFunctional.cc:72-129 (HM, bM algebra only)
    // immature-point activation: ImmaturePoint::linearizeResidual (ImmaturePoint.cc:316-383) and
    // FullSystem::optimizeImmaturePoint (FullSystem.cc:892-978). res_state: one entry per frame (255 for the host itself).
    struct ImmatureCand { float u, v, idepth_min, idepth_max, energyTH; float color[8], weights[8]; int host; };
    bool optimizeImmaturePoint(const ImmatureCand &c, int minObs, float &idepth_out, unsigned char *res_state);
};

// projectPoint (pattern) — include/internal/ResidualProjections.h:24-33
inline bool projectPointA(float u_pt, float v_pt, float idepth, const float *KRKi, const float *Kt,
                                 float wM3G, float hM3G, float &Ku, float &Kv) {
    float ptp[3];
    for (int i = 0; i < 3; i++) {
        float s = KRKi[i * 3 + 0] * u_pt;
        s += KRKi[i * 3 + 1] * v_pt;
        s += KRKi[i * 3 + 2] * 1.0f;
        ptp[i] = s + Kt[i] * idepth;
    }
    Ku = ptp[0] / ptp[2];
    Kv = ptp[1] / ptp[2];
    return Ku > 1.1f && Kv > 1.1f && Ku < wM3G && Kv < hM3G;
}
// projectPoint (centre, eval point) — ResidualProjections.h:57-84
inline bool projectPointB(float u_pt, float v_pt, float idepth, int dx, int dy, const Calib &HCalib,
                                 const float *R, const float *t, float wM3G, float hM3G,
                                 float &drescale, float &u, float &v, float &Ku, float &Kv, float KliP[3],
                                 float &new_idepth) {
    KliP[0] = (u_pt + dx - HCalib.cxl()) * HCalib.fxli();
    KliP[1] = (v_pt + dy - HCalib.cyl()) * HCalib.fyli();
    KliP[2] = 1;
    float ptp[3];
    for (int i = 0; i < 3; i++) {
        float s = R[i * 3 + 0] * KliP[0];
        s += R[i * 3 + 1] * KliP[1];
        s += R[i * 3 + 2] * KliP[2];
        ptp[i] = s + t[i] * idepth;
    }
    drescale = 1.0f / ptp[2];
    new_idepth = idepth * drescale;
    if (!(drescale > 0)) return false;
    u = ptp[0] * drescale;
    v = ptp[1] * drescale;
    Ku = u * HCalib.fxl() + HCalib.cxl();
    Kv = v * HCalib.fyl() + HCalib.cyl();
    return Ku > 1.1f && Kv > 1.1f && Ku < wM3G && Kv < hM3G;
}

// derive_idepth — ResidualProjections.h:12-18
inline float derive_idepth(const float t[3], float u, float v, int dx, int dy, float dxInterp, float dyInterp, float drescale) {
    (void) dx; (void) dy;
    return (dxInterp * drescale * (t[0] - t[2] * u) + dyInterp * drescale * (t[1] - t[2] * v)) * SCALE_IDEPTH;
}

// ImmaturePointTemporaryResidual (ImmaturePoint.h:17-27) and ImmaturePoint::linearizeResidual (ImmaturePoint.cc:316-383)
struct TmpRes { int state_state, state_NewState; float state_energy, state_NewEnergy; };
double immatureLinearizeResidual(const Window &W, const Window::ImmatureCand &c, int target, float outlierTHSlack, TmpRes &tr,
                                 float &Hdd, float &bd, float idepth);

// bilinear sampler, GlobalFuncs.h:89-103
inline void getInterpolatedElement33(const float *mat, float x, float y, int width, float out[3]) {
    int ix = (int) x;
    int iy = (int) y;
    float dx = x - ix;
    float dy = y - iy;
    float dxdy = dx * dy;
    const float *bp = mat + 3 * (ix + iy * width);
    const float *p11 = bp + 3 * (1 + width), *p01 = bp + 3 * width, *p10 = bp + 3;
    for (int k = 0; k < 3; k++)
        out[k] = dxdy * p11[k] + (dy - dxdy) * p01[k] + (dx - dxdy) * p10[k] + (1 - dx - dy + dxdy) * bp[k];
}

// GlobalFuncs.h:185-207
inline void getInterpolatedElement33BiLin(const float *mat, float x, float y, int width, float out[3]) {
    int ix = (int) x;
    int iy = (int) y;
    const float *bp = mat + 3 * (ix + iy * width);
    float tl = bp[0], tr = bp[3], bl = bp[3 * width], br = bp[3 * width + 3];
    float dx = x - ix;
    float dy = y - iy;
    float topInt = dx * tr + (1 - dx) * tl;
    float botInt = dx * br + (1 - dx) * bl;
    float leftInt = dy * bl + (1 - dy) * tl;
    float rightInt = dy * br + (1 - dy) * tr;
    out[0] = dx * rightInt + (1 - dx) * leftInt;
    out[1] = rightInt - leftInt;
    out[2] = botInt - topInt;
}

// AffLight::fromToVecExposure, include/AffLight.h:27-35
inline void fromToVecExposure(float exposureF, float exposureT, float aF, float bF, float aT, float bT, double out[2]) {
    if (exposureF == 0 || exposureT == 0) exposureT = exposureF = 1;
    float a = expf(aT - aF) * exposureT / exposureF;  // reference: exp(float) -> double -> float
    float b = bT - a * bF;
    out[0] = a;
    out[1] = b;
}

// FrameHessian::makeImages pyramid + gradients, src/internal/FrameHessian.cc:44-98
// (without the gamma weighting of absSquaredGrad, which feeds pixel selection only).
void makeImages(const float *color, int w, int h, int levels, float **dIp /*out, preallocated w_l*h_l*3*/);

}  // namespace oracle
