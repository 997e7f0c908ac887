// ORACLE — TEST INFRASTRUCTURE ONLY (see omath.h header). PARITY PINNED AGAINST THE REFERENCE'S OWN SOURCES, EXCEPT EIGEN / SOPHUS INTERNALS (oracle/ref_pin compiles, unmodified, the reference's Residuals.cc, ImmaturePoint.cc, PointHessian.cc, FrameHessian.cc, FrameFramePrecalc.cc, AccumulatedTopHessian.cc, AccumulatedSCHessian.cc, EnergyFunctional.cc, CoarseTracker.cc, Setting.cc and their headers against stand-in Eigen/Sophus headers and matches the oracle bit for bit on 108 checks; the float SSE / scalar code is pinned completely, the Eigen-expression code structurally - what is added where, in which order - with LDLT, PartialPivLU, JacobiSVD, SE3 exp/log and the product kernels being the oracle's own restatements on both sides; FullSystem.cc's driver loop needs the whole front end and stays restated from the cited lines).
// Flat extern "C" surface over the oracle so tests/ (ctypes), __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs can drive it. Nothing in ldso_b200/ may load this.
#include "ba.h"
#include "tracker.h"
#include "trace.h"
#include "initializer.h"
#include <chrono>

using namespace oracle;

extern "C" {

// ---------------------------------------------------------------------------- BA window
void *oracle_ba_create(int w, int h, int threads_mode) { return new Window(w, h, threads_mode); }
void oracle_ba_destroy(void *o) { delete (Window *) o; }

void oracle_ba_set_calib(void *o, const double value_scaled[4]) {
    Window *W = (Window *) o;
    W->HCalib.setValueScaled(value_scaled);
    for (int i = 0; i < 4; i++) W->HCalib.value_zero[i] = W->HCalib.value[i];  // CalibHessian ctor :29-31
    W->HCalib.setValueScaled(value_scaled);
}
// shift the calibration state away from its linearisation point (value = value_zero + delta)
void oracle_ba_set_calib_delta(void *o, const double delta[4]) {
    Window *W = (Window *) o;
    double v[4];
    for (int i = 0; i < 4; i++) v[i] = W->HCalib.value_zero[i] + delta[i];
    W->HCalib.setValue(v);
}

int oracle_ba_add_frame(void *o, const double R[9], const double t[3], const double state_zero[10],
                        const double state[10], float ab_exposure, int frame_id, const float *dI) {
    Window *W = (Window *) o;
    Frame f;
    f.w = W->wG0;
    f.h = W->hG0;
    f.dI = dI;
    f.ab_exposure = ab_exposure;
    f.id = frame_id;
    f.frameID = (int) W->frames.size();
    f.worldToCam_evalPT = SE3::fromRt(R, t);
    // setEvalPT(worldToCam_evalPT, state_zero) then setState(state)  (FrameHessian.h:107-112)
    f.setState(state_zero);
    f.setStateZero(state_zero);
    f.setState(state);
    W->frames.push_back(f);
    return (int) W->frames.size() - 1;
}

int oracle_ba_add_point(void *o, int host, float u, float v, float idepth_zero, float idepth, int hasDepthPrior,
                        const float color[8], const float weights[8]) {
    Window *W = (Window *) o;
    Point p;
    p.host = host;
    p.u = u;
    p.v = v;
    p.hasDepthPrior = hasDepthPrior != 0;
    p.setIdepthZero(idepth_zero);
    p.setIdepth(idepth);
    for (int i = 0; i < 8; i++) { p.color[i] = color[i]; p.weights[i] = weights[i]; }
    // PointHessian::takeData (PointHessian.h:112-117)
    p.priorF = p.hasDepthPrior ? W->S.idepthFixPrior * SCALE_IDEPTH * SCALE_IDEPTH : 0;
    p.deltaF = p.idepth - p.idepth_zero;
    W->points.push_back(p);
    return (int) W->points.size() - 1;
}

int oracle_ba_add_residual(void *o, int point, int target) {
    Window *W = (Window *) o;
    Residual r;
    r.point = point;
    r.host = W->points[point].host;
    r.target = target;
    r.resetOOB();
    W->residuals.push_back(r);
    W->points[point].residuals.push_back((int) W->residuals.size() - 1);
    return (int) W->residuals.size() - 1;
}

void oracle_ba_set_frame_energy_th(void *o, int frame, float th) { ((Window *) o)->frames[frame].frameEnergyTH = th; }

void oracle_ba_finalize(void *o) {
    Window *W = (Window *) o;
    W->insertFrames();
    W->setPrecalcValues();
}

void oracle_ba_set_marg_prior(void *o, const double *HM_colmajor, const double *bM) {
    Window *W = (Window *) o;
    int n = 8 * W->nFrames + CPARS;
    W->HM = MatX(n, n);
    W->bM.assign(n, 0.0);
    for (int i = 0; i < n * n; i++) W->HM.d[i] = HM_colmajor[i];
    for (int i = 0; i < n; i++) W->bM[i] = bM[i];
}

double oracle_ba_optimize_begin(void *o) {
    Window *W = (Window *) o;
    W->optimizeBegin();
    return W->lastEnergyP;
}
int oracle_ba_gn_iteration(void *o, int iteration) { return ((Window *) o)->gnIteration(iteration) ? 1 : 0; }

// finer-grained steps
double oracle_ba_linearize_all(void *o, int fix) { return ((Window *) o)->linearizeAll(fix != 0); }
void oracle_ba_apply_res(void *o) { ((Window *) o)->applyResAll(); }
void oracle_ba_solve_system(void *o, int iteration) {
    Window *W = (Window *) o;
    W->backupState();
    W->getNullspaces();
    W->solveSystemF(iteration, 1e-1);
}
int oracle_ba_do_step(void *o) { return ((Window *) o)->doStepFromBackup(1, 1, 1, 1, 1) ? 1 : 0; }
double oracle_ba_last_energy(void *o) { return ((Window *) o)->lastEnergyP; }
double oracle_ba_calc_m_energy(void *o) { return ((Window *) o)->calcMEnergyF(); }
double oracle_ba_calc_l_energy(void *o) { return ((Window *) o)->calcLEnergyF_MT(); }

// fixLinearizationF on every active residual of the listed points, then marginalizePointsF on them
void oracle_ba_marginalize_points(void *o, int n, const int *pointIdx, float priorFac) {
    Window *W = (Window *) o;
    std::vector<int> idx(pointIdx, pointIdx + n);
    for (int pi : idx) {
        Point &p = W->points[pi];
        for (int ri : p.residuals) {
            Residual &r = W->residuals[ri];
            r.resetOOB();
            W->linearize(r);
            r.isLinearized = false;
            W->applyRes(r, true);
            if (r.isActive()) W->fixLinearizationF(r);
        }
        p.priorF *= priorFac;
    }
    W->marginalizePointsF(idx);
}

// EnergyFunctional::marginalizeFrame's prior algebra for frame idx; HM, bM shrink to 8(nF-1)+4 (get_marg_prior then
// returns that size). Returns the new dimension.
int oracle_ba_marginalize_frame(void *o, int idx) {
    Window *W = (Window *) o;
    W->marginalizeFramePrior(idx);
    return W->HM.r;
}

int oracle_ba_dims(void *o, int *nFrames, int *nPoints, int *nResiduals) {
    Window *W = (Window *) o;
    *nFrames = (int) W->frames.size();
    *nPoints = (int) W->points.size();
    *nResiduals = (int) W->residuals.size();
    return 0;
}

static void copy_mat(const MatX &M, double *out) { if (out && M.r > 0) memcpy(out, M.d.data(), sizeof(double) * M.d.size()); }
static void copy_vec(const VecXd &v, double *out) { if (out && !v.empty()) memcpy(out, v.data(), sizeof(double) * v.size()); }

// all matrices column-major (8nF+4)^2, vectors (8nF+4)
void oracle_ba_get_system(void *o, double *HA, double *bA, double *Hsc, double *bsc, double *lastHS, double *lastbS,
                          double *lastX, double *HL, double *bL) {
    Window *W = (Window *) o;
    copy_mat(W->last_HA, HA); copy_vec(W->last_bA, bA);
    copy_mat(W->last_Hsc, Hsc); copy_vec(W->last_bsc, bsc);
    copy_mat(W->lastHS, lastHS); copy_vec(W->lastbS, lastbS); copy_vec(W->lastX, lastX);
    copy_mat(W->last_HL, HL); copy_vec(W->last_bL, bL);
}
void oracle_ba_get_marg_prior(void *o, double *HM, double *bM) {
    Window *W = (Window *) o;
    copy_mat(W->HM, HM);
    copy_vec(W->bM, bM);
}
int oracle_ba_res_counts(void *o, int *resInA, int *resInL, int *resInM) {
    Window *W = (Window *) o;
    *resInA = W->resInA; *resInL = W->resInL; *resInM = W->resInM;
    return 0;
}

void oracle_ba_get_points(void *o, float *idepth, float *idepth_zero, float *step, float *HdiF, float *bdSumF,
                          float *Hdd_accAF, float *bd_accAF, float *Hcd_accAF /*4 per*/, float *deltaF) {
    Window *W = (Window *) o;
    for (size_t i = 0; i < W->points.size(); i++) {
        const Point &p = W->points[i];
        if (idepth) idepth[i] = p.idepth;
        if (idepth_zero) idepth_zero[i] = p.idepth_zero;
        if (step) step[i] = p.step;
        if (HdiF) HdiF[i] = p.HdiF;
        if (bdSumF) bdSumF[i] = p.bdSumF;
        if (Hdd_accAF) Hdd_accAF[i] = p.Hdd_accAF;
        if (bd_accAF) bd_accAF[i] = p.bd_accAF;
        if (Hcd_accAF) for (int k = 0; k < 4; k++) Hcd_accAF[4 * i + k] = p.Hcd_accAF[k];
        if (deltaF) deltaF[i] = p.deltaF;
    }
}

// J layout per residual (74 floats): resF[8] Jpdxi[12] Jpdc[8] Jpdd[2] JIdx[16] JabF[16] JIdx2[4] JabJIdx[4] Jab2[4]
void oracle_ba_get_residuals(void *o, int *state_state, int *state_NewState, double *state_energy,
                             double *state_NewEnergy, double *state_NewEnergyWithOutlier, float *J74,
                             float *JpJdF, float *projectedTo /*16*/, float *centerProjectedTo /*3*/,
                             unsigned char *isActive, unsigned char *isLinearized, float *res_toZeroF) {
    Window *W = (Window *) o;
    for (size_t i = 0; i < W->residuals.size(); i++) {
        const Residual &r = W->residuals[i];
        if (state_state) state_state[i] = r.state_state;
        if (state_NewState) state_NewState[i] = r.state_NewState;
        if (state_energy) state_energy[i] = r.state_energy;
        if (state_NewEnergy) state_NewEnergy[i] = r.state_NewEnergy;
        if (state_NewEnergyWithOutlier) state_NewEnergyWithOutlier[i] = r.state_NewEnergyWithOutlier;
        if (J74) {
            float *d = J74 + 74 * i;
            const RawResidualJacobian &J = r.J;
            memcpy(d, J.resF, 32); d += 8;
            memcpy(d, J.Jpdxi, 48); d += 12;
            memcpy(d, J.Jpdc, 32); d += 8;
            memcpy(d, J.Jpdd, 8); d += 2;
            memcpy(d, J.JIdx, 64); d += 16;
            memcpy(d, J.JabF, 64); d += 16;
            memcpy(d, J.JIdx2, 16); d += 4;
            memcpy(d, J.JabJIdx, 16); d += 4;
            memcpy(d, J.Jab2, 16);
        }
        if (JpJdF) memcpy(JpJdF + 8 * i, r.JpJdF, 32);
        if (projectedTo) memcpy(projectedTo + 16 * i, r.projectedTo, 64);
        if (centerProjectedTo) memcpy(centerProjectedTo + 3 * i, r.centerProjectedTo, 12);
        if (isActive) isActive[i] = r.isActiveAndIsGoodNEW ? 1 : 0;
        if (isLinearized) isLinearized[i] = r.isLinearized ? 1 : 0;
        if (res_toZeroF) memcpy(res_toZeroF + 8 * i, r.res_toZeroF, 32);
    }
}

// per-frame: state[10], frameEnergyTH; per pair (h + nF*t): precalc 40 floats
//   [RTll_0(9) tTll_0(3) RTll(9) tTll(3) KRKiTll(9) KtTll(3) aff(2) b0(1) distanceLL(1)],
//   adHost/adTarget 64 doubles row-major, adHTdeltaF 8 floats
void oracle_ba_get_frames(void *o, double *state, double *step, float *frameEnergyTH, float *precalc40, double *adHost,
                          double *adTarget, float *adHTdeltaF, double *calib_value, double *prior8, double *delta_prior8,
                          double *delta8) {
    Window *W = (Window *) o;
    int nF = (int) W->frames.size();
    for (int h = 0; h < nF; h++) {
        const Frame &f = W->frames[h];
        if (state) memcpy(state + 10 * h, f.state, 80);
        if (step) memcpy(step + 10 * h, f.step, 80);
        if (frameEnergyTH) frameEnergyTH[h] = f.frameEnergyTH;
        if (prior8) memcpy(prior8 + 8 * h, f.prior, 64);
        if (delta_prior8) memcpy(delta_prior8 + 8 * h, f.delta_prior, 64);
        if (delta8) memcpy(delta8 + 8 * h, f.delta, 64);
        if (precalc40)
            for (int t = 0; t < nF; t++) {
                const FramePrecalc &pc = f.targetPrecalc[t];
                float *d = precalc40 + 40 * (h + nF * t);
                memcpy(d, pc.PRE_RTll_0, 36); d += 9;
                memcpy(d, pc.PRE_tTll_0, 12); d += 3;
                memcpy(d, pc.PRE_RTll, 36); d += 9;
                memcpy(d, pc.PRE_tTll, 12); d += 3;
                memcpy(d, pc.PRE_KRKiTll, 36); d += 9;
                memcpy(d, pc.PRE_KtTll, 12); d += 3;
                d[0] = pc.PRE_aff_mode[0]; d[1] = pc.PRE_aff_mode[1]; d[2] = pc.PRE_b0_mode; d[3] = pc.distanceLL;
            }
    }
    if (adHost) memcpy(adHost, W->adHost.data(), sizeof(double) * W->adHost.size());
    if (adTarget) memcpy(adTarget, W->adTarget.data(), sizeof(double) * W->adTarget.size());
    if (adHTdeltaF) memcpy(adHTdeltaF, W->adHTdeltaF.data(), sizeof(float) * W->adHTdeltaF.size());
    if (calib_value) memcpy(calib_value, W->HCalib.value, 32);
}

// nullspace projector NNpiTS (n x n col-major) that orthogonalize() applies: x -= P x
void oracle_ba_get_nullspace_projector(void *o, double *P) {
    Window *W = (Window *) o;
    W->getNullspaces();
    int n = 8 * W->nFrames + CPARS;
    for (int c = 0; c < n; c++) {
        VecXd e(n, 0.0);
        e[c] = 1.0;
        VecXd x = e;
        W->orthogonalize(&x, 0);
        for (int r = 0; r < n; r++) P[(size_t) c * n + r] = e[r] - x[r];
    }
}

// time `iters` GN iterations (after `warmup`), returns seconds per iteration (median)
double oracle_ba_time_gn(void *o, int iters, int warmup) {
    Window *W = (Window *) o;
    for (int i = 0; i < warmup; i++) W->gnIteration(3);
    std::vector<double> ts;
    for (int i = 0; i < iters; i++) {
        auto t0 = std::chrono::steady_clock::now();
        W->gnIteration(3);
        auto t1 = std::chrono::steady_clock::now();
        ts.push_back(std::chrono::duration<double>(t1 - t0).count());
    }
    std::sort(ts.begin(), ts.end());
    return ts.empty() ? 0.0 : ts[ts.size() / 2];
}

// ---------------------------------------------------------------------------- misc primitives
void oracle_make_images(const float *color, int w, int h, int levels, float **dIp) { makeImages(color, w, h, levels, dIp); }

void oracle_se3_exp(const double a[6], double R[9], double t[3]) {
    SE3 T = SE3::exp(a);
    M3 m = T.rotationMatrix();
    memcpy(R, m.m, 72);
    for (int i = 0; i < 3; i++) t[i] = T.t[i];
}
void oracle_se3_log(const double R[9], const double t[3], double a[6]) { SE3::fromRt(R, t).log(a); }

void oracle_ldlt_solve(int n, const double *A_colmajor, const double *b, double *x) {
    MatX A(n, n);
    memcpy(A.d.data(), A_colmajor, sizeof(double) * n * n);
    VecXd bb(b, b + n);
    VecXd xx = ldlt_solve(A, bb);
    memcpy(x, xx.data(), sizeof(double) * n);
}

void oracle_sample33(const float *dI, int width, float x, float y, float out[3]) { getInterpolatedElement33(dI, x, y, width, out); }
void oracle_sample33_bilin(const float *dI, int width, float x, float y, float out[3]) { getInterpolatedElement33BiLin(dI, x, y, width, out); }

// ---------------------------------------------------------------------------- coarse tracker
void *oracle_tracker_create(int w, int h, int levels) { return new CoarseTracker(w, h, levels); }
void oracle_tracker_destroy(void *o) { delete (CoarseTracker *) o; }
void oracle_tracker_make_k(void *o, float fx, float fy, float cx, float cy) { ((CoarseTracker *) o)->makeK(fx, fy, cx, cy); }
void oracle_tracker_set_ref(void *o, const float **refDIp, float aff_a, float aff_b, float ab_exposure, int n,
                            const float *cpt, const float *HdiF) {
    CoarseTracker *T = (CoarseTracker *) o;
    for (int l = 0; l < T->pyrLevelsUsed; l++) T->refDIp[l] = refDIp[l];
    T->lastRef_aff_a = aff_a;
    T->lastRef_aff_b = aff_b;
    T->lastRef_ab_exposure = ab_exposure;
    T->makeCoarseDepthL0(n, cpt, HdiF);
}
void oracle_tracker_set_new_frame(void *o, const float **newDIp, float ab_exposure) {
    CoarseTracker *T = (CoarseTracker *) o;
    for (int l = 0; l < T->pyrLevelsUsed; l++) T->newDIp[l] = newDIp[l];
    T->newFrame_ab_exposure = ab_exposure;
}
int oracle_tracker_pc_n(void *o, int lvl) { return ((CoarseTracker *) o)->pc_n[lvl]; }
void oracle_tracker_get_pc(void *o, int lvl, float *u, float *v, float *idepth, float *color) {
    CoarseTracker *T = (CoarseTracker *) o;
    int n = T->pc_n[lvl];
    memcpy(u, T->pc_u[lvl].data(), 4 * n);
    memcpy(v, T->pc_v[lvl].data(), 4 * n);
    memcpy(idepth, T->pc_idepth[lvl].data(), 4 * n);
    memcpy(color, T->pc_color[lvl].data(), 4 * n);
}
// one calcRes (+ calcGSSSE) evaluation; H row-major 8x8
void oracle_tracker_eval(void *o, int lvl, const double R[9], const double t[3], float aff_a, float aff_b,
                         float cutoffTH, double res6[6], double H[64], double b[8]) {
    CoarseTracker *T = (CoarseTracker *) o;
    SE3 P = SE3::fromRt(R, t);
    T->calcRes(lvl, P, aff_a, aff_b, cutoffTH, res6);
    if (H && b) T->calcGSSSE(lvl, H, b, P, aff_a, aff_b);
}
int oracle_tracker_track(void *o, double R[9], double t[3], float *aff_a, float *aff_b, int coarsestLvl,
                         const double minResForAbort[5], double lastResiduals[5], double lastFlowIndicators[3],
                         int *n_evals) {
    CoarseTracker *T = (CoarseTracker *) o;
    SE3 P = SE3::fromRt(R, t);
    bool ok = T->trackNewestCoarse(P, *aff_a, *aff_b, coarsestLvl, minResForAbort);
    M3 m = P.rotationMatrix();
    memcpy(R, m.m, 72);
    for (int i = 0; i < 3; i++) t[i] = P.t[i];
    memcpy(lastResiduals, T->lastResiduals, 40);
    memcpy(lastFlowIndicators, T->lastFlowIndicators, 24);
    if (n_evals) *n_evals = T->lm_iterations_total;
    return ok ? 1 : 0;
}

// ---- immature-point trace (ImmaturePoint ctor + traceOn, FullSystem::traceNewCoarse's loop) ----------------------------
// candidate construction on the host keyframe: colour[8], weights[8], gradH[4], energyTH per point
void oracle_trace_init(const float *dI_host, int w, int n, const float *u, const float *v, float *color8, float *weights8,
                       float *gradH4, float *energyTH) {
    TraceSettings S;
    for (int i = 0; i < n; i++) {
        ImmaturePt p;
        immature_init(p, dI_host, w, u[i], v[i], S);
        memcpy(color8 + 8 * i, p.color, 32); memcpy(weights8 + 8 * i, p.weights, 32); memcpy(gradH4 + 4 * i, p.gradH, 16);
        energyTH[i] = p.energyTH;
    }
}
// one traceNewCoarse pass over n candidates; KRKi9/Kt3/aff2 are per host keyframe (row-major 3x3). In/out: idepth_min,
// idepth_max, quality, status; out: lastTraceUV[2], lastTracePixelInterval.
void oracle_trace_on(const float *dI, int w, int h, int n, const float *u, const float *v, const float *color8, const float *weights8,
                     const float *gradH4, const float *energyTH, const int *host, const float *KRKi9, const float *Kt3,
                     const float *aff2, float *idepth_min, float *idepth_max, float *quality, int *status, float *uv2, float *interval) {
    TraceSettings S;
    for (int i = 0; i < n; i++) {
        ImmaturePt p;
        p.u = u[i]; p.v = v[i];
        memcpy(p.color, color8 + 8 * i, 32); memcpy(p.weights, weights8 + 8 * i, 32); memcpy(p.gradH, gradH4 + 4 * i, 16);
        p.energyTH = energyTH[i]; p.quality = quality[i]; p.idepth_min = idepth_min[i]; p.idepth_max = idepth_max[i];
        p.lastTraceStatus = status[i];
        p.lastTraceUV[0] = uv2[2 * i]; p.lastTraceUV[1] = uv2[2 * i + 1]; p.lastTracePixelInterval = interval[i];
        const int hh = host[i];
        trace_on(p, dI, w, h, KRKi9 + 9 * hh, Kt3 + 3 * hh, aff2 + 2 * hh, S);
        idepth_min[i] = p.idepth_min; idepth_max[i] = p.idepth_max; quality[i] = p.quality; status[i] = p.lastTraceStatus;
        uv2[2 * i] = p.lastTraceUV[0]; uv2[2 * i + 1] = p.lastTraceUV[1]; interval[i] = p.lastTracePixelInterval;
    }
}

// FullSystem::optimizeImmaturePoint for n candidates against the window's frames (current states). ok[n], idepth[n],
// res_state[n * nFrames] (ResState per target frame, 255 for the host).
void oracle_ba_optimize_immature(void *o, int n, const float *u, const float *v, const int *host, const float *idepth_min,
                                 const float *idepth_max, const float *color8, const float *weights8, const float *energyTH, int minObs,
                                 int *ok, float *idepth, unsigned char *res_state) {
    Window *W = (Window *) o;
    const int nF = (int) W->frames.size();
    for (int i = 0; i < n; i++) {
        Window::ImmatureCand c;
        c.u = u[i]; c.v = v[i]; c.host = host[i]; c.idepth_min = idepth_min[i]; c.idepth_max = idepth_max[i]; c.energyTH = energyTH[i];
        memcpy(c.color, color8 + 8 * i, 32); memcpy(c.weights, weights8 + 8 * i, 32);
        ok[i] = W->optimizeImmaturePoint(c, minObs, idepth[i], res_state + (size_t) i * nF) ? 1 : 0;
    }
}

// FullSystem::activatePointsMT's selection over the window's frames and points: distance map of the ACTIVE points (all points of the
// window whose host is not `newest`) in the newest keyframe, then the greedy pass over the candidates in the order given.
// action[n] (0 stays, 1 activate, 2 delete), dist_map[(w/2)*(h/2)] (optional).
void oracle_ba_select_activation(void *o, int levels, int newest, float currentMinActDist, float minTraceQuality, int n, const float *u, const float *v,
                                 const int *host, const float *idepth_min, const float *idepth_max, const int *status, const float *interval,
                                 const float *quality, const float *my_type, const unsigned char *flagged, unsigned char *action, float *dist_map) {
    Window *W = (Window *) o;
    const int nF = (int) W->frames.size();
    CoarseDistanceMap M(W->wG0, W->hG0, levels);
    M.makeK(W->HCalib.fxl(), W->HCalib.fyl(), W->HCalib.cxl(), W->HCalib.cyl());
    std::vector<float> R(9 * nF), T(3 * nF);
    for (int f = 0; f < nF; f++) {       // fhToNew = newest.PRE_worldToCam * host.PRE_camToWorld, cast to float (CoarseTracker.cc:704-706)
        SE3 fhToNew = W->frames[newest].PRE_worldToCam * W->frames[f].PRE_camToWorld;
        M3 Rd = fhToNew.rotationMatrix();
        for (int i = 0; i < 9; i++) R[9 * f + i] = (float) Rd.m[i];
        for (int i = 0; i < 3; i++) T[3 * f + i] = (float) fhToNew.t[i];
    }
    M.beginDistanceMap();
    for (int f = 0; f < nF; f++) {
        if (f == newest) continue;
        std::vector<float> pu, pv, pid;
        for (const Point &p : W->points) if (p.host == f) { pu.push_back(p.u); pv.push_back(p.v); pid.push_back(p.idepth_scaled); }
        M.addFramePoints(&R[9 * f], &T[3 * f], (int) pu.size(), pu.data(), pv.data(), pid.data());
    }
    M.finishDistanceMap();
    for (int i = 0; i < n; i++) {        // one call per candidate keeps the caller's order whatever the host sequence is
        ActivationCand c{u[i], v[i], idepth_min[i], idepth_max[i], quality[i], interval[i], my_type[i], status[i]};
        selectActivation(M, &R[9 * host[i]], &T[3 * host[i]], flagged[host[i]] != 0, currentMinActDist, minTraceQuality, 1, &c, action + i);
    }
    if (dist_map) memcpy(dist_map, M.fwdWarpedIDDistFinal.data(), sizeof(float) * (size_t) M.w[1] * M.h[1]);
}

// CoarseInitializer::calcResAndGS for the n points of pyramid level `lvl` (firstDIp / newDIp: `levels` pointers to (I, dx, dy) AoS images).
// Per-point outputs like the reference's Pnt fields / JbBuffer_new; H, b, Hsc, bsc row-major 8x8 / 8; res3 = (E.A, alphaEnergy, E.num).
void oracle_init_calc_res(int w, int h, int levels, const double K4[4], const float **firstDIp, const float **newDIp, int lvl, const double R[9],
                          const double t[3], float aff_a, float aff_b, int n, const float *u, const float *v, const float *idepth_new, const float *iR,
                          const unsigned char *isGood, const float *energy2, const float *outlierTH, float alphaK, float alphaW, float couplingWeight,
                          unsigned char *isGood_new, float *energy_new2, float *maxstep, float *lastHessian_new, float *Jb10, float *H64, float *b8,
                          float *Hsc64, float *bsc8, float *res3) {
    CoarseInitializer I(w, h, levels);
    I.makeK((float) K4[0], (float) K4[1], (float) K4[2], (float) K4[3]);
    for (int l = 0; l < levels; l++) { I.firstDIp[l] = firstDIp[l]; I.newDIp[l] = newDIp[l]; }
    I.alphaK = alphaK; I.alphaW = alphaW; I.couplingWeight = couplingWeight;
    I.points[lvl].assign(n, InitPnt());
    for (int i = 0; i < n; i++) {
        InitPnt &p = I.points[lvl][i];
        p.u = u[i]; p.v = v[i]; p.idepth_new = idepth_new[i]; p.iR = iR[i]; p.isGood = isGood[i] != 0;
        p.energy[0] = energy2[2 * i]; p.energy[1] = energy2[2 * i + 1]; p.outlierTH = outlierTH[i];
    }
    SE3 P = SE3::fromRt(R, t);
    I.calcResAndGS(lvl, H64, b8, Hsc64, bsc8, P, aff_a, aff_b, res3);
    for (int i = 0; i < n; i++) {
        const InitPnt &p = I.points[lvl][i];
        isGood_new[i] = p.isGood_new ? 1 : 0; energy_new2[2 * i] = p.energy_new[0]; energy_new2[2 * i + 1] = p.energy_new[1];
        maxstep[i] = p.maxstep; lastHessian_new[i] = p.lastHessian_new;
        for (int k = 0; k < 10; k++) Jb10[10 * i + k] = I.JbBuffer_new[i][k];
    }
}

}  // extern "C"
