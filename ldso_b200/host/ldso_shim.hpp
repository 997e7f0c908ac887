// ldso_shim.hpp — host side of the drop-in boundary, in the reference's own language (C++).
//
// Link-compatible stand-ins for the reference classes that FullSystem drives on this path, with the reference's
// member names, argument meaning and error behaviour (no exceptions: bool / NaN / state flags), forwarding to the
// C ABI of include/ldso_b200.h:
//     ldso::internal::PointFrameResidual   include/internal/Residuals.h:40-130
//     ldso::internal::EnergyFunctional     include/internal/OptimizationBackend/EnergyFunctional.h:54-178
//     ldso::internal::FrameHessian / PointHessian / CalibHessian (the fields the path reads and writes)
//     ldso::CoarseTracker                  include/frontend/CoarseTracker.h:17-127
// Eigen / Sophus / glog are not available in this image, so the few dense types the interfaces expose are tiny
// structs with the same names (Vec8, Vec10, MatXX, VecX, SE3, AffLight); an integrator building inside LDSO replaces
// this block by `#include "NumTypes.h"`.
//
// How an UNCHANGED FullSystem::optimize runs on the GPU: it calls r->linearize(HCalib) per residual from 6 threads
// (FullSystem.cc:1494-1543). The first call after any state change (setState / setIdepth / setValue bump a global
// epoch) uploads the frame states and the window once and launches ONE batched linearize on the device; every other
// call of that generation just returns its cached result. applyRes / solveSystemF are batched the same way.
#pragma once
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/ldso_b200.h"

namespace ldso {

using std::shared_ptr;
using std::weak_ptr;

// ---- the dense types the interfaces mention (NumTypes.h) ------------------------------------------------
const int CPARS = 4;
const int MAX_RES_PER_POINT = 8;
template<int N> struct VecN { double v[N]; double &operator[](int i) { return v[i]; } const double &operator[](int i) const { return v[i]; }
    void setZero() { for (int i = 0; i < N; i++) v[i] = 0; } static VecN Zero() { VecN r; r.setZero(); return r; } };
typedef VecN<3> Vec3; typedef VecN<5> Vec5; typedef VecN<8> Vec8; typedef VecN<10> Vec10; typedef VecN<4> VecC;
template<int N> struct VecNf { float v[N]; float &operator[](int i) { return v[i]; } const float &operator[](int i) const { return v[i]; } };
typedef VecNf<8> Vec8f; typedef VecNf<3> Vec3f; typedef VecNf<2> Vec2f;
struct VecX { std::vector<double> d; int size() const { return (int) d.size(); } double &operator[](int i) { return d[i]; } const double &operator[](int i) const { return d[i]; } };
struct MatXX { int r = 0, c = 0; std::vector<double> d;   // column-major like Eigen
    void resize(int r_, int c_) { r = r_; c = c_; d.assign((size_t) r_ * c_, 0.0); }
    double &operator()(int i, int j) { return d[(size_t) j * r + i]; } double operator()(int i, int j) const { return d[(size_t) j * r + i]; }
    int rows() const { return r; } int cols() const { return c; } };
struct SE3 { double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; double t[3] = {0, 0, 0}; };   // rotationMatrix() row-major, translation()
struct AffLight { float a = 0, b = 0; AffLight() {} AffLight(float a_, float b_) : a(a_), b(b_) {} };

const float SCALE_XI_ROT = 1.0f, SCALE_XI_TRANS = 0.5f, SCALE_F = 50.0f, SCALE_C = 50.0f, SCALE_A = 10.0f, SCALE_B = 1000.0f;

namespace internal {

inline unsigned long &stateEpoch() { static unsigned long e = 1; return e; }   // bumped by every state mutation

enum ResLocation { ACTIVE = 0, LINEARIZED, MARGINALIZED, NONE };
enum ResState { IN = 0, OOB, OUTLIER };

class PointHessian; class FrameHessian; class EnergyFunctional;

struct RawResidualJacobian {    // RawResidualJacobian.h:13-39 (filled on request from the device copy)
    float resF[8], Jpdxi[2][6], Jpdc[2][4], Jpdd[2], JIdx[2][8], JabF[2][8], JIdx2[4], JabJIdx[4], Jab2[4];
};

class CalibHessian {            // CalibHessian.h:16-140
public:
    CalibHessian(double fx, double fy, double cx, double cy) {
        value_zero.setZero();
        VecC v; v[0] = fx; v[1] = fy; v[2] = cx; v[3] = cy;
        setValueScaled(v);
        value_zero = value;
        value_minus_value_zero.setZero();
    }
    float fxl() const { return (float) value_scaled[0]; } float fyl() const { return (float) value_scaled[1]; }
    float cxl() const { return (float) value_scaled[2]; } float cyl() const { return (float) value_scaled[3]; }
    void setValue(const VecC &v) {
        value = v;
        value_scaled[0] = SCALE_F * v[0]; value_scaled[1] = SCALE_F * v[1]; value_scaled[2] = SCALE_C * v[2]; value_scaled[3] = SCALE_C * v[3];
        for (int i = 0; i < 4; i++) value_minus_value_zero[i] = value[i] - value_zero[i];
        stateEpoch()++;
    }
    void setValueScaled(const VecC &vs) {
        value_scaled = vs;
        value[0] = (double) (1.0f / SCALE_F) * vs[0]; value[1] = (double) (1.0f / SCALE_F) * vs[1];
        value[2] = (double) (1.0f / SCALE_C) * vs[2]; value[3] = (double) (1.0f / SCALE_C) * vs[3];
        for (int i = 0; i < 4; i++) value_minus_value_zero[i] = value[i] - value_zero[i];
        stateEpoch()++;
    }
    VecC value_zero, value_scaled, value, step, step_backup, value_backup, value_minus_value_zero;
};

struct Frame { int id = 0; };   // Frame::id == 0 carries the gauge prior (FrameHessian.h:129)

class FrameHessian {            // FrameHessian.h:27-214
public:
    explicit FrameHessian(shared_ptr<Frame> f) : frame(f) { state.setZero(); state_zero.setZero(); state_scaled.setZero(); step.setZero(); state_backup.setZero(); }
    const SE3 &get_worldToCam_evalPT() const { return worldToCam_evalPT; }
    const Vec10 &get_state() const { return state; }
    const Vec10 &get_state_zero() const { return state_zero; }
    AffLight aff_g2l() const { return AffLight((float) state_scaled[6], (float) state_scaled[7]); }
    void setState(const Vec10 &s) {
        state = s;
        for (int i = 0; i < 3; i++) state_scaled[i] = SCALE_XI_TRANS * s[i];
        for (int i = 3; i < 6; i++) state_scaled[i] = SCALE_XI_ROT * s[i];
        state_scaled[6] = SCALE_A * s[6]; state_scaled[7] = SCALE_B * s[7]; state_scaled[8] = SCALE_A * s[8]; state_scaled[9] = SCALE_B * s[9];
        stateEpoch()++;       // PRE_worldToCam / PRE_camToWorld are recomputed on the device (FrameHessian.h:89-90)
    }
    void setStateZero(const Vec10 &sz) { state_zero = sz; stateEpoch()++; }
    void setEvalPT(const SE3 &w2c, const Vec10 &s) { worldToCam_evalPT = w2c; setState(s); setStateZero(s); }
    // dIp[lvl]: (I,dx,dy) Eigen::Vector3f arrays as makeImages leaves them; uploaded once with uploadImages()
    const float *dIp[LDSO_B200_MAX_LEVELS] = {nullptr};
    int frameID = 0;
    shared_ptr<Frame> frame;
    float frameEnergyTH = 8 * 8 * 8;
    float ab_exposure = 1;
    SE3 worldToCam_evalPT;
    Vec10 state, state_zero, state_scaled, step, step_backup, state_backup;
    int idx = 0;
    int imageSlot = -1;       // device slot (ldso_b200_upload_frame)
    std::vector<shared_ptr<PointHessian>> pointHessians;   // ACTIVE points hosted here, in feature order
};

class PointFrameResidual {      // Residuals.h:40-130
public:
    PointFrameResidual(shared_ptr<PointHessian> point_, shared_ptr<FrameHessian> host_, shared_ptr<FrameHessian> target_)
        : point(point_), host(host_), target(target_) { resetOOB(); }
    double linearize(shared_ptr<CalibHessian> &HCalib);      // Residuals.cc:13-214, batched on the device
    void resetOOB() { state_NewEnergy = state_energy = 0; state_NewState = ResState::OUTLIER; state_state = ResState::IN; stateEpoch()++; }
    void applyRes(bool copyJacobians);                        // Residuals.h:70-87
    bool isActive() const { return isActiveAndIsGoodNEW; }
    ResState state_state = ResState::OUTLIER;
    double state_energy = 0;
    ResState state_NewState = ResState::OUTLIER;
    double state_NewEnergy = 0, state_NewEnergyWithOutlier = 0;
    weak_ptr<PointHessian> point; weak_ptr<FrameHessian> host; weak_ptr<FrameHessian> target;
    bool isNew = true;
    float projectedTo[MAX_RES_PER_POINT][2];
    Vec3f centerProjectedTo;
    int hostIDX = 0, targetIDX = 0;
    Vec8f JpJdF;
    bool isLinearized = false, isActiveAndIsGoodNEW = false;
    EnergyFunctional *ef = nullptr;   // set by EnergyFunctional::insertResidual
    int devIndex = -1;                // position in the flattened window of the current generation
};

class PointHessian {            // PointHessian.h:19-132
public:
    void setIdepth(float id) { idepth = id; idepth_scaled = id; stateEpoch()++; }
    void setIdepthZero(float id) { idepth_zero = id; idepth_zero_scaled = id; stateEpoch()++; }
    float u = 0, v = 0;
    bool hasDepthPrior = false;
    float idepth_scaled = 0, idepth_zero_scaled = 0, idepth_zero = 0, idepth = 0, step = 0, step_backup = 0, idepth_backup = 0;
    float idepth_hessian = 0, maxRelBaseline = 0;
    std::vector<shared_ptr<PointFrameResidual>> residuals;
    float color[MAX_RES_PER_POINT], weights[MAX_RES_PER_POINT];
    float priorF = 0, deltaF = 0, bdSumF = 0, HdiF = 0;
    weak_ptr<FrameHessian> hostFrame;
    int devIndex = -1;
};

// ---------------------------------------------------------------------------------------------------------
class EnergyFunctional {        // EnergyFunctional.h:54-178
public:
    // `w,h,levels` as GlobalCalib holds them; the device context is owned here.
    EnergyFunctional(int w, int h, int levels, int device = 0) : W(w), Hh(h), L(levels) {
        ctx = ldso_b200_create(device, w, h, levels, nullptr);
    }
    ~EnergyFunctional() { if (ctx) ldso_b200_destroy(ctx); }
    bool ok() const { return ctx != nullptr; }
    const char *lastError() const { return ldso_b200_last_error(ctx); }
    ldso_b200_ctx *context() { return ctx; }

    void insertResidual(shared_ptr<PointFrameResidual> r) { r->ef = this; nResiduals++; topologyEpoch++; }
    void insertFrame(shared_ptr<FrameHessian> fh, shared_ptr<CalibHessian> Hcalib) {
        frames.push_back(fh); fh->idx = (int) frames.size() - 1; nFrames++;
        const int n = 8 * nFrames + CPARS;
        MatXX HMn; HMn.resize(n, n);
        for (int j = 0; j < HM.c; j++) for (int i = 0; i < HM.r; i++) HMn(i, j) = HM(i, j);
        HM = HMn; bM.d.resize(n, 0.0);
        if (fh->imageSlot < 0) uploadImages(fh);
        topologyEpoch++; stateEpoch()++;
        (void) Hcalib;
    }
    void dropResidual(shared_ptr<PointFrameResidual> r) {
        shared_ptr<PointHessian> p = r->point.lock();
        for (size_t i = 0; i < p->residuals.size(); i++) if (p->residuals[i] == r) { p->residuals.erase(p->residuals.begin() + i); break; }
        nResiduals--; topologyEpoch++;
    }
    void removePoint(shared_ptr<PointHessian> ph) {
        nResiduals -= (int) ph->residuals.size(); ph->residuals.clear();
        shared_ptr<FrameHessian> h = ph->hostFrame.lock();
        if (h) for (size_t i = 0; i < h->pointHessians.size(); i++) if (h->pointHessians[i] == ph) { h->pointHessians.erase(h->pointHessians.begin() + i); break; }
        nPoints--; topologyEpoch++;
    }
    // EnergyFunctional::marginalizeFrame (EnergyFunctional.cc:72-150): the HM/bM algebra runs on the device-resident prior
    // (ldso_b200_marginalize_frame), the frame-list bookkeeping (:131-150) here. Points/residuals of the frame must
    // already have been dropped or marginalised by the caller, as FullSystem::marginalizeFrame does.
    bool marginalizeFrame(shared_ptr<FrameHessian> fh) {
        if (!lastCalib || !syncToDevice(lastCalib)) return false;       // current states, delta_prior and HM on the device
        int nd = 0;
        if (ldso_b200_marginalize_frame(ctx, fh->idx, &nd)) return false;
        HM.resize(nd, nd); bM.d.assign(nd, 0.0);
        if (ldso_b200_get_marg_prior(ctx, HM.d.data(), bM.d.data())) return false;
        for (size_t i = fh->idx; i + 1 < frames.size(); i++) { frames[i] = frames[i + 1]; frames[i]->idx = (int) i; }
        frames.pop_back();
        nFrames--;
        makeIDX();
        stateEpoch()++;
        return true;
    }
    void makeIDX() {            // EnergyFunctional.cc:385-401
        for (size_t i = 0; i < frames.size(); i++) frames[i]->idx = (int) i;
        allPoints.clear();
        for (auto &f : frames) for (auto &p : f->pointHessians) {
            allPoints.push_back(p);
            for (auto &r : p->residuals) { r->hostIDX = r->host.lock()->idx; r->targetIDX = r->target.lock()->idx; }
        }
        topologyEpoch++;
    }
    void setDeltaF(shared_ptr<CalibHessian>) { stateEpoch()++; }     // adHTdeltaF / delta are derived on the device
    void setAdjointsF(shared_ptr<CalibHessian>) { stateEpoch()++; }  // adjoints + null spaces are derived in ldso_b200_set_frames
    // EnergyFunctional::solveSystemF (EnergyFunctional.cc:240-351): fills lastHS, lastbS, lastX, frame/calib/point steps.
    void solveSystemF(int iteration, double lambda, shared_ptr<CalibHessian> HCalib);
    // the fused, device-resident loop (one call instead of FullSystem::optimize's loop body)
    bool optimizeOnDevice(int firstIteration, int nIterations, shared_ptr<CalibHessian> HCalib);

    std::vector<shared_ptr<FrameHessian>> frames;
    int nPoints = 0, nFrames = 0, nResiduals = 0;
    MatXX HM; VecX bM;
    int resInA = 0, resInL = 0, resInM = 0;
    MatXX lastHS; VecX lastbS, lastX;
    double lastEnergy = 0;

    // ---- batching machinery (not part of the reference interface)
    bool syncToDevice(shared_ptr<CalibHessian> &HCalib);      // upload states (+ window when the topology changed)
    bool deviceLinearizeAll(shared_ptr<CalibHessian> &HCalib);
    void uploadImages(shared_ptr<FrameHessian> fh) {
        int slot = 0;
        std::vector<bool> used(2 * LDSO_B200_MAX_FRAMES, false);
        for (auto &f : frames) if (f->imageSlot >= 0) used[f->imageSlot] = true;
        while (used[slot]) slot++;
        if (ldso_b200_upload_frame(ctx, slot, fh->dIp, L) == 0) fh->imageSlot = slot;
    }
    std::vector<shared_ptr<PointHessian>> allPoints;
    shared_ptr<CalibHessian> lastCalib;          // the CalibHessian of the last sync (marginalizeFrame has no such argument)
    std::vector<shared_ptr<PointFrameResidual>> flatResiduals;
    unsigned long topologyEpoch = 1, uploadedTopology = 0, uploadedState = 0, linearizedState = 0;
    bool applyPending = false;
    std::mutex devMutex;
    std::vector<uint8_t> cacheNewState; std::vector<float> cacheNewEnergy, cacheNewEnergyWO, cacheCpt, cacheProj;
private:
    ldso_b200_ctx *ctx = nullptr;
    int W, Hh, L;
};

// ---------------------------------------------------------------------------------------------------------
inline bool EnergyFunctional::syncToDevice(shared_ptr<CalibHessian> &HCalib) {
    if (!ctx) return false;
    lastCalib = HCalib;
    if (uploadedState == stateEpoch() && uploadedTopology == topologyEpoch) return true;
    std::vector<ldso_b200_frame_state> fs(frames.size());
    for (size_t i = 0; i < frames.size(); i++) {
        FrameHessian &f = *frames[i];
        memcpy(fs[i].evalR, f.worldToCam_evalPT.R, 72); memcpy(fs[i].evalT, f.worldToCam_evalPT.t, 24);
        memcpy(fs[i].state_zero, f.state_zero.v, 80); memcpy(fs[i].state, f.state.v, 80);
        fs[i].ab_exposure = f.ab_exposure; fs[i].frameEnergyTH = f.frameEnergyTH; fs[i].frame_id = f.frame->id; fs[i].image_slot = f.imageSlot;
    }
    if (ldso_b200_set_frames(ctx, (int) frames.size(), fs.data(), HCalib->value_scaled.v, HCalib->value_zero.v)) return false;
    // the host mirror owns HM, bM in this shim: push them (or an explicit zero prior) after every set_frames
    if (HM.r == 8 * nFrames + CPARS) ldso_b200_set_marg_prior(ctx, HM.d.data(), bM.d.data());
    else ldso_b200_set_marg_prior(ctx, nullptr, nullptr);
    // flatten allPoints / residuals (EnergyFunctional::makeIDX order)
    const int nP = (int) allPoints.size();
    std::vector<int32_t> host(nP), rb(nP + 1, 0), tgt;
    std::vector<float> u(nP), v(nP), id(nP), idz(nP), col(8 * (size_t) nP), wts(8 * (size_t) nP);
    std::vector<uint8_t> prior(nP), st, lin;
    flatResiduals.clear();
    for (int p = 0; p < nP; p++) {
        PointHessian &P = *allPoints[p];
        P.devIndex = p;
        host[p] = P.hostFrame.lock()->idx; u[p] = P.u; v[p] = P.v; id[p] = P.idepth; idz[p] = P.idepth_zero; prior[p] = P.hasDepthPrior;
        memcpy(&col[8 * (size_t) p], P.color, 32); memcpy(&wts[8 * (size_t) p], P.weights, 32);
        for (auto &r : P.residuals) {
            r->devIndex = (int) flatResiduals.size();
            flatResiduals.push_back(r);
            tgt.push_back(r->targetIDX); st.push_back((uint8_t) r->state_state); lin.push_back(r->isLinearized);
        }
        rb[p + 1] = (int) flatResiduals.size();
    }
    ldso_b200_window w;
    memset(&w, 0, sizeof(w));
    w.nPoints = nP; w.nResiduals = (int) flatResiduals.size();
    w.pt_host = host.data(); w.pt_u = u.data(); w.pt_v = v.data(); w.pt_idepth = id.data(); w.pt_idepth_zero = idz.data();
    w.pt_has_prior = prior.data(); w.pt_color = col.data(); w.pt_weights = wts.data(); w.res_begin = rb.data(); w.res_target = tgt.data();
    w.res_state = st.data(); w.res_is_linearized = lin.data();
    if (ldso_b200_set_window(ctx, &w)) return false;
    uploadedState = stateEpoch(); uploadedTopology = topologyEpoch;
    linearizedState = 0;
    return true;
}

inline bool EnergyFunctional::deviceLinearizeAll(shared_ptr<CalibHessian> &HCalib) {
    std::lock_guard<std::mutex> lock(devMutex);     // FullSystem::linearizeAll_Reductor calls in from 6 threads
    if (linearizedState == stateEpoch() && uploadedTopology == topologyEpoch) return true;
    if (!syncToDevice(HCalib)) return false;
    if (ldso_b200_linearize_all(ctx, 0, 1, &lastEnergy)) return false;
    const size_t nR = flatResiduals.size();
    cacheNewState.resize(nR); cacheNewEnergy.resize(nR); cacheNewEnergyWO.resize(nR); cacheCpt.resize(3 * nR); cacheProj.resize(16 * nR);
    if (ldso_b200_get_residuals(ctx, nullptr, cacheNewState.data(), nullptr, cacheNewEnergy.data(), cacheNewEnergyWO.data(), nullptr, nullptr,
                                nullptr, cacheProj.data(), cacheCpt.data())) return false;
    // setNewFrameEnergyTH ran on the device: mirror the newest frame's threshold (FullSystem.cc:1762-1793)
    std::vector<float> th(frames.size());
    ldso_b200_get_frames(ctx, nullptr, nullptr, th.data(), nullptr, nullptr, nullptr, nullptr, nullptr);
    if (!frames.empty()) frames.back()->frameEnergyTH = th.back();
    linearizedState = stateEpoch();
    applyPending = false;
    return true;
}

inline double PointFrameResidual::linearize(shared_ptr<CalibHessian> &HCalib) {
    state_NewEnergyWithOutlier = -1;
    if (state_state == ResState::OOB) { state_NewState = ResState::OOB; return state_energy; }
    if (!ef || !ef->deviceLinearizeAll(HCalib) || devIndex < 0) { state_NewState = ResState::OOB; return NAN; }   // isLost path (FullSystem.cc:845-849)
    state_NewState = (ResState) ef->cacheNewState[devIndex];
    state_NewEnergyWithOutlier = ef->cacheNewEnergyWO[devIndex];
    for (int i = 0; i < 3; i++) centerProjectedTo[i] = ef->cacheCpt[3 * devIndex + i];
    memcpy(projectedTo, &ef->cacheProj[16 * (size_t) devIndex], 64);
    if (state_NewState == ResState::OOB) return state_energy;
    state_NewEnergy = ef->cacheNewEnergy[devIndex];
    return state_NewEnergy;
}

inline void PointFrameResidual::applyRes(bool copyJacobians) {
    if (copyJacobians) {
        if (state_state == ResState::OOB) return;
        isActiveAndIsGoodNEW = (state_NewState == ResState::IN);     // takeData happens on the device (JpJdF stays there)
    }
    state_state = state_NewState;
    state_energy = state_NewEnergy;
    if (ef) ef->applyPending = true;
}

inline void EnergyFunctional::solveSystemF(int iteration, double lambda, shared_ptr<CalibHessian> HCalib) {
    (void) lambda;      // SOLVER_FIX_LAMBDA: the reference overwrites it with 1e-5 (EnergyFunctional.cc:243)
    std::lock_guard<std::mutex> lock(devMutex);
    const int n = 8 * nFrames + CPARS;
    lastHS.resize(n, n); lastbS.d.assign(n, 0.0); lastX.d.assign(n, NAN);
    if (!ctx || linearizedState != stateEpoch()) return;             // nothing linearized at this state: x stays NaN (caller's isLost)
    if (applyPending) { ldso_b200_apply_res(ctx); applyPending = false; }
    if (ldso_b200_backup_state(ctx) || ldso_b200_solve_system(ctx, iteration, lastHS.d.data(), lastbS.d.data(), lastX.d.data())) return;
    ldso_b200_get_system(ctx, nullptr, nullptr, nullptr, nullptr, &resInA);
    // resubstituteF_MT: frame / calib / point steps back into the host objects (EnergyFunctional.cc:491-547)
    for (int i = 0; i < 4; i++) HCalib->step[i] = -lastX[i];
    for (auto &f : frames) { for (int i = 0; i < 8; i++) f->step[i] = -lastX[CPARS + 8 * f->idx + i]; f->step[8] = f->step[9] = 0; }
    const size_t nP = allPoints.size();
    std::vector<float> step(nP), HdiF(nP), bdSumF(nP), Hdd(nP);
    ldso_b200_get_points(ctx, nullptr, nullptr, step.data(), HdiF.data(), bdSumF.data(), Hdd.data(), nullptr, nullptr);
    for (size_t p = 0; p < nP; p++) {
        PointHessian &P = *allPoints[p];
        P.step = step[p]; P.HdiF = HdiF[p]; P.bdSumF = bdSumF[p];
        P.idepth_hessian = (HdiF[p] > 0) ? 1.0f / HdiF[p] : 0;
    }
}

inline bool EnergyFunctional::optimizeOnDevice(int firstIteration, int nIterations, shared_ptr<CalibHessian> HCalib) {
    std::lock_guard<std::mutex> lock(devMutex);
    if (!syncToDevice(HCalib)) return false;
    if (ldso_b200_optimize_begin(ctx, &lastEnergy)) return false;
    if (ldso_b200_gn_iterations(ctx, firstIteration, nIterations)) return false;
    // FullSystem::optimize ends with linearizeAll(true) (FullSystem.cc:843): fixes the states and leaves the projections
    // (centerProjectedTo / projectedTo) that makeCoarseDepthL0 and LoopClosing read
    if (ldso_b200_linearize_all(ctx, 1, 1, &lastEnergy)) return false;
    const int n = 8 * nFrames + CPARS;
    lastHS.resize(n, n); lastbS.d.assign(n, 0.0); lastX.d.assign(n, 0.0);
    if (ldso_b200_get_last_solution(ctx, lastHS.d.data(), lastbS.d.data(), lastX.d.data())) return false;
    int cb = 0;
    ldso_b200_get_energy(ctx, &lastEnergy, &cb);
    // mirror the optimised state back (what doStepFromBackup leaves in the host objects)
    std::vector<double> st(10 * frames.size()); std::vector<float> th(frames.size()); double cv[4];
    ldso_b200_get_frames(ctx, st.data(), nullptr, th.data(), nullptr, nullptr, nullptr, nullptr, cv);
    for (size_t i = 0; i < frames.size(); i++) { Vec10 s; memcpy(s.v, &st[10 * i], 80); frames[i]->setState(s); frames[i]->frameEnergyTH = th[i]; }
    VecC v; memcpy(v.v, cv, 32); HCalib->setValue(v);
    const size_t nP = allPoints.size(), nR = flatResiduals.size();
    std::vector<float> id(nP), step(nP), HdiF(nP);
    ldso_b200_get_points(ctx, id.data(), nullptr, step.data(), HdiF.data(), nullptr, nullptr, nullptr, nullptr);
    for (size_t p = 0; p < nP; p++) { allPoints[p]->setIdepth(id[p]); allPoints[p]->setIdepthZero(id[p]); allPoints[p]->step = step[p]; allPoints[p]->HdiF = HdiF[p]; }
    std::vector<uint8_t> ss(nR), act(nR); std::vector<float> en(nR), cpt(3 * nR), proj(16 * nR);
    ldso_b200_get_residuals(ctx, ss.data(), nullptr, en.data(), nullptr, nullptr, act.data(), nullptr, nullptr, proj.data(), cpt.data());
    for (size_t r = 0; r < nR; r++) {
        PointFrameResidual &R = *flatResiduals[r];
        R.state_state = (ResState) ss[r]; R.state_energy = en[r]; R.isActiveAndIsGoodNEW = act[r];
        for (int i = 0; i < 3; i++) R.centerProjectedTo[i] = cpt[3 * r + i];
        memcpy(R.projectedTo, &proj[16 * r], 64);
    }
    uploadedState = stateEpoch();      // host and device agree again
    return true;
}

}  // namespace internal

// ---------------------------------------------------------------------------------------------------------
class CoarseTracker {           // include/frontend/CoarseTracker.h:17-127
public:
    CoarseTracker(int w_, int h_, int levels, int device = 0) : levelsUsed(levels) {
        ctx = ldso_b200_create(device, w_, h_, levels, nullptr);
        for (int i = 0; i < 5; i++) lastResiduals[i] = NAN;
        for (int i = 0; i < 3; i++) lastFlowIndicators[i] = 1000;
    }
    ~CoarseTracker() { if (ctx) ldso_b200_destroy(ctx); }
    void makeK(shared_ptr<internal::CalibHessian> HCalib) { ldso_b200_tracker_make_k(ctx, HCalib->fxl(), HCalib->fyl(), HCalib->cxl(), HCalib->cyl()); }
    // setCoarseTrackingRef (CoarseTracker.cc:248-256): lastRef = frameHessians.back(); makeCoarseDepthL0 on the device from
    // every ACTIVE point whose newest residual targets lastRef and is IN (centerProjectedTo, HdiF).
    void setCoarseTrackingRef(std::vector<shared_ptr<internal::FrameHessian>> &frameHessians) {
        lastRef = frameHessians.back();
        std::vector<float> cpt, hdi;
        for (auto &fh : frameHessians) for (auto &ph : fh->pointHessians) for (auto &r : ph->residuals)
            if (r->target.lock() == lastRef && r->state_state == internal::ResState::IN && r->isActive()) {
                for (int i = 0; i < 3; i++) cpt.push_back(r->centerProjectedTo[i]);
                hdi.push_back(ph->HdiF);
            }
        if (refSlot < 0) refSlot = 0;
        ldso_b200_upload_frame(ctx, refSlot, lastRef->dIp, levelsUsed);
        ldso_b200_tracker_make_coarse_depth(ctx, refSlot, (int) hdi.size(), cpt.data(), hdi.data());
        lastRef_aff_g2l = lastRef->aff_g2l();
        refFrameID = lastRef->frameID;
        firstCoarseRMSE = -1;
    }
    // trackNewestCoarse (CoarseTracker.cc:61-217): true if tracking is good; lastToNew_out / aff_g2l_out updated in place.
    bool trackNewestCoarse(shared_ptr<internal::FrameHessian> newFrameHessian, SE3 &lastToNew_out, AffLight &aff_g2l_out, int coarsestLvl,
                           Vec5 minResForAbort) {
        newFrame = newFrameHessian;
        const int slot = 1;
        if (ldso_b200_upload_frame(ctx, slot, newFrame->dIp, levelsUsed)) return false;
        ldso_b200_tracker_set_frames(ctx, lastRef_aff_g2l.a, lastRef_aff_g2l.b, lastRef->ab_exposure, slot, newFrame->ab_exposure);
        int ok = 0;
        if (ldso_b200_tracker_track(ctx, lastToNew_out.R, lastToNew_out.t, &aff_g2l_out.a, &aff_g2l_out.b, coarsestLvl, minResForAbort.v,
                                    lastResiduals.v, lastFlowIndicators.v, &ok)) return false;
        return ok != 0;
    }
    shared_ptr<internal::FrameHessian> lastRef, newFrame;
    AffLight lastRef_aff_g2l;
    int refFrameID = -1;
    Vec5 lastResiduals;
    Vec3 lastFlowIndicators;
    double firstCoarseRMSE = 0;
private:
    ldso_b200_ctx *ctx = nullptr;
    int levelsUsed, refSlot = -1;
};

}  // namespace ldso
