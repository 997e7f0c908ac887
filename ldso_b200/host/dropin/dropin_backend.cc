// Drop-in replacement for FOUR of the reference's translation units, written against the reference's OWN headers:
//     src/internal/OptimizationBackend/EnergyFunctional.cc      (class EnergyFunctional,  include/internal/OptimizationBackend/EnergyFunctional.h:54-259)
//     src/internal/Residuals.cc                                  (PointFrameResidual::linearize / fixLinearizationF, include/internal/Residuals.h:40-130)
//     src/internal/OptimizationBackend/AccumulatedTopHessian.cc  (AccumulatedTopHessianSSE::addPoint<mode> / stitchDouble*, AccumulatedTopHessian.h:20-125)
//     src/internal/OptimizationBackend/AccumulatedSCHessian.cc   (AccumulatedSCHessianSSE::addPoint / stitchDouble*,        AccumulatedSCHessian.h:17-118)
// Same class declarations, same member signatures, same public data (frames, nPoints / nFrames / nResiduals, HM, bM, resInA/L/M,
// lastHS, lastbS, lastX, lastNullspaces_*, red, connectivityMap): an unchanged FullSystem.cc links against this file instead of the
// four above. Every arithmetic member forwards to the C ABI of include/ldso_b200.h (libldso_b200.so, CUDA, no CPU fallback);
// what stays here is the bookkeeping the reference also does on the host (frame / point / residual lists, connectivity map, indices).
//
// Batching: FullSystem drives the back end one object at a time from 6 threads (r->linearize(HCalib) per residual,
// FullSystem.cc:1494-1543). The first call after a state change uploads frames + window once and runs ONE batched device pass;
// the other calls of that generation return their cached slice. "State changed" is what the reference itself signals:
// FullSystem::setPrecalcValues ends every state change with ef->setDeltaF (FullSystem.cc:1423-1431), topology changes go through
// insertResidual / dropResidual / insertFrame / removePoint / marginalizeFrame / makeIDX.
//
// Build (only where the reference tree and an Eigen are available; in this repository: oracle/Makefile target `dropin`, against
// the stand-in Eigen of oracle/ref_shim): g++ -I<LDSO>/include ... dropin_backend.cc dropin_tracker.cc -lldso_b200
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "Feature.h"
#include "Point.h"
#include "internal/GlobalCalib.h"
#include "internal/GlobalFuncs.h"
#include "internal/OptimizationBackend/EnergyFunctional.h"
#include "internal/Residuals.h"

#include "../../../include/ldso_b200.h"

namespace ldso {
namespace internal {

bool EFAdjointsValid = false;
bool EFIndicesValid = false;
bool EFDeltaValid = false;
int PointFrameResidual::instanceCounter = 0;

namespace {
// Device-side companion of one EnergyFunctional (the class declaration is the reference's and cannot grow members).
struct B200Side {
    ldso_b200_ctx *ctx = nullptr;
    unsigned long stateEpoch = 1, topoEpoch = 1, uploadedState = 0, uploadedTopo = 0, linearizedState = 0;
    bool applyPending = false;
    std::vector<PointFrameResidual *> flat;                  // residual order of the uploaded window
    std::map<const PointFrameResidual *, int> devIndex;
    std::map<const FrameHessian *, int> imageSlot;
    std::vector<uint8_t> newState;
    std::vector<float> newEnergy, newEnergyWO, J, proj, cpt;
    double lastEnergy = 0;
    std::mutex mtx;
};
std::mutex g_regMutex;
std::map<const EnergyFunctional *, B200Side *> g_side;
EnergyFunctional *g_activeEF = nullptr;                      // LDSO runs one FullSystem, hence one EnergyFunctional, per process

B200Side *sideOf(const EnergyFunctional *ef) {
    std::lock_guard<std::mutex> l(g_regMutex);
    auto it = g_side.find(ef);
    return it == g_side.end() ? nullptr : it->second;
}
uint64_t pairKey(const shared_ptr<PointFrameResidual> &r) {
    return (((uint64_t) r->host.lock()->frameID) << 32) + ((uint64_t) r->target.lock()->frameID);
}
}  // namespace

// -------------------------------------------------------------------------------------------------------------------------------
// Helper with access to EnergyFunctional's private members (the reference declares `friend class AccumulatedTopHessian`).
class AccumulatedTopHessian {
public:
    static std::vector<shared_ptr<PointHessian>> &allPoints(EnergyFunctional *ef) { return ef->allPoints; }
    static bool ensureContext(EnergyFunctional *ef, B200Side *S) {
        if (S->ctx) return true;
        S->ctx = ldso_b200_create(0, wG[0], hG[0], pyrLevelsUsed, nullptr);
        return S->ctx != nullptr;
    }
    // upload the frame states (+ HM, bM) and, when anything moved, the window: EnergyFunctional::makeIDX order
    static bool sync(EnergyFunctional *ef, B200Side *S, shared_ptr<CalibHessian> &HCalib) {
        if (!ensureContext(ef, S)) return false;
        if (S->uploadedState == S->stateEpoch && S->uploadedTopo == S->topoEpoch) return true;
        const int nF = (int) ef->frames.size();
        std::vector<ldso_b200_frame_state> fs(nF);
        for (int i = 0; i < nF; i++) {
            FrameHessian &f = *ef->frames[i];
            if (!S->imageSlot.count(&f)) {      // FrameHessian::dIp as makeImages left it: one upload per keyframe
                std::vector<bool> used(2 * LDSO_B200_MAX_FRAMES, false);
                for (auto &fr : ef->frames) { auto it = S->imageSlot.find(fr.get()); if (it != S->imageSlot.end()) used[it->second] = true; }
                int slot = 0;
                while (used[slot]) slot++;
                const float *lv[LDSO_B200_MAX_LEVELS] = {nullptr};
                for (int l = 0; l < pyrLevelsUsed; l++) lv[l] = (const float *) (l == 0 && f.dIp[0] == nullptr ? f.dI : f.dIp[l]);
                if (ldso_b200_upload_frame(S->ctx, slot, lv, pyrLevelsUsed)) return false;
                S->imageSlot[&f] = slot;
            }
            const SE3 &T = f.get_worldToCam_evalPT();
            const Mat33 R = T.rotationMatrix();
            for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) fs[i].evalR[r * 3 + c] = R(r, c); fs[i].evalT[r] = T.translation()[r]; }
            for (int k = 0; k < 10; k++) { fs[i].state_zero[k] = f.get_state_zero()[k]; fs[i].state[k] = f.get_state()[k]; }
            fs[i].ab_exposure = f.ab_exposure; fs[i].frameEnergyTH = f.frameEnergyTH; fs[i].frame_id = (int) f.frame->id;
            fs[i].image_slot = S->imageSlot[&f];
        }
        double vs[4], vz[4];
        for (int k = 0; k < 4; k++) { vs[k] = HCalib->value_scaled[k]; vz[k] = HCalib->value_zero[k]; }
        if (ldso_b200_set_frames(S->ctx, nF, fs.data(), vs, vz)) return false;
        const int n = 8 * nF + CPARS;
        if (ef->HM.rows() == n) {
            std::vector<double> hm((size_t) n * n), bm(n);
            for (int j = 0; j < n; j++) { bm[j] = ef->bM[j]; for (int i = 0; i < n; i++) hm[(size_t) j * n + i] = ef->HM(i, j); }
            if (ldso_b200_set_marg_prior(S->ctx, hm.data(), bm.data())) return false;
        } else if (ldso_b200_set_marg_prior(S->ctx, nullptr, nullptr)) return false;
        auto &pts = ef->allPoints;
        const int nP = (int) pts.size();
        std::vector<int32_t> host(nP), rb(nP + 1, 0), tgt;
        std::vector<float> u(nP), v(nP), id(nP), idz(nP), col(8 * (size_t) nP), wts(8 * (size_t) nP);
        std::vector<uint8_t> prior(nP), st, lin;
        S->flat.clear(); S->devIndex.clear();
        for (int p = 0; p < nP; p++) {
            PointHessian &P = *pts[p];
            host[p] = P.point->mHostFeature.lock()->host.lock()->frameHessian->idx;
            u[p] = P.u; v[p] = P.v; id[p] = P.idepth; idz[p] = P.idepth_zero; prior[p] = P.hasDepthPrior;
            memcpy(&col[8 * (size_t) p], P.color, 32); memcpy(&wts[8 * (size_t) p], P.weights, 32);
            for (auto &r : P.residuals) {
                S->devIndex[r.get()] = (int) S->flat.size();
                S->flat.push_back(r.get());
                tgt.push_back(r->target.lock()->idx); st.push_back((uint8_t) r->state_state); lin.push_back(r->isLinearized);
            }
            rb[p + 1] = (int) S->flat.size();
        }
        ldso_b200_window w;
        memset(&w, 0, sizeof(w));
        w.nPoints = nP; w.nResiduals = (int) S->flat.size();
        w.pt_host = host.data(); w.pt_u = u.data(); w.pt_v = v.data(); w.pt_idepth = id.data(); w.pt_idepth_zero = idz.data();
        w.pt_has_prior = prior.data(); w.pt_color = col.data(); w.pt_weights = wts.data(); w.res_begin = rb.data(); w.res_target = tgt.data();
        w.res_state = st.data(); w.res_is_linearized = lin.data();
        if (ldso_b200_set_window(S->ctx, &w)) return false;
        S->uploadedState = S->stateEpoch; S->uploadedTopo = S->topoEpoch;
        S->linearizedState = 0;
        return true;
    }
    // ONE batched PointFrameResidual::linearize over the window; results cached per residual
    static bool linearizeAll(EnergyFunctional *ef, B200Side *S, shared_ptr<CalibHessian> &HCalib) {
        std::lock_guard<std::mutex> l(S->mtx);     // FullSystem::linearizeAll_Reductor calls in from 6 threads
        if (S->linearizedState == S->stateEpoch && S->uploadedTopo == S->topoEpoch && S->linearizedState != 0) return true;
        if (!sync(ef, S, HCalib)) return false;
        if (ldso_b200_linearize_all(S->ctx, 0, 1, &S->lastEnergy)) return false;
        const size_t nR = S->flat.size();
        S->newState.resize(nR); S->newEnergy.resize(nR); S->newEnergyWO.resize(nR); S->J.resize(74 * nR); S->proj.resize(16 * nR); S->cpt.resize(3 * nR);
        if (ldso_b200_get_residuals(S->ctx, nullptr, S->newState.data(), nullptr, S->newEnergy.data(), S->newEnergyWO.data(), nullptr, nullptr,
                                    S->J.data(), S->proj.data(), S->cpt.data())) return false;
        S->linearizedState = S->stateEpoch;
        S->applyPending = false;
        return true;
    }
};

// -------------------------------------------------------------------------------------------------------------------------------
// PointFrameResidual (Residuals.cc:13-242)
double PointFrameResidual::linearize(shared_ptr<CalibHessian> &HCalib) {
    state_NewEnergyWithOutlier = -1;
    if (state_state == ResState::OOB) { state_NewState = ResState::OOB; return state_energy; }
    EnergyFunctional *ef = g_activeEF;
    B200Side *S = ef ? sideOf(ef) : nullptr;
    if (!S || !AccumulatedTopHessian::linearizeAll(ef, S, HCalib)) { state_NewState = ResState::OOB; return NAN; }   // FullSystem's isLost path (:845-849)
    auto it = S->devIndex.find(this);
    if (it == S->devIndex.end()) { state_NewState = ResState::OOB; return state_energy; }
    const size_t k = (size_t) it->second;
    state_NewState = (ResState) S->newState[k];
    state_NewEnergyWithOutlier = S->newEnergyWO[k];
    for (int i = 0; i < 3; i++) centerProjectedTo[i] = S->cpt[3 * k + i];
    for (int i = 0; i < MAX_RES_PER_POINT; i++) projectedTo[i] = Eigen::Vector2f(S->proj[16 * k + 2 * i], S->proj[16 * k + 2 * i + 1]);
    if (state_NewState == ResState::OOB) return state_energy;
    // RawResidualJacobian (RawResidualJacobian.h:13-39): the header-inline applyRes / takeData read it on the host
    const float *j = &S->J[74 * k];
    for (int i = 0; i < 8; i++) { J->resF[i] = j[i]; J->JIdx[0][i] = j[30 + i]; J->JIdx[1][i] = j[38 + i]; J->JabF[0][i] = j[46 + i]; J->JabF[1][i] = j[54 + i]; }
    for (int i = 0; i < 6; i++) { J->Jpdxi[0][i] = j[8 + i]; J->Jpdxi[1][i] = j[14 + i]; }
    for (int i = 0; i < 4; i++) { J->Jpdc[0][i] = j[20 + i]; J->Jpdc[1][i] = j[24 + i]; }
    J->Jpdd[0] = j[28]; J->Jpdd[1] = j[29];
    J->JIdx2(0, 0) = j[62]; J->JIdx2(0, 1) = j[63]; J->JIdx2(1, 0) = j[64]; J->JIdx2(1, 1) = j[65];
    J->JabJIdx(0, 0) = j[66]; J->JabJIdx(0, 1) = j[67]; J->JabJIdx(1, 0) = j[68]; J->JabJIdx(1, 1) = j[69];
    J->Jab2(0, 0) = j[70]; J->Jab2(0, 1) = j[71]; J->Jab2(1, 0) = j[72]; J->Jab2(1, 1) = j[73];
    state_NewEnergy = S->newEnergy[k];
    if (S) S->applyPending = true;                // the caller's applyRes(true) follows; the device applies its copy before the next solve
    return state_NewEnergy;
}

// Residuals.cc:216-242. The linearised residual (res_toZeroF) lives on the device: ldso_b200_marginalize_points redoes
// flagPointsForRemoval's resetOOB / linearize / applyRes / fixLinearizationF sequence for the points it marginalises. Here the flag.
void PointFrameResidual::fixLinearizationF(shared_ptr<EnergyFunctional> ef) {
    (void) ef;
    isLinearized = true;
}

// -------------------------------------------------------------------------------------------------------------------------------
// EnergyFunctional (EnergyFunctional.cc)
EnergyFunctional::EnergyFunctional() : accSSE_top_L(new AccumulatedTopHessianSSE), accSSE_top_A(new AccumulatedTopHessianSSE), accSSE_bot(new AccumulatedSCHessianSSE) {
    std::lock_guard<std::mutex> l(g_regMutex);
    g_side[this] = new B200Side();
    g_activeEF = this;
}
EnergyFunctional::~EnergyFunctional() {
    std::lock_guard<std::mutex> l(g_regMutex);
    auto it = g_side.find(this);
    if (it != g_side.end()) { if (it->second->ctx) ldso_b200_destroy(it->second->ctx); delete it->second; g_side.erase(it); }
    if (g_activeEF == this) g_activeEF = nullptr;
}

void EnergyFunctional::insertResidual(shared_ptr<PointFrameResidual> r) {       // :26-30
    r->takeData();
    connectivityMap[pairKey(r)][0]++;
    nResiduals++;
    sideOf(this)->topoEpoch++;
}

void EnergyFunctional::insertFrame(shared_ptr<FrameHessian> fh, shared_ptr<CalibHessian> Hcalib) {      // :32-61
    fh->takeData();
    frames.push_back(fh);
    fh->idx = (int) frames.size();
    nFrames++;
    const int n = 8 * nFrames + CPARS;
    bM.conservativeResize(n);
    HM.conservativeResize(n, n);
    bM.tail<8>().setZero();
    HM.rightCols<8>().setZero();
    HM.bottomRows<8>().setZero();
    EFIndicesValid = false; EFAdjointsValid = false; EFDeltaValid = false;
    setAdjointsF(Hcalib);
    makeIDX();
    for (auto fh2 : frames) {
        connectivityMap[(((uint64_t) fh->frameID) << 32) + ((uint64_t) fh2->frameID)] = Eigen::Vector2i(0, 0);
        if (fh2 != fh) connectivityMap[(((uint64_t) fh2->frameID) << 32) + ((uint64_t) fh->frameID)] = Eigen::Vector2i(0, 0);
    }
}

void EnergyFunctional::dropResidual(shared_ptr<PointFrameResidual> r) {         // :63-71
    shared_ptr<PointHessian> p = r->point.lock();
    deleteOut<PointFrameResidual>(p->residuals, r);
    connectivityMap[pairKey(r)][0]--;
    nResiduals--;
    sideOf(this)->topoEpoch++;
}

// :72-150: the prior algebra (move the frame's block to the end, add its prior, scale, 8x8 inverse, Schur complement, unscale,
// symmetrise) runs on the device-resident HM, bM (k_marginalize_frame); the frame list bookkeeping here.
void EnergyFunctional::marginalizeFrame(shared_ptr<FrameHessian> fh) {
    B200Side *S = sideOf(this);
    {
        std::lock_guard<std::mutex> l(S->mtx);
        // delta_prior of the frame enters the algebra: the device needs the current states
        S->stateEpoch++;
        shared_ptr<CalibHessian> none;
        if (S->ctx) {
            int nd = 0;
            // (states were uploaded by the last sync of this generation; HM, bM are pushed again from the host members)
            const int n = 8 * nFrames + CPARS;
            std::vector<double> hm((size_t) n * n), bm(n);
            for (int j = 0; j < n; j++) { bm[j] = bM[j]; for (int i = 0; i < n; i++) hm[(size_t) j * n + i] = HM(i, j); }
            ldso_b200_set_marg_prior(S->ctx, hm.data(), bm.data());
            if (ldso_b200_marginalize_frame(S->ctx, fh->idx, &nd) == 0) {
                std::vector<double> h2((size_t) nd * nd), b2(nd);
                ldso_b200_get_marg_prior(S->ctx, h2.data(), b2.data());
                HM = MatXX::Zero(nd, nd); bM = VecX::Zero(nd);
                for (int j = 0; j < nd; j++) { bM[j] = b2[j]; for (int i = 0; i < nd; i++) HM(i, j) = h2[(size_t) j * nd + i]; }
            }
        }
        S->imageSlot.erase(fh.get());
    }
    for (unsigned int i = fh->idx; i + 1 < frames.size(); i++) { frames[i] = frames[i + 1]; frames[i]->idx = i; }
    frames.pop_back();
    nFrames--;
    EFIndicesValid = false; EFAdjointsValid = false; EFDeltaValid = false;
    makeIDX();
}

void EnergyFunctional::removePoint(shared_ptr<PointHessian> ph) {                // :152-163
    for (auto &r : ph->residuals) { connectivityMap[pairKey(r)][0]--; nResiduals--; }
    ph->residuals.clear();
    if (!ph->alreadyRemoved) nPoints--;
    EFIndicesValid = false;
    sideOf(this)->topoEpoch++;
}

// :165-222: addPoint<2> + Schur addPoint(p, false) over the MARGINALIZED points, stitch without priors, HM += w (M - Msc): one
// device call (ldso_b200_marginalize_points); the points are then removed like the reference does.
void EnergyFunctional::marginalizePointsF() {
    B200Side *S = sideOf(this);
    allPointsToMarg.clear();
    for (auto f : frames)
        for (shared_ptr<Feature> feat : f->frame->features)
            if (feat->status == Feature::FeatureStatus::VALID && feat->point->status == Point::PointStatus::MARGINALIZED) {
                shared_ptr<PointHessian> p = feat->point->mpPH;
                p->priorF *= setting_idepthFixPriorMargFac;
                for (auto r : p->residuals) if (r->isActive()) connectivityMap[pairKey(r)][1]++;
                allPointsToMarg.push_back(p);
            }
    {
        std::lock_guard<std::mutex> l(S->mtx);
        // indices of the points in the uploaded window (they are still ACTIVE there: FullSystem flags them between two syncs)
        std::vector<int32_t> idx;
        for (size_t p = 0; p < allPoints.size(); p++)
            for (auto &q : allPointsToMarg) if (q == allPoints[p]) { idx.push_back((int32_t) p); break; }
        if (S->ctx && !idx.empty()) {
            int rm = 0;
            if (ldso_b200_marginalize_points(S->ctx, (int) idx.size(), idx.data(), setting_idepthFixPriorMargFac, &rm) == 0) {
                resInM += rm;
                const int n = 8 * nFrames + CPARS;
                std::vector<double> h2((size_t) n * n), b2(n);
                ldso_b200_get_marg_prior(S->ctx, h2.data(), b2.data());
                for (int j = 0; j < n; j++) { bM[j] = b2[j]; for (int i = 0; i < n; i++) HM(i, j) = h2[(size_t) j * n + i]; }
            }
        }
    }
    for (auto p : allPointsToMarg) removePoint(p);
    EFIndicesValid = false;
    makeIDX();
}

void EnergyFunctional::dropPointsF() {                                          // :224-238
    for (auto f : frames)
        for (shared_ptr<Feature> feat : f->frame->features)
            if (feat->point && (feat->point->status == Point::PointStatus::OUTLIER || feat->point->status == Point::PointStatus::OUT) &&
                feat->point->mpPH->alreadyRemoved == false)
                removePoint(feat->point->mpPH);
    EFIndicesValid = false;
    makeIDX();
}

// :240-351 (default solver mode): accumulate A / L / Schur, stitch, damp, scale, LDLT, orthogonalise, resubstitute -- one device call
void EnergyFunctional::solveSystemF(int iteration, double lambda, shared_ptr<CalibHessian> HCalib) {
    (void) lambda;      // SOLVER_FIX_LAMBDA: the reference overwrites it with 1e-5 (:243)
    B200Side *S = sideOf(this);
    std::lock_guard<std::mutex> l(S->mtx);
    const int n = 8 * nFrames + CPARS;
    lastHS = MatXX::Zero(n, n); lastbS = VecX::Zero(n); lastX = VecX::Constant(n, NAN);
    if (!S->ctx || S->linearizedState != S->stateEpoch) return;       // nothing linearised at this state: x stays NaN (the caller's isLost)
    if (S->applyPending) { ldso_b200_apply_res(S->ctx); S->applyPending = false; }
    std::vector<double> hs((size_t) n * n), bs(n), x(n);
    if (ldso_b200_backup_state(S->ctx) || ldso_b200_solve_system(S->ctx, iteration, hs.data(), bs.data(), x.data())) return;
    for (int j = 0; j < n; j++) { lastbS[j] = bs[j]; lastX[j] = x[j]; for (int i = 0; i < n; i++) lastHS(i, j) = hs[(size_t) j * n + i]; }
    ldso_b200_get_system(S->ctx, nullptr, nullptr, nullptr, nullptr, &resInA);
    currentLambda = 1e-5f;
    // resubstituteF_MT (:491-516): frame / calibration / point steps back into the host objects
    for (int i = 0; i < 4; i++) HCalib->step[i] = -lastX[i];
    for (auto &f : frames) { for (int i = 0; i < 8; i++) f->step[i] = -lastX[CPARS + 8 * f->idx + i]; f->step[8] = f->step[9] = 0; }
    const size_t nP = allPoints.size();
    std::vector<float> step(nP), HdiF(nP), bdSumF(nP), Hdd(nP), bd(nP), Hcd(4 * nP);
    ldso_b200_get_points(S->ctx, nullptr, nullptr, step.data(), HdiF.data(), bdSumF.data(), Hdd.data(), bd.data(), Hcd.data());
    for (size_t p = 0; p < nP; p++) {
        PointHessian &P = *allPoints[p];
        P.step = step[p]; P.HdiF = HdiF[p]; P.bdSumF = bdSumF[p]; P.Hdd_accAF = Hdd[p]; P.bd_accAF = bd[p];
        for (int k = 0; k < 4; k++) P.Hcd_accAF[k] = Hcd[4 * p + k];
        P.idepth_hessian = (HdiF[p] > 0) ? 1.0f / HdiF[p] : 0;       // AccumulatedSCHessian.cc:24-26
    }
}

double EnergyFunctional::calcMEnergyF() {                                       // :353-359
    B200Side *S = sideOf(this);
    std::lock_guard<std::mutex> l(S->mtx);
    double el = 0, em = 0;
    if (!S->ctx || ldso_b200_calc_energies(S->ctx, &el, &em)) return NAN;
    return em;
}
double EnergyFunctional::calcLEnergyF_MT() {                                    // :361-378
    B200Side *S = sideOf(this);
    std::lock_guard<std::mutex> l(S->mtx);
    double el = 0, em = 0;
    if (!S->ctx || ldso_b200_calc_energies(S->ctx, &el, &em)) return NAN;
    return el;
}

void EnergyFunctional::makeIDX() {                                              // :385-401
    for (unsigned int idx = 0; idx < frames.size(); idx++) frames[idx]->idx = idx;
    allPoints.clear();
    for (auto f : frames)
        for (shared_ptr<Feature> feat : f->frame->features)
            if (feat->status == Feature::FeatureStatus::VALID && feat->point->status == Point::PointStatus::ACTIVE) {
                shared_ptr<PointHessian> p = feat->point->mpPH;
                allPoints.push_back(p);
                for (auto &r : p->residuals) { r->hostIDX = r->host.lock()->idx; r->targetIDX = r->target.lock()->idx; }
            }
    EFIndicesValid = true;
    sideOf(this)->topoEpoch++;
}

// :403-429 / :431-489: delta, adHTdeltaF, the adjoints and the null spaces are derived on the device from the uploaded states
// (ldso_b200_set_frames); the calls mark the device copy stale. setDeltaF is what FullSystem::setPrecalcValues ends every state
// change with (FullSystem.cc:1423-1431).
void EnergyFunctional::setDeltaF(shared_ptr<CalibHessian> HCalib) {
    cDeltaF = HCalib->value_minus_value_zero.cast<float>();
    for (auto &f : frames) { f->delta = f->get_state_minus_stateZero().head<8>(); f->delta_prior = (f->get_state() - f->getPriorZero()).head<8>(); }
    for (auto &p : allPoints) p->deltaF = p->idepth - p->idepth_zero;
    EFDeltaValid = true;
    sideOf(this)->stateEpoch++;
}
void EnergyFunctional::setAdjointsF(shared_ptr<CalibHessian> Hcalib) {
    cPrior = VecC::Constant(setting_initialCalibHessian);
    cPriorF = cPrior.cast<float>();
    (void) Hcalib;
    EFAdjointsValid = true;
    sideOf(this)->stateEpoch++;
}

// private members the reference declares; the device path does not call them
void EnergyFunctional::resubstituteF_MT(const VecX &, shared_ptr<CalibHessian>, bool) {}
void EnergyFunctional::resubstituteFPt(const VecCf &, Mat18f *, int, int, Vec10 *, int) {}
void EnergyFunctional::accumulateAF_MT(MatXX &H, VecX &b, bool) { accSSE_top_A->stitchDouble(H, b, this, true, false); }
void EnergyFunctional::accumulateLF_MT(MatXX &H, VecX &b, bool) { accSSE_top_L->stitchDouble(H, b, this, true, false); }
void EnergyFunctional::accumulateSCF_MT(MatXX &H, VecX &b, bool) { accSSE_bot->stitchDouble(H, b, this); }
void EnergyFunctional::calcLEnergyPt(int, int, Vec10 *, int) {}
void EnergyFunctional::orthogonalize(VecX *, MatXX *) {}

// -------------------------------------------------------------------------------------------------------------------------------
// AccumulatedTopHessianSSE / AccumulatedSCHessianSSE: the per-point calls collect the point set, the stitch runs the device pass
// (ldso_b200_accumulate: addPoint<mode> over the set + stitchDouble). acc[0][0] is not used; the set lives beside the object.
namespace {
struct AccSet { std::vector<const PointHessian *> pts; int mode = 0; bool shiftPrior = true; };
std::mutex g_accMutex;
std::map<const void *, AccSet> g_acc;
bool runAccumulate(const void *self, EnergyFunctional const *EF, MatXX *Htop, VecX *btop, MatXX *Hsc, VecX *bsc, int *nres) {
    EnergyFunctional *ef = const_cast<EnergyFunctional *>(EF);
    B200Side *S = sideOf(ef);
    if (!S || !S->ctx) return false;
    AccSet set;
    { std::lock_guard<std::mutex> l(g_accMutex); set = g_acc[self]; g_acc[self].pts.clear(); }
    auto &all = AccumulatedTopHessian::allPoints(ef);
    std::vector<int32_t> idx;
    for (size_t p = 0; p < all.size(); p++) for (auto q : set.pts) if (q == all[p].get()) { idx.push_back((int32_t) p); break; }
    const int n = 8 * (int) ef->frames.size() + CPARS;
    std::vector<double> ha((size_t) n * n), ba(n), hs((size_t) n * n), bs(n);
    int nr = 0;
    std::lock_guard<std::mutex> l(S->mtx);
    if (ldso_b200_accumulate(S->ctx, set.mode, (int) idx.size(), idx.data(), set.shiftPrior ? 1 : 0, ha.data(), ba.data(), hs.data(), bs.data(), &nr)) return false;
    if (Htop) { *Htop = MatXX::Zero(n, n); *btop = VecX::Zero(n); for (int j = 0; j < n; j++) { (*btop)[j] = ba[j]; for (int i = 0; i < n; i++) (*Htop)(i, j) = ha[(size_t) j * n + i]; } }
    if (Hsc) { *Hsc = MatXX::Zero(n, n); *bsc = VecX::Zero(n); for (int j = 0; j < n; j++) { (*bsc)[j] = bs[j]; for (int i = 0; i < n; i++) (*Hsc)(i, j) = hs[(size_t) j * n + i]; } }
    if (nres) *nres = nr;
    return true;
}
}  // namespace

template<int mode>
void AccumulatedTopHessianSSE::addPoint(shared_ptr<PointHessian> p, EnergyFunctional const *const ef, int tid) {
    (void) ef; (void) tid;
    std::lock_guard<std::mutex> l(g_accMutex);
    AccSet &s = g_acc[this];
    s.mode = mode;
    s.pts.push_back(p.get());
}
template void AccumulatedTopHessianSSE::addPoint<0>(shared_ptr<PointHessian> p, EnergyFunctional const *const ef, int tid);
template void AccumulatedTopHessianSSE::addPoint<1>(shared_ptr<PointHessian> p, EnergyFunctional const *const ef, int tid);
template void AccumulatedTopHessianSSE::addPoint<2>(shared_ptr<PointHessian> p, EnergyFunctional const *const ef, int tid);

void AccumulatedTopHessianSSE::stitchDouble(MatXX &H, VecX &b, EnergyFunctional const *const EF, bool usePrior, bool useDelta, int tid) {
    (void) useDelta; (void) tid;
    int nr = 0;
    if (!runAccumulate(this, EF, &H, &b, nullptr, nullptr, &nr)) { const int n = 8 * nframes[0] + CPARS; H = MatXX::Zero(n, n); b = VecX::Zero(n); return; }
    nres[0] = nr;
    if (usePrior) {       // AccumulatedTopHessian.cc:176-184
        for (int i = 0; i < CPARS; i++) { H(i, i) += EF->cPrior[i]; b[i] += EF->cPrior[i] * EF->cDeltaF.cast<double>()[i]; }
        for (size_t h = 0; h < EF->frames.size(); h++)
            for (int i = 0; i < 8; i++) { H(CPARS + 8 * h + i, CPARS + 8 * h + i) += EF->frames[h]->prior[i]; b[CPARS + 8 * h + i] += EF->frames[h]->prior[i] * EF->frames[h]->delta_prior[i]; }
    }
}
void AccumulatedTopHessianSSE::stitchDoubleInternal(MatXX *H, VecX *b, EnergyFunctional const *const EF, bool usePrior, int min, int max, Vec10 *stats, int tid) {
    (void) max; (void) stats;
    if (tid == -1) tid = 0;
    if (min != 0) return;       // the 6 reduce workers split [0, nF^2): the worker that owns block 0 delivers the (whole) device result
    stitchDouble(H[tid], b[tid], EF, usePrior, false, 0);
    // the header-inline stitchDoubleMT that follows ADDS each lower frame block's transpose onto the upper one and mirrors
    // (AccumulatedTopHessian.h:95-104): hand it the upper blocks only
    const int nF = (int) EF->frames.size();
    for (int h = 0; h < nF; h++) for (int t = h + 1; t < nF; t++) H[tid].block<8, 8>(CPARS + 8 * t, CPARS + 8 * h).setZero();
}

void AccumulatedSCHessianSSE::addPoint(shared_ptr<PointHessian> p, bool shiftPriorToZero, int tid) {
    (void) tid;
    std::lock_guard<std::mutex> l(g_accMutex);
    AccSet &s = g_acc[this];
    s.mode = 3;                  // the Schur terms use the point sums of A and L together (AccumulatedSCHessian.cc:24-29)
    s.shiftPrior = shiftPriorToZero;
    s.pts.push_back(p.get());
}
void AccumulatedSCHessianSSE::stitchDouble(MatXX &H, VecX &b, EnergyFunctional const *const EF, int tid) {
    (void) tid;
    if (!runAccumulate(this, EF, nullptr, nullptr, &H, &b, nullptr)) { const int n = 8 * nframes[0] + CPARS; H = MatXX::Zero(n, n); b = VecX::Zero(n); }
}
void AccumulatedSCHessianSSE::stitchDoubleInternal(MatXX *H, VecX *b, EnergyFunctional const *const EF, int min, int max, Vec10 *stats, int tid) {
    (void) max; (void) stats;
    if (tid == -1) tid = 0;
    if (min != 0) return;
    stitchDouble(H[tid], b[tid], EF, 0);
}

}  // namespace internal
}  // namespace ldso
