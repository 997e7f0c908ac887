// oracle/ref_pin/ref_bench.cc — the REFERENCE ARM of bench.py (TEST / MEASUREMENT INFRASTRUCTURE ONLY, never part of the product path).
// Builds, into the git-ignored oracle/_ref/libref_ba.so, a small C interface over the reference's OWN back-end translation units
// (src/internal/Residuals.cc, FrameHessian.cc, FrameFramePrecalc.cc, PointHessian.cc, OptimizationBackend/AccumulatedTopHessian.cc,
// AccumulatedSCHessian.cc, EnergyFunctional.cc, src/Setting.cc and their headers, incl. IndexThreadReduce.h), compiled UNMODIFIED where
// they lie under /root/reference with -O3 -march=native (the reference's Release flags), against the stand-in headers of
// oracle/ref_shim (no Eigen / Sophus / OpenCV / glog on this machine). Everything SURVEY section 8(a) names — PointFrameResidual::linearize,
// both addPoint's, the stitchers, solveSystemF, resubstituteF — therefore runs as the reference's own code on the reference's own
// 6-thread IndexThreadReduce. What is restated here is only the thin driver around it, which lives in src/frontend/FullSystem.cc (that
// file includes the whole front end and cannot be compiled): optimize()'s prologue and loop body (:734-831), linearizeAll (:1442-1530),
// applyRes_Reductor (:1706-1709), setNewFrameEnergyTH (:1762-1793), backupState (:1662-1676), doStepFromBackup (:1587-1622),
// setPrecalcValues (:1423-1431), solveSystem / getNullspaces (:1433-1440, :1711-1760). The dense 68x68 algebra inside solveSystemF
// (LDLT, scalings) goes through the stand-in's plain loops instead of Eigen's kernels; it is a few percent of an iteration.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include "../ref_shim/NumTypes.h"
#include "Settings.h"
#include "../ref_shim/ref_classes.h"
#include "internal/OptimizationBackend/EnergyFunctional.h"
#include "internal/GlobalCalib.h"
#include "frontend/CoarseTracker.h"
#include "ref_hooks.h"

namespace ldso { namespace internal { float wM3G, hM3G; int wG[PYR_LEVELS], hG[PYR_LEVELS]; } }
ldso::Camera::Camera(double fx_, double fy_, double cx_, double cy_) { fx = fx_; fy = fy_; cx = cx_; cy = cy_; }       // src/Camera.cc:7-12
ldso::Point::Point() {}                                                                                               // src/Point.cc:23-25 (numbering only)

using namespace ldso;
using namespace ldso::internal;

namespace {
struct RefWindow {
    int w = 0, h = 0;
    shared_ptr<CalibHessian> HC;
    std::vector<shared_ptr<Frame>> frames;                 // FullSystem::frames
    std::vector<shared_ptr<PointHessian>> points;
    std::vector<shared_ptr<Point>> pts;
    std::vector<shared_ptr<PointFrameResidual>> activeResiduals;
    shared_ptr<EnergyFunctional> ef;
    IndexThreadReduce<Vec10> *threadReduce = nullptr;      // leaked on purpose: its destructor prints to stdout
    double lastEnergyP = 0;

    void setPrecalcValues() {                              // FullSystem.cc:1423-1431
        for (auto &fr : frames) {
            fr->frameHessian->targetPrecalc.resize(frames.size());
            for (size_t i = 0; i < frames.size(); i++) fr->frameHessian->targetPrecalc[i].Set(fr->frameHessian, frames[i]->frameHessian, HC);
        }
        ef->setDeltaF(HC);
    }
    void linearizeAll_Reductor(int min, int max, Vec10 *stats, int) {      // :1494-1530 with fixLinearization = false
        for (int k = min; k < max; k++) (*stats)[0] += activeResiduals[k]->linearize(HC);
    }
    void applyRes_Reductor(int min, int max, Vec10 *, int) { for (int k = min; k < max; k++) activeResiduals[k]->applyRes(true); }   // :1706-1709
    void setNewFrameEnergyTH() {                           // :1762-1793
        std::vector<float> allResVec;
        allResVec.reserve(activeResiduals.size() * 2);
        shared_ptr<FrameHessian> newFrame = frames.back()->frameHessian;
        for (auto &r : activeResiduals)
            if (r->state_NewEnergyWithOutlier >= 0 && r->target.lock() == newFrame) allResVec.push_back(r->state_NewEnergyWithOutlier);
        if (allResVec.size() == 0) { newFrame->frameEnergyTH = 12 * 12 * patternNum; return; }
        int nthIdx = setting_frameEnergyTHN * allResVec.size();
        std::nth_element(allResVec.begin(), allResVec.begin() + nthIdx, allResVec.end());
        float nthElement = sqrtf(allResVec[nthIdx]);
        newFrame->frameEnergyTH = nthElement * setting_frameEnergyTHFacMedian;
        newFrame->frameEnergyTH = 26.0f * setting_frameEnergyTHConstWeight + newFrame->frameEnergyTH * (1 - setting_frameEnergyTHConstWeight);
        newFrame->frameEnergyTH = newFrame->frameEnergyTH * newFrame->frameEnergyTH;
        newFrame->frameEnergyTH *= setting_overallEnergyTHWeight * setting_overallEnergyTHWeight;
    }
    double linearizeAll() {                                // :1442-1492 with fixLinearization = false
        if (multiThreading) {
            threadReduce->reduce(std::bind(&RefWindow::linearizeAll_Reductor, this, _1, _2, _3, _4), 0, activeResiduals.size(), 0);
            lastEnergyP = threadReduce->stats[0];
        } else {
            Vec10 stats; stats.setZero();
            linearizeAll_Reductor(0, activeResiduals.size(), &stats, 0);
            lastEnergyP = stats[0];
        }
        setNewFrameEnergyTH();
        return lastEnergyP;
    }
    void applyResAll() {                                   // :762-766
        if (multiThreading) threadReduce->reduce(std::bind(&RefWindow::applyRes_Reductor, this, _1, _2, _3, _4), 0, activeResiduals.size(), 50);
        else applyRes_Reductor(0, activeResiduals.size(), 0, 0);
    }
    void getNullspaces() {                                 // :1711-1760
        auto &np = ef->lastNullspaces_pose, &ns = ef->lastNullspaces_scale, &na = ef->lastNullspaces_affA, &nb = ef->lastNullspaces_affB;
        np.clear(); ns.clear(); na.clear(); nb.clear();
        const int n = CPARS + frames.size() * 8;
        for (int i = 0; i < 6; i++) {
            VecX x0 = VecX::Zero(n);
            for (auto &fr : frames) {
                auto fh = fr->frameHessian;
                for (int r = 0; r < 6; r++) x0[CPARS + fh->idx * 8 + r] = fh->nullspaces_pose(r, i);
                for (int r = 0; r < 3; r++) x0[CPARS + fh->idx * 8 + r] *= SCALE_XI_TRANS_INVERSE;
                for (int r = 3; r < 6; r++) x0[CPARS + fh->idx * 8 + r] *= SCALE_XI_ROT_INVERSE;
            }
            np.push_back(x0);
        }
        for (int i = 0; i < 2; i++) {
            VecX x0 = VecX::Zero(n);
            for (auto &fr : frames) {
                auto fh = fr->frameHessian;
                x0[CPARS + fh->idx * 8 + 6] = fh->nullspaces_affine(0, i); x0[CPARS + fh->idx * 8 + 7] = fh->nullspaces_affine(1, i);
                x0[CPARS + fh->idx * 8 + 6] *= SCALE_A_INVERSE; x0[CPARS + fh->idx * 8 + 7] *= SCALE_B_INVERSE;
            }
            if (i == 0) na.push_back(x0); else nb.push_back(x0);
        }
        VecX x0 = VecX::Zero(n);
        for (auto &fr : frames) {
            auto fh = fr->frameHessian;
            for (int r = 0; r < 6; r++) x0[CPARS + fh->idx * 8 + r] = fh->nullspaces_scale[r];
            for (int r = 0; r < 3; r++) x0[CPARS + fh->idx * 8 + r] *= SCALE_XI_TRANS_INVERSE;
            for (int r = 3; r < 6; r++) x0[CPARS + fh->idx * 8 + r] *= SCALE_XI_ROT_INVERSE;
        }
        ns.push_back(x0);
    }
    void backupState() {                                   // :1662-1676 (no SOLVER_MOMENTUM)
        HC->value_backup = HC->value;
        for (auto &fr : frames) {
            auto fh = fr->frameHessian;
            fh->state_backup = fh->get_state();
            for (auto &feat : fr->features)
                if (feat->status == Feature::FeatureStatus::VALID && feat->point->status == Point::PointStatus::ACTIVE) feat->point->mpPH->idepth_backup = feat->point->mpPH->idepth;
        }
    }
    bool doStepFromBackup(float stepfacC, float stepfacT, float stepfacR, float stepfacA, float stepfacD) {      // :1587-1622
        Vec10 pstepfac;
        for (int i = 0; i < 3; i++) pstepfac[i] = stepfacT;
        for (int i = 3; i < 6; i++) pstepfac[i] = stepfacR;
        for (int i = 6; i < 10; i++) pstepfac[i] = stepfacA;
        float sumA = 0, sumB = 0, sumT = 0, sumR = 0, sumID = 0, numID = 0, sumNID = 0;
        HC->setValue(HC->value_backup + stepfacC * HC->step);
        for (auto &fr : frames) {
            auto fh = fr->frameHessian;
            fh->setState(fh->state_backup + pstepfac.cwiseProduct(fh->step));
            sumA += fh->step[6] * fh->step[6];
            sumB += fh->step[7] * fh->step[7];
            sumT += fh->step.segment<3>(0).squaredNorm();
            sumR += fh->step.segment<3>(3).squaredNorm();
            for (auto &feat : fr->features)
                if (feat->status == Feature::FeatureStatus::VALID && feat->point->status == Point::PointStatus::ACTIVE) {
                    auto ph = feat->point->mpPH;
                    ph->setIdepth(ph->idepth_backup + stepfacD * ph->step);
                    sumID += ph->step * ph->step;
                    sumNID += fabsf(ph->idepth_backup);
                    numID++;
                    ph->setIdepthZero(ph->idepth_backup + stepfacD * ph->step);
                }
        }
        sumA /= frames.size(); sumB /= frames.size(); sumR /= frames.size(); sumT /= frames.size(); sumID /= numID; sumNID /= numID;
        setPrecalcValues();
        return sqrtf(sumA) < 0.0005 * setting_thOptIterations && sqrtf(sumB) < 0.00005 * setting_thOptIterations &&
               sqrtf(sumR) < 0.00005 * setting_thOptIterations && sqrtf(sumT) * sumNID < 0.00005 * setting_thOptIterations;
    }
    double optimizeBegin() {                               // :734-766
        activeResiduals.clear();
        for (auto &fr : frames)
            for (auto &feat : fr->features) {
                shared_ptr<Point> p = feat->point;
                if (feat->status == Feature::FeatureStatus::VALID && p && p->status == Point::PointStatus::ACTIVE)
                    for (auto &r : p->mpPH->residuals) if (!r->isLinearized) { activeResiduals.push_back(r); r->resetOOB(); }
            }
        linearizeAll();
        applyResAll();
        return lastEnergyP;
    }
    bool gnIteration(int iteration) {                      // body of the loop :777-831 (setting_forceAceptStep, no SOLVER_STEPMOMENTUM)
        backupState();
        getNullspaces();
        ef->solveSystemF(iteration, 1e-1, HC);
        const bool canbreak = doStepFromBackup(1, 1, 1, 1, 1);
        linearizeAll();
        applyResAll();
        return canbreak;
    }
};
}  // namespace

extern "C" {
void *ref_ba_create(int w, int h, int multithreaded) {
    RefWindow *W = new RefWindow();
    W->w = w; W->h = h;
    wG[0] = w; hG[0] = h; wM3G = w - 3; hM3G = h - 3;
    pyrLevelsUsed = 1;                                     // FrameHessian's destructor frees pyrLevelsUsed pyramid levels; the harness owns the images
    multiThreading = multithreaded != 0;
    W->ef = std::make_shared<EnergyFunctional>();
    W->threadReduce = new IndexThreadReduce<Vec10>();
    W->ef->red = W->threadReduce;
    return W;
}
// value_scaled = (fx, fy, cx, cy); delta = value - value_zero (CalibHessian.h:22-36, :71-100)
void ref_ba_set_calib(void *o, const double K[4], const double *delta) {
    RefWindow *W = (RefWindow *) o;
    W->HC = std::make_shared<CalibHessian>(std::make_shared<Camera>(K[0], K[1], K[2], K[3]));
    if (delta) { VecC v; for (int i = 0; i < 4; i++) v[i] = W->HC->value_zero[i] + delta[i]; W->HC->setValue(v); }
}
int ref_ba_add_frame(void *o, const double R[9], const double t[3], const double state_zero[10], const double state[10], float ab_exposure,
                     int frame_id, const float *dI) {
    RefWindow *W = (RefWindow *) o;
    auto fr = std::make_shared<Frame>(); fr->id = frame_id;
    FrameHessian *p = new FrameHessian(fr);
    for (int i = 0; i < PYR_LEVELS; i++) { p->dIp[i] = nullptr; p->absSquaredGrad[i] = nullptr; }
    shared_ptr<FrameHessian> fh(p, [](FrameHessian *q) { for (int i = 0; i < PYR_LEVELS; i++) { q->dIp[i] = nullptr; q->absSquaredGrad[i] = nullptr; } delete q; });
    fr->frameHessian = fh;
    fh->frameID = (int) W->frames.size(); fh->ab_exposure = ab_exposure;
    fh->dI = (Vec3f *) dI; fh->dIp[0] = nullptr;
    Vec10 sz, st; for (int i = 0; i < 10; i++) { sz[i] = state_zero[i]; st[i] = state[i]; }
    fh->setEvalPT(SE3(oracle::SE3::fromRt(R, t)), sz);     // FrameHessian.h:107-112
    fh->setState(st);
    W->frames.push_back(fr);
    return (int) W->frames.size() - 1;
}
int ref_ba_add_point(void *o, int host, float u, float v, float idepth_zero, float idepth, int hasDepthPrior, const float color[8], const float weights[8]) {
    RefWindow *W = (RefWindow *) o;
    auto feat = std::make_shared<Feature>(u, v, W->frames[host]); feat->status = Feature::FeatureStatus::VALID;
    auto pt = std::make_shared<Point>(); pt->status = Point::PointStatus::ACTIVE; pt->mHostFeature = feat; feat->point = pt;
    auto ph = std::make_shared<PointHessian>(); pt->mpPH = ph; ph->point = pt;
    ph->u = u; ph->v = v; ph->hasDepthPrior = hasDepthPrior != 0;
    ph->setIdepthZero(idepth_zero); ph->setIdepth(idepth);
    memcpy(ph->color, color, 32); memcpy(ph->weights, weights, 32);
    ph->takeData();
    W->frames[host]->features.push_back(feat);
    W->points.push_back(ph); W->pts.push_back(pt);
    return (int) W->points.size() - 1;
}
int ref_ba_add_residual(void *o, int point, int target) {
    RefWindow *W = (RefWindow *) o;
    auto ph = W->points[point];
    auto host = ph->point->mHostFeature.lock()->host.lock()->frameHessian;
    ph->residuals.push_back(std::make_shared<PointFrameResidual>(ph, host, W->frames[target]->frameHessian));
    return 0;
}
void ref_ba_finalize(void *o) {                            // insertFrame per keyframe (EnergyFunctional.cc:30-61), then setPrecalcValues
    RefWindow *W = (RefWindow *) o;
    for (auto &fr : W->frames) W->ef->insertFrame(fr->frameHessian, W->HC);
    W->ef->makeIDX();
    W->setPrecalcValues();
}
double ref_ba_optimize_begin(void *o) { return ((RefWindow *) o)->optimizeBegin(); }
int ref_ba_gn_iteration(void *o, int iteration) { return ((RefWindow *) o)->gnIteration(iteration) ? 1 : 0; }
double ref_ba_energy(void *o) { return ((RefWindow *) o)->lastEnergyP; }
int ref_ba_num_active(void *o) { return (int) ((RefWindow *) o)->activeResiduals.size(); }
void ref_ba_last_x(void *o, double *x) { RefWindow *W = (RefWindow *) o; for (int i = 0; i < W->ef->lastX.size(); i++) x[i] = W->ef->lastX[i]; }
void ref_ba_point_idepths(void *o, float *idepth) { RefWindow *W = (RefWindow *) o; for (size_t i = 0; i < W->points.size(); i++) idepth[i] = W->points[i]->idepth; }
// seconds per GN iteration (median of `iters` after `warmup`)
double ref_ba_time_gn(void *o, int iters, int warmup) {
    RefWindow *W = (RefWindow *) o;
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
// ---- the reference's own CoarseTracker (src/frontend/CoarseTracker.cc, compiled unmodified): setCoarseTrackingRef + trackNewestCoarse
struct RefTracker {
    shared_ptr<CalibHessian> HC;
    std::vector<shared_ptr<Frame>> FR; std::vector<shared_ptr<FrameHessian>> FH;
    shared_ptr<FrameHessian> newFH;
    std::vector<shared_ptr<PointFrameResidual>> keep;
    CoarseTracker *T = nullptr;
};
static shared_ptr<FrameHessian> ref_make_fh(shared_ptr<Frame> fr) {
    FrameHessian *p = new FrameHessian(fr);
    for (int i = 0; i < PYR_LEVELS; i++) { p->dIp[i] = nullptr; p->absSquaredGrad[i] = nullptr; }
    return shared_ptr<FrameHessian>(p, [](FrameHessian *q) { for (int i = 0; i < PYR_LEVELS; i++) { q->dIp[i] = nullptr; q->absSquaredGrad[i] = nullptr; } delete q; });
}
// refDIp / newDIp: `levels` pointers to (I, dx, dy) AoS pyramids; the n reference contributions are (centerProjectedTo[3], HdiF) of the
// ACTIVE points whose newest residual targets the reference keyframe and is IN (CoarseTracker.cc:258-283)
void *ref_tracker_create(int w, int h, int levels, const double K[4], const float **refDIp, float ref_aff_a, float ref_aff_b, float ref_exposure,
                         int n, const float *cpt3, const float *HdiF, const float **newDIp, float new_exposure) {
    RefTracker *R = new RefTracker();
    pyrLevelsUsed = levels;
    for (int l = 0; l < levels; l++) { wG[l] = w >> l; hG[l] = h >> l; }
    wM3G = w - 3; hM3G = h - 3;
    R->HC = std::make_shared<CalibHessian>(std::make_shared<Camera>(K[0], K[1], K[2], K[3]));
    auto fr = std::make_shared<Frame>(); auto fh = ref_make_fh(fr); fr->frameHessian = fh; fr->id = 1;
    for (int l = 0; l < levels; l++) fh->dIp[l] = (Vec3f *) refDIp[l];
    fh->dI = fh->dIp[0]; fh->ab_exposure = ref_exposure;
    fh->setEvalPT_scaled(SE3(), AffLight(ref_aff_a, ref_aff_b));
    for (int i = 0; i < n; i++) {
        auto feat = std::make_shared<Feature>(0.f, 0.f, fr); auto pt = std::make_shared<Point>(); auto ph = std::make_shared<PointHessian>();
        feat->point = pt; pt->mpPH = ph; feat->status = Feature::FeatureStatus::VALID; pt->status = Point::PointStatus::ACTIVE;
        auto r = std::make_shared<PointFrameResidual>(ph, fh, fh);
        r->isActiveAndIsGoodNEW = true;
        r->centerProjectedTo = Vec3f(cpt3[3 * i], cpt3[3 * i + 1], cpt3[3 * i + 2]);
        ph->HdiF = HdiF[i];
        ph->lastResiduals[0] = std::make_pair(r, ResState::IN);
        fr->features.push_back(feat); R->keep.push_back(r);
    }
    R->FR.push_back(fr); R->FH.push_back(fh);
    R->newFH = ref_make_fh(nullptr);
    for (int l = 0; l < levels; l++) R->newFH->dIp[l] = (Vec3f *) newDIp[l];
    R->newFH->dI = R->newFH->dIp[0]; R->newFH->ab_exposure = new_exposure;
    R->T = new CoarseTracker(w, h);
    R->T->makeK(R->HC);
    R->T->setCoarseTrackingRef(R->FH);
    return R;
}
// trackNewestCoarse from (R, t), (a, b); returns its bool, the pose / brightness it found and seconds per call (median of `reps`)
int ref_tracker_track(void *o, double Rm[9], double t[3], float *aff_a, float *aff_b, int coarsestLvl, int reps, double *seconds) {
    RefTracker *R = (RefTracker *) o;
    const SE3 start(oracle::SE3::fromRt(Rm, t));
    const AffLight aff0(*aff_a, *aff_b);
    Vec5 minRes; for (int i = 0; i < 5; i++) minRes[i] = NAN;
    std::vector<double> ts; bool ok = false; SE3 T; AffLight aff;
    for (int k = 0; k < std::max(reps, 1); k++) {
        T = start; aff = aff0;
        auto t0 = std::chrono::steady_clock::now();
        ok = R->T->trackNewestCoarse(R->newFH, T, aff, coarsestLvl, minRes);
        ts.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    std::sort(ts.begin(), ts.end());
    if (seconds) *seconds = ts[ts.size() / 2];
    const Mat33 Ro = T.rotationMatrix();
    for (int i = 0; i < 3; i++) { t[i] = T.translation()[i]; for (int j = 0; j < 3; j++) Rm[i * 3 + j] = Ro(i, j); }
    *aff_a = aff.a; *aff_b = aff.b;
    return ok ? 1 : 0;
}

// Host-side bookkeeping of the back end after a scripted piece of FullSystem's window maintenance -- no arithmetic member is called, so
// the drop-in translation units (libdropin_ba.so) can be compared with the reference's own (libref_ba.so) WITHOUT a GPU:
//   insertResidual for every residual (FullSystem.cc:1003-1005, activatePointsMT), dropResidual of every residual whose target is frame
//   `drop_target` (flagFramesForMarginalization, FullSystem.cc:1312-1330), removePoint of every `remove_every`-th point after marking it
//   OUT (flagPointsForRemoval + EnergyFunctional::dropPointsF's effect on the lists), makeIDX.
// out: nFrames, nPoints, nResiduals, allPoints.size(), EFIndicesValid, resInA; per frame idx, frameID; per ordered pair (h, t) the two
// connectivityMap counters; per remaining residual of every remaining point hostIDX * 64 + targetIDX. Returns the number of values.
int ref_ba_bookkeeping(void *o, int drop_target, int remove_every, long long *out, int cap) {
    RefWindow *W = (RefWindow *) o;
    EnergyFunctional &ef = *W->ef;
    for (auto &ph : W->points) { for (auto &r : ph->residuals) ef.insertResidual(r); ef.nPoints++; }       // insertPoint's counter (EnergyFunctional.h) + insertResidual
    if (drop_target >= 0 && drop_target < (int) W->frames.size()) {
        auto fh = W->frames[drop_target]->frameHessian;
        for (auto &ph : W->points)
            for (size_t i = 0; i < ph->residuals.size();) {
                if (ph->residuals[i]->target.lock() == fh) ef.dropResidual(ph->residuals[i]);      // erases the entry from ph->residuals
                else i++;
            }
    }
    if (remove_every > 0)
        for (size_t p = 0; p < W->points.size(); p += remove_every) {
            W->pts[p]->status = Point::PointStatus::OUTLIER;
            ef.removePoint(W->points[p]);
        }
    ef.makeIDX();
    int k = 0;
    auto put = [&](long long v) { if (k < cap) out[k] = v; k++; };
    put(ef.nFrames); put(ef.nPoints); put(ef.nResiduals); put((long long) ef.allPoints.size()); put(EFIndicesValid ? 1 : 0); put(ef.resInA);
    for (auto &f : ef.frames) { put(f->idx); put(f->frameID); }
    for (auto &fh : ef.frames)
        for (auto &ft : ef.frames) {
            auto it = ef.connectivityMap.find((((uint64_t) fh->frameID) << 32) + ((uint64_t) ft->frameID));
            put(it == ef.connectivityMap.end() ? -1 : it->second[0]);
            put(it == ef.connectivityMap.end() ? -1 : it->second[1]);
        }
    for (auto &ph : ef.allPoints)
        for (auto &r : ph->residuals) put(r->hostIDX * 64 + r->targetIDX);
    return k;
}

// CoarseTracker(w, h) + makeK (CoarseTracker.cc:219-246): the public per-level intrinsics FullSystem and LoopClosing read. out: per level
// w, h, fx, fy, cx, cy, fxi, fyi, cxi, cyi (10 values). Host logic only: comparable between the two libraries without a GPU.
int ref_tracker_make_k(int w, int h, int levels, const double K[4], double *out) {
    pyrLevelsUsed = levels;
    for (int l = 0; l < levels; l++) { wG[l] = w >> l; hG[l] = h >> l; }
    wM3G = w - 3; hM3G = h - 3;
    auto HC = std::make_shared<CalibHessian>(std::make_shared<Camera>(K[0], K[1], K[2], K[3]));
    CoarseTracker *T = new CoarseTracker(w, h);          // (leaked: the harness never frees trackers)
    T->makeK(HC);
    for (int l = 0; l < levels; l++) {
        double *o = out + 10 * l;
        o[0] = T->w[l]; o[1] = T->h[l]; o[2] = T->fx[l]; o[3] = T->fy[l]; o[4] = T->cx[l]; o[5] = T->cy[l];
        o[6] = T->fxi[l]; o[7] = T->fyi[l]; o[8] = T->cxi[l]; o[9] = T->cyi[l];
    }
    return levels;
}

// development aid: seconds spent in the phases of `iters` GN iterations: [backup + nullspaces, solveSystemF, doStepFromBackup, linearizeAll, applyRes]
void ref_ba_profile(void *o, int iters, double out[5]) {
    RefWindow *W = (RefWindow *) o;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto sec = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    for (int i = 0; i < 5; i++) out[i] = 0;
    for (int k = 0; k < iters; k++) {
        auto t0 = now(); W->backupState(); W->getNullspaces();
        auto t1 = now(); W->ef->solveSystemF(3, 1e-1, W->HC);
        auto t2 = now(); W->doStepFromBackup(1, 1, 1, 1, 1);
        auto t3 = now(); W->linearizeAll();
        auto t4 = now(); W->applyResAll();
        auto t5 = now();
        out[0] += sec(t0, t1); out[1] += sec(t1, t2); out[2] += sec(t2, t3); out[3] += sec(t3, t4); out[4] += sec(t4, t5);
    }
}
}  // extern "C"
