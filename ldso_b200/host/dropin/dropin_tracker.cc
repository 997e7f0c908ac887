// Drop-in replacement for the reference's src/frontend/CoarseTracker.cc, written against the reference's OWN header
// (include/frontend/CoarseTracker.h:17-169): same constructors, same members, same public data. CoarseTracker forwards to the C ABI
// of include/ldso_b200.h (calcRes / calcGSSSE / the LM loop of trackNewestCoarse and makeCoarseDepthL0 run on the device; one device
// context per instance, as FullSystem keeps two trackers alive on two threads, FullSystem.h:300-301). CoarseDistanceMap is host
// bookkeeping that FullSystem::activatePointsMT consults per candidate (FullSystem.cc:1076-1150): it stays on the host here, written
// as a table-driven frontier BFS; the batched device version of the whole selection is ldso_b200_select_activation.
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "Feature.h"
#include "Point.h"
#include "frontend/CoarseTracker.h"
#include "internal/GlobalCalib.h"

#include "../../../include/ldso_b200.h"

namespace ldso {

namespace {
struct TrackerSide { ldso_b200_ctx *ctx = nullptr; int refSlot = 0, newSlot = 1; float refExposure = 1; };
std::mutex g_trkMutex;
std::map<const CoarseTracker *, TrackerSide> g_trk;
TrackerSide &trkSide(const CoarseTracker *t) { std::lock_guard<std::mutex> l(g_trkMutex); return g_trk[t]; }
}  // namespace

CoarseTracker::CoarseTracker(int ww, int hh) {
    TrackerSide &S = trkSide(this);
    S.ctx = ldso_b200_create(0, ww, hh, pyrLevelsUsed, nullptr);
    for (int l = 0; l < PYR_LEVELS; l++) { pc_u[l] = pc_v[l] = pc_idepth[l] = pc_color[l] = nullptr; pc_n[l] = 0; idepth[l] = weightSums[l] = weightSums_bak[l] = nullptr; }
    buf_warped_idepth = buf_warped_u = buf_warped_v = buf_warped_dx = buf_warped_dy = buf_warped_residual = buf_warped_weight = buf_warped_refColor = nullptr;
    buf_warped_n = 0;
    w[0] = h[0] = 0;
}

void CoarseTracker::makeK(shared_ptr<CalibHessian> HCalib) {       // CoarseTracker.cc:219-246
    w[0] = wG[0]; h[0] = hG[0];
    fx[0] = HCalib->fxl(); fy[0] = HCalib->fyl(); cx[0] = HCalib->cxl(); cy[0] = HCalib->cyl();
    for (int level = 1; level < pyrLevelsUsed; ++level) {
        w[level] = w[0] >> level; h[level] = h[0] >> level;
        fx[level] = fx[level - 1] * 0.5; fy[level] = fy[level - 1] * 0.5;
        cx[level] = (cx[0] + 0.5) / ((int) 1 << level) - 0.5; cy[level] = (cy[0] + 0.5) / ((int) 1 << level) - 0.5;
    }
    for (int level = 0; level < pyrLevelsUsed; ++level) {
        K[level] << fx[level], 0.0, cx[level], 0.0, fy[level], cy[level], 0.0, 0.0, 1.0;
        Ki[level] = K[level].inverse();
        fxi[level] = Ki[level](0, 0); fyi[level] = Ki[level](1, 1); cxi[level] = Ki[level](0, 2); cyi[level] = Ki[level](1, 2);
    }
    ldso_b200_tracker_make_k(trkSide(this).ctx, fx[0], fy[0], cx[0], cy[0]);
}

// CoarseTracker.cc:248-256 with makeCoarseDepthL0 (:258-438) on the device: the contributions are the ACTIVE points whose newest
// residual is IN (it targets lastRef): centerProjectedTo and HdiF.
void CoarseTracker::setCoarseTrackingRef(std::vector<shared_ptr<FrameHessian>> &frameHessians) {
    TrackerSide &S = trkSide(this);
    lastRef = frameHessians.back();
    std::vector<float> cpt, hdi;
    for (shared_ptr<FrameHessian> fh : frameHessians)
        for (shared_ptr<Feature> feat : fh->frame->features)
            if (feat->status == Feature::FeatureStatus::VALID && feat->point->status == Point::PointStatus::ACTIVE) {
                shared_ptr<PointHessian> ph = feat->point->mpPH;
                if (ph->lastResiduals[0].first != 0 && ph->lastResiduals[0].second == ResState::IN) {
                    shared_ptr<PointFrameResidual> r = ph->lastResiduals[0].first;
                    for (int i = 0; i < 3; i++) cpt.push_back(r->centerProjectedTo[i]);
                    hdi.push_back(ph->HdiF);
                }
            }
    const float *lv[LDSO_B200_MAX_LEVELS] = {nullptr};
    for (int l = 0; l < pyrLevelsUsed; l++) lv[l] = (const float *) lastRef->dIp[l];
    ldso_b200_upload_frame(S.ctx, S.refSlot, lv, pyrLevelsUsed);
    ldso_b200_tracker_make_coarse_depth(S.ctx, S.refSlot, (int) hdi.size(), cpt.data(), hdi.data());
    S.refExposure = lastRef->ab_exposure;
    refFrameID = lastRef->frame->id;
    lastRef_aff_g2l = lastRef->aff_g2l();
    firstCoarseRMSE = -1;
}

// CoarseTracker.cc:61-217: the whole coarse-to-fine LM loop in one device launch
bool CoarseTracker::trackNewestCoarse(shared_ptr<FrameHessian> newFrameHessian, SE3 &lastToNew_out, AffLight &aff_g2l_out, int coarsestLvl,
                                      Vec5 minResForAbort) {
    TrackerSide &S = trkSide(this);
    newFrame = newFrameHessian;
    const float *lv[LDSO_B200_MAX_LEVELS] = {nullptr};
    for (int l = 0; l < pyrLevelsUsed; l++) lv[l] = (const float *) newFrame->dIp[l];
    if (ldso_b200_upload_frame(S.ctx, S.newSlot, lv, pyrLevelsUsed)) return false;
    ldso_b200_tracker_set_frames(S.ctx, (float) lastRef_aff_g2l.a, (float) lastRef_aff_g2l.b, S.refExposure, S.newSlot, newFrame->ab_exposure);
    double R[9], t[3], mra[5], lr[5], lf[3];
    const Mat33 Rm = lastToNew_out.rotationMatrix();
    for (int i = 0; i < 3; i++) { t[i] = lastToNew_out.translation()[i]; for (int j = 0; j < 3; j++) R[i * 3 + j] = Rm(i, j); }
    for (int i = 0; i < 5; i++) mra[i] = minResForAbort[i];
    float a = (float) aff_g2l_out.a, b = (float) aff_g2l_out.b;
    int ok = 0;
    if (ldso_b200_tracker_track(S.ctx, R, t, &a, &b, coarsestLvl, mra, lr, lf, &ok)) return false;
    for (int i = 0; i < 5; i++) lastResiduals[i] = lr[i];
    for (int i = 0; i < 3; i++) lastFlowIndicators[i] = lf[i];
    Mat33 Ro; Vec3 to;
    for (int i = 0; i < 3; i++) { to[i] = t[i]; for (int j = 0; j < 3; j++) Ro(i, j) = R[i * 3 + j]; }
    lastToNew_out = SE3(Ro, to);
    aff_g2l_out = AffLight(a, b);
    return ok != 0;
}

// private members of the reference's class: their work happens inside the device calls above
void CoarseTracker::makeCoarseDepthL0(std::vector<shared_ptr<FrameHessian>>) {}
Vec6 CoarseTracker::calcRes(int lvl, const SE3 &refToNew, AffLight aff_g2l, float cutoffTH) {
    TrackerSide &S = trkSide(this);
    double R[9], t[3], res6[6] = {0, 0, 0, 0, 0, 0};
    const Mat33 Rm = refToNew.rotationMatrix();
    for (int i = 0; i < 3; i++) { t[i] = refToNew.translation()[i]; for (int j = 0; j < 3; j++) R[i * 3 + j] = Rm(i, j); }
    ldso_b200_tracker_eval(S.ctx, lvl, R, t, (float) aff_g2l.a, (float) aff_g2l.b, cutoffTH, res6, nullptr, nullptr);
    Vec6 r; for (int i = 0; i < 6; i++) r[i] = res6[i];
    return r;
}
void CoarseTracker::calcGSSSE(int lvl, Mat88 &H_out, Vec8 &b_out, const SE3 &refToNew, AffLight aff_g2l) {
    TrackerSide &S = trkSide(this);
    double R[9], t[3], res6[6], H[64], b[8];
    const Mat33 Rm = refToNew.rotationMatrix();
    for (int i = 0; i < 3; i++) { t[i] = refToNew.translation()[i]; for (int j = 0; j < 3; j++) R[i * 3 + j] = Rm(i, j); }
    ldso_b200_tracker_eval(S.ctx, lvl, R, t, (float) aff_g2l.a, (float) aff_g2l.b, setting_coarseCutoffTH, res6, H, b);
    for (int i = 0; i < 8; i++) { b_out[i] = b[i]; for (int j = 0; j < 8; j++) H_out(i, j) = H[j * 8 + i]; }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// CoarseDistanceMap (CoarseTracker.cc:634-819), host side
CoarseDistanceMap::CoarseDistanceMap(int ww, int hh) {
    fwdWarpedIDDistFinal = new float[ww * hh / 4];
    bfsList1 = new Eigen::Vector2i[ww * hh / 4];
    bfsList2 = new Eigen::Vector2i[ww * hh / 4];
    coarseProjectionGrid = nullptr;          // unused by any caller
    coarseProjectionGridNum = nullptr;
    w[0] = h[0] = 0;
}
CoarseDistanceMap::~CoarseDistanceMap() {
    delete[] fwdWarpedIDDistFinal;
    delete[] bfsList1;
    delete[] bfsList2;
}
void CoarseDistanceMap::makeK(shared_ptr<CalibHessian> HCalib) {
    w[0] = wG[0]; h[0] = hG[0];
    fx[0] = HCalib->fxl(); fy[0] = HCalib->fyl(); cx[0] = HCalib->cxl(); cy[0] = HCalib->cyl();
    for (int level = 1; level < pyrLevelsUsed; ++level) {
        w[level] = w[0] >> level; h[level] = h[0] >> level;
        fx[level] = fx[level - 1] * 0.5; fy[level] = fy[level - 1] * 0.5;
        cx[level] = (cx[0] + 0.5) / ((int) 1 << level) - 0.5; cy[level] = (cy[0] + 0.5) / ((int) 1 << level) - 0.5;
    }
    for (int level = 0; level < pyrLevelsUsed; ++level) {
        K[level] << fx[level], 0.0, cx[level], 0.0, fy[level], cy[level], 0.0, 0.0, 1.0;
        Ki[level] = K[level].inverse();
        fxi[level] = Ki[level](0, 0); fyi[level] = Ki[level](1, 1); cxi[level] = Ki[level](0, 2); cyi[level] = Ki[level](1, 2);
    }
}
// every ACTIVE point of the other keyframes projected into `frame` at level 1 seeds the map with distance 0
void CoarseDistanceMap::makeDistanceMap(std::vector<shared_ptr<FrameHessian>> &frameHessians, shared_ptr<FrameHessian> frame) {
    const int w1 = w[1], h1 = h[1];
    for (int i = 0; i < w1 * h1; i++) fwdWarpedIDDistFinal[i] = 1000;
    int seeds = 0;
    for (auto fh : frameHessians) {
        if (frame == fh) continue;
        const SE3 fhToNew = frame->PRE_worldToCam * fh->PRE_camToWorld;
        const Mat33f KRKi = (K[1] * fhToNew.rotationMatrix().cast<float>() * Ki[0]);
        const Vec3f Kt = (K[1] * fhToNew.translation().cast<float>());
        for (auto feat : fh->frame->features) {
            if (!(feat->point && feat->point->status == Point::PointStatus::ACTIVE)) continue;
            auto ph = feat->point->mpPH;
            const Vec3f ptp = KRKi * Vec3f(ph->u, ph->v, 1) + Kt * ph->idepth_scaled;
            const int u = ptp[0] / ptp[2] + 0.5f, v = ptp[1] / ptp[2] + 0.5f;
            if (!(u > 0 && v > 0 && u < w1 && v < h1)) continue;
            fwdWarpedIDDistFinal[u + w1 * v] = 0;
            bfsList1[seeds++] = Eigen::Vector2i(u, v);
        }
    }
    growDistBFS(seeds);
}
// 39 frontier steps; even steps reach the 4-neighbourhood, odd steps the 8-neighbourhood; a cell keeps the first (smallest) step that reaches it
void CoarseDistanceMap::growDistBFS(int bfsNum) {
    static const int dx[8] = {1, -1, 0, 0, 1, -1, -1, 1}, dy[8] = {0, 0, 1, -1, 1, 1, -1, -1};
    const int w1 = w[1], h1 = h[1];
    for (int k = 1; k < 40; k++) {
        const int nPrev = bfsNum, nNb = (k % 2 == 0) ? 4 : 8;
        std::swap(bfsList1, bfsList2);
        bfsNum = 0;
        for (int i = 0; i < nPrev; i++) {
            const int x = bfsList2[i][0], y = bfsList2[i][1];
            if (x == 0 || y == 0 || x == w1 - 1 || y == h1 - 1) continue;
            for (int q = 0; q < nNb; q++) {
                const int xx = x + dx[q], yy = y + dy[q];
                float &cell = fwdWarpedIDDistFinal[xx + yy * w1];
                if (cell > k) { cell = k; bfsList1[bfsNum++] = Eigen::Vector2i(xx, yy); }
            }
        }
    }
}
void CoarseDistanceMap::addIntoDistFinal(int u, int v) {
    if (w[0] == 0) return;
    bfsList1[0] = Eigen::Vector2i(u, v);
    fwdWarpedIDDistFinal[u + w[1] * v] = 0;
    growDistBFS(1);
}

}  // namespace ldso
