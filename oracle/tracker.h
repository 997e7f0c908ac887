// ORACLE — TEST INFRASTRUCTURE ONLY (see omath.h header). PARITY PINNED AGAINST THE REFERENCE'S OWN SOURCES, EXCEPT EIGEN / SOPHUS INTERNALS (oracle/ref_pin compiles, unmodified, the reference's Residuals.cc, ImmaturePoint.cc, PointHessian.cc, FrameHessian.cc, FrameFramePrecalc.cc, AccumulatedTopHessian.cc, AccumulatedSCHessian.cc, EnergyFunctional.cc, CoarseTracker.cc, Setting.cc and their headers against stand-in Eigen/Sophus headers and matches the oracle bit for bit on 108 checks; the float SSE / scalar code is pinned completely, the Eigen-expression code structurally - what is added where, in which order - with LDLT, PartialPivLU, JacobiSVD, SE3 exp/log and the product kernels being the oracle's own restatements on both sides; FullSystem.cc's driver loop needs the whole front end and stays restated from the cited lines).
// CPU restatement of the reference's coarse direct tracker:
//   include/frontend/CoarseTracker.h:17-127, src/frontend/CoarseTracker.cc:61-246,258-632
#pragma once
#include "ba.h"

namespace oracle {

static const int PYR_LEVELS = 6;

struct CoarseTracker {
    Settings S;
    int pyrLevelsUsed = 0;
    int w[PYR_LEVELS], h[PYR_LEVELS];
    float fx[PYR_LEVELS], fy[PYR_LEVELS], cx[PYR_LEVELS], cy[PYR_LEVELS];
    float fxi[PYR_LEVELS], fyi[PYR_LEVELS], cxi[PYR_LEVELS], cyi[PYR_LEVELS];
    float K[PYR_LEVELS][9], Ki[PYR_LEVELS][9];

    // reference frame data
    const float *refDIp[PYR_LEVELS];   // lastRef->dIp[lvl], AoS (I,dx,dy), not owned
    float lastRef_aff_a = 0, lastRef_aff_b = 0;
    float lastRef_ab_exposure = 1;
    // new frame data
    const float *newDIp[PYR_LEVELS];
    float newFrame_ab_exposure = 1;

    std::vector<float> idepth[PYR_LEVELS], weightSums[PYR_LEVELS], weightSums_bak[PYR_LEVELS];
    std::vector<float> pc_u[PYR_LEVELS], pc_v[PYR_LEVELS], pc_idepth[PYR_LEVELS], pc_color[PYR_LEVELS];
    int pc_n[PYR_LEVELS];

    std::vector<float> buf_warped_idepth, buf_warped_u, buf_warped_v, buf_warped_dx, buf_warped_dy,
            buf_warped_residual, buf_warped_weight, buf_warped_refColor;
    int buf_warped_n = 0;
    Accumulator9 acc;

    double lastResiduals[5];
    double lastFlowIndicators[3];
    int lm_iterations_total = 0;  // diagnostic: number of calcRes evaluations in the last track

    CoarseTracker(int ww, int hh, int levels);
    void makeK(float fxl, float fyl, float cxl, float cyl);                         // CoarseTracker.cc:219-246
    // makeCoarseDepthL0 (:258-438): contributions = (centerProjectedTo[0..2], HdiF) of every ACTIVE
    // point whose newest residual targets lastRef and is IN.
    void makeCoarseDepthL0(int n, const float *cpt /*n*3*/, const float *HdiF /*n*/);
    void calcRes(int lvl, const SE3 &refToNew, float aff_a, float aff_b, float cutoffTH, double res[6]);  // :440-572
    void calcGSSSE(int lvl, double H_out[64] /*row-major*/, double b_out[8], const SE3 &refToNew, float aff_a, float aff_b);  // :574-632
    bool trackNewestCoarse(SE3 &lastToNew_out, float &aff_a_out, float &aff_b_out, int coarsestLvl,
                           const double minResForAbort[5]);                          // :61-217
};

// CoarseDistanceMap — include/frontend/CoarseTracker.h:128-170, src/frontend/CoarseTracker.cc:634-870: for every pixel of pyramid
// level 1 of the newest keyframe, the (alternating 4-/8-neighbourhood) distance to the nearest projected ACTIVE point, capped at 40.
struct CoarseDistanceMap {
    int pyrLevelsUsed = 0;
    int w[PYR_LEVELS], h[PYR_LEVELS];
    float fx[PYR_LEVELS], fy[PYR_LEVELS], cx[PYR_LEVELS], cy[PYR_LEVELS];
    float K[PYR_LEVELS][9], Ki[PYR_LEVELS][9];
    std::vector<float> fwdWarpedIDDistFinal;     // w[1] * h[1]
    std::vector<int> bfsList1, bfsList2;         // (x, y) pairs
    int numItems = 0;
    CoarseDistanceMap(int ww, int hh, int levels);                                 // :637-647
    void makeK(float fxl, float fyl, float cxl, float cyl);                        // :657-685
    // makeDistanceMap (:687-726) in three parts: reset, the points of one host keyframe (R, t = rotation / translation of
    // newest.PRE_worldToCam * host.PRE_camToWorld cast to float, as :704-706 does), grow
    void beginDistanceMap();
    void addFramePoints(const float R[9], const float t[3], int n, const float *u, const float *v, const float *idepth_scaled);
    void finishDistanceMap() { growDistBFS(numItems); }
    void growDistBFS(int bfsNum);                                                  // :728-812
    void addIntoDistFinal(int u, int v);                                           // :814-819
    void hostToNewest(const float R[9], const float t[3], float KRKi[9], float Kt[3]) const;   // :705-706, FullSystem.cc:1093-1094
};

// The selection loop of FullSystem::activatePointsMT (FullSystem.cc:1088-1150) over candidates of one host keyframe, in order.
// action: 0 = stays immature, 1 = goes to optimizeImmaturePoint (and was added to the distance map), 2 = deleted.
struct ActivationCand { float u, v, idepth_min, idepth_max, quality, lastTracePixelInterval, my_type; int lastTraceStatus; };
void selectActivation(CoarseDistanceMap &M, const float R[9], const float t[3], bool hostFlaggedForMarginalization, float currentMinActDist,
                      float minTraceQuality, int n, const ActivationCand *c, unsigned char *action);

}  // namespace oracle
