// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/ba.h). PARITY PINNED AGAINST THE REFERENCE'S OWN SOURCES, EXCEPT EIGEN / SOPHUS INTERNALS (oracle/ref_pin compiles, unmodified, the reference's Residuals.cc, ImmaturePoint.cc, PointHessian.cc, FrameHessian.cc, FrameFramePrecalc.cc, AccumulatedTopHessian.cc, AccumulatedSCHessian.cc, EnergyFunctional.cc, CoarseTracker.cc, Setting.cc and their headers against stand-in Eigen/Sophus headers and matches the oracle bit for bit on 108 checks; the float SSE / scalar code is pinned completely, the Eigen-expression code structurally - what is added where, in which order - with LDLT, PartialPivLU, JacobiSVD, SE3 exp/log and the product kernels being the oracle's own restatements on both sides; FullSystem.cc's driver loop needs the whole front end and stays restated from the cited lines).
// CPU restatement of the immature-point epipolar trace of tum-vision/LDSO:
//   ImmaturePoint::ImmaturePoint   src/internal/ImmaturePoint.cc:14-38   (colour, weights, gradH, energyTH of a candidate)
//   ImmaturePoint::traceOn         src/internal/ImmaturePoint.cc:46-314  (epipolar search + 1-D Gauss-Newton refinement)
//   FullSystem::traceNewCoarse     src/frontend/FullSystem.cc:1012-1050  (per-host KRKi, Kt, affine; the loop over candidates)
// Every statement keeps the reference's float arithmetic and evaluation order.
#pragma once
#include <cmath>
#include <cstring>

namespace oracle {

enum ImmaturePointStatus { IPS_GOOD = 0, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED };   // ImmaturePoint.h:31-38

struct TraceSettings {       // Setting.cc:28,39,41,52,76,81,89-94
    float maxPixSearch = 0.027f, outlierTH = 12 * 12, outlierTHSumComponent = 50 * 50, huberTH = 9, overallEnergyTHWeight = 1;
    int minTraceTestRadius = 2;
    float trace_stepsize = 1.0f; int trace_GNIterations = 3; float trace_GNThreshold = 0.1f, trace_extraSlackOnTH = 1.2f,
          trace_slackInterval = 1.5f, trace_minImprovementFactor = 2;
};

struct ImmaturePt {          // ImmaturePoint.h:103-121
    float u = 0, v = 0;
    float color[8], weights[8];
    float gradH[4] = {0, 0, 0, 0};   // row-major 2x2
    float energyTH = 0;
    float quality = 10000;
    float idepth_min = 0, idepth_max = NAN;
    int lastTraceStatus = IPS_UNINITIALIZED;
    float lastTraceUV[2] = {0, 0};
    float lastTracePixelInterval = 0;
};

// dI: level-0 (I, dx, dy) AoS of the HOST keyframe
void immature_init(ImmaturePt &p, const float *dI_host, int w, float u, float v, const TraceSettings &S);
// dI: level-0 (I, dx, dy) AoS of the frame traced on; KRKi row-major 3x3
int trace_on(ImmaturePt &p, const float *dI, int w, int h, const float KRKi[9], const float Kt[3], const float aff[2], const TraceSettings &S);

}  // namespace oracle
