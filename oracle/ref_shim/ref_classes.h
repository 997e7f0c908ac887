// STAND-INS for the reference classes PointFrameResidual::linearize / fixLinearizationF and ImmaturePoint::* reach into — TEST INFRASTRUCTURE ONLY
// (see NumTypes.h in this directory). The reference's own Residuals.cc, Residuals.h, RawResidualJacobian.h, FrameFramePrecalc.h,
// ResidualProjections.h, GlobalFuncs.h, ImmaturePoint.cc/.h, Feature.h and Setting.cc are compiled UNMODIFIED; the classes below
// replace Frame.h, internal/FrameHessian.h, PointHessian.h, CalibHessian.h and OptimizationBackend/EnergyFunctional.h (whose real definitions pull
// in OpenCV-backed Frame, Sophus, IndexThreadReduce ...) with the members those two functions read, same names and types.
#pragma once
#include "NumTypes.h"
#define LDSO_FRAME_HESSIAN_H_
#define LDSO_POINT_HESSIAN_H_
#define LDSO_CALIB_HESSIAN_H_
#define LDSO_ENERGY_FUNCTIONAL_H_
#define LDSO_FRAME_H_
namespace ldso { namespace internal {
class FrameHessian;
class CalibHessian;
} }
#include "internal/FrameFramePrecalc.h"      // the reference's own struct (PRE_* members)
#include "Settings.h"
#include "AffLight.h"
#include "Feature.h"                         // the reference's own Feature and Point structs (plain data; their .cc files are not needed)
#include "Point.h"
namespace ldso { namespace internal {
class PointFrameResidual;
} }
namespace ldso { namespace internal {
class CalibHessian {          // include/internal/CalibHessian.h:39-69 (accessors only)
public:
    float fx = 0, fy = 0, cx = 0, cy = 0, fxi = 0, fyi = 0;
    float fxl() const { return fx; } float fyl() const { return fy; } float cxl() const { return cx; } float cyl() const { return cy; }
    float fxli() const { return fxi; } float fyli() const { return fyi; }
};
class FrameHessian {          // include/internal/FrameHessian.h:163-201 (members read by linearize)
public:
    int idx = 0;
    Eigen::Vector3f *dI = nullptr;
    std::vector<FrameFramePrecalc> targetPrecalc;
    float frameEnergyTH = 8 * 8 * 8;
    // read by CoarseTracker.cc (FrameHessian.h:32,68,169,179,200-201)
    shared_ptr<Frame> frame;
    Vec3f *dIp[PYR_LEVELS] = {};
    float ab_exposure = 0;
    AffLight aff;
    AffLight aff_g2l() { return aff; }
    SE3 PRE_worldToCam, PRE_camToWorld;
    // read by AccumulatedTopHessianSSE::stitchDouble* (FrameHessian.h: prior, delta_prior, set by takeData())
    Vec8 prior = Vec8::Zero(), delta_prior = Vec8::Zero();
};
class PointHessian {          // include/internal/PointHessian.h:83-107
public:
    float u = 0, v = 0, idepth_scaled = 0, idepth_zero_scaled = 0, deltaF = 0;
    float color[MAX_RES_PER_POINT], weights[MAX_RES_PER_POINT];
    std::pair<shared_ptr<PointFrameResidual>, int /*ResState*/> lastResiduals[2];      // PointHessian.h:103 (ResState is a plain enum: compares with int)
    float HdiF = 0;                                                                     // PointHessian.h:124
    // read / written by AccumulatedTopHessianSSE::addPoint and AccumulatedSCHessianSSE::addPoint (PointHessian.h:100,110-131)
    std::vector<shared_ptr<PointFrameResidual>> residuals;
    float priorF = 0, bdSumF = 0, idepth_hessian = 0, maxRelBaseline = 0;
    float Hdd_accLF = 0, bd_accLF = 0, Hdd_accAF = 0, bd_accAF = 0;
    VecCf Hcd_accLF = VecCf::Zero(), Hcd_accAF = VecCf::Zero();
};
class EnergyFunctional {      // include/internal/OptimizationBackend/EnergyFunctional.h:152,213,222
public:
    int nFrames = 0;
    Mat18f *adHTdeltaF = nullptr;
    VecCf cDeltaF;
    // read by the stitchDouble* functions (EnergyFunctional.h: adHost, adTarget, cPrior, frames)
    Mat88 *adHost = nullptr, *adTarget = nullptr;
    VecC cPrior;
    std::vector<shared_ptr<FrameHessian>> frames;
};
} }
namespace ldso {
struct Frame {                // include/Frame.h (the members ImmaturePoint.cc and CoarseTracker.cc read)
    shared_ptr<internal::FrameHessian> frameHessian;
    unsigned long id = 0;
    std::vector<shared_ptr<Feature>> features;
};
}
// the real PointHessian.h pulls in the residual class: the reference's own Residuals.h (+ RawResidualJacobian.h), unmodified
#include "internal/Residuals.h"
