// STAND-INS for the parts of the reference's object model that cannot be compiled here — TEST INFRASTRUCTURE ONLY (see NumTypes.h in
// this directory). The back-end classes are the reference's OWN headers, unmodified: internal/FrameHessian.h, PointHessian.h,
// CalibHessian.h, Residuals.h, RawResidualJacobian.h, FrameFramePrecalc.h, ImmaturePoint.h, OptimizationBackend/*.h, Feature.h, Point.h,
// Camera.h, frontend/CoarseTracker.h. Replaced here: Frame.h (the real one needs DBoW3, Sim3 and OpenCV members; the back-end reads
// id, features, frameHessian and, in one debug branch, imgDisplay) and glog's LOG() macro.
#pragma once
#include "NumTypes.h"
#include "opencv2/opencv.hpp"
#define LDSO_FRAME_H_
struct RefShimNullLog {
    template<typename T> RefShimNullLog &operator<<(const T &) { return *this; }
    RefShimNullLog &operator<<(std::ostream &(*)(std::ostream &)) { return *this; }
};
#define LOG(severity) RefShimNullLog()
namespace ldso {
struct Feature;
struct Point;
namespace internal { class FrameHessian; }
struct Frame {                // include/Frame.h: id, features, frameHessian, imgDisplay
    unsigned long id = 0;
    std::vector<shared_ptr<Feature>> features;
    shared_ptr<internal::FrameHessian> frameHessian;
    cv::Mat imgDisplay;
};
}
#include "internal/FrameHessian.h"
#include "internal/CalibHessian.h"
#include "internal/PointHessian.h"
