// ORACLE — TEST INFRASTRUCTURE ONLY (see omath.h header). CoarseInitializer::makeK / calcResAndGS restated; pinned bit for bit against the
// reference's own src/frontend/CoarseInitializer.cc by oracle/ref_pin (the rest of the initializer — point selection, kNN, the LM loop,
// propagateUp/Down — is not restated).
//   include/frontend/CoarseInitializer.h:22-56 (Pnt), :60-160; src/frontend/CoarseInitializer.cc:181-405 (calcResAndGS), :689-715 (makeK)
#pragma once
#include <vector>
#include <array>
#include "tracker.h"

namespace oracle {

struct InitPnt {                 // CoarseInitializer.h:22-56 (the fields calcResAndGS touches)
    float u = 0, v = 0;
    float idepth = 1; bool isGood = true; float energy[2] = {0, 0};
    bool isGood_new = false; float idepth_new = 1; float energy_new[2] = {0, 0};
    float iR = 1, lastHessian = 0, lastHessian_new = 0, maxstep = 0, outlierTH = 0;
};

struct CoarseInitializer {
    Settings S;
    int pyrLevelsUsed = 0;
    int w[PYR_LEVELS], h[PYR_LEVELS];
    double fx[PYR_LEVELS], fy[PYR_LEVELS], cx[PYR_LEVELS], cy[PYR_LEVELS];
    double K[PYR_LEVELS][9], Ki[PYR_LEVELS][9];          // Mat33 (double) in the reference
    const float *firstDIp[PYR_LEVELS], *newDIp[PYR_LEVELS];   // (I, dx, dy) AoS pyramids of firstFrame / newFrame
    std::vector<InitPnt> points[PYR_LEVELS];
    std::vector<std::array<float, 10>> JbBuffer_new;
    float alphaK = 2.5f * 2.5f, alphaW = 150 * 150, couplingWeight = 1;      // trackFrame :44-47
    Accumulator9 acc9, acc9SC;
    CoarseInitializer(int ww, int hh, int levels);
    void makeK(float fxl, float fyl, float cxl, float cyl);                  // :689-715
    // H / b row-major 8x8 / 8; res3 = (E.A, alphaEnergy, E.num)
    void calcResAndGS(int lvl, float H_out[64], float b_out[8], float H_out_sc[64], float b_out_sc[8], const SE3 &refToNew, float aff_a, float aff_b,
                      float res3[3]);                                        // :181-405
};

}  // namespace oracle
