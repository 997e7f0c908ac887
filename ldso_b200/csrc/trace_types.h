// Types and launchers of the immature-point kernels (trace.cu). trace.cu is its own translation unit because it is compiled
// with -fmad=false: ImmaturePoint::traceOn takes discrete decisions (best epipolar step, status thresholds) on float sums, so
// the kernel keeps the reference's separate multiply / add roundings instead of nvcc's contracted FMAs.
#pragma once
#include "common.cuh"

struct TraceSettingsDev {      // Setting.cc:28,39,41,52,76,81,89-94
    float maxPixSearch, outlierTH, outlierTHSumComponent, huberTH, overallEnergyTHWeight;
    int minTraceTestRadius, trace_GNIterations;
    float trace_stepsize, trace_GNThreshold, trace_extraSlackOnTH, trace_slackInterval, trace_minImprovementFactor;
};
enum { IPS_GOOD = 0, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED };   // ImmaturePoint.h:31-38

struct TraceArgs {
    int n, w, h;
    const float4 *img;                 // level 0 of the frame traced on: (I, dx, dy, 0)
    const float *u, *v, *color8, *weights8, *gradH4, *energyTH;
    const int *host;
    const float *KRKi9, *Kt3, *aff2;   // per host keyframe
    float *idepth_min, *idepth_max, *quality;
    int *status;
    float *uv2, *interval;
    TraceSettingsDev S;
};

void launch_immature_init(int n, const float4 *img, int w, const float *u, const float *v, const TraceSettingsDev &S, float *color8,
                          float *weights8, float *gradH4, float *energyTH, cudaStream_t stream);
void launch_trace_on(const TraceArgs &A, cudaStream_t stream);
void launch_optimize_immature(int n, const WinState *ws, const float *u, const float *v, const int *host, const float *idmin, const float *idmax,
                              const float *color8, const float *weights8, const float *energyTH, int minObs, int *ok, float *idepth,
                              unsigned char *res_state, cudaStream_t stream);

// Activation selection (FullSystem::activatePointsMT, FullSystem.cc:1076-1150, with CoarseDistanceMap, CoarseTracker.cc:634-870)
struct ActSelArgs {
    const WinState *ws; int newest;            // index of the newest keyframe in the window
    int w1, h1;                                // size of pyramid level 1
    int nP; const int *pt_host; const float *pt_u, *pt_v, *pt_idepth;       // ACTIVE points of the window: distance-map seeds
    int n;                                     // candidates, in the order the reference visits them
    const float *u, *v, *idmin, *idmax, *quality, *interval, *my_type; const int *status, *host;
    const unsigned char *flagged;              // [nFrames] FrameHessian::flaggedForMarginalization
    float currentMinActDist, minTraceQuality;
    unsigned char *action;                     // [n] out: 0 stays immature, 1 activate, 2 delete
    unsigned char *map;                        // [map_bytes] distance map, one byte per level-1 pixel (255 = the reference's 1000)
    int map_bytes;                             // w1*h1 rounded up to a multiple of 4
    int *front0, *front1;                      // [w1*h1] BFS frontiers of the initial (multi-source) growth
    int *pre_idx; float *pre_frac, *pre_thresh;   // [n] scratch
    int use_smem;                              // the map fits in shared memory
    long long *dbg;                            // clock64() phase stamps (development aid; null = off)
};
#define ACTSEL_THREADS 1024
#define ACTSEL_LOCAL_CAP 512                   // one seed improves at most 8k cells at step k <= 39
size_t actsel_smem_bytes(const ActSelArgs &A);
void launch_activation_select(const ActSelArgs &A, cudaStream_t stream);

// CoarseInitializer::calcResAndGS (src/frontend/CoarseInitializer.cc:181-405) for the points of one pyramid level
#define INIT_NACC 91            // 45 entries of acc9.H (upper triangle, row by row), 45 of acc9SC.H, the energy sum
#define INIT_THREADS 256        // 8 lanes (= pattern pixels) per point, 32 points per CTA
struct InitArgs {
    int n, w, h;
    const float4 *imgRef, *imgNew;            // firstFrame->dIp[lvl], newFrame->dIp[lvl] as (I, dx, dy, 0)
    float RKi[9], t[3], aff0, aff1;           // (R * Ki).cast<float>(), t.cast<float>(), exp(a), b
    float fx, fy, cx, cy, huberTH;
    const float *u, *v, *idepth_new, *iR, *energy2, *outlierTH; const unsigned char *isGood;
    float alphaOpt, couplingWeight;
    unsigned char *isGood_new; float *energy_new2, *maxstep, *lastHessian_new, *Jb;      // per point out (Jb: n*10)
    float *partials; unsigned *counter; double *out;                                      // [grid][INIT_NACC], 1, [INIT_NACC]
};
void launch_init_calc_res(const InitArgs &A, cudaStream_t stream);
