// ORACLE — TEST INFRASTRUCTURE ONLY (see omath.h header). PARITY PINNED AGAINST THE REFERENCE'S OWN SOURCES, EXCEPT EIGEN / SOPHUS INTERNALS (oracle/ref_pin compiles, unmodified, the reference's Residuals.cc, ImmaturePoint.cc, PointHessian.cc, FrameHessian.cc, FrameFramePrecalc.cc, AccumulatedTopHessian.cc, AccumulatedSCHessian.cc, EnergyFunctional.cc, CoarseTracker.cc, Setting.cc and their headers against stand-in Eigen/Sophus headers and matches the oracle bit for bit on 108 checks; the float SSE / scalar code is pinned completely, the Eigen-expression code structurally - what is added where, in which order - with LDLT, PartialPivLU, JacobiSVD, SE3 exp/log and the product kernels being the oracle's own restatements on both sides; FullSystem.cc's driver loop needs the whole front end and stays restated from the cited lines).
// CPU restatement of the reference's windowed photometric bundle-adjustment path.
// Every function cites the reference file:line it follows (paths relative to /root/reference).
#include "ba.h"
#include <cassert>
#include <cstdio>

namespace oracle {

// ==========================================================================================
// small float 3x3 helpers (Eigen Mat33f semantics, row-major storage here)
static void m33f_mul(const float *A, const float *B, float *C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float s = A[i * 3 + 0] * B[0 * 3 + j];
            s += A[i * 3 + 1] * B[1 * 3 + j];
            s += A[i * 3 + 2] * B[2 * 3 + j];
            C[i * 3 + j] = s;
        }
}
static void m33f_inverse(const float *m, float *inv) {  // Eigen 3x3 inverse: cofactors * (1/det)
    float c00 = m[4] * m[8] - m[5] * m[7];
    float c01 = m[5] * m[6] - m[3] * m[8];
    float c02 = m[3] * m[7] - m[4] * m[6];
    float det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    float invdet = 1.0f / det;
    inv[0] = c00 * invdet;
    inv[3] = c01 * invdet;
    inv[6] = c02 * invdet;
    inv[1] = (m[2] * m[7] - m[1] * m[8]) * invdet;
    inv[4] = (m[0] * m[8] - m[2] * m[6]) * invdet;
    inv[7] = (m[1] * m[6] - m[0] * m[7]) * invdet;
    inv[2] = (m[1] * m[5] - m[2] * m[4]) * invdet;
    inv[5] = (m[2] * m[3] - m[0] * m[5]) * invdet;
    inv[8] = (m[0] * m[4] - m[1] * m[3]) * invdet;
}

// ==========================================================================================
// CalibHessian::setValue / setValueScaled — include/internal/CalibHessian.h:71-100
void Calib::setValue(const double v[4]) {
    for (int i = 0; i < 4; i++) value[i] = v[i];
    value_scaled[0] = SCALE_F * value[0];
    value_scaled[1] = SCALE_F * value[1];
    value_scaled[2] = SCALE_C * value[2];
    value_scaled[3] = SCALE_C * value[3];
    for (int i = 0; i < 4; i++) value_scaledf[i] = (float) value_scaled[i];
    value_scaledi[0] = 1.0f / value_scaledf[0];
    value_scaledi[1] = 1.0f / value_scaledf[1];
    value_scaledi[2] = -value_scaledf[2] / value_scaledf[0];
    value_scaledi[3] = -value_scaledf[3] / value_scaledf[1];
    for (int i = 0; i < 4; i++) value_minus_value_zero[i] = value[i] - value_zero[i];
}
void Calib::setValueScaled(const double vs[4]) {
    for (int i = 0; i < 4; i++) { value_scaled[i] = vs[i]; value_scaledf[i] = (float) vs[i]; }
    value[0] = SCALE_F_INVERSE * value_scaled[0];
    value[1] = SCALE_F_INVERSE * value_scaled[1];
    value[2] = SCALE_C_INVERSE * value_scaled[2];
    value[3] = SCALE_C_INVERSE * value_scaled[3];
    for (int i = 0; i < 4; i++) value_minus_value_zero[i] = value[i] - value_zero[i];
    value_scaledi[0] = 1.0f / value_scaledf[0];
    value_scaledi[1] = 1.0f / value_scaledf[1];
    value_scaledi[2] = -value_scaledf[2] / value_scaledf[0];
    value_scaledi[3] = -value_scaledf[3] / value_scaledf[1];
}

// ==========================================================================================
// FrameHessian — include/internal/FrameHessian.h:78-154, src/internal/FrameHessian.cc:11-42,108-112
void Frame::setState(const double s[10]) {  // FrameHessian.h:78-91
    for (int i = 0; i < 10; i++) state[i] = s[i];
    for (int i = 0; i < 3; i++) state_scaled[i] = SCALE_XI_TRANS * state[i];
    for (int i = 3; i < 6; i++) state_scaled[i] = SCALE_XI_ROT * state[i];
    state_scaled[6] = SCALE_A * state[6];
    state_scaled[7] = SCALE_B * state[7];
    state_scaled[8] = SCALE_A * state[8];
    state_scaled[9] = SCALE_B * state[9];
    PRE_worldToCam = SE3::exp(state_scaled) * worldToCam_evalPT;
    PRE_camToWorld = PRE_worldToCam.inverse();
}
void Frame::setStateScaled(const double ss[10]) {  // :93-105
    for (int i = 0; i < 10; i++) state_scaled[i] = ss[i];
    for (int i = 0; i < 3; i++) state[i] = SCALE_XI_TRANS_INVERSE * state_scaled[i];
    for (int i = 3; i < 6; i++) state[i] = SCALE_XI_ROT_INVERSE * state_scaled[i];
    state[6] = SCALE_A_INVERSE * state_scaled[6];
    state[7] = SCALE_B_INVERSE * state_scaled[7];
    state[8] = SCALE_A_INVERSE * state_scaled[8];
    state[9] = SCALE_B_INVERSE * state_scaled[9];
    PRE_worldToCam = SE3::exp(state_scaled) * worldToCam_evalPT;
    PRE_camToWorld = PRE_worldToCam.inverse();
}
void Frame::setStateZero(const double sz[10]) {  // FrameHessian.cc:11-42
    for (int i = 0; i < 10; i++) state_zero[i] = sz[i];
    SE3 evInv = worldToCam_evalPT.inverse();
    for (int i = 0; i < 6; i++) {
        double eps[6] = {0, 0, 0, 0, 0, 0}, meps[6] = {0, 0, 0, 0, 0, 0};
        eps[i] = 1e-3;
        meps[i] = -1e-3;
        SE3 P = (worldToCam_evalPT * SE3::exp(eps)) * evInv;
        SE3 M = (worldToCam_evalPT * SE3::exp(meps)) * evInv;
        double lp[6], lm[6];
        P.log(lp);
        M.log(lm);
        for (int r = 0; r < 6; r++) nullspaces_pose[r][i] = (lp[r] - lm[r]) / (2e-3);
    }
    SE3 P = worldToCam_evalPT;
    for (int k = 0; k < 3; k++) P.t[k] *= 1.00001;
    P = P * evInv;
    SE3 M = worldToCam_evalPT;
    for (int k = 0; k < 3; k++) M.t[k] /= 1.00001;
    M = M * evInv;
    double lp[6], lm[6];
    P.log(lp);
    M.log(lm);
    for (int r = 0; r < 6; r++) nullspaces_scale[r] = (lp[r] - lm[r]) / (2e-3);
    memset(nullspaces_affine, 0, sizeof(nullspaces_affine));
    nullspaces_affine[0][0] = 1;
    nullspaces_affine[1][0] = 0;
    float a0, b0;
    aff_g2l_0(a0, b0);
    nullspaces_affine[0][1] = 0;
    nullspaces_affine[1][1] = expf(a0) * ab_exposure;
}
void Frame::getPrior(const Settings &S, double p[10]) const {  // FrameHessian.h:125-150
    for (int i = 0; i < 10; i++) p[i] = 0;
    if (id == 0) {
        p[0] = p[1] = p[2] = S.initialTransPrior;
        p[3] = p[4] = p[5] = S.initialRotPrior;
        p[6] = S.initialAffAPrior;
        p[7] = S.initialAffBPrior;
    } else {
        p[6] = (S.affineOptModeA < 0) ? S.initialAffAPrior : S.affineOptModeA;
        p[7] = (S.affineOptModeB < 0) ? S.initialAffBPrior : S.affineOptModeB;
    }
    p[8] = S.initialAffAPrior;
    p[9] = S.initialAffBPrior;
}
void Frame::takeData(const Settings &S) {  // FrameHessian.cc:108-112
    double p[10];
    getPrior(S, p);
    for (int i = 0; i < 8; i++) {
        prior[i] = p[i];
        delta[i] = state[i] - state_zero[i];
        delta_prior[i] = state[i] - 0.0;  // getPriorZero() == 0
    }
}

// ==========================================================================================
// makeImages — src/internal/FrameHessian.cc:44-98
void makeImages(const float *color, int w, int h, int levels, float **dIp) {
    for (int i = 0; i < w * h; i++) { dIp[0][3 * i] = color[i]; dIp[0][3 * i + 1] = 0; dIp[0][3 * i + 2] = 0; }
    for (int lvl = 0; lvl < levels; lvl++) {
        int wl = w >> lvl, hl = h >> lvl;
        float *dI_l = dIp[lvl];
        if (lvl > 0) {
            int wlm1 = w >> (lvl - 1);
            const float *dI_lm = dIp[lvl - 1];
            for (int y = 0; y < hl; y++)
                for (int x = 0; x < wl; x++) {
                    dI_l[3 * (x + y * wl)] = 0.25f * (dI_lm[3 * (2 * x + 2 * y * wlm1)] +
                                                     dI_lm[3 * (2 * x + 1 + 2 * y * wlm1)] +
                                                     dI_lm[3 * (2 * x + 2 * y * wlm1 + wlm1)] +
                                                     dI_lm[3 * (2 * x + 1 + 2 * y * wlm1 + wlm1)]);
                    dI_l[3 * (x + y * wl) + 1] = 0;
                    dI_l[3 * (x + y * wl) + 2] = 0;
                }
        }
        for (int idx = wl; idx < wl * (hl - 1); idx++) {
            float dx = 0.5f * (dI_l[3 * (idx + 1)] - dI_l[3 * (idx - 1)]);
            float dy = 0.5f * (dI_l[3 * (idx + wl)] - dI_l[3 * (idx - wl)]);
            if (std::isnan(dx) || std::fabs(dx) > 255.0) dx = 0;
            if (std::isnan(dy) || std::fabs(dy) > 255.0) dy = 0;
            dI_l[3 * idx + 1] = dx;
            dI_l[3 * idx + 2] = dy;
        }
    }
}

// ==========================================================================================
// ThreadReduce
ThreadReduce::ThreadReduce(bool spawn) : threaded(spawn) {
    memset(stats, 0, sizeof(stats));
    if (threaded)
        for (int i = 0; i < NUM_THREADS; i++) workers[i] = std::thread(&ThreadReduce::loop, this, i);
}
ThreadReduce::~ThreadReduce() {
    if (threaded) {
        {
            std::unique_lock<std::mutex> lk(mtx);
            running = false;
            generation++;
        }
        cv_go.notify_all();
        for (int i = 0; i < NUM_THREADS; i++) workers[i].join();
    }
}
void ThreadReduce::run_tid(int tid) {
    // worker `tid` executes chunks tid, tid+6, ... ; a worker that gets no chunk still calls
    // fn(0,0,..) once (IndexThreadReduce.h:145-153: used for the per-thread setZero calls).
    int nchunks = (cur_end - cur_first + cur_step - 1) / cur_step;
    if (cur_end <= cur_first) nchunks = 0;
    bool got = false;
    for (int c = tid; c < nchunks; c += NUM_THREADS) {
        int todo = cur_first + c * cur_step;
        double s[10];
        memset(s, 0, sizeof(s));
        (*cur)(todo, std::min(todo + cur_step, cur_end), s, tid);
        for (int k = 0; k < 10; k++) chunk_stats[(size_t) c * 10 + k] = s[k];
        got = true;
    }
    if (!got) {
        double s[10];
        memset(s, 0, sizeof(s));
        (*cur)(0, 0, s, tid);
    }
}
void ThreadReduce::loop(int tid) {
    unsigned long seen = 0;
    while (true) {
        {
            std::unique_lock<std::mutex> lk(mtx);
            cv_go.wait(lk, [&] { return generation != seen; });
            seen = generation;
            if (!running) return;
        }
        run_tid(tid);
        {
            std::unique_lock<std::mutex> lk(mtx);
            n_done++;
            if (n_done == NUM_THREADS) cv_done.notify_all();
        }
    }
}
void ThreadReduce::reduce(const Fn &fn, int first, int end, int stepSize) {
    memset(stats, 0, sizeof(stats));
    if (stepSize == 0) stepSize = ((end - first) + NUM_THREADS - 1) / NUM_THREADS;
    if (stepSize <= 0) stepSize = 1;
    cur = &fn;
    cur_first = first;
    cur_end = end;
    cur_step = stepSize;
    int nchunks = end > first ? (end - first + stepSize - 1) / stepSize : 0;
    chunk_stats.assign((size_t) nchunks * 10, 0.0);
    if (threaded) {
        {
            std::unique_lock<std::mutex> lk(mtx);
            n_done = 0;
            generation++;
        }
        cv_go.notify_all();
        {
            std::unique_lock<std::mutex> lk(mtx);
            cv_done.wait(lk, [&] { return n_done == NUM_THREADS; });
        }
    } else {
        for (int tid = 0; tid < NUM_THREADS; tid++) run_tid(tid);
    }
    for (int c = 0; c < nchunks; c++)
        for (int k = 0; k < 10; k++) stats[k] += chunk_stats[(size_t) c * 10 + k];
    cur = nullptr;
}

// ==========================================================================================
Window::Window(int w, int h, int nthreads_mode) {
    wG0 = w;
    hG0 = h;
    wM3G = w - 3;  // GlobalCalib.cc:42-43
    hM3G = h - 3;
    S.multiThreading = nthreads_mode != 1;
    red = new ThreadReduce(nthreads_mode == NUM_THREADS);  // 6: real threads; 0: serial emulation of 6; 1: MT off
    HM = MatX(CPARS, CPARS);
    bM.assign(CPARS, 0.0);
}
Window::~Window() { delete red; }

// ------------------------------------------------------------------------------------------
// FrameFramePrecalc::Set — src/internal/FrameFramePrecalc.cc:6-35
void Window::precalcSet(FramePrecalc &pc, const Frame &host, const Frame &target) {
    SE3 leftToLeft_0 = target.worldToCam_evalPT * host.worldToCam_evalPT.inverse();
    M3 R0 = leftToLeft_0.rotationMatrix();
    for (int i = 0; i < 9; i++) pc.PRE_RTll_0[i] = (float) R0.m[i];
    for (int i = 0; i < 3; i++) pc.PRE_tTll_0[i] = (float) leftToLeft_0.t[i];

    SE3 leftToLeft = target.PRE_worldToCam * host.PRE_camToWorld;
    M3 R = leftToLeft.rotationMatrix();
    for (int i = 0; i < 9; i++) pc.PRE_RTll[i] = (float) R.m[i];
    for (int i = 0; i < 3; i++) pc.PRE_tTll[i] = (float) leftToLeft.t[i];
    pc.distanceLL = (float) std::sqrt(leftToLeft.t[0] * leftToLeft.t[0] + leftToLeft.t[1] * leftToLeft.t[1] +
                                      leftToLeft.t[2] * leftToLeft.t[2]);

    float K[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    K[0] = HCalib.fxl();
    K[4] = HCalib.fyl();
    K[2] = HCalib.cxl();
    K[5] = HCalib.cyl();
    K[8] = 1;
    float Ki[9], tmp[9];
    m33f_inverse(K, Ki);
    m33f_mul(K, pc.PRE_RTll, tmp);
    m33f_mul(tmp, Ki, pc.PRE_KRKiTll);
    m33f_mul(pc.PRE_RTll, Ki, pc.PRE_RKiTll);
    for (int i = 0; i < 3; i++) {
        float s = K[i * 3 + 0] * pc.PRE_tTll[0];
        s += K[i * 3 + 1] * pc.PRE_tTll[1];
        s += K[i * 3 + 2] * pc.PRE_tTll[2];
        pc.PRE_KtTll[i] = s;
    }
    float ah, bh, at, bt;
    host.aff_g2l(ah, bh);
    target.aff_g2l(at, bt);
    double ab[2];
    fromToVecExposure(host.ab_exposure, target.ab_exposure, ah, bh, at, bt, ab);
    pc.PRE_aff_mode[0] = (float) ab[0];
    pc.PRE_aff_mode[1] = (float) ab[1];
    float a0, b0;
    host.aff_g2l_0(a0, b0);
    pc.PRE_b0_mode = b0;
}

// FullSystem::setPrecalcValues — src/frontend/FullSystem.cc:1423-1431
void Window::setPrecalcValues() {
    for (auto &fr : frames) {
        fr.targetPrecalc.resize(frames.size());
        for (size_t i = 0; i < frames.size(); i++) precalcSet(fr.targetPrecalc[i], fr, frames[i]);
    }
    setDeltaF();
}

// ------------------------------------------------------------------------------------------
// projectPointA / projectPointB (ResidualProjections.h:24-33, :57-84) live in ba.h (also used by oracle/ref_pin)

// PointFrameResidual::linearize — src/internal/Residuals.cc:13-214
double Window::linearize(Residual &r) {
    r.state_NewEnergyWithOutlier = -1;
    if (r.state_state == RS_OOB) {
        r.state_NewState = RS_OOB;
        return r.state_energy;
    }
    const Frame &f = frames[r.host];
    const Frame &ftarget = frames[r.target];
    const Point &p = points[r.point];
    const FramePrecalc *precalc = &f.targetPrecalc[ftarget.idx];

    float energyLeft = 0;
    const float *dIl = ftarget.dI;
    const float *PRE_KRKiTll = precalc->PRE_KRKiTll;
    const float *PRE_KtTll = precalc->PRE_KtTll;
    const float *PRE_RTll_0 = precalc->PRE_RTll_0;
    const float *PRE_tTll_0 = precalc->PRE_tTll_0;
    const float *color = p.color;
    const float *weights = p.weights;
    const float affLL[2] = {precalc->PRE_aff_mode[0], precalc->PRE_aff_mode[1]};
    float b0 = precalc->PRE_b0_mode;
    RawResidualJacobian *J = &r.J;

    float d_xi_x[6], d_xi_y[6], d_C_x[4], d_C_y[4], d_d_x, d_d_y;
    {
        float drescale, u, v, new_idepth, Ku, Kv, KliP[3];
        if (!projectPointB(p.u, p.v, p.idepth_zero_scaled, 0, 0, HCalib, PRE_RTll_0, PRE_tTll_0, wM3G, hM3G,
                           drescale, u, v, Ku, Kv, KliP, new_idepth)) {
            r.state_NewState = RS_OOB;
            return r.state_energy;
        }
        r.centerProjectedTo[0] = Ku;
        r.centerProjectedTo[1] = Kv;
        r.centerProjectedTo[2] = new_idepth;
#define R0(i, j) PRE_RTll_0[(i) * 3 + (j)]
        d_d_x = drescale * (PRE_tTll_0[0] - PRE_tTll_0[2] * u) * SCALE_IDEPTH * HCalib.fxl();
        d_d_y = drescale * (PRE_tTll_0[1] - PRE_tTll_0[2] * v) * SCALE_IDEPTH * HCalib.fyl();

        d_C_x[2] = drescale * (R0(2, 0) * u - R0(0, 0));
        d_C_x[3] = HCalib.fxl() * drescale * (R0(2, 1) * u - R0(0, 1)) * HCalib.fyli();
        d_C_x[0] = KliP[0] * d_C_x[2];
        d_C_x[1] = KliP[1] * d_C_x[3];

        d_C_y[2] = HCalib.fyl() * drescale * (R0(2, 0) * v - R0(1, 0)) * HCalib.fxli();
        d_C_y[3] = drescale * (R0(2, 1) * v - R0(1, 1));
        d_C_y[0] = KliP[0] * d_C_y[2];
        d_C_y[1] = KliP[1] * d_C_y[3];
#undef R0
        d_C_x[0] = (d_C_x[0] + u) * SCALE_F;
        d_C_x[1] *= SCALE_F;
        d_C_x[2] = (d_C_x[2] + 1) * SCALE_C;
        d_C_x[3] *= SCALE_C;

        d_C_y[0] *= SCALE_F;
        d_C_y[1] = (d_C_y[1] + v) * SCALE_F;
        d_C_y[2] *= SCALE_C;
        d_C_y[3] = (d_C_y[3] + 1) * SCALE_C;

        d_xi_x[0] = new_idepth * HCalib.fxl();
        d_xi_x[1] = 0;
        d_xi_x[2] = -new_idepth * u * HCalib.fxl();
        d_xi_x[3] = -u * v * HCalib.fxl();
        d_xi_x[4] = (1 + u * u) * HCalib.fxl();
        d_xi_x[5] = -v * HCalib.fxl();

        d_xi_y[0] = 0;
        d_xi_y[1] = new_idepth * HCalib.fyl();
        d_xi_y[2] = -new_idepth * v * HCalib.fyl();
        d_xi_y[3] = -(1 + v * v) * HCalib.fyl();
        d_xi_y[4] = u * v * HCalib.fyl();
        d_xi_y[5] = u * HCalib.fyl();
    }
    for (int i = 0; i < 6; i++) { J->Jpdxi[0][i] = d_xi_x[i]; J->Jpdxi[1][i] = d_xi_y[i]; }
    for (int i = 0; i < 4; i++) { J->Jpdc[0][i] = d_C_x[i]; J->Jpdc[1][i] = d_C_y[i]; }
    J->Jpdd[0] = d_d_x;
    J->Jpdd[1] = d_d_y;

    float JIdxJIdx_00 = 0, JIdxJIdx_11 = 0, JIdxJIdx_10 = 0;
    float JabJIdx_00 = 0, JabJIdx_01 = 0, JabJIdx_10 = 0, JabJIdx_11 = 0;
    float JabJab_00 = 0, JabJab_01 = 0, JabJab_11 = 0;
    float wJI2_sum = 0;

    for (int idx = 0; idx < patternNum; idx++) {
        float Ku, Kv;
        if (!projectPointA(p.u + patternP[idx][0], p.v + patternP[idx][1], p.idepth_scaled, PRE_KRKiTll, PRE_KtTll,
                           wM3G, hM3G, Ku, Kv)) {
            r.state_NewState = RS_OOB;
            return r.state_energy;
        }
        r.projectedTo[idx][0] = Ku;
        r.projectedTo[idx][1] = Kv;

        float hitColor[3];
        getInterpolatedElement33(dIl, Ku, Kv, wG0, hitColor);
        float residual = hitColor[0] - (float) (affLL[0] * color[idx] + affLL[1]);

        float drdA = (color[idx] - b0);
        if (!std::isfinite((float) hitColor[0])) {
            r.state_NewState = RS_OOB;
            return r.state_energy;
        }
        float w = sqrtf(S.outlierTHSumComponent /
                        (S.outlierTHSumComponent + (hitColor[1] * hitColor[1] + hitColor[2] * hitColor[2])));
        w = 0.5f * (w + weights[idx]);

        float hw = fabsf(residual) < S.huberTH ? 1 : S.huberTH / fabsf(residual);
        energyLeft += w * w * hw * residual * residual * (2 - hw);
        {
            if (hw < 1) hw = sqrtf(hw);
            hw = hw * w;

            hitColor[1] *= hw;
            hitColor[2] *= hw;

            J->resF[idx] = residual * hw;
            J->JIdx[0][idx] = hitColor[1];
            J->JIdx[1][idx] = hitColor[2];
            J->JabF[0][idx] = drdA * hw;
            J->JabF[1][idx] = hw;

            JIdxJIdx_00 += hitColor[1] * hitColor[1];
            JIdxJIdx_11 += hitColor[2] * hitColor[2];
            JIdxJIdx_10 += hitColor[1] * hitColor[2];

            JabJIdx_00 += drdA * hw * hitColor[1];
            JabJIdx_01 += drdA * hw * hitColor[2];
            JabJIdx_10 += hw * hitColor[1];
            JabJIdx_11 += hw * hitColor[2];

            JabJab_00 += drdA * drdA * hw * hw;
            JabJab_01 += drdA * hw * hw;
            JabJab_11 += hw * hw;

            wJI2_sum += hw * hw * (hitColor[1] * hitColor[1] + hitColor[2] * hitColor[2]);

            if (S.affineOptModeA < 0) J->JabF[0][idx] = 0;
            if (S.affineOptModeB < 0) J->JabF[1][idx] = 0;
        }
    }
    J->JIdx2[0] = JIdxJIdx_00; J->JIdx2[1] = JIdxJIdx_10; J->JIdx2[2] = JIdxJIdx_10; J->JIdx2[3] = JIdxJIdx_11;
    J->JabJIdx[0] = JabJIdx_00; J->JabJIdx[1] = JabJIdx_01; J->JabJIdx[2] = JabJIdx_10; J->JabJIdx[3] = JabJIdx_11;
    J->Jab2[0] = JabJab_00; J->Jab2[1] = JabJab_01; J->Jab2[2] = JabJab_01; J->Jab2[3] = JabJab_11;

    r.state_NewEnergyWithOutlier = energyLeft;

    if (energyLeft > std::max<float>(f.frameEnergyTH, ftarget.frameEnergyTH) || wJI2_sum < 2) {
        energyLeft = std::max<float>(f.frameEnergyTH, ftarget.frameEnergyTH);
        r.state_NewState = RS_OUTLIER;
    } else {
        r.state_NewState = RS_IN;
    }
    r.state_NewEnergy = energyLeft;
    return energyLeft;
}

// PointFrameResidual::takeData — include/internal/Residuals.h:123-128
void Window::takeData(Residual &r) {
    const RawResidualJacobian *J = &r.J;
    float JI_JI_Jd[2];
    JI_JI_Jd[0] = J->JIdx2[0] * J->Jpdd[0] + J->JIdx2[1] * J->Jpdd[1];
    JI_JI_Jd[1] = J->JIdx2[2] * J->Jpdd[0] + J->JIdx2[3] * J->Jpdd[1];
    for (int i = 0; i < 6; i++) r.JpJdF[i] = J->Jpdxi[0][i] * JI_JI_Jd[0] + J->Jpdxi[1][i] * JI_JI_Jd[1];
    r.JpJdF[6] = J->JabJIdx[0] * J->Jpdd[0] + J->JabJIdx[1] * J->Jpdd[1];
    r.JpJdF[7] = J->JabJIdx[2] * J->Jpdd[0] + J->JabJIdx[3] * J->Jpdd[1];
}

// PointFrameResidual::applyRes — include/internal/Residuals.h:70-87
void Window::applyRes(Residual &r, bool copyJacobians) {
    if (copyJacobians) {
        if (r.state_state == RS_OOB) return;
        if (r.state_NewState == RS_IN) {
            r.isActiveAndIsGoodNEW = true;
            takeData(r);
        } else {
            r.isActiveAndIsGoodNEW = false;
        }
    }
    r.state_state = r.state_NewState;
    r.state_energy = r.state_NewEnergy;
}

// PointFrameResidual::fixLinearizationF — src/internal/Residuals.cc:216-242
void Window::fixLinearizationF(Residual &r) {
    const float *dp = &adHTdeltaF[(size_t) (r.hostIDX + nFrames * r.targetIDX) * 8];
    const RawResidualJacobian *J = &r.J;
    const Point &p = points[r.point];
    float dx = 0, dy = 0;
    {
        float a = 0, b = 0;
        for (int i = 0; i < 6; i++) a += J->Jpdxi[0][i] * dp[i];
        for (int i = 0; i < 4; i++) b += J->Jpdc[0][i] * cDeltaF[i];
        dx = a + b + J->Jpdd[0] * p.deltaF;
        a = 0; b = 0;
        for (int i = 0; i < 6; i++) a += J->Jpdxi[1][i] * dp[i];
        for (int i = 0; i < 4; i++) b += J->Jpdc[1][i] * cDeltaF[i];
        dy = a + b + J->Jpdd[1] * p.deltaF;
    }
    float delta_a = dp[6], delta_b = dp[7];
    for (int i = 0; i < patternNum; i++) {
        float rtz = J->resF[i];
        rtz = rtz - J->JIdx[0][i] * dx;
        rtz = rtz - J->JIdx[1][i] * dy;
        rtz = rtz - J->JabF[0][i] * delta_a;
        rtz = rtz - J->JabF[1][i] * delta_b;
        r.res_toZeroF[i] = rtz;
    }
    r.isLinearized = true;
}

// ------------------------------------------------------------------------------------------
// FullSystem::linearizeAll (+ _Reductor) — src/frontend/FullSystem.cc:1442-1543
// (fixLinearization: the removal of non-IN residuals from the point is reported by marking them
//  dropped: they are erased from Point::residuals like EnergyFunctional::dropResidual does.)
double Window::linearizeAll(bool fixLinearization) {
    std::vector<int> toRemove[NUM_THREADS];
    auto reductor = [&](int min, int max, double *stats, int tid) {
        for (int k = min; k < max; k++) {
            Residual &r = residuals[activeResiduals[k]];
            stats[0] += linearize(r);
            if (fixLinearization) {
                applyRes(r, true);
                if (r.isActive()) {
                    if (r.isNew) {
                        Point &p = points[r.point];
                        const FramePrecalc &pc = frames[r.host].targetPrecalc[frames[r.target].idx];
                        float inf[3], ptp[3];
                        for (int i = 0; i < 3; i++) {
                            float s = pc.PRE_KRKiTll[i * 3] * p.u;
                            s += pc.PRE_KRKiTll[i * 3 + 1] * p.v;
                            s += pc.PRE_KRKiTll[i * 3 + 2] * 1.0f;
                            inf[i] = s;
                            ptp[i] = s + pc.PRE_KtTll[i] * p.idepth_scaled;
                        }
                        float ex = inf[0] / inf[2] - ptp[0] / ptp[2], ey = inf[1] / inf[2] - ptp[1] / ptp[2];
                        float relBS = 0.01 * std::sqrt(ex * ex + ey * ey);
                        if (relBS > p.maxRelBaseline) p.maxRelBaseline = relBS;
                    }
                } else {
                    toRemove[tid].push_back(activeResiduals[k]);
                }
            }
        }
    };
    if (S.multiThreading) {
        red->reduce(reductor, 0, (int) activeResiduals.size(), 0);
        lastEnergyP = red->stats[0];
    } else {
        double stats[10];
        memset(stats, 0, sizeof(stats));
        reductor(0, (int) activeResiduals.size(), stats, 0);
        lastEnergyP = stats[0];
    }
    setNewFrameEnergyTH();
    if (fixLinearization) {
        for (int i = 0; i < NUM_THREADS; i++)
            for (int ri : toRemove[i]) {
                Point &p = points[residuals[ri].point];
                p.residuals.erase(std::find(p.residuals.begin(), p.residuals.end(), ri));  // ef->dropResidual
            }
    }
    return lastEnergyP;
}

// FullSystem::applyRes_Reductor — FullSystem.cc:1706-1709
void Window::applyResAll() {
    for (int k : activeResiduals) applyRes(residuals[k], true);
}

// FullSystem::setNewFrameEnergyTH — FullSystem.cc:1762-1793
void Window::setNewFrameEnergyTH() {
    std::vector<float> allResVec;
    allResVec.reserve(activeResiduals.size() * 2);
    Frame &newFrame = frames.back();
    int newIdx = (int) frames.size() - 1;
    for (int k : activeResiduals) {
        const Residual &r = residuals[k];
        if (r.state_NewEnergyWithOutlier >= 0 && r.target == newIdx) allResVec.push_back((float) r.state_NewEnergyWithOutlier);
    }
    if (allResVec.size() == 0) {
        newFrame.frameEnergyTH = 12 * 12 * patternNum;
        return;
    }
    int nthIdx = S.frameEnergyTHN * allResVec.size();
    std::nth_element(allResVec.begin(), allResVec.begin() + nthIdx, allResVec.end());
    float nthElement = sqrtf(allResVec[nthIdx]);
    newFrame.frameEnergyTH = nthElement * S.frameEnergyTHFacMedian;
    newFrame.frameEnergyTH = 26.0f * S.frameEnergyTHConstWeight + newFrame.frameEnergyTH * (1 - S.frameEnergyTHConstWeight);
    newFrame.frameEnergyTH = newFrame.frameEnergyTH * newFrame.frameEnergyTH;
    newFrame.frameEnergyTH *= S.overallEnergyTHWeight * S.overallEnergyTHWeight;
}

// FullSystem::backupState (non-momentum branch) — FullSystem.cc:1662-1676
void Window::backupState() {
    for (int i = 0; i < 4; i++) HCalib.value_backup[i] = HCalib.value[i];
    for (auto &fh : frames) for (int i = 0; i < 10; i++) fh.state_backup[i] = fh.state[i];
    for (auto &ph : points) ph.idepth_backup = ph.idepth;
}

// FullSystem::doStepFromBackup (non-momentum branch) — FullSystem.cc:1587-1622
bool Window::doStepFromBackup(float stepfacC, float stepfacT, float stepfacR, float stepfacA, float stepfacD) {
    double pstepfac[10];
    for (int i = 0; i < 3; i++) pstepfac[i] = stepfacT;
    for (int i = 3; i < 6; i++) pstepfac[i] = stepfacR;
    for (int i = 6; i < 10; i++) pstepfac[i] = stepfacA;
    float sumA = 0, sumB = 0, sumT = 0, sumR = 0, sumID = 0, numID = 0;
    float sumNID = 0;
    {
        double nv[4];
        for (int i = 0; i < 4; i++) nv[i] = HCalib.value_backup[i] + stepfacC * HCalib.step[i];
        HCalib.setValue(nv);
    }
    for (auto &fh : frames) {
        double ns[10];
        for (int i = 0; i < 10; i++) ns[i] = fh.state_backup[i] + pstepfac[i] * fh.step[i];
        fh.setState(ns);
        sumA += fh.step[6] * fh.step[6];
        sumB += fh.step[7] * fh.step[7];
        sumT += fh.step[0] * fh.step[0] + fh.step[1] * fh.step[1] + fh.step[2] * fh.step[2];
        sumR += fh.step[3] * fh.step[3] + fh.step[4] * fh.step[4] + fh.step[5] * fh.step[5];
    }
    // points are visited frame by frame in the reference; `points` is already in that order.
    for (auto &ph : points) {
        ph.setIdepth(ph.idepth_backup + stepfacD * ph.step);
        sumID += ph.step * ph.step;
        sumNID += fabsf(ph.idepth_backup);
        numID++;
        ph.setIdepthZero(ph.idepth_backup + stepfacD * ph.step);
    }
    sumA /= frames.size();
    sumB /= frames.size();
    sumR /= frames.size();
    sumT /= frames.size();
    sumID /= numID;
    sumNID /= numID;
    setPrecalcValues();
    return sqrtf(sumA) < 0.0005 * S.thOptIterations &&
           sqrtf(sumB) < 0.00005 * S.thOptIterations &&
           sqrtf(sumR) < 0.00005 * S.thOptIterations &&
           sqrtf(sumT) * sumNID < 0.00005 * S.thOptIterations;
}

// FullSystem::getNullspaces — FullSystem.cc:1711-1760
void Window::getNullspaces() {
    lastNullspaces_pose.clear();
    lastNullspaces_scale.clear();
    lastNullspaces_affA.clear();
    lastNullspaces_affB.clear();
    int n = CPARS + (int) frames.size() * 8;
    for (int i = 0; i < 6; i++) {
        VecXd ns(n, 0.0);
        for (auto &fh : frames) {
            for (int r = 0; r < 6; r++) ns[CPARS + fh.idx * 8 + r] = fh.nullspaces_pose[r][i];
            for (int r = 0; r < 3; r++) ns[CPARS + fh.idx * 8 + r] *= SCALE_XI_TRANS_INVERSE;
            for (int r = 3; r < 6; r++) ns[CPARS + fh.idx * 8 + r] *= SCALE_XI_ROT_INVERSE;
        }
        lastNullspaces_pose.push_back(ns);
    }
    for (int i = 0; i < 2; i++) {
        VecXd ns(n, 0.0);
        for (auto &fh : frames) {
            ns[CPARS + fh.idx * 8 + 6] = fh.nullspaces_affine[0][i];
            ns[CPARS + fh.idx * 8 + 7] = fh.nullspaces_affine[1][i];
            ns[CPARS + fh.idx * 8 + 6] *= SCALE_A_INVERSE;
            ns[CPARS + fh.idx * 8 + 7] *= SCALE_B_INVERSE;
        }
        if (i == 0) lastNullspaces_affA.push_back(ns);
        if (i == 1) lastNullspaces_affB.push_back(ns);
    }
    VecXd ns(n, 0.0);
    for (auto &fh : frames) {
        for (int r = 0; r < 6; r++) ns[CPARS + fh.idx * 8 + r] = fh.nullspaces_scale[r];
        for (int r = 0; r < 3; r++) ns[CPARS + fh.idx * 8 + r] *= SCALE_XI_TRANS_INVERSE;
        for (int r = 3; r < 6; r++) ns[CPARS + fh.idx * 8 + r] *= SCALE_XI_ROT_INVERSE;
    }
    lastNullspaces_scale.push_back(ns);
}

// FullSystem::optimize prologue — FullSystem.cc:734-771
void Window::optimizeBegin() {
    activeResiduals.clear();
    for (auto &ph : points)
        for (int ri : ph.residuals) {
            Residual &r = residuals[ri];
            if (!r.isLinearized) {
                activeResiduals.push_back(ri);
                r.resetOOB();
            }
        }
    linearizeAll(false);
    applyResAll();
}

// FullSystem::optimize loop body — FullSystem.cc:777-831 with setting_forceAceptStep=true
// (Setting.cc:73) and no SOLVER_STEPMOMENTUM (stepsize = 1).
bool Window::gnIteration(int iteration) {
    backupState();
    getNullspaces();  // FullSystem::solveSystem :1433-1440
    solveSystemF(iteration, 1e-1);
    bool canbreak = doStepFromBackup(1, 1, 1, 1, 1);
    linearizeAll(false);
    applyResAll();
    return canbreak;
}

// ==========================================================================================
// EnergyFunctional

// insertFrame x nF — EnergyFunctional.cc:30-61 (HM/bM are resized with zeros)
void Window::insertFrames() {
    nFrames = (int) frames.size();
    for (int i = 0; i < nFrames; i++) {
        frames[i].idx = i;
        frames[i].takeData(S);
    }
    int n = 8 * nFrames + CPARS;
    if (HM.r != n) {
        HM = MatX(n, n);
        bM.assign(n, 0.0);
    }
    setAdjointsF();
    makeIDX();
}

// EnergyFunctional::makeIDX — :385-401
void Window::makeIDX() {
    for (size_t i = 0; i < frames.size(); i++) frames[i].idx = (int) i;
    for (auto &p : points)
        for (int ri : p.residuals) {
            residuals[ri].hostIDX = frames[residuals[ri].host].idx;
            residuals[ri].targetIDX = frames[residuals[ri].target].idx;
        }
}

// EnergyFunctional::setAdjointsF — :431-489
void Window::setAdjointsF() {
    adHost.assign((size_t) nFrames * nFrames * 64, 0.0);
    adTarget.assign((size_t) nFrames * nFrames * 64, 0.0);
    for (int h = 0; h < nFrames; h++)
        for (int t = 0; t < nFrames; t++) {
            const Frame &host = frames[h], &target = frames[t];
            SE3 hostToTarget = target.worldToCam_evalPT * host.worldToCam_evalPT.inverse();
            double AH[64], AT[64];
            memset(AH, 0, sizeof(AH));
            memset(AT, 0, sizeof(AT));
            for (int i = 0; i < 8; i++) AH[i * 8 + i] = AT[i * 8 + i] = 1;
            double Adj[36];
            hostToTarget.Adj(Adj);
            for (int i = 0; i < 6; i++)
                for (int j = 0; j < 6; j++) AH[i * 8 + j] = -Adj[j * 6 + i];  // -Adj^T
            float a0h, b0h, a0t, b0t;
            host.aff_g2l_0(a0h, b0h);
            target.aff_g2l_0(a0t, b0t);
            double ab[2];
            fromToVecExposure(host.ab_exposure, target.ab_exposure, a0h, b0h, a0t, b0t, ab);
            float affLL0 = (float) ab[0];
            AT[6 * 8 + 6] = -affLL0;
            AH[6 * 8 + 6] = affLL0;
            AT[7 * 8 + 7] = -1;
            AH[7 * 8 + 7] = affLL0;
            for (int j = 0; j < 8; j++) {
                for (int i = 0; i < 3; i++) { AH[i * 8 + j] *= SCALE_XI_TRANS; AT[i * 8 + j] *= SCALE_XI_TRANS; }
                for (int i = 3; i < 6; i++) { AH[i * 8 + j] *= SCALE_XI_ROT; AT[i * 8 + j] *= SCALE_XI_ROT; }
                AH[6 * 8 + j] *= SCALE_A; AT[6 * 8 + j] *= SCALE_A;
                AH[7 * 8 + j] *= SCALE_B; AT[7 * 8 + j] *= SCALE_B;
            }
            memcpy(&adHost[(size_t) (h + t * nFrames) * 64], AH, sizeof(AH));
            memcpy(&adTarget[(size_t) (h + t * nFrames) * 64], AT, sizeof(AT));
        }
    for (int i = 0; i < 4; i++) { cPrior[i] = S.initialCalibHessian; cPriorF[i] = (float) cPrior[i]; }
    adHostF.resize(adHost.size());
    adTargetF.resize(adTarget.size());
    for (size_t i = 0; i < adHost.size(); i++) { adHostF[i] = (float) adHost[i]; adTargetF[i] = (float) adTarget[i]; }
}

// EnergyFunctional::setDeltaF — :403-429
void Window::setDeltaF() {
    adHTdeltaF.assign((size_t) nFrames * nFrames * 8, 0.0f);
    for (int h = 0; h < nFrames; h++)
        for (int t = 0; t < nFrames; t++) {
            int idx = h + t * nFrames;
            float dh[8], dt[8];
            for (int i = 0; i < 8; i++) {
                dh[i] = (float) (frames[h].state[i] - frames[h].state_zero[i]);
                dt[i] = (float) (frames[t].state[i] - frames[t].state_zero[i]);
            }
            const float *AH = &adHostF[(size_t) idx * 64], *AT = &adTargetF[(size_t) idx * 64];
            for (int j = 0; j < 8; j++) {
                float a = 0, b = 0;
                for (int i = 0; i < 8; i++) a += dh[i] * AH[i * 8 + j];
                for (int i = 0; i < 8; i++) b += dt[i] * AT[i * 8 + j];
                adHTdeltaF[(size_t) idx * 8 + j] = a + b;
            }
        }
    for (int i = 0; i < 4; i++) cDeltaF[i] = (float) HCalib.value_minus_value_zero[i];
    for (auto &f : frames) {
        for (int i = 0; i < 8; i++) {
            f.delta[i] = f.state[i] - f.state_zero[i];
            f.delta_prior[i] = f.state[i];
        }
    }
    for (auto &p : points) p.deltaF = p.idepth - p.idepth_zero;
}

VecXd Window::getStitchedDeltaF() const {  // EnergyFunctional.h:178-184
    VecXd d(CPARS + nFrames * 8);
    for (int i = 0; i < CPARS; i++) d[i] = cDeltaF[i];
    for (int h = 0; h < nFrames; h++)
        for (int i = 0; i < 8; i++) d[CPARS + 8 * h + i] = frames[h].delta[i];
    return d;
}

// ------------------------------------------------------------------------------------------
// AccumulatedTopHessianSSE::addPoint<mode> — AccumulatedTopHessian.cc:9-118
template<int mode>
void Window::topAddPoint(AccumulatedTopHessianSSE &A, Point &p, int tid) {
    const float *dc = cDeltaF;
    float dd = p.deltaF;
    float bd_acc = 0, Hdd_acc = 0;
    float Hcd_acc[4] = {0, 0, 0, 0};

    for (int ri : p.residuals) {
        Residual *r = &residuals[ri];
        if (mode == 0) { if (r->isLinearized || !r->isActive()) continue; }
        if (mode == 1) { if (!r->isLinearized || !r->isActive()) continue; }
        if (mode == 2) { if (!r->isActive()) continue; }
        const RawResidualJacobian *rJ = &r->J;
        int htIDX = r->hostIDX + r->targetIDX * A.nframes[tid];
        const float *dp = &adHTdeltaF[(size_t) htIDX * 8];

        float resApprox[8];
        if (mode == 0) for (int i = 0; i < 8; i++) resApprox[i] = rJ->resF[i];
        if (mode == 2) for (int i = 0; i < 8; i++) resApprox[i] = r->res_toZeroF[i];
        if (mode == 1) {
            float a = 0, b = 0;
            for (int i = 0; i < 6; i++) a += rJ->Jpdxi[0][i] * dp[i];
            for (int i = 0; i < 4; i++) b += rJ->Jpdc[0][i] * dc[i];
            float Jp_delta_x = a + b + rJ->Jpdd[0] * dd;
            a = 0; b = 0;
            for (int i = 0; i < 6; i++) a += rJ->Jpdxi[1][i] * dp[i];
            for (int i = 0; i < 4; i++) b += rJ->Jpdc[1][i] * dc[i];
            float Jp_delta_y = a + b + rJ->Jpdd[1] * dd;
            float delta_a = dp[6], delta_b = dp[7];
            for (int i = 0; i < patternNum; i++) {
                float rtz = r->res_toZeroF[i];
                rtz = rtz + rJ->JIdx[0][i] * Jp_delta_x;
                rtz = rtz + rJ->JIdx[1][i] * Jp_delta_y;
                rtz = rtz + rJ->JabF[0][i] * delta_a;
                rtz = rtz + rJ->JabF[1][i] * delta_b;
                resApprox[i] = rtz;
            }
        }
        float JI_r[2] = {0, 0}, Jab_r[2] = {0, 0}, rr = 0;
        for (int i = 0; i < patternNum; i++) {
            JI_r[0] += resApprox[i] * rJ->JIdx[0][i];
            JI_r[1] += resApprox[i] * rJ->JIdx[1][i];
            Jab_r[0] += resApprox[i] * rJ->JabF[0][i];
            Jab_r[1] += resApprox[i] * rJ->JabF[1][i];
            rr += resApprox[i] * resApprox[i];
        }
        AccumulatorApprox &acc = A.acc[tid][htIDX];
        acc.update(rJ->Jpdc[0], rJ->Jpdxi[0], rJ->Jpdc[1], rJ->Jpdxi[1], rJ->JIdx2[0], rJ->JIdx2[1], rJ->JIdx2[3]);
        acc.updateBotRight(rJ->Jab2[0], rJ->Jab2[1], Jab_r[0], rJ->Jab2[3], Jab_r[1], rr);
        acc.updateTopRight(rJ->Jpdc[0], rJ->Jpdxi[0], rJ->Jpdc[1], rJ->Jpdxi[1],
                           rJ->JabJIdx[0], rJ->JabJIdx[1], rJ->JabJIdx[2], rJ->JabJIdx[3], JI_r[0], JI_r[1]);

        float Ji2_Jpdd[2];
        Ji2_Jpdd[0] = rJ->JIdx2[0] * rJ->Jpdd[0] + rJ->JIdx2[1] * rJ->Jpdd[1];
        Ji2_Jpdd[1] = rJ->JIdx2[2] * rJ->Jpdd[0] + rJ->JIdx2[3] * rJ->Jpdd[1];
        bd_acc += JI_r[0] * rJ->Jpdd[0] + JI_r[1] * rJ->Jpdd[1];
        Hdd_acc += Ji2_Jpdd[0] * rJ->Jpdd[0] + Ji2_Jpdd[1] * rJ->Jpdd[1];
        for (int i = 0; i < 4; i++) Hcd_acc[i] += rJ->Jpdc[0][i] * Ji2_Jpdd[0] + rJ->Jpdc[1][i] * Ji2_Jpdd[1];
        A.nres[tid]++;
    }
    if (mode == 0) {
        p.Hdd_accAF = Hdd_acc;
        p.bd_accAF = bd_acc;
        for (int i = 0; i < 4; i++) p.Hcd_accAF[i] = Hcd_acc[i];
    }
    if (mode == 1 || mode == 2) {
        p.Hdd_accLF = Hdd_acc;
        p.bd_accLF = bd_acc;
        for (int i = 0; i < 4; i++) p.Hcd_accLF[i] = Hcd_acc[i];
    }
    if (mode == 2) {
        for (int i = 0; i < 4; i++) p.Hcd_accAF[i] = 0;
        p.Hdd_accAF = 0;
        p.bd_accAF = 0;
    }
}

// 8x8 helpers for the stitch (double, row-major)
static inline void mm8(const double *A, const double *B, double *C, int bc /*cols of B*/, int ldb) {
    // C(8 x bc) = A(8x8) * B(8 x bc), B row stride ldb
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < bc; j++) {
            double s = 0;
            for (int k = 0; k < 8; k++) s += A[i * 8 + k] * B[k * ldb + j];
            C[i * bc + j] = s;
        }
}
static inline void addABAt(MatX &H, int r0, int c0, const double *A, const double *M /*8x8 row-major*/, const double *B) {
    // H.block<8,8>(r0,c0) += A * M * B^T
    double T[64];
    mm8(A, M, T, 8, 8);
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
            double s = 0;
            for (int k = 0; k < 8; k++) s += T[i * 8 + k] * B[j * 8 + k];
            H(r0 + i, c0 + j) += s;
        }
}

// AccumulatedTopHessianSSE::stitchDoubleInternal — AccumulatedTopHessian.cc:193-255
void Window::topStitchDoubleInternal(AccumulatedTopHessianSSE &A, MatX *H, VecXd *b, bool usePrior, int min, int max,
                                     int tid, bool hostOuter) {
    int toAggregate = NUM_THREADS;
    if (tid == -1) { toAggregate = 1; tid = 0; }
    if (min == max) return;
    int nf = A.nframes[0];
    for (int k = min; k < max; k++) {
        // stitchDoubleInternal walks the pairs target-major (k = h + nf*t); the single-accumulator stitchDouble (.cc:133-134) walks
        // them host-major (for h, for t), which changes the order the diagonal blocks receive their terms
        int h = hostOuter ? k / nf : k % nf, t = hostOuter ? k % nf : k / nf;
        int hIdx = CPARS + h * 8, tIdx = CPARS + t * 8;
        int aidx = h + nf * t;
        double accH[13 * 13];
        memset(accH, 0, sizeof(accH));
        for (int tid2 = 0; tid2 < toAggregate; tid2++) {
            A.acc[tid2][aidx].finish();
            if (A.acc[tid2][aidx].num == 0) continue;
            for (int i = 0; i < 169; i++) accH[i] += (double) A.acc[tid2][aidx].H[i];
        }
        double M88[64], M8C[32], bcol[8];
        for (int i = 0; i < 8; i++) {
            for (int j = 0; j < 8; j++) M88[i * 8 + j] = accH[(CPARS + i) * 13 + CPARS + j];
            for (int j = 0; j < 4; j++) M8C[i * 4 + j] = accH[(CPARS + i) * 13 + j];
            bcol[i] = accH[(CPARS + i) * 13 + CPARS + 8];
        }
        const double *AH = &adHost[(size_t) aidx * 64], *AT = &adTarget[(size_t) aidx * 64];
        addABAt(H[tid], hIdx, hIdx, AH, M88, AH);
        addABAt(H[tid], tIdx, tIdx, AT, M88, AT);
        addABAt(H[tid], hIdx, tIdx, AH, M88, AT);
        double T[32];
        mm8(AH, M8C, T, 4, 4);
        for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) H[tid](hIdx + i, j) += T[i * 4 + j];
        mm8(AT, M8C, T, 4, 4);
        for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) H[tid](tIdx + i, j) += T[i * 4 + j];
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) H[tid](i, j) += accH[i * 13 + j];
        for (int i = 0; i < 8; i++) {
            double s = 0, s2 = 0;
            for (int k2 = 0; k2 < 8; k2++) { s += AH[i * 8 + k2] * bcol[k2]; s2 += AT[i * 8 + k2] * bcol[k2]; }
            b[tid][hIdx + i] += s;
            b[tid][tIdx + i] += s2;
        }
        for (int i = 0; i < 4; i++) b[tid][i] += accH[i * 13 + CPARS + 8];
    }
    if (min == 0 && usePrior) {
        for (int i = 0; i < 4; i++) {
            H[tid](i, i) += cPrior[i];
            b[tid][i] += cPrior[i] * (double) cDeltaF[i];
        }
        for (int h = 0; h < nf; h++)
            for (int i = 0; i < 8; i++) {
                H[tid](CPARS + h * 8 + i, CPARS + h * 8 + i) += frames[h].prior[i];
                b[tid][CPARS + h * 8 + i] += frames[h].prior[i] * frames[h].delta_prior[i];
            }
    }
}

// AccumulatedTopHessianSSE::stitchDoubleMT — AccumulatedTopHessian.h:64-105
void Window::topStitchDoubleMT(AccumulatedTopHessianSSE &A, MatX &H, VecXd &b, bool usePrior, bool MT) {
    int nf = A.nframes[0];
    int n = nf * 8 + CPARS;
    if (MT) {
        MatX Hs[NUM_THREADS];
        VecXd bs[NUM_THREADS];
        for (int i = 0; i < NUM_THREADS; i++) { Hs[i] = MatX(n, n); bs[i].assign(n, 0.0); }
        red->reduce([&](int min, int max, double *, int tid) { topStitchDoubleInternal(A, Hs, bs, usePrior, min, max, tid); },
                    0, nf * nf, 0);
        H = Hs[0];
        b = bs[0];
        for (int i = 1; i < NUM_THREADS; i++) {
            for (size_t k = 0; k < H.d.size(); k++) H.d[k] += Hs[i].d[k];
            for (int k = 0; k < n; k++) b[k] += bs[i][k];
            A.nres[0] += A.nres[i];
        }
    } else {
        H = MatX(n, n);
        b.assign(n, 0.0);
        topStitchDoubleInternal(A, &H, &b, usePrior, 0, nf * nf, -1);
    }
    for (int h = 0; h < nf; h++) {
        int hIdx = CPARS + h * 8;
        for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) H(i, hIdx + j) = H(hIdx + j, i);
        for (int t = h + 1; t < nf; t++) {
            int tIdx = CPARS + t * 8;
            for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) H(hIdx + i, tIdx + j) += H(tIdx + j, hIdx + i);
            for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) H(tIdx + i, hIdx + j) = H(hIdx + j, tIdx + i);
        }
    }
}

// AccumulatedTopHessianSSE::stitchDouble (single accumulator, used by marginalizePointsF) — .cc:129-191
void Window::topStitchDouble(AccumulatedTopHessianSSE &A, MatX &H, VecXd &b, bool usePrior, int tid) {
    int nf = A.nframes[tid];
    int n = nf * 8 + CPARS;
    H = MatX(n, n);
    b.assign(n, 0.0);
    // identical block algebra to stitchDoubleInternal with toAggregate == 1
    int save = A.nframes[0];
    A.nframes[0] = nf;
    topStitchDoubleInternal(A, &H, &b, usePrior, 0, nf * nf, -1, true);
    A.nframes[0] = save;
    for (int h = 0; h < nf; h++) {
        int hIdx = CPARS + h * 8;
        for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) H(i, hIdx + j) = H(hIdx + j, i);
        for (int t = h + 1; t < nf; t++) {
            int tIdx = CPARS + t * 8;
            for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) H(hIdx + i, tIdx + j) += H(tIdx + j, hIdx + i);
            for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) H(tIdx + i, hIdx + j) = H(hIdx + j, tIdx + i);
        }
    }
}

// AccumulatedSCHessianSSE::addPoint — AccumulatedSCHessian.cc:9-51
void Window::scAddPoint(Point &p, bool shiftPriorToZero, int tid) {
    AccumulatedSCHessianSSE &B = accSSE_bot;
    int ngoodres = 0;
    for (int ri : p.residuals) if (residuals[ri].isActive()) ngoodres++;
    if (ngoodres == 0) {
        p.HdiF = 0;
        p.bdSumF = 0;
        p.idepth_hessian = 0;
        p.maxRelBaseline = 0;
        return;
    }
    float H = p.Hdd_accAF + p.Hdd_accLF + p.priorF;
    if (H < 1e-10) H = 1e-10;
    p.idepth_hessian = H;
    p.HdiF = 1.0 / H;
    p.bdSumF = p.bd_accAF + p.bd_accLF;
    if (shiftPriorToZero) p.bdSumF += p.priorF * p.deltaF;
    float Hcd[4];
    for (int i = 0; i < 4; i++) Hcd[i] = p.Hcd_accAF[i] + p.Hcd_accLF[i];
    B.accHcc[tid].update(Hcd, Hcd, p.HdiF);
    B.accbc[tid].update(Hcd, p.bdSumF * p.HdiF);

    int nf = B.nframes[tid];
    int nFrames2 = nf * nf;
    for (int r1i : p.residuals) {
        Residual &r1 = residuals[r1i];
        if (!r1.isActive()) continue;
        int r1ht = r1.hostIDX + r1.targetIDX * nf;
        for (int r2i : p.residuals) {
            Residual &r2 = residuals[r2i];
            if (!r2.isActive()) continue;
            B.accD[tid][r1ht + r2.targetIDX * nFrames2].update(r1.JpJdF, r2.JpJdF, p.HdiF);
        }
        B.accE[tid][r1ht].update(r1.JpJdF, Hcd, p.HdiF);
        B.accEB[tid][r1ht].update(r1.JpJdF, p.HdiF * p.bdSumF);
    }
}

// AccumulatedSCHessianSSE::stitchDoubleInternal — AccumulatedSCHessian.cc:53-119
void Window::scStitchDoubleInternal(MatX *H, VecXd *b, int min, int max, int tid, bool hostOuter) {
    AccumulatedSCHessianSSE &B = accSSE_bot;
    int toAggregate = NUM_THREADS;
    if (tid == -1) { toAggregate = 1; tid = 0; }
    if (min == max) return;
    int nf = B.nframes[0];
    int nframes2 = nf * nf;
    for (int k = min; k < max; k++) {
        // same remark as the top stitch: stitchDouble (.cc:130-131) walks (i, j) with i outermost
        int i = hostOuter ? k / nf : k % nf, j = hostOuter ? k % nf : k / nf;
        int iIdx = CPARS + i * 8, jIdx = CPARS + j * 8;
        int ijIdx = i + nf * j;
        double Hpc[32], bp[8];  // Hpc 8x4 row-major
        memset(Hpc, 0, sizeof(Hpc));
        memset(bp, 0, sizeof(bp));
        for (int tid2 = 0; tid2 < toAggregate; tid2++) {
            B.accE[tid2][ijIdx].finish();
            B.accEB[tid2][ijIdx].finish();
            for (int r = 0; r < 8; r++) for (int c = 0; c < 4; c++) Hpc[r * 4 + c] += (double) B.accE[tid2][ijIdx].A1m[c * 8 + r];
            for (int r = 0; r < 8; r++) bp[r] += (double) B.accEB[tid2][ijIdx].A1m[r];
        }
        const double *AHij = &adHost[(size_t) ijIdx * 64], *ATij = &adTarget[(size_t) ijIdx * 64];
        double T[32];
        mm8(AHij, Hpc, T, 4, 4);
        for (int r = 0; r < 8; r++) for (int c = 0; c < 4; c++) H[tid](iIdx + r, c) += T[r * 4 + c];
        mm8(ATij, Hpc, T, 4, 4);
        for (int r = 0; r < 8; r++) for (int c = 0; c < 4; c++) H[tid](jIdx + r, c) += T[r * 4 + c];
        for (int r = 0; r < 8; r++) {
            double s = 0, s2 = 0;
            for (int q = 0; q < 8; q++) { s += AHij[r * 8 + q] * bp[q]; s2 += ATij[r * 8 + q] * bp[q]; }
            b[tid][iIdx + r] += s;
            b[tid][jIdx + r] += s2;
        }
        for (int kk = 0; kk < nf; kk++) {
            int kIdx = CPARS + kk * 8;
            int ijkIdx = ijIdx + kk * nframes2;
            int ikIdx = i + nf * kk;
            double accDM[64];
            memset(accDM, 0, sizeof(accDM));
            for (int tid2 = 0; tid2 < toAggregate; tid2++) {
                B.accD[tid2][ijkIdx].finish();
                if (B.accD[tid2][ijkIdx].num == 0) continue;
                for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) accDM[r * 8 + c] += (double) B.accD[tid2][ijkIdx].A1m[c * 8 + r];
            }
            const double *AHik = &adHost[(size_t) ikIdx * 64], *ATik = &adTarget[(size_t) ikIdx * 64];
            addABAt(H[tid], iIdx, iIdx, AHij, accDM, AHik);
            addABAt(H[tid], jIdx, kIdx, ATij, accDM, ATik);
            addABAt(H[tid], jIdx, iIdx, ATij, accDM, AHik);
            addABAt(H[tid], iIdx, kIdx, AHij, accDM, ATik);
        }
    }
    if (min == 0) {
        for (int tid2 = 0; tid2 < toAggregate; tid2++) {
            B.accHcc[tid2].finish();
            B.accbc[tid2].finish();
            for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) H[tid](r, c) += (double) B.accHcc[tid2].A1m[c * 4 + r];
            for (int r = 0; r < 4; r++) b[tid][r] += (double) B.accbc[tid2].A1m[r];
        }
    }
}

// AccumulatedSCHessianSSE::stitchDoubleMT — AccumulatedSCHessian.h:64-98
void Window::scStitchDoubleMT(MatX &H, VecXd &b, bool MT) {
    int nf = accSSE_bot.nframes[0];
    int n = nf * 8 + CPARS;
    if (MT) {
        MatX Hs[NUM_THREADS];
        VecXd bs[NUM_THREADS];
        for (int i = 0; i < NUM_THREADS; i++) { Hs[i] = MatX(n, n); bs[i].assign(n, 0.0); }
        red->reduce([&](int min, int max, double *, int tid) { scStitchDoubleInternal(Hs, bs, min, max, tid); }, 0, nf * nf, 0);
        H = Hs[0];
        b = bs[0];
        for (int i = 1; i < NUM_THREADS; i++) {
            for (size_t k = 0; k < H.d.size(); k++) H.d[k] += Hs[i].d[k];
            for (int k = 0; k < n; k++) b[k] += bs[i][k];
        }
    } else {
        H = MatX(n, n);
        b.assign(n, 0.0);
        scStitchDoubleInternal(&H, &b, 0, nf * nf, -1);
    }
    for (int h = 0; h < nf; h++) {
        int hIdx = CPARS + h * 8;
        for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) H(i, hIdx + j) = H(hIdx + j, i);
    }
}

// AccumulatedSCHessianSSE::stitchDouble (single accumulator) — AccumulatedSCHessian.cc:121-177
void Window::scStitchDouble(MatX &H, VecXd &b, int tid) {
    (void) tid;
    int nf = accSSE_bot.nframes[0];
    int n = nf * 8 + CPARS;
    H = MatX(n, n);
    b.assign(n, 0.0);
    scStitchDoubleInternal(&H, &b, 0, nf * nf, -1, true);
    for (int h = 0; h < nf; h++) {
        int hIdx = CPARS + h * 8;
        for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) H(i, hIdx + j) = H(hIdx + j, i);
    }
}

// ------------------------------------------------------------------------------------------
// EnergyFunctional::accumulate{AF,LF,SCF}_MT — EnergyFunctional.cc:550-625
void Window::accumulateAF_MT(MatX &H, VecXd &b, bool MT) {
    if (MT) {
        red->reduce([&](int, int, double *, int tid) { accSSE_top_A.setZero(nFrames, tid); }, 0, 0, 0);
        red->reduce([&](int min, int max, double *, int tid) {
            for (int i = min; i < max; i++) topAddPoint<0>(accSSE_top_A, points[i], tid);
        }, 0, (int) points.size(), 50);
        topStitchDoubleMT(accSSE_top_A, H, b, false, true);
        resInA = accSSE_top_A.nres[0];
    } else {
        accSSE_top_A.setZero(nFrames, 0);
        for (auto &p : points) topAddPoint<0>(accSSE_top_A, p, 0);
        topStitchDoubleMT(accSSE_top_A, H, b, false, false);
        resInA = accSSE_top_A.nres[0];
    }
}
void Window::accumulateLF_MT(MatX &H, VecXd &b, bool MT) {
    if (MT) {
        red->reduce([&](int, int, double *, int tid) { accSSE_top_L.setZero(nFrames, tid); }, 0, 0, 0);
        red->reduce([&](int min, int max, double *, int tid) {
            for (int i = min; i < max; i++) topAddPoint<1>(accSSE_top_L, points[i], tid);
        }, 0, (int) points.size(), 50);
        topStitchDoubleMT(accSSE_top_L, H, b, true, true);
        resInL = accSSE_top_L.nres[0];
    } else {
        accSSE_top_L.setZero(nFrames, 0);
        for (auto &p : points) topAddPoint<1>(accSSE_top_L, p, 0);
        topStitchDoubleMT(accSSE_top_L, H, b, true, false);
        resInL = accSSE_top_L.nres[0];
    }
}
void Window::accumulateSCF_MT(MatX &H, VecXd &b, bool MT) {
    if (MT) {
        red->reduce([&](int, int, double *, int tid) { accSSE_bot.setZero(nFrames, tid); }, 0, 0, 0);
        red->reduce([&](int min, int max, double *, int tid) {
            for (int i = min; i < max; i++) scAddPoint(points[i], true, tid);
        }, 0, (int) points.size(), 50);
        scStitchDoubleMT(H, b, true);
    } else {
        accSSE_bot.setZero(nFrames, 0);
        for (auto &p : points) scAddPoint(p, true, 0);
        scStitchDoubleMT(H, b, false);
    }
}

// EnergyFunctional::orthogonalize — :685-717
void Window::orthogonalize(VecXd *b, MatX *H) {
    std::vector<VecXd> ns;
    ns.insert(ns.end(), lastNullspaces_pose.begin(), lastNullspaces_pose.end());
    ns.insert(ns.end(), lastNullspaces_scale.begin(), lastNullspaces_scale.end());
    int dim = (int) ns[0].size(), k = (int) ns.size();
    MatX N(dim, k);
    for (int i = 0; i < k; i++) {
        double nn = 0;
        for (int r = 0; r < dim; r++) nn += ns[i][r] * ns[i][r];
        nn = std::sqrt(nn);
        for (int r = 0; r < dim; r++) N(r, i) = ns[i][r] / nn;
    }
    MatX U, V;
    VecXd SNN;
    jacobi_svd(N, U, SNN, V);
    double maxSv = 0;
    for (int i = 0; i < k; i++) if (SNN[i] > maxSv) maxSv = SNN[i];
    for (int i = 0; i < k; i++) {
        if (SNN[i] > S.solverModeDelta * maxSv) SNN[i] = 1.0 / SNN[i];
        else SNN[i] = 0;
    }
    // Npi = U * diag(SNN) * V^T   [dim x k]
    MatX US(dim, k);
    for (int i = 0; i < k; i++) for (int r = 0; r < dim; r++) US(r, i) = U(r, i) * SNN[i];
    MatX Npi = matmul(US, transpose(V));
    MatX NNpiT = matmul(N, transpose(Npi));
    MatX NNpiTS(dim, dim);
    for (int r = 0; r < dim; r++) for (int c = 0; c < dim; c++) NNpiTS(r, c) = 0.5 * (NNpiT(r, c) + NNpiT(c, r));
    if (b != 0) {
        VecXd nb(dim, 0.0);
        for (int r = 0; r < dim; r++) for (int c = 0; c < dim; c++) nb[r] += NNpiTS(r, c) * (*b)[c];
        for (int r = 0; r < dim; r++) (*b)[r] -= nb[r];
    }
    if (H != 0) {
        MatX T = matmul(matmul(NNpiTS, *H), NNpiTS);
        for (size_t i = 0; i < H->d.size(); i++) H->d[i] -= T.d[i];
    }
}

// EnergyFunctional::solveSystemF — :240-351 (default solver mode:
// SOLVER_FIX_LAMBDA | SOLVER_ORTHOGONALIZE_X_LATER, Setting.cc:23)
void Window::solveSystemF(int iteration, double lambda) {
    lambda = 1e-5;  // SOLVER_FIX_LAMBDA
    MatX HL_top, HA_top, H_sc;
    VecXd bL_top, bA_top, bM_top, b_sc;
    accumulateAF_MT(HA_top, bA_top, S.multiThreading);
    accumulateLF_MT(HL_top, bL_top, S.multiThreading);
    accumulateSCF_MT(H_sc, b_sc, S.multiThreading);
    int n = 8 * nFrames + CPARS;
    VecXd delta = getStitchedDeltaF();
    bM_top.assign(n, 0.0);
    for (int r = 0; r < n; r++) {
        double s = 0;
        for (int c = 0; c < n; c++) s += HM(r, c) * delta[c];
        bM_top[r] = bM[r] + s;
    }
    MatX HFinal_top(n, n);
    VecXd bFinal_top(n);
    for (size_t i = 0; i < HFinal_top.d.size(); i++) HFinal_top.d[i] = HL_top.d[i] + HM.d[i] + HA_top.d[i];
    for (int i = 0; i < n; i++) bFinal_top[i] = bL_top[i] + bM_top[i] + bA_top[i] - b_sc[i];
    lastHS = MatX(n, n);
    for (size_t i = 0; i < lastHS.d.size(); i++) lastHS.d[i] = HFinal_top.d[i] - H_sc.d[i];
    lastbS = bFinal_top;
    for (int i = 0; i < n; i++) HFinal_top(i, i) *= (1 + lambda);
    {
        double f = (double) (1.0f / (1 + lambda));  // `H_sc * (1.0f / (1 + lambda))`: double expression
        f = 1.0 / (1 + lambda);
        for (size_t i = 0; i < HFinal_top.d.size(); i++) HFinal_top.d[i] -= H_sc.d[i] * f;
    }
    last_HA = HA_top; last_bA = bA_top; last_Hsc = H_sc; last_bsc = b_sc; last_HL = HL_top; last_bL = bL_top;

    VecXd SVecI(n);
    for (int i = 0; i < n; i++) SVecI[i] = 1.0 / std::sqrt(HFinal_top(i, i) + 10);
    MatX HFinalScaled(n, n);
    for (int r = 0; r < n; r++) for (int c = 0; c < n; c++) HFinalScaled(r, c) = SVecI[r] * HFinal_top(r, c) * SVecI[c];
    VecXd bs(n);
    for (int i = 0; i < n; i++) bs[i] = SVecI[i] * bFinal_top[i];
    VecXd x = ldlt_solve(HFinalScaled, bs);
    for (int i = 0; i < n; i++) x[i] *= SVecI[i];

    if (iteration >= 2) orthogonalize(&x, 0);  // SOLVER_ORTHOGONALIZE_X_LATER
    lastX = x;
    resubstituteF_MT(x, S.multiThreading);
}

// EnergyFunctional::resubstituteF_MT — :491-516
void Window::resubstituteF_MT(const VecXd &x, bool MT) {
    int n = CPARS + nFrames * 8;
    std::vector<float> xF(n);
    for (int i = 0; i < n; i++) xF[i] = (float) x[i];
    for (int i = 0; i < 4; i++) HCalib.step[i] = -x[i];
    std::vector<float> xAd((size_t) nFrames * nFrames * 8);
    float cstep[4] = {xF[0], xF[1], xF[2], xF[3]};
    for (auto &h : frames) {
        for (int i = 0; i < 8; i++) h.step[i] = -x[CPARS + 8 * h.idx + i];
        h.step[8] = h.step[9] = 0;
        for (auto &t : frames) {
            const float *AH = &adHostF[(size_t) (h.idx + nFrames * t.idx) * 64];
            const float *AT = &adTargetF[(size_t) (h.idx + nFrames * t.idx) * 64];
            for (int j = 0; j < 8; j++) {
                float a = 0, b = 0;
                for (int i = 0; i < 8; i++) a += xF[CPARS + 8 * h.idx + i] * AH[i * 8 + j];
                for (int i = 0; i < 8; i++) b += xF[CPARS + 8 * t.idx + i] * AT[i * 8 + j];
                xAd[(size_t) (nFrames * h.idx + t.idx) * 8 + j] = a + b;
            }
        }
    }
    if (MT)
        red->reduce([&](int min, int max, double *, int) { resubstituteFPt(cstep, xAd.data(), min, max); }, 0,
                    (int) points.size(), 50);
    else
        resubstituteFPt(cstep, xAd.data(), 0, (int) points.size());
}

// EnergyFunctional::resubstituteFPt — :518-547
void Window::resubstituteFPt(const float xc[4], const float *xAd, int min, int max) {
    for (int k = min; k < max; k++) {
        Point &p = points[k];
        int ngoodres = 0;
        for (int ri : p.residuals) if (residuals[ri].isActive()) ngoodres++;
        if (ngoodres == 0) {
            p.step = 0;
            continue;
        }
        float b = p.bdSumF;
        {
            float s = 0;
            for (int i = 0; i < 4; i++) s += xc[i] * (p.Hcd_accAF[i] + p.Hcd_accLF[i]);
            b -= s;
        }
        for (int ri : p.residuals) {
            const Residual &r = residuals[ri];
            if (!r.isActive()) continue;
            const float *xa = &xAd[(size_t) (r.hostIDX * nFrames + r.targetIDX) * 8];
            float s = 0;
            for (int i = 0; i < 8; i++) s += xa[i] * r.JpJdF[i];
            b -= s;
        }
        if (!std::isfinite(b) || std::isnan(b)) return;
        p.step = -b * p.HdiF;
    }
}

// EnergyFunctional::calcMEnergyF — :353-359
double Window::calcMEnergyF() {
    VecXd delta = getStitchedDeltaF();
    int n = (int) delta.size();
    double e = 0;
    for (int r = 0; r < n; r++) {
        double s = 0;
        for (int c = 0; c < n; c++) s += HM(r, c) * delta[c];
        e += delta[r] * (2 * bM[r] + s);
    }
    return e;
}

// EnergyFunctional::calcLEnergyPt — :627-682
void Window::calcLEnergyPt(int min, int max, double *stats, int tid) {
    (void) tid;
    Accumulator11 E;
    E.initialize();
    const float *dc = cDeltaF;
    for (int i = min; i < max; i++) {
        Point &p = points[i];
        float dd = p.deltaF;
        for (int ri : p.residuals) {
            Residual &r = residuals[ri];
            if (!r.isLinearized || !r.isActive()) continue;
            const float *dp = &adHTdeltaF[(size_t) (r.hostIDX + nFrames * r.targetIDX) * 8];
            const RawResidualJacobian *rJ = &r.J;
            float a = 0, b = 0;
            for (int k = 0; k < 6; k++) a += rJ->Jpdxi[0][k] * dp[k];
            for (int k = 0; k < 4; k++) b += rJ->Jpdc[0][k] * dc[k];
            float Jp_delta_x_1 = a + b + rJ->Jpdd[0] * dd;
            a = 0; b = 0;
            for (int k = 0; k < 6; k++) a += rJ->Jpdxi[1][k] * dp[k];
            for (int k = 0; k < 4; k++) b += rJ->Jpdc[1][k] * dc[k];
            float Jp_delta_y_1 = a + b + rJ->Jpdd[1] * dd;
            for (int k = 0; k + 3 < patternNum; k += 4) {
                float Jd[4];
                for (int l = 0; l < 4; l++) {
                    float Jdelta = rJ->JIdx[0][k + l] * Jp_delta_x_1;
                    Jdelta = Jdelta + rJ->JIdx[1][k + l] * Jp_delta_y_1;
                    Jdelta = Jdelta + rJ->JabF[0][k + l] * dp[6];
                    Jdelta = Jdelta + rJ->JabF[1][k + l] * dp[7];
                    float r0 = r.res_toZeroF[k + l];
                    r0 = r0 + r0;
                    r0 = r0 + Jdelta;
                    Jd[l] = Jdelta * r0;
                }
                E.updateSSENoShift(Jd);
            }
        }
        E.updateSingle(p.deltaF * p.deltaF * p.priorF);
    }
    E.finish();
    stats[0] += E.A;
}

// EnergyFunctional::calcLEnergyF_MT — :361-378
double Window::calcLEnergyF_MT() {
    double E = 0;
    for (auto &f : frames) for (int i = 0; i < 8; i++) E += f.delta_prior[i] * f.prior[i] * f.delta_prior[i];
    {
        float s = 0;
        for (int i = 0; i < 4; i++) s += cDeltaF[i] * cPriorF[i] * cDeltaF[i];
        E += s;
    }
    red->reduce([&](int min, int max, double *stats, int tid) { calcLEnergyPt(min, max, stats, tid); }, 0,
                (int) points.size(), 50);
    return E + red->stats[0];
}

// EnergyFunctional::marginalizePointsF — :165-222 (algebra only: the points listed are
// accumulated in mode 2 + SC(shiftPriorToZero=false) and folded into HM,bM with
// setting_margWeightFac = 0.25; the caller has already multiplied priorF and called
// fixLinearizationF like FullSystem::flagPointsForRemoval does).
void Window::marginalizePointsF(const std::vector<int> &pointIdx) {
    accSSE_bot.setZero(nFrames, 0);
    accSSE_top_A.setZero(nFrames, 0);
    for (int pi : pointIdx) {
        topAddPoint<2>(accSSE_top_A, points[pi], 0);
        scAddPoint(points[pi], false, 0);
    }
    MatX M, Msc;
    VecXd Mb, Mbsc;
    topStitchDouble(accSSE_top_A, M, Mb, false, 0);
    scStitchDouble(Msc, Mbsc, 0);
    resInM += accSSE_top_A.nres[0];
    const double margWeightFac = 0.5 * 0.5;
    for (size_t i = 0; i < HM.d.size(); i++) HM.d[i] += margWeightFac * (M.d[i] - Msc.d[i]);
    for (size_t i = 0; i < bM.size(); i++) bM[i] += margWeightFac * (Mb[i] - Mbsc[i]);
    last_HA = M; last_bA = Mb; last_Hsc = Msc; last_bsc = Mbsc;
}

// EnergyFunctional::marginalizeFrame — :72-129, the prior algebra: move the frame's 8 rows/columns of HM, bM to the end,
// add its own prior there, scale by (|diag| + 10)^1/2, Schur-complement the 8x8 block out, unscale, symmetrise.
// HM, bM shrink by 8. (The bookkeeping half, :131-150, belongs to the caller, who rebuilds the window without the frame.)
void Window::marginalizeFramePrior(int idx) {
    const int odim = nFrames * 8 + CPARS, ndim = odim - 8;
    const Frame &fh = frames[idx];
    std::vector<int> p(odim);                       // new position -> old position (the block moves of :81-99)
    const int io = idx * 8 + CPARS;
    for (int i = 0; i < io; i++) p[i] = i;
    for (int i = io; i < ndim; i++) p[i] = i + 8;
    for (int k = 0; k < 8; k++) p[ndim + k] = io + k;
    MatX H(odim, odim);
    VecXd b(odim);
    for (int j = 0; j < odim; j++) { b[j] = bM[p[j]]; for (int i = 0; i < odim; i++) H(i, j) = HM(p[i], p[j]); }
    for (int k = 0; k < 8; k++) {                   // :103-104
        H(ndim + k, ndim + k) += fh.prior[k];
        b[ndim + k] += fh.prior[k] * fh.delta_prior[k];
    }
    VecXd SVec(odim), SVecI(odim);                  // :106-107
    for (int i = 0; i < odim; i++) { SVec[i] = std::sqrt(std::fabs(H(i, i)) + 10.0); SVecI[i] = 1.0 / SVec[i]; }
    MatX Hs(odim, odim);                            // :110-111
    VecXd bs(odim);
    for (int j = 0; j < odim; j++) { bs[j] = SVecI[j] * b[j]; for (int i = 0; i < odim; i++) Hs(i, j) = (SVecI[i] * H(i, j)) * SVecI[j]; }
    MatX hpi(8, 8);                                 // :114-117 (the two 0.5f*(hpi+hpi) statements are exact no-ops)
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) hpi(i, j) = 0.5f * (Hs(ndim + i, ndim + j) + Hs(ndim + i, ndim + j));
    hpi = inverse_partial_piv_lu(hpi);
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) hpi(i, j) = 0.5f * (hpi(i, j) + hpi(i, j));
    MatX bli(ndim, 8);                              // :120  bli = bottomLeft^T * hpi
    for (int i = 0; i < ndim; i++)
        for (int k = 0; k < 8; k++) {
            double s = 0.0;
            for (int m = 0; m < 8; m++) s += Hs(ndim + m, i) * hpi(m, k);
            bli(i, k) = s;
        }
    for (int j = 0; j < ndim; j++)                  // :121-122
        for (int i = 0; i < ndim; i++) {
            double s = 0.0;
            for (int k = 0; k < 8; k++) s += bli(i, k) * Hs(ndim + k, j);
            Hs(i, j) -= s;
        }
    for (int i = 0; i < ndim; i++) {
        double s = 0.0;
        for (int k = 0; k < 8; k++) s += bli(i, k) * bs[ndim + k];
        bs[i] -= s;
    }
    for (int j = 0; j < odim; j++) { bs[j] = SVec[j] * bs[j]; for (int i = 0; i < odim; i++) Hs(i, j) = (SVec[i] * Hs(i, j)) * SVec[j]; }   // :125-126
    MatX Hn(ndim, ndim);                            // :129-130
    VecXd bn(ndim);
    for (int j = 0; j < ndim; j++) { bn[j] = bs[j]; for (int i = 0; i < ndim; i++) Hn(i, j) = 0.5 * (Hs(i, j) + Hs(j, i)); }
    HM = Hn;
    bM = bn;
}

// ImmaturePoint::linearizeResidual — src/internal/ImmaturePoint.cc:316-383 (the temporary residual is (state, energy, newState, newEnergy))
double immatureLinearizeResidual(const Window &W, const Window::ImmatureCand &c, int target, float outlierTHSlack, TmpRes &tr,
                                 float &Hdd, float &bd, float idepth) {
    if (tr.state_state == RS_OOB) { tr.state_NewState = RS_OOB; return tr.state_energy; }
    const FramePrecalc &pc = W.frames[c.host].targetPrecalc[target];
    const float *dIl = W.frames[target].dI;
    float energyLeft = 0;
    const float *affLL = pc.PRE_aff_mode;
    for (int idx = 0; idx < patternNum; idx++) {
        int dx = patternP[idx][0], dy = patternP[idx][1];
        float drescale, u, v, new_idepth, Ku, Kv, KliP[3];
        if (!projectPointB(c.u, c.v, idepth, dx, dy, W.HCalib, pc.PRE_RTll, pc.PRE_tTll, W.wM3G, W.hM3G, drescale, u, v, Ku, Kv, KliP, new_idepth)) {
            tr.state_NewState = RS_OOB;
            return tr.state_energy;
        }
        float hitColor[3];
        getInterpolatedElement33(dIl, Ku, Kv, W.frames[target].w, hitColor);
        if (!std::isfinite((float) hitColor[0])) { tr.state_NewState = RS_OOB; return tr.state_energy; }
        float residual = hitColor[0] - (affLL[0] * c.color[idx] + affLL[1]);
        float hw = fabsf(residual) < W.S.huberTH ? 1 : W.S.huberTH / fabsf(residual);
        energyLeft += c.weights[idx] * c.weights[idx] * hw * residual * residual * (2 - hw);
        // depth derivatives (derive_idepth, ResidualProjections.h:12-18)
        float dxInterp = hitColor[1] * W.HCalib.fxl();
        float dyInterp = hitColor[2] * W.HCalib.fyl();
        float d_idepth = derive_idepth(pc.PRE_tTll, u, v, dx, dy, dxInterp, dyInterp, drescale);
        hw *= c.weights[idx] * c.weights[idx];
        Hdd += (hw * d_idepth) * d_idepth;
        bd += (hw * residual) * d_idepth;
    }
    if (energyLeft > c.energyTH * outlierTHSlack) {
        energyLeft = c.energyTH * outlierTHSlack;
        tr.state_NewState = RS_OUTLIER;
    } else {
        tr.state_NewState = RS_IN;
    }
    tr.state_NewEnergy = energyLeft;
    return energyLeft;
}

// FullSystem::optimizeImmaturePoint — src/frontend/FullSystem.cc:892-978 (the part that decides; building the PointHessian and its
// residual objects, :980-1009, is the caller's bookkeeping: every residual left IN becomes a PointFrameResidual)
bool Window::optimizeImmaturePoint(const ImmatureCand &c, int minObs, float &idepth_out, unsigned char *res_state) {
    const float setting_minIdepthH_act = 100;          // Setting.cc:25
    const int setting_GNItsOnPointActivation = 3;      // Setting.cc:47
    const int nFr = (int) frames.size();
    std::vector<TmpRes> residuals;
    std::vector<int> targets;
    for (int t = 0; t < nFr; t++) {
        res_state[t] = 255;
        if (t != c.host) { residuals.push_back({RS_IN, RS_OUTLIER, 0.f, 0.f}); targets.push_back(t); }
    }
    const int nres = (int) residuals.size();
    auto publish = [&]() { for (int i = 0; i < nres; i++) res_state[targets[i]] = (unsigned char) residuals[i].state_state; };
    float lastEnergy = 0, lastHdd = 0, lastbd = 0;
    float currentIdepth = (c.idepth_max + c.idepth_min) * 0.5f;
    for (int i = 0; i < nres; i++) {
        lastEnergy += immatureLinearizeResidual(*this, c, targets[i], 1000, residuals[i], lastHdd, lastbd, currentIdepth);
        residuals[i].state_state = residuals[i].state_NewState;
        residuals[i].state_energy = residuals[i].state_NewEnergy;
    }
    idepth_out = currentIdepth;
    if (!std::isfinite(lastEnergy) || lastHdd < setting_minIdepthH_act) { publish(); return false; }
    float lambda = 0.1;
    for (int iteration = 0; iteration < setting_GNItsOnPointActivation; iteration++) {
        float H = lastHdd;
        H *= 1 + lambda;
        float step = (1.0 / H) * lastbd;
        float newIdepth = currentIdepth - step;
        float newHdd = 0, newbd = 0, newEnergy = 0;
        for (int i = 0; i < nres; i++)
            newEnergy += immatureLinearizeResidual(*this, c, targets[i], 1, residuals[i], newHdd, newbd, newIdepth);
        if (!std::isfinite(lastEnergy) || newHdd < setting_minIdepthH_act) { idepth_out = currentIdepth; publish(); return false; }
        if (newEnergy < lastEnergy) {
            currentIdepth = newIdepth;
            lastHdd = newHdd; lastbd = newbd; lastEnergy = newEnergy;
            for (int i = 0; i < nres; i++) {
                residuals[i].state_state = residuals[i].state_NewState;
                residuals[i].state_energy = residuals[i].state_NewEnergy;
            }
            lambda *= 0.5;
        } else {
            lambda *= 5;
        }
        if (fabsf(step) < 0.0001 * currentIdepth) break;
    }
    idepth_out = currentIdepth;
    publish();
    if (!std::isfinite(currentIdepth)) return false;
    int numGoodRes = 0;
    for (int i = 0; i < nres; i++) if (residuals[i].state_state == RS_IN) numGoodRes++;
    return numGoodRes >= minObs;
}

template void Window::topAddPoint<0>(AccumulatedTopHessianSSE &, Point &, int);
template void Window::topAddPoint<1>(AccumulatedTopHessianSSE &, Point &, int);
template void Window::topAddPoint<2>(AccumulatedTopHessianSSE &, Point &, int);

}  // namespace oracle
