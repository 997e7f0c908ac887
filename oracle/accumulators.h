// ORACLE — TEST INFRASTRUCTURE ONLY (see omath.h header). PARITY PINNED AGAINST THE REFERENCE'S OWN SOURCES, EXCEPT EIGEN / SOPHUS INTERNALS (oracle/ref_pin compiles, unmodified, the reference's Residuals.cc, ImmaturePoint.cc, PointHessian.cc, FrameHessian.cc, FrameFramePrecalc.cc, AccumulatedTopHessian.cc, AccumulatedSCHessian.cc, EnergyFunctional.cc, CoarseTracker.cc, Setting.cc and their headers against stand-in Eigen/Sophus headers and matches the oracle bit for bit on 108 checks; the float SSE / scalar code is pinned completely, the Eigen-expression code structurally - what is added where, in which order - with LDLT, PartialPivLU, JacobiSVD, SE3 exp/log and the product kernels being the oracle's own restatements on both sides; FullSystem.cc's driver loop needs the whole front end and stays restated from the cited lines).
// Restates include/internal/OptimizationBackend/MatrixAccumulators.h of the reference:
//   AccumulatorXX<i,j> :20-66, Accumulator11 :68-142, AccumulatorX<i> :145-197,
//   AccumulatorApprox :749-1101, Accumulator9 :1104-1135,1250-1369,1624-1642.
// The unrolled SSE/scalar lines of the reference are written as loops here, keeping the
// per-entry float expression (operand order) and the 1 / 1k / 1M tiered "shiftUp" sums.
#pragma once
#include <cstring>
#include <cstddef>

namespace oracle {

// AccumulatorXX<i,j>: A += w * L * R^T   (column-major i x j like Eigen)
template<int I, int J>
struct AccumulatorXX {
    float A[I * J], A1k[I * J], A1m[I * J];
    size_t num;
    float numIn1, numIn1k, numIn1m;

    void initialize() {
        memset(A, 0, sizeof(A)); memset(A1k, 0, sizeof(A1k)); memset(A1m, 0, sizeof(A1m));
        num = 0; numIn1 = numIn1k = numIn1m = 0;
    }
    void finish() {
        shiftUp(true);
        num = (size_t) (numIn1 + numIn1k + numIn1m);
    }
    // MatrixAccumulators.h:43-47 — Eigen evaluates (w*L) * R^T entrywise
    void update(const float *L, const float *R, float w) {
        for (int c = 0; c < J; c++)
            for (int r = 0; r < I; r++)
                A[c * I + r] += (w * L[r]) * R[c];
        numIn1++;
        shiftUp(false);
    }
    void shiftUp(bool force) {  // :52-65
        if (numIn1 > 1000 || force) {
            for (int k = 0; k < I * J; k++) { A1k[k] += A[k]; A[k] = 0; }
            numIn1k += numIn1; numIn1 = 0;
        }
        if (numIn1k > 1000 || force) {
            for (int k = 0; k < I * J; k++) { A1m[k] += A1k[k]; A1k[k] = 0; }
            numIn1m += numIn1k; numIn1k = 0;
        }
    }
};

// AccumulatorX<i>: A += w * L
template<int I>
struct AccumulatorX {
    float A[I], A1k[I], A1m[I];
    size_t num;
    float numIn1, numIn1k, numIn1m;
    void initialize() {
        memset(A, 0, sizeof(A)); memset(A1k, 0, sizeof(A1k)); memset(A1m, 0, sizeof(A1m));
        num = 0; numIn1 = numIn1k = numIn1m = 0;
    }
    void finish() {
        shiftUp(true);
        num = (size_t) (numIn1 + numIn1k + numIn1m);
    }
    void update(const float *L, float w) {  // :166-170
        for (int r = 0; r < I; r++) A[r] += w * L[r];
        numIn1++;
        shiftUp(false);
    }
    void shiftUp(bool force) {
        if (numIn1 > 1000 || force) {
            for (int k = 0; k < I; k++) { A1k[k] += A[k]; A[k] = 0; }
            numIn1k += numIn1; numIn1 = 0;
        }
        if (numIn1k > 1000 || force) {
            for (int k = 0; k < I; k++) { A1m[k] += A1k[k]; A1k[k] = 0; }
            numIn1m += numIn1k; numIn1k = 0;
        }
    }
};

// Accumulator11 (:68-142): 4-lane float sum
struct Accumulator11 {
    float A;
    size_t num;
    float SSEData[4], SSEData1k[4], SSEData1m[4];
    float numIn1, numIn1k, numIn1m;
    void initialize() {
        A = 0;
        memset(SSEData, 0, sizeof(SSEData)); memset(SSEData1k, 0, sizeof(SSEData1k));
        memset(SSEData1m, 0, sizeof(SSEData1m));
        num = 0; numIn1 = numIn1k = numIn1m = 0;
    }
    void finish() {
        shiftUp(true);
        A = SSEData1m[0] + SSEData1m[1] + SSEData1m[2] + SSEData1m[3];
    }
    void updateSingle(float val) { SSEData[0] += val; num++; numIn1++; shiftUp(false); }
    void updateSingleNoShift(float val) { SSEData[0] += val; num++; numIn1++; }
    void updateSSENoShift(const float v[4]) {
        for (int k = 0; k < 4; k++) SSEData[k] += v[k];
        num += 4; numIn1++;
    }
    void shiftUp(bool force) {
        if (numIn1 > 1000 || force) {
            for (int k = 0; k < 4; k++) { SSEData1k[k] = SSEData[k] + SSEData1k[k]; SSEData[k] = 0; }
            numIn1k += numIn1; numIn1 = 0;
        }
        if (numIn1k > 1000 || force) {
            for (int k = 0; k < 4; k++) { SSEData1m[k] = SSEData1k[k] + SSEData1m[k]; SSEData1k[k] = 0; }
            numIn1m += numIn1k; numIn1k = 0;
        }
    }
};

// AccumulatorApprox (:749-1101): 13x13 symmetric [C(4) | xi(6) | ab(2) | r(1)]
struct AccumulatorApprox {
    float H[13 * 13];  // row-major (symmetric anyway)
    size_t num;
    float Data[60], Data1k[60], Data1m[60];
    float TopRight_Data[32], TopRight_Data1k[32], TopRight_Data1m[32];
    float BotRight_Data[8], BotRight_Data1k[8], BotRight_Data1m[8];
    float numIn1, numIn1k, numIn1m;

    void initialize() {
        memset(Data, 0, sizeof(Data)); memset(Data1k, 0, sizeof(Data1k)); memset(Data1m, 0, sizeof(Data1m));
        memset(TopRight_Data, 0, sizeof(TopRight_Data)); memset(TopRight_Data1k, 0, sizeof(TopRight_Data1k));
        memset(TopRight_Data1m, 0, sizeof(TopRight_Data1m));
        memset(BotRight_Data, 0, sizeof(BotRight_Data)); memset(BotRight_Data1k, 0, sizeof(BotRight_Data1k));
        memset(BotRight_Data1m, 0, sizeof(BotRight_Data1m));
        num = 0; numIn1 = numIn1k = numIn1m = 0;
    }
    void finish() {  // :771-801
        memset(H, 0, sizeof(H));
        shiftUp(true);
        int idx = 0;
        for (int r = 0; r < 10; r++)
            for (int c = r; c < 10; c++) { H[r * 13 + c] = H[c * 13 + r] = Data1m[idx]; idx++; }
        idx = 0;
        for (int r = 0; r < 10; r++)
            for (int c = 0; c < 3; c++) { H[r * 13 + c + 10] = H[(c + 10) * 13 + r] = TopRight_Data1m[idx]; idx++; }
        H[10 * 13 + 10] = BotRight_Data1m[0];
        H[10 * 13 + 11] = H[11 * 13 + 10] = BotRight_Data1m[1];
        H[10 * 13 + 12] = H[12 * 13 + 10] = BotRight_Data1m[2];
        H[11 * 13 + 11] = BotRight_Data1m[3];
        H[11 * 13 + 12] = H[12 * 13 + 11] = BotRight_Data1m[4];
        H[12 * 13 + 12] = BotRight_Data1m[5];
        num = (size_t) (numIn1 + numIn1k + numIn1m);
    }
    // :893-981 — x = [x4;x6], y = [y4;y6]; Data[(r,c>=r)] += a*x[c]*x[r] + c*y[c]*y[r] + b*(x[c]*y[r] + y[c]*x[r])
    void update(const float *x4, const float *x6, const float *y4, const float *y6, float a, float b, float c) {
        float x[10], y[10];
        for (int i = 0; i < 4; i++) { x[i] = x4[i]; y[i] = y4[i]; }
        for (int i = 0; i < 6; i++) { x[4 + i] = x6[i]; y[4 + i] = y6[i]; }
        int idx = 0;
        for (int r = 0; r < 10; r++)
            for (int cc = r; cc < 10; cc++) {
                Data[idx] += a * x[cc] * x[r] + c * y[cc] * y[r] + b * (x[cc] * y[r] + y[cc] * x[r]);
                idx++;
            }
        num++;
        numIn1++;
        shiftUp(false);
    }
    // :984-1030 — TopRight[3*i+k] += x[i]*TRk0 + y[i]*TRk1
    void updateTopRight(const float *x4, const float *x6, const float *y4, const float *y6,
                        float TR00, float TR10, float TR01, float TR11, float TR02, float TR12) {
        float x[10], y[10];
        for (int i = 0; i < 4; i++) { x[i] = x4[i]; y[i] = y4[i]; }
        for (int i = 0; i < 6; i++) { x[4 + i] = x6[i]; y[4 + i] = y6[i]; }
        for (int i = 0; i < 10; i++) {
            TopRight_Data[3 * i + 0] += x[i] * TR00 + y[i] * TR10;
            TopRight_Data[3 * i + 1] += x[i] * TR01 + y[i] * TR11;
            TopRight_Data[3 * i + 2] += x[i] * TR02 + y[i] * TR12;
        }
    }
    void updateBotRight(float a00, float a01, float a02, float a11, float a12, float a22) {  // :1032-1045
        BotRight_Data[0] += a00; BotRight_Data[1] += a01; BotRight_Data[2] += a02;
        BotRight_Data[3] += a11; BotRight_Data[4] += a12; BotRight_Data[5] += a22;
    }
    void shiftUp(bool force) {  // :1065-1100
        if (numIn1 > 1000 || force) {
            for (int i = 0; i < 60; i++) { Data1k[i] = Data[i] + Data1k[i]; Data[i] = 0; }
            for (int i = 0; i < 32; i++) { TopRight_Data1k[i] = TopRight_Data[i] + TopRight_Data1k[i]; TopRight_Data[i] = 0; }
            for (int i = 0; i < 8; i++) { BotRight_Data1k[i] = BotRight_Data[i] + BotRight_Data1k[i]; BotRight_Data[i] = 0; }
            numIn1k += numIn1; numIn1 = 0;
        }
        if (numIn1k > 1000 || force) {
            for (int i = 0; i < 60; i++) { Data1m[i] = Data1k[i] + Data1m[i]; Data1k[i] = 0; }
            for (int i = 0; i < 32; i++) { TopRight_Data1m[i] = TopRight_Data1k[i] + TopRight_Data1m[i]; TopRight_Data1k[i] = 0; }
            for (int i = 0; i < 8; i++) { BotRight_Data1m[i] = BotRight_Data1k[i] + BotRight_Data1m[i]; BotRight_Data1k[i] = 0; }
            numIn1m += numIn1k; numIn1k = 0;
        }
    }
};

// Accumulator9 (:1104-1135, updateSSE_eighted :1250-1369, shiftUp :1624-1642)
// 45 upper-triangle entries x 4 SSE lanes.
struct Accumulator9 {
    float H[9 * 9];
    size_t num;
    float SSEData[4 * 45], SSEData1k[4 * 45], SSEData1m[4 * 45];
    float numIn1, numIn1k, numIn1m;
    void initialize() {
        memset(H, 0, sizeof(H));
        memset(SSEData, 0, sizeof(SSEData)); memset(SSEData1k, 0, sizeof(SSEData1k));
        memset(SSEData1m, 0, sizeof(SSEData1m));
        num = 0; numIn1 = numIn1k = numIn1m = 0;
    }
    void finish() {
        memset(H, 0, sizeof(H));
        shiftUp(true);
        int idx = 0;
        for (int r = 0; r < 9; r++)
            for (int c = r; c < 9; c++) {
                float d = SSEData1m[idx + 0] + SSEData1m[idx + 1] + SSEData1m[idx + 2] + SSEData1m[idx + 3];
                H[r * 9 + c] = H[c * 9 + r] = d;
                idx += 4;
            }
    }
    // J[k][lane], w[lane]: SSEData[(r,c>=r)][lane] += (J[r]*w) * J[c]
    void updateSSE_eighted(const float J[9][4], const float w[4]) {
        float *pt = SSEData;
        for (int r = 0; r < 9; r++) {
            float Jw[4];
            for (int l = 0; l < 4; l++) Jw[l] = J[r][l] * w[l];
            for (int c = r; c < 9; c++) {
                for (int l = 0; l < 4; l++) pt[l] = pt[l] + Jw[l] * J[c][l];
                pt += 4;
            }
        }
        num += 4;
        numIn1++;
        shiftUp(false);
    }
    // updateSSE (MatrixAccumulators.h:1138-1248): SSEData[(r, c >= r)][lane] += J[r] * J[c], four residuals at a time
    void updateSSE(const float J[9][4]) {
        float *pt = SSEData;
        for (int r = 0; r < 9; r++)
            for (int c = r; c < 9; c++) {
                for (int l = 0; l < 4; l++) pt[l] = pt[l] + J[r][l] * J[c][l];
                pt += 4;
            }
        num += 4;
        numIn1++;
        shiftUp(false);
    }
    // updateSingle (:1372-1487): one residual into lane `off`
    void updateSingle(const float J[9], int off = 0) {
        float *pt = SSEData + off;
        for (int r = 0; r < 9; r++)
            for (int c = r; c < 9; c++) { *pt += J[c] * J[r]; pt += 4; }
        num++;
        numIn1++;
        shiftUp(false);
    }
    // updateSingleWeighted (:1489-1604): the diagonal term is (J_r * J_r) * w, then J_r *= w feeds the rest of its row
    void updateSingleWeighted(const float Jin[9], float w, int off = 0) {
        float J[9];
        for (int i = 0; i < 9; i++) J[i] = Jin[i];
        float *pt = SSEData + off;
        for (int r = 0; r < 9; r++) {
            *pt += J[r] * J[r] * w; pt += 4;
            if (r < 8) J[r] *= w;
            for (int c = r + 1; c < 9; c++) { *pt += J[c] * J[r]; pt += 4; }
        }
        num++;
        numIn1++;
        shiftUp(false);
    }
    void shiftUp(bool force) {
        if (numIn1 > 1000 || force) {
            for (int i = 0; i < 4 * 45; i++) { SSEData1k[i] = SSEData[i] + SSEData1k[i]; SSEData[i] = 0; }
            numIn1k += numIn1; numIn1 = 0;
        }
        if (numIn1k > 1000 || force) {
            for (int i = 0; i < 4 * 45; i++) { SSEData1m[i] = SSEData1k[i] + SSEData1m[i]; SSEData1k[i] = 0; }
            numIn1m += numIn1k; numIn1k = 0;
        }
    }
};

}  // namespace oracle
