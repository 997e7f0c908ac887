// Development aid: cycles of the solver's panel routine (k3_panel_row) for 1 and 3 warps, and the single-warp DFMA issue interval.
//   nvcc -arch=sm_100a -O3 -I ldso_b200/csrc tools/panel_bench.cu -o tools/panel_bench.bin
#include <cstdio>
#include <cuda_runtime.h>
#include "../ldso_b200/csrc/ba_k3.cuh"

__device__ __forceinline__ double lds64(unsigned a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory"); }
// the same block step with explicit 32-bit shared addresses (no generic pointers), fully unrolled
__device__ __forceinline__ void panel_row_s(unsigned sA, unsigned sWp, unsigned sV, int k0, int i) {
    double D[36], a[K3_NB];
#pragma unroll
    for (int r = 0; r < K3_NB; r++)
#pragma unroll
        for (int c = 0; c <= r; c++) D[K3_TRI(r, c)] = lds64(sA + 8 * ((k0 + r) * K3_LD + k0 + c));
#pragma unroll
    for (int c = 0; c < K3_NB; c++) a[c] = lds64(sA + 8 * (i * K3_LD + k0 + c));
    const bool below = i >= k0 + K3_NB;
#pragma unroll
    for (int C = 0; C < K3_NB; C++) {
        const double dk = D[K3_TRI(C, C)];
        double sq = 0.0;
        if (C + 1 < K3_NB) sq = D[K3_TRI((C + 1) & 7, C)] * D[K3_TRI((C + 1) & 7, C)];
        const double inv = (fabs(dk) > 0.0) ? k3_rcp(dk) : 1.0;
        if (C + 1 < K3_NB) D[K3_TRI((C + 1) & 7, (C + 1) & 7)] = fma(-sq, inv, D[K3_TRI((C + 1) & 7, (C + 1) & 7)]);
        const double w = a[C], l = w * inv;
#pragma unroll
        for (int r = C + 1; r < K3_NB; r++) {
            const double lr = D[K3_TRI(r, C)] * inv;
#pragma unroll
            for (int q = C + 1; q <= r; q++)
                if (!(r == C + 1 && q == C + 1)) D[K3_TRI(r, q)] = fma(-lr, D[K3_TRI(q, C)], D[K3_TRI(r, q)]);
            a[r] = fma(-l, D[K3_TRI(r, C)], a[r]);
        }
        if (k0 + C < i) {
            sts64(sA + 8 * (i * K3_LD + k0 + C), l);
            if (below) sts64(sWp + 8 * (C * K3_WPLD + i), w);
        } else if (k0 + C == i) sts64(sA + 8 * (i * K3_LD + i), w);
        if (i == k0) sts64(sV + 8 * (k0 + C), inv);
    }
}
__global__ void kpanel_s(int nthreads, long long *cyc, double *out) {
    extern __shared__ double sm[];
    double *A = sm, *Wp = A + K3_A_DOUBLES, *vinv = Wp + 2 * K3_NB * K3_WPLD;
    for (int e = threadIdx.x; e < K3_A_DOUBLES; e += blockDim.x) { const int r = e / K3_LD, c = e % K3_LD; A[e] = (r == c) ? 4.0 + 0.01 * r : 0.01 / (1 + ((r * 7 + c * 3) % 11)); }
    __syncthreads();
    const unsigned sA = (unsigned) __cvta_generic_to_shared(A), sWp = (unsigned) __cvta_generic_to_shared(Wp), sV = (unsigned) __cvta_generic_to_shared(vinv);
    for (int rep = 0; rep < 4; rep++) {
        __syncthreads();
        const long long t0 = clock64();
        if ((int) threadIdx.x < nthreads) panel_row_s(sA, sWp, sV, 0, threadIdx.x);
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0 && threadIdx.x < 96) cyc[rep * 4 + (threadIdx.x >> 5)] = t1 - t0;
    }
    out[threadIdx.x] = A[threadIdx.x] + Wp[threadIdx.x] + vinv[threadIdx.x & 7];
}
__global__ void kpanel(int nthreads, int full, long long *cyc, double *out) {
    extern __shared__ double sm[];
    double *A = sm, *Wp = A + K3_A_DOUBLES, *vinv = Wp + 2 * K3_NB * K3_WPLD;
    for (int e = threadIdx.x; e < K3_A_DOUBLES; e += blockDim.x) { const int r = e / K3_LD, c = e % K3_LD; A[e] = (r == c) ? 4.0 + 0.01 * r : 0.01 / (1 + ((r * 7 + c * 3) % 11)); }
    __syncthreads();
    for (int rep = 0; rep < 4; rep++) {
        __syncthreads();
        const long long t0 = clock64();
        if ((int) threadIdx.x < nthreads) k3_panel_row(A, Wp, vinv, 0, threadIdx.x, 0, full ? K3_NB : 1);
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0 && threadIdx.x < 96) cyc[rep * 4 + (threadIdx.x >> 5)] = t1 - t0;
    }
    out[threadIdx.x] = A[threadIdx.x] + Wp[threadIdx.x] + vinv[threadIdx.x & 7];
}
template<int CH> __global__ void kdfma(long long *cyc, double *out, double m) {
    double a[CH];
    for (int i = 0; i < CH; i++) a[i] = 1.0 + i + threadIdx.x * 1e-12;
    const long long t0 = clock64();
    for (int i = 0; i < 256; i++)
#pragma unroll
        for (int q = 0; q < CH; q++) a[q] = fma(a[q], m, 1e-9);
    const long long t1 = clock64();
    double s = 0; for (int i = 0; i < CH; i++) s += a[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void ks2r(long long *cyc, unsigned *out) {
    unsigned acc = 0;
    const long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < 64; i++) { unsigned v; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(v)); acc += v; }
    const long long t1 = clock64();
    unsigned acc2 = 0;
#pragma unroll
    for (int i = 0; i < 64; i++) { unsigned v; asm volatile("mov.u32 %0, %%smid;" : "=r"(v)); acc2 += v; }
    const long long t2 = clock64();
    out[threadIdx.x] = acc + acc2;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; }
}
// generic pointers into shared memory handed to a noinline function: what address arithmetic does the compiler emit?
__device__ __noinline__ double gsum(double *p, int n, int stride) { double s = 0; for (int i = 0; i < n; i++) { s += p[i * stride]; p[i * stride] = s; } return s; }
__global__ void kgen(long long *cyc, double *out) {
    extern __shared__ double sm[];
    for (int e = threadIdx.x; e < 4096; e += blockDim.x) sm[e] = e;
    __syncthreads();
    const long long t0 = clock64();
    double s = 0;
    for (int r = 0; r < 16; r++) s += gsum(sm + 64 * r + threadIdx.x, 8, 33);
    const long long t1 = clock64();
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template<int CH> __global__ void kdfma3(long long *cyc, double *out, double m) {
    double a[CH], b[CH], c[CH];
    for (int i = 0; i < CH; i++) { a[i] = 1.0 + i + threadIdx.x * 1e-12; b[i] = m + i * 1e-9; c[i] = 1e-9 * (i + 1) + threadIdx.x * 1e-15; }
    const long long t0 = clock64();
    for (int i = 0; i < 256; i++)
#pragma unroll
        for (int q = 0; q < CH; q++) a[q] = fma(b[q], c[(q + 1) % CH], a[q]);
    const long long t1 = clock64();
    double s = 0; for (int i = 0; i < CH; i++) s += a[i] + b[i] + c[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void kpanel_loop(int reps, double *out) {
    extern __shared__ double sm[];
    double *A = sm, *Wp = A + K3_A_DOUBLES, *vinv = Wp + 2 * K3_NB * K3_WPLD;
    for (int e = threadIdx.x; e < K3_A_DOUBLES; e += blockDim.x) { const int r = e / K3_LD, c = e % K3_LD; A[e] = (r == c) ? 4.0 + 0.01 * r : 0.01 / (1 + ((r * 7 + c * 3) % 11)); }
    __syncthreads();
    const unsigned sA = (unsigned) __cvta_generic_to_shared(A), sWp = (unsigned) __cvta_generic_to_shared(Wp), sV = (unsigned) __cvta_generic_to_shared(vinv);
    if (threadIdx.x < 32) for (int rep = 0; rep < reps; rep++) panel_row_s(sA, sWp, sV, 0, threadIdx.x);
    out[threadIdx.x] = A[threadIdx.x] + Wp[threadIdx.x] + vinv[threadIdx.x & 7];
}
int main(int argc, char **argv) {
    long long *cyc; double *out;
    cudaMallocManaged(&cyc, 64 * 8); cudaMalloc(&out, 8192);
    const size_t smem = (K3_A_DOUBLES + 2 * K3_NB * K3_WPLD + 128) * 8;
    cudaFuncSetAttribute(kpanel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    for (int full = 1; full >= 0; full--)
        for (int nt : {32, 73}) {
            kpanel<<<1, 512, smem>>>(nt, full, cyc, out); cudaDeviceSynchronize();
            printf("panel %s, %2d row threads: cycles per call (4 reps) warp0:", full ? "8 columns" : "1 column ", nt);
            for (int r = 0; r < 4; r++) printf(" %lld", cyc[4 * r]);
            if (nt > 64) { printf("  warp2:"); for (int r = 0; r < 4; r++) printf(" %lld", cyc[4 * r + 2]); }
            printf("\n");
        }
    cudaFuncSetAttribute(kpanel_s, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    for (int nt : {32, 73}) {
        kpanel_s<<<1, 512, smem>>>(nt, cyc, out); cudaDeviceSynchronize();
        printf("panel (explicit shared addresses, unrolled) 8 columns, %2d row threads: warp0:", nt);
        for (int r = 0; r < 4; r++) printf(" %lld", cyc[4 * r]);
        printf("\n");
    }
    if (argc > 1) { cudaFuncSetAttribute(kpanel_loop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem); kpanel_loop<<<1, 512, smem>>>(3000, out); cudaDeviceSynchronize(); printf("loop done %s\n", cudaGetErrorString(cudaGetLastError())); return 0; }
    kdfma3<4><<<1, 32>>>(cyc, out, 1.0000001); cudaDeviceSynchronize(); printf("1 warp, 4 DFMA chains, 3 distinct register operands: %.2f cycles/DFMA\n", cyc[0] / 1024.0);
    kdfma3<8><<<1, 32>>>(cyc, out, 1.0000001); cudaDeviceSynchronize(); printf("1 warp, 8 DFMA chains, 3 distinct register operands: %.2f cycles/DFMA\n", cyc[0] / 2048.0);
    kdfma<1><<<1, 32>>>(cyc, out, 1.0000001); cudaDeviceSynchronize(); printf("1 warp, 1 DFMA chain : %.2f cycles/DFMA\n", cyc[0] / 256.0);
    kdfma<2><<<1, 32>>>(cyc, out, 1.0000001); cudaDeviceSynchronize(); printf("1 warp, 2 DFMA chains: %.2f cycles/DFMA\n", cyc[0] / 512.0);
    kdfma<4><<<1, 32>>>(cyc, out, 1.0000001); cudaDeviceSynchronize(); printf("1 warp, 4 DFMA chains: %.2f cycles/DFMA\n", cyc[0] / 1024.0);
    kdfma<8><<<1, 32>>>(cyc, out, 1.0000001); cudaDeviceSynchronize(); printf("1 warp, 8 DFMA chains: %.2f cycles/DFMA\n", cyc[0] / 2048.0);
    kdfma<8><<<1, 128>>>(cyc, out, 1.0000001); cudaDeviceSynchronize(); printf("4 warps, 8 DFMA chains: %.2f cycles/DFMA per warp\n", cyc[0] / 2048.0);
    kdfma<8><<<1, 512>>>(cyc, out, 1.0000001); cudaDeviceSynchronize(); printf("16 warps, 8 DFMA chains: %.2f cycles/DFMA per warp\n", cyc[0] / 2048.0);
    unsigned *ou; cudaMalloc(&ou, 4096);
    ks2r<<<1, 32>>>(cyc, ou); cudaDeviceSynchronize(); printf("S2R cluster_ctarank: %.1f cycles each; S2R smid: %.1f cycles each (64 back to back)\n", cyc[0] / 64.0, cyc[1] / 64.0);
    kgen<<<1, 32, 65536>>>(cyc, out); cudaDeviceSynchronize(); printf("16 calls of a noinline 8-element shared RMW chain through generic pointers: %lld cycles\n", cyc[0]);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
