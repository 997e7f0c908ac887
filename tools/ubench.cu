// Dependent-issue latency micro-benchmarks (one warp) for the FP64 / shuffle / shared-memory ops the solver kernels
// chain on their critical paths. Development aid: nvcc -arch=sm_100a -O3 tools/ubench.cu -o gpurun_out/ubench
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
template<int OP> __global__ void k(double *out, long long *cyc, double seed, double m) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double x = seed + threadIdx.x * 1e-12, y = m;
    int idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = fma(x, y, 1e-9);
        if (OP == 1) x = x + y;
        if (OP == 2) x = x * y;
        if (OP == 3) x = __drcp_rn(x) + 0.5;
        if (OP == 4) x = 1.0 / x + 0.5;
        if (OP == 5) x = sqrt(x) + 0.5;
        if (OP == 6) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
        if (OP == 7) { idx = (int) sm[idx & 1023 ] + (idx & 1); x += idx; }   // LDS + cvt chain
        if (OP == 8) { float f = (float) x; f = fmaf(f, 1.0000001f, 1e-9f); x = f; }
        if (OP == 9) x = sm[((int) __double2loint(x)) & 1023];                      // LDS dependent (addr from data)
        if (OP == 10) x = rsqrt(x) + 0.5;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + idx;
    if (threadIdx.x == 0) cyc[OP] = (t1 - t0);
}
template<int OP> __global__ void kf(float *out, long long *cyc, float seed, float m) {
    float x = seed + threadIdx.x * 1e-6f, y = m;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = fmaf(x, y, 1e-9f);
        if (OP == 1) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
        if (OP == 2) x = __frcp_rn(x) + 0.5f;
        if (OP == 3) x = __fdividef(1.0f, x) + 0.5f;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[16 + OP] = (t1 - t0);
}
// throughput: many warps of independent DFMA chains
__global__ void kthr(double *out, long long *cyc, double seed, double m) {
    double a[8];
    for (int i = 0; i < 8; i++) a[i] = seed + i + threadIdx.x * 1e-12;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int q = 0; q < 8; q++) a[q] = fma(a[q], m, 1e-9);
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 8; i++) s += a[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[24 + (blockDim.x >> 8)] = t1 - t0;
}
int main() {
    double *out; long long *cyc; float *outf;
    cudaMalloc(&out, 8192); cudaMalloc(&outf, 8192); cudaMallocManaged(&cyc, 64 * 8);
    for (int r = 0; r < 2; r++) {
        k<0><<<1, 32>>>(out, cyc, 1.0, 1.0000001); k<1><<<1, 32>>>(out, cyc, 1.0, 1e-9); k<2><<<1, 32>>>(out, cyc, 1.0, 1.0000001);
        k<3><<<1, 32>>>(out, cyc, 1.3, 1.0); k<4><<<1, 32>>>(out, cyc, 1.3, 1.0); k<5><<<1, 32>>>(out, cyc, 1.3, 1.0);
        k<6><<<1, 32>>>(out, cyc, 1.3, 1.0); k<7><<<1, 32>>>(out, cyc, 1.3, 1.0); k<8><<<1, 32>>>(out, cyc, 1.3, 1.0);
        k<9><<<1, 32>>>(out, cyc, 1.3, 1.0); k<10><<<1, 32>>>(out, cyc, 1.3, 1.0);
        kf<0><<<1, 32>>>(outf, cyc, 1.0f, 1.0000001f); kf<1><<<1, 32>>>(outf, cyc, 1.0f, 1.0f); kf<2><<<1, 32>>>(outf, cyc, 1.3f, 1.0f); kf<3><<<1, 32>>>(outf, cyc, 1.3f, 1.0f);
        kthr<<<1, 256>>>(out, cyc, 1.0, 1.0000001); kthr<<<1, 512>>>(out, cyc, 1.0, 1.0000001); kthr<<<1, 1024>>>(out, cyc, 1.0, 1.0000001);
        cudaDeviceSynchronize();
    }
    const char *nm[] = {"DFMA", "DADD", "DMUL", "__drcp_rn+DADD", "1.0/x+DADD", "sqrt+DADD", "SHFL f64", "LDS+cvt chain", "f64->f32 FFMA f32->f64", "LDS f64 dependent", "rsqrt+DADD"};
    for (int i = 0; i < 11; i++) printf("%-28s %7.1f cycles/op\n", nm[i], (double) cyc[i] / N);
    const char *nf[] = {"FFMA", "SHFL f32", "__frcp_rn+FADD", "__fdividef+FADD"};
    for (int i = 0; i < 4; i++) printf("%-28s %7.1f cycles/op\n", nf[i], (double) cyc[16 + i] / N);
    for (int w = 1; w <= 4; w <<= 1) printf("DFMA throughput %4d threads: %7.2f DFMA/clk/SM\n", 256 * w, 256.0 * w * 8 * N / (double) cyc[24 + w]);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
