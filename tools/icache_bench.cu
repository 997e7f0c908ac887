// Instruction-fetch micro-benchmark for single-CTA latency-critical kernels (development aid).
//   nvcc -arch=sm_100a -O3 tools/icache_bench.cu -o gpurun_out/icache_bench
// Questions: (1) what does a first-touch 128-byte line of straight-line code cost one warp (SM-cold, L2-warm)? (2) do the fetches of
// DIFFERENT warps overlap (can a gang of warps pull a routine in concurrently)? (3) what does a line cost once another warp of the SM
// has fetched it (L1.5 hit)? (4) how large may a loop body be before every iteration re-fetches (L0 capacity)?
#include <cstdio>
#include <cuda_runtime.h>

template<int ID, int N> struct Seg {       // N x 4 independent FFMAs with distinct immediates: 4N instructions of straight-line code
    __device__ __forceinline__ static void run(float &a, float &b, float &c, float &d) {
        a = fmaf(a, 1.0001f + N * 1e-7f + ID * 1e-5f, 0.5f);
        b = fmaf(b, 1.0002f + N * 1e-7f + ID * 1e-5f, 0.25f);
        c = fmaf(c, 1.0003f + N * 1e-7f + ID * 1e-5f, 0.125f);
        d = fmaf(d, 1.0004f + N * 1e-7f + ID * 1e-5f, 0.0625f);
        Seg<ID, N - 1>::run(a, b, c, d);
    }
};
template<int ID> struct Seg<ID, 0> { __device__ __forceinline__ static void run(float &, float &, float &, float &) {} };

#define SEGN 64          // 64 x 4 = 256 FFMA = 4 KB = 32 lines per segment
template<int ID> __device__ __noinline__ float4 segv(float4 v) { float a = v.x, b = v.y, c = v.z, d = v.w; Seg<ID, SEGN>::run(a, b, c, d); return make_float4(a, b, c, d); }
template<int ID> __device__ __forceinline__ void seg(float &a, float &b, float &c, float &d) { const float4 r = segv<ID>(make_float4(a, b, c, d)); a = r.x; b = r.y; c = r.z; d = r.w; }
__device__ __forceinline__ void run_seg(int id, float &a, float &b, float &c, float &d) {
    switch (id) {
        case 0: seg<0>(a, b, c, d); break; case 1: seg<1>(a, b, c, d); break; case 2: seg<2>(a, b, c, d); break; case 3: seg<3>(a, b, c, d); break;
        case 4: seg<4>(a, b, c, d); break; case 5: seg<5>(a, b, c, d); break; case 6: seg<6>(a, b, c, d); break; case 7: seg<7>(a, b, c, d); break;
        case 8: seg<8>(a, b, c, d); break; case 9: seg<9>(a, b, c, d); break; case 10: seg<10>(a, b, c, d); break; case 11: seg<11>(a, b, c, d); break;
        case 12: seg<12>(a, b, c, d); break; case 13: seg<13>(a, b, c, d); break; case 14: seg<14>(a, b, c, d); break; default: seg<15>(a, b, c, d); break;
    }
}
// mode 0: warp 0 runs segments [0, nseg) one after the other. mode 1: warp w runs segment w (w < nseg), all at once.
// mode 2: warp 0 loops `reps` times over segments [0, nseg). Reports cycles of warp 0 (mode 0, 2) or of every warp (mode 1).
// others: 0 = every CTA runs the test; 1 = CTAs > 0 exit at once; 2 = CTAs > 0 spin until CTA 0 is done
__global__ void kcode(int mode, int nseg, int reps, float *out, long long *cyc, int others = 0, volatile int *flag = nullptr) {
    float a = threadIdx.x * 1e-3f, b = 1.f, c = 2.f, d = 3.f;
    const int w = threadIdx.x >> 5;
    if (blockIdx.x > 0) {
        if (others == 1) return;
        if (others == 2) { if (threadIdx.x == 0) while (*flag == 0) { } return; }
        cyc = cyc + 64;      // scratch
    }
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) { if (w == 0) for (int s = 0; s < nseg; s++) run_seg(s, a, b, c, d); }
    else if (mode == 1) { if (w < nseg) run_seg(w, a, b, c, d); }
    else { if (w == 0) for (int r = 0; r < reps; r++) { const long long t = clock64(); for (int s = 0; s < nseg; s++) run_seg(s, a, b, c, d); if ((threadIdx.x & 31) == 0) cyc[32 + r] = clock64() - t; } }
    const long long t1 = clock64();
    out[threadIdx.x] = a + b + c + d;
    if ((threadIdx.x & 31) == 0) cyc[w] = t1 - t0;
    if (flag && blockIdx.x == 0 && threadIdx.x == 0) { __threadfence(); *flag = 1; }
}
// evictor: a different big body on every SM (what the other kernels of the iteration do to the solver's code)
template<int ID> __global__ void kevict(float *out) {
    float a = threadIdx.x, b = 1.f, c = 2.f, d = 3.f;
    Seg<100 + 8 * ID, 150>::run(a, b, c, d); Seg<101 + 8 * ID, 150>::run(a, b, c, d); Seg<102 + 8 * ID, 150>::run(a, b, c, d); Seg<103 + 8 * ID, 150>::run(a, b, c, d);
    Seg<104 + 8 * ID, 150>::run(a, b, c, d); Seg<105 + 8 * ID, 150>::run(a, b, c, d); Seg<106 + 8 * ID, 150>::run(a, b, c, d); Seg<107 + 8 * ID, 150>::run(a, b, c, d);       // ~77 KB of code
    out[threadIdx.x + blockIdx.x * blockDim.x] = a + b + c + d;
}
int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 1 << 22); cudaMallocManaged(&cyc, 128 * 8);
    auto evict = [&]() { kevict<0><<<296, 256>>>(out); kevict<1><<<296, 256>>>(out); cudaDeviceSynchronize(); };
    auto show = [&](const char *what, int nw) { cudaDeviceSynchronize(); printf("%-70s", what); for (int i = 0; i < nw; i++) printf(" %lld", cyc[i]); printf("\n"); };
    printf("segment = %d FFMA = %d lines of 128 B\n", 4 * SEGN, 4 * SEGN / 8);
    for (int rep = 0; rep < 2; rep++) {
        evict(); kcode<<<1, 512>>>(0, 12, 1, out, cyc); show("SM-cold, 1 warp x 12 segments sequentially (cycles)", 1);
        kcode<<<1, 512>>>(0, 12, 1, out, cyc); show("  same again (hot in L1.5? 48 KB > 32 KB)", 1);
        evict(); kcode<<<1, 512>>>(0, 6, 1, out, cyc); show("SM-cold, 1 warp x 6 segments (24 KB)", 1);
        kcode<<<1, 512>>>(0, 6, 1, out, cyc); show("  same again (24 KB < L1.5)", 1);
        evict(); kcode<<<1, 512>>>(1, 12, 1, out, cyc); show("SM-cold, 12 warps x 1 segment each, concurrently (cycles per warp)", 12);
        kcode<<<1, 512>>>(0, 12, 1, out, cyc); show("  then 1 warp x 12 segments sequentially", 1);
        evict(); kcode<<<1, 512>>>(1, 6, 1, out, cyc); show("SM-cold, 6 warps x 1 segment each, concurrently", 6);
        kcode<<<1, 512>>>(0, 6, 1, out, cyc); show("  then 1 warp x 6 segments sequentially (all in L1.5, other SMSPs' L0)", 1);
        for (int ns = 1; ns <= 4; ns++) {
            evict(); kcode<<<1, 512>>>(2, ns, 6, out, cyc); cudaDeviceSynchronize();
            printf("loop body %2d KB, 6 iterations (cycles each):", 4 * ns); for (int r = 0; r < 6; r++) printf(" %lld", cyc[32 + r]); printf("\n");
        }
    }
    int *flag; cudaMalloc(&flag, 4);
    for (int others = 0; others < 3; others++)
        for (int grid : {1, 2, 8, 74, 148, 296}) {
            cudaMemset(flag, 0, 4);
            evict(); kcode<<<grid, 512>>>(2, 2, 6, out, cyc, others, flag); cudaDeviceSynchronize();
            printf("grid %3d others-mode %d: loop body 8 KB, 6 iterations (cycles each):", grid, others); for (int r = 0; r < 6; r++) printf(" %lld", cyc[32 + r]); printf("\n");
        }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
