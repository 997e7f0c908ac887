// Keyframe image pyramids on the device: 16-byte texels (I, dI/dx, dI/dy, 0) so that one bilinear tap is one
// LDG.128. Either repacked from the reference's Eigen::Vector3f arrays (FrameHessian::dIp) or built on the device
// from the raw irradiance image exactly as FrameHessian::makeImages does (src/internal/FrameHessian.cc:44-98).
#pragma once
#include "common.cuh"

__global__ void k_repack_aos3(const float *__restrict__ src, float4 *__restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
}

__global__ void k_unpack_aos3(const float4 *__restrict__ src, float *__restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = src[i];
    dst[3 * i] = v.x; dst[3 * i + 1] = v.y; dst[3 * i + 2] = v.z;
}

// level 0: intensity from the raw image; level l>0: 2x2 box filter of level l-1 (FrameHessian.cc:69-81)
__global__ void k_pyr_intensity(const float *__restrict__ color, const float4 *__restrict__ prev, float4 *__restrict__ dst,
                                int wl, int hl, int wlm1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= wl * hl) return;
    float v;
    if (prev == nullptr) v = color[i];
    else {
        const int x = i % wl, y = i / wl;
        const float4 *b = prev + 2 * x + 2 * y * wlm1;
        v = 0.25f * (b[0].x + b[1].x + b[wlm1].x + b[wlm1 + 1].x);
    }
    dst[i] = make_float4(v, 0.f, 0.f, 0.f);
}

// central differences over the flat index range [wl, wl*(hl-1)) (FrameHessian.cc:83-92); dx at a row border reads
// the neighbouring row's pixel exactly like the reference's flat indexing does.
__global__ void k_pyr_gradients(float4 *__restrict__ img, int wl, int hl) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < wl || idx >= wl * (hl - 1)) return;
    float dx = 0.5f * (img[idx + 1].x - img[idx - 1].x);
    float dy = 0.5f * (img[idx + wl].x - img[idx - wl].x);
    if (isnan(dx) || fabsf(dx) > 255.0f) dx = 0;
    if (isnan(dy) || fabsf(dy) > 255.0f) dy = 0;
    float *px = (float *) (img + idx);   // only dx,dy are written: neighbours read .x concurrently
    px[1] = dx;
    px[2] = dy;
}
