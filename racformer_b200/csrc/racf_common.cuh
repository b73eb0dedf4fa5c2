// Shared device helpers for the sampling kernels (sm_100a).
//
// One "tap" = one bilinear sample of one feature level. Both operators on the hot path (MSMV sampling
// and multi-scale deformable attention) reduce to: decide which of the 4 corner pixels exist, fetch
// them, blend. The integer decisions are made in ONE place (tap_geometry) so the forward kernel, the
// backward kernel and the mask-dump entry point can never disagree.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace racf {

// Corner bits of TapGeom::mask (bit k set <=> corner k is inside the map and is read).
constexpr unsigned kTL = 1u, kTR = 2u, kBL = 4u, kBR = 8u;

struct TapGeom {
    int h_low, w_low;  // top-left corner (may be -1)
    float lh, lw;      // fractional position inside the cell
    unsigned mask;     // corner bits; 0 <=> tap contributes nothing
    bool in_range;     // the reference's outer predicate (h>-1 && w>-1 && h<H && w<W)
};

// Pixel-space geometry of one tap. Follows ms_deform_attn_im2col_bilinear
// (models/csrc/msmv_sampling/msmv_sampling_forward.cu:27-73) and the range predicate at :126.
// NaN coordinates fail every comparison and give mask 0.
__device__ __forceinline__ TapGeom tap_geometry(float h_im, float w_im, int H, int W) {
    TapGeom g;
    g.in_range = (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)H) && (w_im < (float)W);
    const float hf = floorf(h_im), wf = floorf(w_im);
    g.h_low = g.in_range ? (int)hf : 0;
    g.w_low = g.in_range ? (int)wf : 0;
    g.lh = h_im - hf;
    g.lw = w_im - wf;
    unsigned m = 0;
    if (g.in_range) {
        const bool t = g.h_low >= 0, b = g.h_low + 1 <= H - 1;
        const bool l = g.w_low >= 0, r = g.w_low + 1 <= W - 1;
        m = (t && l ? kTL : 0u) | (t && r ? kTR : 0u) | (b && l ? kBL : 0u) | (b && r ? kBR : 0u);
    }
    g.mask = m;
    return g;
}

// MSMV pixel coordinate: align_corners=True, a single fp32 multiply (msmv_sampling_forward.cu:123-124).
__device__ __forceinline__ float msmv_pixel(float loc, int size) { return loc * (float)(size - 1); }

// MSDA pixel coordinate: align_corners=False. nvcc contracts mmcv's `loc * size - 0.5` into one FMA; we
// make that explicit so the CPU oracle (fmaf) takes the same floor() at half-pixel boundaries.
__device__ __forceinline__ float msda_pixel(float loc, int size) { return fmaf(loc, (float)size, -0.5f); }

// View selection of the MSMV op: round-half-away-from-zero like CUDA round() (msmv_sampling_forward.cu:110).
__device__ __forceinline__ int msmv_view(float z, int num_views) { return (int)roundf(z * (float)(num_views - 1)); }

// 128-bit read-only load of feature data.
__device__ __forceinline__ float4 ldg128(const float4* p) { return __ldg(p); }

// 128-bit vector reduction into global memory (sm_90+): one L2 atomic transaction for 4 floats.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

// Number of SMs of the current device (B200: 148), queried per launch -- a driver attribute read, no synchronisation,
// legal during stream capture. Grid caps and persistent grids are sized from it instead of a hard-coded 148.
inline int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
        n = 148;
    return n;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}

}  // namespace racf
