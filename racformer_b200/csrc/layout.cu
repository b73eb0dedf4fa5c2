// Channel-last re-layout of one FPN level for the sampling op -- SURVEY.md 8f-3.
//
// The reference decoder does `feat.reshape(B,T,N,G,C,H,W).permute(0,1,3,2,5,6,4).reshape(B*T*G,N,H,W,C).contiguous()`
// (models/racformer_transformer.py:112-124) through PyTorch's generic strided copy (measured 0.87 ms for the 735 MB
// f8 pyramid on B200, 1.7 TB/s). This is the same copy as a tiled transpose: each (b,t,n,g) slab is a [C][H*W] matrix
// that becomes [H*W][C]; 64 channels x 32 pixels go through shared memory so that both the reads (along pixels) and
// the writes (along channels, 256 B per pixel) are coalesced.
#include <cuda_runtime.h>

#include "racformer_ops.h"

namespace racf {

constexpr int kTilePix = 32;

__global__ void __launch_bounds__(256) to_sampling_layout_c64_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                     int T, int N, int G, int HW) {
    __shared__ float tile[64][kTilePix + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tiles = (HW + kTilePix - 1) / kTilePix;
    const long long slab = blockIdx.x / tiles;            // ((b*T + t)*N + n)*G + g  in INPUT order
    const int p0 = (blockIdx.x % tiles) * kTilePix;
    const int g = (int)(slab % G);
    const long long btn = slab / G;
    const int n = (int)(btn % N);
    const long long bt = btn / N;
    const float* src = in + (btn * G + g) * 64 * (long long)HW;                 // [C][HW]
    float* dst = out + ((bt * G + g) * N + n) * (long long)HW * 64;             // [HW][C]
    const int p = p0 + lane;
#pragma unroll
    for (int c = warp; c < 64; c += 8) tile[c][lane] = (p < HW) ? __ldg(src + (long long)c * HW + p) : 0.f;
    __syncthreads();
#pragma unroll
    for (int q = warp; q < kTilePix; q += 8) {
        if (p0 + q < HW) {
            float* row = dst + (long long)(p0 + q) * 64;
            row[lane] = tile[lane][q];
            row[lane + 32] = tile[lane + 32][q];
        }
    }
}

// in [batch][C][S] -> out[batch][S][ld] (first C of ld floats per pixel) and, optionally, the same into out2 with its
// own row length: the NCHW -> NHWC conversion in front of the radar temporal encoder's convolutions, written once into
// the dense channel-last tensor and once into the first C channels of the [C + hidden]-channel concatenation buffer
// (PyTorch's generic strided copy takes 153 us for the 134 MB f8 BEV queue, this tiled transpose about a third).
__global__ void __launch_bounds__(256) chw_to_hwc_kernel(const float* __restrict__ in, float* __restrict__ out, int ld,
                                                         float* __restrict__ out2, int ld2, int C, int S) {
    __shared__ float tile[32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const long long b = blockIdx.z;
    const float* src = in + b * (long long)C * S;
#pragma unroll
    for (int c = warp; c < 32; c += 8)
        tile[c][lane] = (c0 + c < C && s0 + lane < S) ? __ldg(src + (long long)(c0 + c) * S + s0 + lane) : 0.f;
    __syncthreads();
#pragma unroll
    for (int q = warp; q < 32; q += 8) {
        if (s0 + q < S && c0 + lane < C) {
            const float v = tile[lane][q];
            out[(b * S + s0 + q) * ld + c0 + lane] = v;
            if (out2 != nullptr) out2[(b * S + s0 + q) * ld2 + c0 + lane] = v;
        }
    }
}

}  // namespace racf

extern "C" int racf_chw_to_hwc(const float* in, int batch, int channels, int positions, float* out, int ld, float* out2,
                               int ld2, racf_stream_t stream) {
    if (!in || !out) return RACF_ERR_NULL_POINTER;
    if (batch <= 0 || channels <= 0 || positions <= 0 || batch > 65535 || ld < channels || (out2 && ld2 < channels))
        return RACF_ERR_BAD_SHAPE;
    const dim3 grid((unsigned)((positions + 31) / 32), (unsigned)((channels + 31) / 32), (unsigned)batch);
    if (grid.y > 65535) return RACF_ERR_BAD_SHAPE;
    racf::chw_to_hwc_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, ld, out2, ld2, channels, positions);
    return (int)cudaGetLastError();
}

// in: [B, T*N, G*64, H, W] contiguous -> out: [B*T*G, N, H, W, 64] contiguous. Only C == 64 is implemented.
extern "C" int racf_to_sampling_layout(const float* in, float* out, int batch, int num_frames, int num_views,
                                       int num_groups, int channels, int height, int width, racf_stream_t stream) {
    if (!in || !out) return RACF_ERR_NULL_POINTER;
    if (batch <= 0 || num_frames <= 0 || num_views <= 0 || num_groups <= 0 || height <= 0 || width <= 0 || channels != 64)
        return RACF_ERR_BAD_SHAPE;
    const int HW = height * width;
    const long long slabs = (long long)batch * num_frames * num_views * num_groups;
    const long long blocks = slabs * ((HW + racf::kTilePix - 1) / racf::kTilePix);
    if (blocks >= (1LL << 31)) return RACF_ERR_BAD_SHAPE;
    racf::to_sampling_layout_c64_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, out, num_frames, num_views, num_groups, HW);
    return (int)cudaGetLastError();
}
