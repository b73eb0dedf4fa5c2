// Measurement aid (not on the product path): the speed of light of the sampling kernels' access pattern.
// Every warp issues independent, fully coalesced 512-byte reads (32 lanes x 128 bit) at pseudo-random 512-byte-aligned
// offsets of a buffer -- the same request shape as one bilinear cell row of the channel-last pyramid -- with `ilp`
// loads in flight per warp and no arithmetic besides an add. bench/tools time it for an HBM-sized and an L2-sized
// footprint; the result is the denominator the gather kernels are compared with in DESIGN.md.
#include <cuda_runtime.h>
#include <stdint.h>

#include "racformer_tools.h"

namespace racf {

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int ILP>
__global__ void __launch_bounds__(256) gather_ceiling_kernel(const float4* __restrict__ buf, uint32_t num_rows,
                                                             int rows_per_warp, float* __restrict__ sink) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < rows_per_warp; i += ILP) {
        float4 v[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            const uint32_t row = mix(warp * 9781u + (uint32_t)(i + u) * 2654435761u) % num_rows;
            v[u] = __ldg(buf + (size_t)row * 32 + lane);
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;   // keep the loads alive
}

// Same request shape for the scatter: every warp issues red.global.add.v4.f32 on random 512-byte rows.
__global__ void __launch_bounds__(256) scatter_ceiling_kernel(float* __restrict__ buf, uint32_t num_rows, int rows_per_warp) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int i = 0; i < rows_per_warp; ++i) {
        const uint32_t row = mix(warp * 9781u + (uint32_t)i * 2654435761u) % num_rows;
        float* p = buf + ((size_t)row * 32 + lane) * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(1.0f) : "memory");
    }
}

}  // namespace racf

extern "C" int racf_bench_scatter_ceiling(float* buf, long long num_rows, long long total_rows, racf_stream_t stream) {
    if (!buf) return RACF_ERR_NULL_POINTER;
    if (num_rows <= 0 || num_rows >= (1LL << 32) || total_rows <= 0) return RACF_ERR_BAD_SHAPE;
    const int rows_per_warp = 64;
    const long long warps = (total_rows + rows_per_warp - 1) / rows_per_warp;
    racf::scatter_ceiling_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        buf, (uint32_t)num_rows, rows_per_warp);
    return (int)cudaGetLastError();
}

// Reads total_rows x 512 B at random 512-byte rows of buf[0 : num_rows*512 B]. ilp in {1,2,4,8,16}.
extern "C" int racf_bench_gather_ceiling(const float* buf, long long num_rows, long long total_rows, int ilp,
                                         float* sink, racf_stream_t stream) {
    if (!buf || !sink) return RACF_ERR_NULL_POINTER;
    if (num_rows <= 0 || num_rows >= (1LL << 32) || total_rows <= 0) return RACF_ERR_BAD_SHAPE;
    const int rows_per_warp = 64;
    const long long warps = (total_rows + rows_per_warp - 1) / rows_per_warp;
    const unsigned grid = (unsigned)((warps + 7) / 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float4* b4 = reinterpret_cast<const float4*>(buf);
    switch (ilp) {
        case 1: racf::gather_ceiling_kernel<1><<<grid, 256, 0, st>>>(b4, (uint32_t)num_rows, rows_per_warp, sink); break;
        case 2: racf::gather_ceiling_kernel<2><<<grid, 256, 0, st>>>(b4, (uint32_t)num_rows, rows_per_warp, sink); break;
        case 4: racf::gather_ceiling_kernel<4><<<grid, 256, 0, st>>>(b4, (uint32_t)num_rows, rows_per_warp, sink); break;
        case 8: racf::gather_ceiling_kernel<8><<<grid, 256, 0, st>>>(b4, (uint32_t)num_rows, rows_per_warp, sink); break;
        case 16: racf::gather_ceiling_kernel<16><<<grid, 256, 0, st>>>(b4, (uint32_t)num_rows, rows_per_warp, sink); break;
        default: return RACF_ERR_BAD_SHAPE;
    }
    return (int)cudaGetLastError();
}
