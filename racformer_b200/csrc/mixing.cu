// Fused core of AdaptiveMixing (AdaMixer) for inference -- SURVEY.md 8f-4, consumer of the MSMV output.
//
// Reference: AdaptiveMixing.inner_forward, models/racformer_transformer.py:580-610. Per (query, group):
//     out = x @ M                      x [P_in, C] sampled features, M [C, C]      (adaptive channel mixing)
//     out = relu(layer_norm(out))      normalised over all P_in * C elements, no affine, eps 1e-5
//     out = S @ out                    S [P_out, P_in]                             (adaptive point mixing)
//     out = relu(layer_norm(out))      over all P_out * C elements
// where M and S are slices of `parameter_generator(query)`. PyTorch runs this as 2 batched SGEMMs, 2 layer norms and
// 2 ReLUs: six launches and ~0.9 GB of intermediate traffic per decoder iteration (0.43 ms on B200). Here one CTA
// keeps x, M, S and the intermediate in shared memory (88 KB at P_in = 96) and does both products on the CUDA cores
// in fp32 with register tiles -- the same arithmetic as SGEMM (fp32 FMA), only the summation order differs.
//
// Layouts: x [BQ, G, P_in, C] (what racf_msmv_forward_grouped writes), params [BQ, G, C*C + P_out*P_in] (M row-major
// [C][C], then S row-major [P_out][P_in]), out [BQ, G, P_out, C]. C == 64, P_out == 128, P_in % 4 == 0, P_in <= 128.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "linear_tiled.cuh"
#include "racformer_ops.h"

namespace racf {

constexpr int kMixC = 64;        // channels per group
constexpr int kMixPout = 128;    // out_points
constexpr int kMixThreads = 256;
constexpr int kMixMaxRows = 8;   // rows per thread in either product (P_in, P_out <= 128)

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// Two-pass mean / variance over the values held in the CTA's register tiles, like F.layer_norm in fp32.
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();               // `red` may still be read by the previous reduction
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kMixThreads / 32; ++w) t += red[w];
    return t;
}

// element (row, c) of a [rows][64] tile whose 16-byte chunks are XOR-swizzled by the row, so that the two rows a warp
// touches in the same instruction fall into different banks although the row stride is a multiple of 32 words
__device__ __forceinline__ int swz64(int row, int chunk) { return row * kMixC + ((chunk ^ (row & 7)) << 2); }

// kSplitOut: write the result as three bf16 pieces (out3 [3][QG][128][64], value == p0 + p1 + p2 exactly) -- the A
// operand of the tensor-core out_proj (csrc/linear.cu) -- instead of fp32, saving a 118 MB read + 177 MB write pass.
template <bool kSplitOut>
__global__ void __launch_bounds__(kMixThreads, 2)
adaptive_mixing_kernel(const float* __restrict__ x, const float* __restrict__ params, float* __restrict__ out,
                       __nv_bfloat16* __restrict__ out3, long long piece_stride, int tiled_groups, int p_in, float eps) {
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                          // [p_in][64] swizzled; reused for the normalised intermediate
    float* ms = xs + 128 * kMixC;              // [64][64]
    float* ss = ms + kMixC * kMixC;            // [128][p_in] (chunks swizzled by row when p_in % 32 == 0)
    __shared__ float red[kMixThreads / 32];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;    // 16 column groups of 4 channels x 16 row groups
    const long long qg = blockIdx.x;
    const float* xg = x + qg * (long long)p_in * kMixC;
    const float* pg = params + qg * (long long)(kMixC * kMixC + kMixPout * p_in);
    const bool swz_s = (p_in & 31) == 0;
    const int in_chunks = p_in >> 2;

    // ---- stage x, M (group 0) and S (group 1) with cp.async ------------------------------------------------------
    for (int i = tid; i < p_in * 16; i += kMixThreads) {
        const int row = i >> 4, ch = i & 15;
        cp_async16(xs + swz64(row, ch), xg + row * kMixC + ch * 4);
    }
    for (int i = tid; i < kMixC * 16; i += kMixThreads) cp_async16(ms + i * 4, pg + i * 4);
    cp_async_commit();
    const float* sg = pg + kMixC * kMixC;
    for (int i = tid; i < kMixPout * in_chunks; i += kMixThreads) {
        const int row = i / in_chunks, ch = i - row * in_chunks;
        const int dch = swz_s ? (ch ^ (row & 7)) : ch;
        cp_async16(ss + row * p_in + dch * 4, sg + row * p_in + ch * 4);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    // ---- channel mixing: t[row][4tx..] = sum_c x[row][c] * M[c][4tx..], rows ty, ty+16, ... -----------------------
    const int rows_a = (p_in - ty + 15) >> 4;   // rows owned by this thread (<= 8)
    float4 acc[kMixMaxRows];
#pragma unroll
    for (int i = 0; i < kMixMaxRows; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k4 = 0; k4 < 16; ++k4) {
        float4 mrow[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) mrow[kk] = *reinterpret_cast<const float4*>(ms + (k4 * 4 + kk) * kMixC + tx * 4);
#pragma unroll
        for (int i = 0; i < kMixMaxRows; ++i) {
            if (i < rows_a) {
                const float4 xv = *reinterpret_cast<const float4*>(xs + swz64(ty + 16 * i, k4));
                acc[i].x = fmaf(xv.x, mrow[0].x, acc[i].x); acc[i].y = fmaf(xv.x, mrow[0].y, acc[i].y);
                acc[i].z = fmaf(xv.x, mrow[0].z, acc[i].z); acc[i].w = fmaf(xv.x, mrow[0].w, acc[i].w);
                acc[i].x = fmaf(xv.y, mrow[1].x, acc[i].x); acc[i].y = fmaf(xv.y, mrow[1].y, acc[i].y);
                acc[i].z = fmaf(xv.y, mrow[1].z, acc[i].z); acc[i].w = fmaf(xv.y, mrow[1].w, acc[i].w);
                acc[i].x = fmaf(xv.z, mrow[2].x, acc[i].x); acc[i].y = fmaf(xv.z, mrow[2].y, acc[i].y);
                acc[i].z = fmaf(xv.z, mrow[2].z, acc[i].z); acc[i].w = fmaf(xv.z, mrow[2].w, acc[i].w);
                acc[i].x = fmaf(xv.w, mrow[3].x, acc[i].x); acc[i].y = fmaf(xv.w, mrow[3].y, acc[i].y);
                acc[i].z = fmaf(xv.w, mrow[3].z, acc[i].z); acc[i].w = fmaf(xv.w, mrow[3].w, acc[i].w);
            }
        }
    }
    // ---- layer norm over p_in * 64 values + ReLU, written back into xs (swizzled) as the B operand of the 2nd product
    {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kMixMaxRows; ++i)
            if (i < rows_a) s += (acc[i].x + acc[i].y) + (acc[i].z + acc[i].w);
        const float n = (float)(p_in * kMixC);
        const float mean = block_sum(s, red) / n;
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < kMixMaxRows; ++i)
            if (i < rows_a) {
                const float a = acc[i].x - mean, b = acc[i].y - mean, c = acc[i].z - mean, d = acc[i].w - mean;
                v += (a * a + b * b) + (c * c + d * d);
            }
        const float rstd = rsqrtf(block_sum(v, red) / n + eps);
        __syncthreads();   // every thread has finished reading xs as the A operand
#pragma unroll
        for (int i = 0; i < kMixMaxRows; ++i)
            if (i < rows_a) {
                float4 o;
                o.x = fmaxf((acc[i].x - mean) * rstd, 0.f);
                o.y = fmaxf((acc[i].y - mean) * rstd, 0.f);
                o.z = fmaxf((acc[i].z - mean) * rstd, 0.f);
                o.w = fmaxf((acc[i].w - mean) * rstd, 0.f);
                *reinterpret_cast<float4*>(xs + swz64(ty + 16 * i, tx)) = o;
            }
    }
    cp_async_wait<0>();
    __syncthreads();

    // ---- point mixing: out[row][4tx..] = sum_p S[row][p] * t[p][4tx..], rows ty, ty+16, ..., ty+112 ----------------
#pragma unroll
    for (int i = 0; i < kMixMaxRows; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k4 = 0; k4 < in_chunks; ++k4) {
        float4 trow[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) trow[kk] = *reinterpret_cast<const float4*>(xs + swz64(k4 * 4 + kk, tx));
#pragma unroll
        for (int i = 0; i < kMixMaxRows; ++i) {
            const int row = ty + 16 * i;
            const int dch = swz_s ? (k4 ^ (row & 7)) : k4;
            const float4 sv = *reinterpret_cast<const float4*>(ss + row * p_in + dch * 4);
            acc[i].x = fmaf(sv.x, trow[0].x, acc[i].x); acc[i].y = fmaf(sv.x, trow[0].y, acc[i].y);
            acc[i].z = fmaf(sv.x, trow[0].z, acc[i].z); acc[i].w = fmaf(sv.x, trow[0].w, acc[i].w);
            acc[i].x = fmaf(sv.y, trow[1].x, acc[i].x); acc[i].y = fmaf(sv.y, trow[1].y, acc[i].y);
            acc[i].z = fmaf(sv.y, trow[1].z, acc[i].z); acc[i].w = fmaf(sv.y, trow[1].w, acc[i].w);
            acc[i].x = fmaf(sv.z, trow[2].x, acc[i].x); acc[i].y = fmaf(sv.z, trow[2].y, acc[i].y);
            acc[i].z = fmaf(sv.z, trow[2].z, acc[i].z); acc[i].w = fmaf(sv.z, trow[2].w, acc[i].w);
            acc[i].x = fmaf(sv.w, trow[3].x, acc[i].x); acc[i].y = fmaf(sv.w, trow[3].y, acc[i].y);
            acc[i].z = fmaf(sv.w, trow[3].z, acc[i].z); acc[i].w = fmaf(sv.w, trow[3].w, acc[i].w);
        }
    }
    // ---- layer norm over 128 * 64 values + ReLU -> global -------------------------------------------------------
    {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kMixMaxRows; ++i) s += (acc[i].x + acc[i].y) + (acc[i].z + acc[i].w);
        const float n = (float)(kMixPout * kMixC);
        const float mean = block_sum(s, red) / n;
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < kMixMaxRows; ++i) {
            const float a = acc[i].x - mean, b = acc[i].y - mean, c = acc[i].z - mean, d = acc[i].w - mean;
            v += (a * a + b * b) + (c * c + d * d);
        }
        const float rstd = rsqrtf(block_sum(v, red) / n + eps);
        float* og = out + qg * (long long)(kMixPout * kMixC);
#pragma unroll
        for (int i = 0; i < kMixMaxRows; ++i) {
            float4 o;
            o.x = fmaxf((acc[i].x - mean) * rstd, 0.f);
            o.y = fmaxf((acc[i].y - mean) * rstd, 0.f);
            o.z = fmaxf((acc[i].z - mean) * rstd, 0.f);
            o.w = fmaxf((acc[i].w - mean) * rstd, 0.f);
            const long long off = (ty + 16 * i) * kMixC + tx * 4;
            if constexpr (kSplitOut) {
                const float f[4] = {o.x, o.y, o.z, o.w};
                __align__(8) __nv_bfloat16 p[3][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    p[0][j] = __float2bfloat16_rn(f[j]);
                    const float r1 = f[j] - __bfloat162float(p[0][j]);
                    p[1][j] = __float2bfloat16_rn(r1);
                    p[2][j] = __float2bfloat16_rn(r1 - __bfloat162float(p[1][j]));
                }
                if (tiled_groups > 0) {
                    // the A operand of out_proj in the pre-tiled format: row = query, k = group * 8192 + point * 64 + channel
                    const long long q = qg / tiled_groups;
                    const int k0 = (int)(qg - q * tiled_groups) * (kMixPout * kMixC) + (int)off;
                    const int kblocks = tiled_groups * (kMixPout * kMixC / kTileK);
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        *reinterpret_cast<uint2*>(out3 + tiled_offset(q, k0, kblocks, k)) = *reinterpret_cast<const uint2*>(p[k]);
                } else {
                    __nv_bfloat16* o3 = out3 + qg * (long long)(kMixPout * kMixC) + off;
#pragma unroll
                    for (int k = 0; k < 3; ++k) *reinterpret_cast<uint2*>(o3 + k * piece_stride) = *reinterpret_cast<const uint2*>(p[k]);
                }
            } else {
                *reinterpret_cast<float4*>(og + off) = o;
            }
        }
    }
}

}  // namespace racf

static int mixing_launch(const float* x, const float* params, int num_query_groups, int in_points, int out_points,
                         int channels, float eps, float* out, void* out3, int tiled_groups, racf_stream_t stream) {
    using namespace racf;
    if (!x || !params || (!out && !out3)) return RACF_ERR_NULL_POINTER;
    if (num_query_groups <= 0) return RACF_ERR_BAD_SHAPE;
    if (channels != kMixC || out_points != kMixPout || in_points <= 0 || in_points > 128 || (in_points & 3) != 0)
        return RACF_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(out) |
         reinterpret_cast<uintptr_t>(out3)) & 15u)
        return RACF_ERR_UNSUPPORTED;
    const size_t smem = sizeof(float) * (size_t)(128 * kMixC + kMixC * kMixC + kMixPout * in_points);
    // opt in to > 48 KB of dynamic shared memory (per device, idempotent, not a stream operation)
    cudaError_t e = cudaFuncSetAttribute(adaptive_mixing_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 114688);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(adaptive_mixing_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 114688);
    if (e != cudaSuccess) return (int)e;
    const long long piece_stride = (long long)num_query_groups * kMixPout * kMixC;
    if (out3 != nullptr)
        adaptive_mixing_kernel<true><<<(unsigned)num_query_groups, kMixThreads, smem, static_cast<cudaStream_t>(stream)>>>(
            x, params, nullptr, static_cast<__nv_bfloat16*>(out3), piece_stride, tiled_groups, in_points, eps);
    else
        adaptive_mixing_kernel<false><<<(unsigned)num_query_groups, kMixThreads, smem, static_cast<cudaStream_t>(stream)>>>(
            x, params, out, nullptr, 0, 0, in_points, eps);
    return (int)cudaGetLastError();
}

extern "C" int racf_adaptive_mixing_forward(const float* x, const float* params, int num_query_groups, int in_points,
                                            int out_points, int channels, float eps, float* out, racf_stream_t stream) {
    return mixing_launch(x, params, num_query_groups, in_points, out_points, channels, eps, out, nullptr, 0, stream);
}

extern "C" int racf_adaptive_mixing_forward_split(const float* x, const float* params, int num_query_groups, int in_points,
                                                  int out_points, int channels, float eps, void* out3,
                                                  int tiled_groups, racf_stream_t stream) {
    if (tiled_groups < 0 || (tiled_groups > 0 && num_query_groups % tiled_groups != 0)) return RACF_ERR_BAD_SHAPE;
    return mixing_launch(x, params, num_query_groups, in_points, out_points, channels, eps, nullptr, out3, tiled_groups, stream);
}
