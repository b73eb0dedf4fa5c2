// Backward of the fused AdaptiveMixing core (training) -- SURVEY.md 8f-4; forward: csrc/mixing_tc.cu / csrc/mixing.cu.
//
// Reference: AdaptiveMixing.inner_forward, models/racformer_transformer.py:592-604, differentiated by autograd there
// (2 batched SGEMMs forward, 4 backward, 2 layer norms, 2 ReLUs -- with their [QG,P,64] intermediates kept in HBM).
// Per (query, group) item, with x [P_in,64], M [64,64], S [128,P_in] (M, S = slices of parameter_generator(query)):
//     o1 = x M ;  n1 = LN(o1) ;  t = relu(n1) ;  o2 = S t ;  n2 = LN(o2) ;  y = relu(n2)          (LN: no affine, whole item)
// Given gy = dL/dy the kernel RECOMPUTES o1, n1, o2, n2 from x and the parameters (nothing but the op's inputs is saved by
// the forward; 7 MFLOP per item against 130 KB of traffic) and produces, in one launch:
//     g_n2 = gy * [n2 > 0]        g_o2 = rstd2 (g_n2 - mean(g_n2) - n2 mean(g_n2 n2))
//     g_S  = g_o2 t^T             g_t  = S^T g_o2
//     g_n1 = g_t * [n1 > 0]       g_o1 = rstd1 (g_n1 - mean(g_n1) - n1 mean(g_n1 n1))
//     g_M  = x^T g_o1             g_x  = g_o1 M^T
// One CTA per item, everything in shared memory (x, M, M^T, S, n1, g_o2 / g_o1: 160 KB at P_in = 96), all six products on the
// CUDA cores in fp32 FMA with register tiles (the arithmetic of the reference's SGEMMs, other summation order).
// Layouts: x [QG,P_in,64], params [QG, 64*64 + 128*P_in] (M row-major, then S row-major), gy [QG,128,64];
// grad_x like x, grad_params like params; both fully overwritten. C == 64, P_out == 128, P_in % 16 == 0, P_in <= 128.
#include <cuda_runtime.h>
#include <stdint.h>

#include "racformer_ops.h"

namespace racf {
namespace mixbwd {

constexpr int kC = 64, kPout = 128;
constexpr int kCols = 8;          // g_S columns per thread: p = tx + 16 j, j < p_in / 16 <= 8
constexpr int kMtStride = 68;     // M^T rows padded: 16-byte aligned, consecutive rows 4 banks apart

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;"); }

template <int kThreads>
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];
    return t;
}

// [rows][64] tile, 16-byte chunks XOR-swizzled by the row (see csrc/mixing.cu)
__device__ __forceinline__ int swz64(int row, int chunk) { return row * kC + ((chunk ^ (row & 7)) << 2); }
__device__ __forceinline__ float ld_swz(const float* t, int row, int col) { return t[swz64(row, col >> 2) + (col & 3)]; }

// acc += a * b as two packed FMAs (fma.rn.f32x2): a 3-register FFMA issues only every second cycle per scheduler on sm_100,
// the packed form does two of them in one issue slot (same rounding, one FMA per element)
__device__ __forceinline__ void fma4(float4& acc, float a, const float4& b) {
    const float2 aa = make_float2(a, a);
    const float2 lo = __ffma2_rn(aa, make_float2(b.x, b.y), make_float2(acc.x, acc.y));
    const float2 hi = __ffma2_rn(aa, make_float2(b.z, b.w), make_float2(acc.z, acc.w));
    acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float comp(const float4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }
__device__ __forceinline__ float4 relu4(const float4& v) {
    return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
}
__device__ __forceinline__ float sum4(const float4& v) { return (v.x + v.y) + (v.z + v.w); }

// TY = row groups: thread (ty, tx) owns rows ty + TY * i of every tile and the 4 columns 4 tx ..; the launcher uses TY = 16
// (256 threads, 8 warps, one CTA per SM). ncu (profiles/r02_mixing_bwd_ncu.json): issue slots 47 % busy, FMA pipe 37 %, top
// stalls "wait" and short scoreboard -- latency-bound at 8 warps per SM.
template <int TY>
__global__ void __launch_bounds__(TY * 16, 1)
adaptive_mixing_bwd_kernel(const float* __restrict__ x, const float* __restrict__ params, const float* __restrict__ gy,
                           float* __restrict__ grad_x, float* __restrict__ grad_params, int p_in, float eps) {
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                          // [p_in][64] swizzled
    float* ms = xs + p_in * kC;                // [64][64]
    float* mt = ms + kC * kC;                  // [64][68]   M^T
    float* ss = mt + kC * kMtStride;           // [128][p_in]
    float* n1s = ss + kPout * p_in;            // [p_in][64] swizzled: LN1 output before the ReLU
    float* gs = n1s + p_in * kC;               // [128][64] swizzled: g_o2, later g_o1
    constexpr int kThreads = TY * 16, kRows = kPout / TY;
    __shared__ float red[kThreads / 32];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long qg = blockIdx.x;
    const long long pstride = kC * kC + (long long)kPout * p_in;
    const float* xg = x + qg * (long long)p_in * kC;
    const float* pg = params + qg * pstride;
    const int in_chunks = p_in >> 2;
    const int rows_a = p_in / TY;               // rows of a [p_in][64] tile owned by this thread (p_in % TY == 0)

    for (int i = tid; i < p_in * 16; i += kThreads) cp_async16(xs + swz64(i >> 4, i & 15), xg + i * 4);
    for (int i = tid; i < kC * 16; i += kThreads) cp_async16(ms + i * 4, pg + i * 4);
    for (int i = tid; i < kPout * in_chunks; i += kThreads) cp_async16(ss + i * 4, pg + kC * kC + i * 4);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    for (int i = tid; i < kC * kC; i += kThreads) mt[(i & 63) * kMtStride + (i >> 6)] = ms[i];   // mt[d][c] = M[c][d]

    // ---- recompute: o1 = x M, LN1 -> n1s -----------------------------------------------------------------------------
    float4 acc[kRows];
#pragma unroll
    for (int i = 0; i < kRows; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k4 = 0; k4 < 16; ++k4) {
        float4 mrow[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) mrow[kk] = *reinterpret_cast<const float4*>(ms + (k4 * 4 + kk) * kC + tx * 4);
        // k outer, rows inner: up to 16 independent packed FMAs between two updates of one accumulator pair (8 warps per SM
        // cannot hide the 4-cycle dependent-issue latency otherwise: ncu "wait" was the top stall)
        float4 xv[kRows];
#pragma unroll
        for (int i = 0; i < kRows; ++i)
            if (i < rows_a) xv[i] = *reinterpret_cast<const float4*>(xs + swz64(ty + TY * i, k4));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int i = 0; i < kRows; ++i)
                if (i < rows_a) fma4(acc[i], comp(xv[i], kk), mrow[kk]);
    }
    float rstd1;
    {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kRows; ++i)
            if (i < rows_a) s += sum4(acc[i]);
        const float n = (float)(p_in * kC);
        const float mean = block_sum<kThreads>(s, red) / n;
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < kRows; ++i)
            if (i < rows_a) {
                const float a = acc[i].x - mean, b = acc[i].y - mean, c = acc[i].z - mean, d = acc[i].w - mean;
                v += (a * a + b * b) + (c * c + d * d);
            }
        rstd1 = rsqrtf(block_sum<kThreads>(v, red) / n + eps);
#pragma unroll
        for (int i = 0; i < kRows; ++i)
            if (i < rows_a)
                *reinterpret_cast<float4*>(n1s + swz64(ty + TY * i, tx)) = make_float4(
                    (acc[i].x - mean) * rstd1, (acc[i].y - mean) * rstd1, (acc[i].z - mean) * rstd1, (acc[i].w - mean) * rstd1);
    }
    __syncthreads();

    // ---- recompute: o2 = S relu(n1), LN2 -> n2 (registers) -------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < kRows; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k4 = 0; k4 < in_chunks; ++k4) {
        float4 trow[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) trow[kk] = relu4(*reinterpret_cast<const float4*>(n1s + swz64(k4 * 4 + kk, tx)));
        float4 sv[kRows];
#pragma unroll
        for (int i = 0; i < kRows; ++i) sv[i] = *reinterpret_cast<const float4*>(ss + (ty + TY * i) * p_in + k4 * 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int i = 0; i < kRows; ++i) fma4(acc[i], comp(sv[i], kk), trow[kk]);
    }
    {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kRows; ++i) s += sum4(acc[i]);
        const float n = (float)(kPout * kC);
        const float mean = block_sum<kThreads>(s, red) / n;
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < kRows; ++i) {
            const float a = acc[i].x - mean, b = acc[i].y - mean, c = acc[i].z - mean, d = acc[i].w - mean;
            v += (a * a + b * b) + (c * c + d * d);
        }
        const float rstd2 = rsqrtf(block_sum<kThreads>(v, red) / n + eps);
        // ---- LN2 + ReLU backward on the thread's own elements -> g_o2 in shared memory ---------------------------------
        const float* gyg = gy + qg * (long long)(kPout * kC);
        float4 g[kRows];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < kRows; ++i) {
            acc[i] = make_float4((acc[i].x - mean) * rstd2, (acc[i].y - mean) * rstd2, (acc[i].z - mean) * rstd2,
                                 (acc[i].w - mean) * rstd2);                                          // n2
            const float4 gv = __ldg(reinterpret_cast<const float4*>(gyg + (ty + TY * i) * kC + tx * 4));
            g[i] = make_float4(acc[i].x > 0.f ? gv.x : 0.f, acc[i].y > 0.f ? gv.y : 0.f, acc[i].z > 0.f ? gv.z : 0.f,
                               acc[i].w > 0.f ? gv.w : 0.f);
            s1 += sum4(g[i]);
            s2 += (g[i].x * acc[i].x + g[i].y * acc[i].y) + (g[i].z * acc[i].z + g[i].w * acc[i].w);
        }
        const float c1 = block_sum<kThreads>(s1, red) / n, c2 = block_sum<kThreads>(s2, red) / n;
#pragma unroll
        for (int i = 0; i < kRows; ++i)
            *reinterpret_cast<float4*>(gs + swz64(ty + TY * i, tx)) = make_float4(
                rstd2 * (g[i].x - c1 - acc[i].x * c2), rstd2 * (g[i].y - c1 - acc[i].y * c2),
                rstd2 * (g[i].z - c1 - acc[i].z * c2), rstd2 * (g[i].w - c1 - acc[i].w * c2));
    }
    __syncthreads();

    float* gpg = grad_params + qg * pstride;
    // ---- g_S[r][p] = sum_c g_o2[r][c] t[p][c]: rows r = ty + 16 i, columns p = tx + 16 j ---------------------------------
    {
        const int cols = p_in >> 4;            // <= 8
        float2 a2[kRows][kCols / 2];           // columns (2 jj, 2 jj + 1) of row i, accumulated with packed FMAs
#pragma unroll
        for (int i = 0; i < kRows; ++i)
#pragma unroll
            for (int j = 0; j < kCols / 2; ++j) a2[i][j] = make_float2(0.f, 0.f);
        for (int c4 = 0; c4 < 16; ++c4) {
            float4 gr[kRows];
#pragma unroll
            for (int i = 0; i < kRows; ++i) gr[i] = *reinterpret_cast<const float4*>(gs + swz64(ty + TY * i, c4));
#pragma unroll
            for (int j = 0; j < kCols / 2; ++j)
                if (2 * j < cols) {            // cols is even (p_in % 32 == 0) or the odd last column is computed and dropped
                    const float4 t0 = relu4(*reinterpret_cast<const float4*>(n1s + swz64(tx + 16 * (2 * j), c4)));
                    const int r1 = min(tx + 16 * (2 * j + 1), p_in - 1);
                    const float4 t1 = relu4(*reinterpret_cast<const float4*>(n1s + swz64(r1, c4)));
#pragma unroll
                    for (int i = 0; i < kRows; ++i) {
                        float2 v = a2[i][j];
                        v = __ffma2_rn(make_float2(gr[i].x, gr[i].x), make_float2(t0.x, t1.x), v);
                        v = __ffma2_rn(make_float2(gr[i].y, gr[i].y), make_float2(t0.y, t1.y), v);
                        v = __ffma2_rn(make_float2(gr[i].z, gr[i].z), make_float2(t0.z, t1.z), v);
                        v = __ffma2_rn(make_float2(gr[i].w, gr[i].w), make_float2(t0.w, t1.w), v);
                        a2[i][j] = v;
                    }
                }
        }
        float* gS = gpg + kC * kC;
#pragma unroll
        for (int i = 0; i < kRows; ++i)
#pragma unroll
            for (int j = 0; j < kCols; ++j)
                if (j < cols) gS[(ty + TY * i) * p_in + tx + 16 * j] = (j & 1) ? a2[i][j >> 1].y : a2[i][j >> 1].x;
    }
    // ---- g_t[p][c] = sum_r S[r][p] g_o2[r][c]: rows p = ty + 16 i, columns 4 tx .. ---------------------------------------
#pragma unroll
    for (int i = 0; i < kRows; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int r = 0; r < kPout; ++r) {
        const float4 gr = *reinterpret_cast<const float4*>(gs + swz64(r, tx));
        const float* srow = ss + r * p_in + ty;
#pragma unroll
        for (int i = 0; i < kRows; ++i)
            if (i < rows_a) fma4(acc[i], srow[TY * i], gr);
    }
    // ---- ReLU + LN1 backward -> g_o1 (overwrites g_o2 once every thread is done with it) ------------------------------
    {
        const float n = (float)(p_in * kC);
        float4 n1[kRows];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < kRows; ++i)
            if (i < rows_a) {
                n1[i] = *reinterpret_cast<const float4*>(n1s + swz64(ty + TY * i, tx));
                acc[i] = make_float4(n1[i].x > 0.f ? acc[i].x : 0.f, n1[i].y > 0.f ? acc[i].y : 0.f,
                                     n1[i].z > 0.f ? acc[i].z : 0.f, n1[i].w > 0.f ? acc[i].w : 0.f);
                s1 += sum4(acc[i]);
                s2 += (acc[i].x * n1[i].x + acc[i].y * n1[i].y) + (acc[i].z * n1[i].z + acc[i].w * n1[i].w);
            }
        const float c1 = block_sum<kThreads>(s1, red) / n, c2 = block_sum<kThreads>(s2, red) / n;   // (its barriers also retire all reads of g_o2)
#pragma unroll
        for (int i = 0; i < kRows; ++i)
            if (i < rows_a)
                *reinterpret_cast<float4*>(gs + swz64(ty + TY * i, tx)) = make_float4(
                    rstd1 * (acc[i].x - c1 - n1[i].x * c2), rstd1 * (acc[i].y - c1 - n1[i].y * c2),
                    rstd1 * (acc[i].z - c1 - n1[i].z * c2), rstd1 * (acc[i].w - c1 - n1[i].w * c2));
    }
    __syncthreads();
    // ---- g_M[c][d] = sum_p x[p][c] g_o1[p][d]: rows c = ty + 16 i (i < 4), columns 4 tx .. ----------------------------
    {
        constexpr int kMRows = kC / TY;
        float4 am[kMRows];
#pragma unroll
        for (int i = 0; i < kMRows; ++i) am[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int p = 0; p < p_in; ++p) {
            const float4 gr = *reinterpret_cast<const float4*>(gs + swz64(p, tx));
#pragma unroll
            for (int i = 0; i < kMRows; ++i) fma4(am[i], ld_swz(xs, p, ty + TY * i), gr);
        }
#pragma unroll
        for (int i = 0; i < kMRows; ++i) *reinterpret_cast<float4*>(gpg + (ty + TY * i) * kC + tx * 4) = am[i];
    }
    // ---- g_x[p][c] = sum_d g_o1[p][d] M[c][d] = sum_d g_o1[p][d] mt[d][c]: rows p = ty + 16 i, columns 4 tx .. ------------
#pragma unroll
    for (int i = 0; i < kRows; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int d4 = 0; d4 < 16; ++d4) {
        float4 mrow[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) mrow[kk] = *reinterpret_cast<const float4*>(mt + (d4 * 4 + kk) * kMtStride + tx * 4);
        float4 gv[kRows];
#pragma unroll
        for (int i = 0; i < kRows; ++i)
            if (i < rows_a) gv[i] = *reinterpret_cast<const float4*>(gs + swz64(ty + TY * i, d4));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int i = 0; i < kRows; ++i)
                if (i < rows_a) fma4(acc[i], comp(gv[i], kk), mrow[kk]);
    }
    float* gxg = grad_x + qg * (long long)p_in * kC;
#pragma unroll
    for (int i = 0; i < kRows; ++i)
        if (i < rows_a) *reinterpret_cast<float4*>(gxg + (ty + TY * i) * kC + tx * 4) = acc[i];
}

}  // namespace mixbwd
}  // namespace racf

int racf_mixbwd_tc_launch(const float* x, const float* params, const float* grad_out, int num_query_groups, int in_points, float eps,
                          float* grad_x, float* grad_params, int sms, cudaStream_t st);   // csrc/mixing_bwd_tc.cu

// Which kernel variant 0 selects where both exist (in_points <= 96): tools/mixing_bwd_check.py measures both.
constexpr bool kDefaultTensorCores = true;     // 1.40 ms vs 2.67 ms for 9760 items at P_in = 96 (profiles/r02f_mixing_bwd_check.json)

// variant: 0 = default, 1 = CUDA-core kernel of this file, 2 = tcgen05 kernel (csrc/mixing_bwd_tc.cu; RACF_ERR_UNSUPPORTED for
// in_points > 96).
extern "C" int racf_adaptive_mixing_backward_variant(const float* x, const float* params, const float* grad_out,
                                                     int num_query_groups, int in_points, int out_points, int channels, float eps,
                                                     float* grad_x, float* grad_params, int variant, racf_stream_t stream) {
    using namespace racf::mixbwd;
    if (!x || !params || !grad_out || !grad_x || !grad_params) return RACF_ERR_NULL_POINTER;
    if (num_query_groups <= 0) return RACF_ERR_BAD_SHAPE;
    if (channels != kC || out_points != kPout || in_points < 16 || in_points > 128 || (in_points & 15) != 0)
        return RACF_ERR_UNSUPPORTED;
    if (variant < 0 || variant > 2 || (variant == 2 && in_points > 96)) return RACF_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grad_out) |
         reinterpret_cast<uintptr_t>(grad_x) | reinterpret_cast<uintptr_t>(grad_params)) & 15u)
        return RACF_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (variant == 2 || (variant == 0 && kDefaultTensorCores && in_points <= 96)) {
        int dev = 0, sms = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return (int)e;
        return racf_mixbwd_tc_launch(x, params, grad_out, num_query_groups, in_points, eps, grad_x, grad_params, sms, st);
    }
    const size_t smem = sizeof(float) * (size_t)(2 * in_points * kC + kC * kC + kC * kMtStride + kPout * in_points + kPout * kC);
    // TY = 16 (256 threads): measured 2.65 ms for 9760 items at P_in = 96 against 2.83 ms with TY = 32 (512 threads halve the
    // rows a thread amortises each shared-memory operand load over; the shared-memory pipe then limits)
    cudaError_t e = cudaFuncSetAttribute(adaptive_mixing_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    adaptive_mixing_bwd_kernel<16><<<(unsigned)num_query_groups, 256, smem, st>>>(x, params, grad_out, grad_x, grad_params, in_points, eps);
    return (int)cudaGetLastError();
}

extern "C" int racf_adaptive_mixing_backward(const float* x, const float* params, const float* grad_out, int num_query_groups,
                                             int in_points, int out_points, int channels, float eps, float* grad_x,
                                             float* grad_params, racf_stream_t stream) {
    return racf_adaptive_mixing_backward_variant(x, params, grad_out, num_query_groups, in_points, out_points, channels, eps, grad_x,
                                                 grad_params, 0, stream);
}
