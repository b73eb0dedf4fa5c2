// Scale-adaptive self-attention core (models/racformer_transformer.py:283-336 ScaleAdaptiveSelfAttention around
// mmcv's MultiheadAttention) for inference, as one launch (sm_100a):
//
//     centre_i          = decode_bbox(theta_d2xy_coods(query_ray_i), pc_range)[:2]
//     score[h, i, j]    = (q_i . k_j) / sqrt(d) - tau[i, h] * |centre_i - centre_j|
//     out[i, h*d : ..]  = softmax_j(score[h, i, :]) . v
//
// The reference materialises the [B, 8, Q, Q] distance mask (26 MB at Q = 900) with ~10 elementwise launches and then
// runs a generic fp32 attention kernel over it; here the bias is recomputed from two floats per key.
//
// fp32 FMA on the CUDA cores (the products are 0.8 GFLOP in total -- far too small for a tensor-core pipeline to pay
// off, and fp32-exact). Lane = query, warp = a contiguous slice of the keys: every lane keeps its q row and its
// accumulator in registers; each warp streams its keys through a private shared-memory double buffer (16-byte cp.async
// copies, one block of eight keys ahead of the arithmetic -- a first version with direct warp-uniform global loads was
// latency-bound at 160 us) and reads k_j / v_j back as warp-uniform 128-bit loads; scores are formed eight keys at a
// time with one running-max update per block, and the kSplits partial (max, sum, acc) triples of a query are merged
// through shared memory at the end.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "racformer_ops.h"

namespace racf {

constexpr int kHeadDim = 32;
constexpr int kSplits = 8;              // warps per CTA = key slices
constexpr int kKeyBlock = 8;
constexpr float kMapSize = 102.4f, kCentre = 51.2f, kRayR = 65.0f, kTwoPi = 6.283185307179586f;

struct SasaArgs {
    const float* qkv;        // [B*Q, 3*E]  (q | k | v), E = H * 32
    const float* tau;        // [B*Q, H]
    const float* query_ray;  // [B*Q, code]  (theta, d, ...)
    float* out;              // [B*Q, E]
    int Q, H, E, code;
    float scale;
    float x0, xs, y0, ys;    // pc_range[0], pc_range[3] - pc_range[0], pc_range[1], pc_range[4] - pc_range[1]
};

__global__ void __launch_bounds__(kSplits * 32, 2)
sasa_attention_kernel(const SasaArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* centres = smem;                                   // [Q][2]
    float* part = smem + 2 * ((a.Q + 1) & ~1);               // [kSplits][kHeadDim + 2][32]
    float* kvbuf = part + kSplits * (kHeadDim + 2) * 32;     // [kSplits][2][kKeyBlock][2 * kHeadDim]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 32;
    const size_t row_base = (size_t)b * a.Q;

    // metric (x, y) centre of every query of this batch element (separate multiply / add like the PyTorch ops)
    for (int j = threadIdx.x; j < a.Q; j += blockDim.x) {
        const float* ray = a.query_ray + (row_base + j) * a.code;
        const float ang = __fmul_rn(ray[0], kTwoPi), rad = __fmul_rn(ray[1], kRayR);
        float x = __fdiv_rn(__fadd_rn(kCentre, __fmul_rn(rad, cosf(ang))), kMapSize);
        float y = __fdiv_rn(__fadd_rn(kCentre, __fmul_rn(rad, sinf(ang))), kMapSize);
        x = fminf(fmaxf(x, 0.f), 1.f);
        y = fminf(fmaxf(y, 0.f), 1.f);
        centres[2 * j] = __fadd_rn(__fmul_rn(x, a.xs), a.x0);
        centres[2 * j + 1] = __fadd_rn(__fmul_rn(y, a.ys), a.y0);
    }
    __syncthreads();

    const int qi = min(q0 + lane, a.Q - 1);                  // lanes past the end repeat the last query (not stored)
    const int E3 = 3 * a.E;
    // q and the accumulator as float2 pairs: the products run as packed FMAs (fma.rn.f32x2, two per issue slot; a
    // 3-register FFMA issues only every second cycle per scheduler on sm_100)
    float2 q2[kHeadDim / 2], acc2[kHeadDim / 2];
    {
        const float4* qp = reinterpret_cast<const float4*>(a.qkv + (row_base + qi) * E3 + h * kHeadDim);
#pragma unroll
        for (int i = 0; i < kHeadDim / 4; ++i) {
            const float4 v = __ldg(qp + i);
            q2[2 * i] = make_float2(v.x * a.scale, v.y * a.scale);
            q2[2 * i + 1] = make_float2(v.z * a.scale, v.w * a.scale);
        }
    }
#pragma unroll
    for (int d = 0; d < kHeadDim / 2; ++d) acc2[d] = make_float2(0.f, 0.f);
    const float cx = centres[2 * qi], cy = centres[2 * qi + 1];
    const float neg_tau = -__ldg(a.tau + (row_base + qi) * a.H + h);
    float m = -INFINITY, l = 0.f;

    const int per = (a.Q + kSplits - 1) / kSplits;
    const int j_begin = warp * per, j_end = min(a.Q, j_begin + per);
    const float* kbase = a.qkv + row_base * E3 + a.E + h * kHeadDim;     // v of the same row is a.E floats further
    // Per-warp double buffer of key blocks: [2][kKeyBlock][k 32 | v 32] floats, filled with 16-byte cp.async copies
    // (4 per lane and block) one block ahead of the arithmetic, read back as warp-uniform 128-bit loads.
    float* kv = kvbuf + (size_t)warp * 2 * kKeyBlock * 2 * kHeadDim;
    const uint32_t kv_s = (uint32_t)__cvta_generic_to_shared(kv);
    auto prefetch = [&](int j0, int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = lane + 32 * i, key = c >> 4, within = c & 15;
            const int j = min(j0 + key, j_end - 1);
            const float* src = kbase + (size_t)j * E3 + (within < 8 ? within * 4 : a.E + (within - 8) * 4);
            const uint32_t dst = kv_s + (uint32_t)(((buf * kKeyBlock + key) * 2 * kHeadDim + within * 4) * 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (j_begin < j_end) prefetch(j_begin, 0);
    int buf = 0;
    for (int j0 = j_begin; j0 < j_end; j0 += kKeyBlock, buf ^= 1) {
        if (j0 + kKeyBlock < j_end) {
            prefetch(j0 + kKeyBlock, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        const float* blk = kv + (size_t)buf * kKeyBlock * 2 * kHeadDim;
        float s[kKeyBlock];
#pragma unroll
        for (int jj = 0; jj < kKeyBlock; ++jj) {
            const int j = min(j0 + jj, j_end - 1);
            const float4* kp = reinterpret_cast<const float4*>(blk + jj * 2 * kHeadDim);
            float2 d0 = make_float2(0.f, 0.f), d1 = make_float2(0.f, 0.f);   // (even, odd) dims of the two halves
#pragma unroll
            for (int i = 0; i < kHeadDim / 4; i += 2) {
                const float4 k0 = kp[i], k1 = kp[i + 1];
                d0 = __ffma2_rn(q2[2 * i], make_float2(k0.x, k0.y), d0);
                d0 = __ffma2_rn(q2[2 * i + 1], make_float2(k0.z, k0.w), d0);
                d1 = __ffma2_rn(q2[2 * i + 2], make_float2(k1.x, k1.y), d1);
                d1 = __ffma2_rn(q2[2 * i + 3], make_float2(k1.z, k1.w), d1);
            }
            const float2 c = *reinterpret_cast<const float2*>(centres + 2 * j);
            const float dx = cx - c.x, dy = cy - c.y;
            const float dist = sqrtf(fmaf(dy, dy, __fmul_rn(dx, dx)));      // torch.norm's reduction: acc + x * x
            s[jj] = (j0 + jj < j_end) ? __fadd_rn((d0.x + d0.y) + (d1.x + d1.y), __fmul_rn(dist, neg_tau)) : -INFINITY;
        }
        float mb = s[0];
#pragma unroll
        for (int jj = 1; jj < kKeyBlock; ++jj) mb = fmaxf(mb, s[jj]);
        const float m_new = fmaxf(m, mb);
        const float corr = expf(m - m_new);                  // 0 on the first block (m = -inf)
        m = m_new;
        l *= corr;
#pragma unroll
        for (int d = 0; d < kHeadDim / 2; ++d) { acc2[d].x *= corr; acc2[d].y *= corr; }
#pragma unroll
        for (int jj = 0; jj < kKeyBlock; ++jj) {
            const float p = expf(s[jj] - m);                 // exp(-inf) = 0 for the padded keys
            l += p;
            const float4* vp = reinterpret_cast<const float4*>(blk + jj * 2 * kHeadDim + kHeadDim);
#pragma unroll
            for (int i = 0; i < kHeadDim / 4; ++i) {
                const float4 v = vp[i];
                const float2 pp = make_float2(p, p);
                acc2[2 * i] = __ffma2_rn(pp, make_float2(v.x, v.y), acc2[2 * i]);
                acc2[2 * i + 1] = __ffma2_rn(pp, make_float2(v.z, v.w), acc2[2 * i + 1]);
            }
        }
        __syncwarp();                                        // the block after next overwrites this buffer
    }

    // merge the kSplits partial results of each query: part[split][0] = m, [1] = l, [2 + d] = acc[d]; lane fastest
    float* mine = part + (size_t)warp * (kHeadDim + 2) * 32 + lane;
    mine[0] = m;
    mine[32] = l;
#pragma unroll
    for (int d = 0; d < kHeadDim / 2; ++d) {
        mine[(2 + 2 * d) * 32] = acc2[d].x;
        mine[(3 + 2 * d) * 32] = acc2[d].y;
    }
    __syncthreads();
    float mx = -INFINITY;
#pragma unroll
    for (int s2 = 0; s2 < kSplits; ++s2) mx = fmaxf(mx, part[(size_t)s2 * (kHeadDim + 2) * 32 + lane]);
    float w[kSplits], den = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < kSplits; ++s2) {
        const float* p2 = part + (size_t)s2 * (kHeadDim + 2) * 32 + lane;
        w[s2] = expf(p2[0] - mx);                            // an empty slice has m = -inf, l = 0
        den = fmaf(p2[32], w[s2], den);
    }
    constexpr int kPerWarp = kHeadDim / kSplits;             // 4 output channels per warp
    float o[kPerWarp];
#pragma unroll
    for (int e = 0; e < kPerWarp; ++e) {
        const int d = warp * kPerWarp + e;
        float v = 0.f;
#pragma unroll
        for (int s2 = 0; s2 < kSplits; ++s2) v = fmaf(part[((size_t)s2 * (kHeadDim + 2) + 2 + d) * 32 + lane], w[s2], v);
        o[e] = v / den;
    }
    if (q0 + lane < a.Q) {
        float4* op = reinterpret_cast<float4*>(a.out + (row_base + q0 + lane) * a.E + h * kHeadDim + warp * kPerWarp);
        *op = make_float4(o[0], o[1], o[2], o[3]);
    }
}

}  // namespace racf

extern "C" int racf_sasa_attention_forward(const float* qkv, const float* tau, const float* query_ray, const double* pc_range,
                                           int batch, int num_query, int num_heads, int head_dim, int code_size,
                                           float* out, racf_stream_t stream) {
    using namespace racf;
    if (!qkv || !tau || !query_ray || !pc_range || !out) return RACF_ERR_NULL_POINTER;
    if (batch <= 0 || num_query <= 0 || num_heads <= 0 || code_size < 2) return RACF_ERR_BAD_SHAPE;
    if (head_dim != kHeadDim || batch > 65535 || num_heads > 65535) return RACF_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15u) return RACF_ERR_UNSUPPORTED;
    SasaArgs a;
    a.qkv = qkv; a.tau = tau; a.query_ray = query_ray; a.out = out;
    a.Q = num_query; a.H = num_heads; a.E = num_heads * head_dim; a.code = code_size;
    a.scale = (float)(1.0 / sqrt((double)head_dim));
    a.x0 = (float)pc_range[0]; a.xs = (float)(pc_range[3] - pc_range[0]);
    a.y0 = (float)pc_range[1]; a.ys = (float)(pc_range[4] - pc_range[1]);
    const size_t smem = (size_t)(2 * ((num_query + 1) & ~1) + kSplits * (kHeadDim + 2) * 32 +
                                 kSplits * 2 * kKeyBlock * 2 * kHeadDim) * sizeof(float);
    if (smem > 200u * 1024u) return RACF_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(sasa_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const dim3 grid((unsigned)((num_query + 31) / 32), (unsigned)num_heads, (unsigned)batch);
    sasa_attention_kernel<<<grid, kSplits * 32, smem, static_cast<cudaStream_t>(stream)>>>(a);
    return (int)cudaGetLastError();
}
