// Scale-adaptive self-attention core for TRAINING (forward with denoising mask + attention dropout, and backward).
//
// Reference: ScaleAdaptiveSelfAttention.inner_forward around mmcv's MultiheadAttention(embed_dims, 8, attn_drop=0.1)
// (models/racformer_transformer.py:296-336), with the query-denoising attention mask of models/racformer_head.py:205-206,
// differentiated by autograd there through a materialised [B, 8, Q, Q] additive mask (masked_fill, a generic fp32
// attention kernel and its backward).  Per (batch b, head h), with d = 32, scale = 1 / sqrt(d):
//
//     s_ij   = scale <q_i, k_j> - tau[i,h] |c_i - c_j|            (c = metric box centres, no gradient; -inf where blocked[i][j])
//     P_ij   = softmax_j(s_ij) ;  A_ij = keep_ij / (1 - p) * P_ij   (keep: counter-based hash of (seed, b, h, i, j))
//     o_i    = sum_j A_ij v_j
//   backward (lse_i = log sum_j exp s_ij saved by the forward, D_i = <do_i, o_i>):
//     dA_ij  = <do_i, v_j> ;  dS_ij = P_ij (keep_ij / (1 - p) dA_ij - D_i)
//     dq_i   = scale sum_j dS_ij k_j ;  dk_j = scale sum_i dS_ij q_i ;  dv_j = sum_i A_ij do_i ;  dtau[i,h] = -sum_j dS_ij |c_i - c_j|
//
// Three kernels, all fp32 FMA on the CUDA cores (5 GFLOP per call at Q = 1220, B = 2: nothing a tensor-core pipeline
// would pay off for), none of which materialises a [Q, Q] tensor:
//   forward        lane = query, the warps of a CTA split the keys and merge (max, sum, acc) through shared memory;
//   backward dq    lane = query, warps split the keys, partial dq / dtau summed through shared memory; also writes D;
//   backward dk/dv lane = key, warps split the queries, partial dk / dv summed through shared memory.
// The blocked-pair mask is passed TRANSPOSED (blocked_t[j][i], uint8) so that a warp's 32 queries read 32 adjacent bytes.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "racformer_ops.h"

namespace racf {
namespace sasat {

constexpr int kD = 32;                  // head dim
constexpr int kWarps = 8;
constexpr float kMapSize = 102.4f, kCentre = 51.2f, kRayR = 65.0f, kTwoPi = 6.283185307179586f;

struct Args {
    const float* qkv;        // [B*Q, 3E] (q | k | v)
    const float* tau;        // [B*Q, H]
    const float* query_ray;  // [B*Q, code]
    const uint8_t* blocked_t;  // [Q][Q] transposed: blocked_t[j * Q + i] != 0 <=> query i may not attend to key j; or NULL
    float* out;              // [B*Q, E]
    float* lse;              // [B, H, Q]
    const float* grad_out;   // [B*Q, E]
    float* dsum;             // [B, H, Q]   D_i
    float* grad_qkv;         // [B*Q, 3E]
    float* grad_tau;         // [B*Q, H]
    int Q, H, E, code;
    float scale, drop_p, keep_scale;
    unsigned seed;
    float x0, xs, y0, ys;
};

__device__ __forceinline__ float keep_factor(const Args& a, int b, int h, int i, int j) {
    if (a.drop_p <= 0.f) return 1.f;
    unsigned x = a.seed ^ ((unsigned)(b * a.H + h) * 0x9E3779B1u) ^ ((unsigned)i * 0x85EBCA77u) ^ ((unsigned)j * 0xC2B2AE3Du);
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return ((float)(x >> 8) * (1.0f / 16777216.0f) >= a.drop_p) ? a.keep_scale : 0.f;
}

__device__ __forceinline__ void compute_centres(const Args& a, size_t row_base, float* centres) {
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
}

__device__ __forceinline__ float dist_of(float cx, float cy, const float* centres, int j) {
    const float dx = cx - centres[2 * j], dy = cy - centres[2 * j + 1];
    return sqrtf(fmaf(dy, dy, __fmul_rn(dx, dx)));
}

// Products run as packed FMAs (fma.rn.f32x2, SASS FFMA2: two per issue slot; a 3-register FFMA issues only every second
// cycle per scheduler on sm_100). Rows live in registers as float2 pairs; dot products use two pair accumulators.
__device__ __forceinline__ void load_row32_f2(const float* p, float2* r) {
#pragma unroll
    for (int i = 0; i < kD / 4; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
        r[2 * i] = make_float2(v.x, v.y);
        r[2 * i + 1] = make_float2(v.z, v.w);
    }
}
// sum_d a[d] * row[d] for a 32-float row in shared memory (warp-uniform 128-bit loads)
__device__ __forceinline__ float dot32(const float2* a2, const float4* row) {
    float2 d0 = make_float2(0.f, 0.f), d1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kD / 4; ++i) {
        const float4 v = row[i];
        d0 = __ffma2_rn(a2[2 * i], make_float2(v.x, v.y), d0);
        d1 = __ffma2_rn(a2[2 * i + 1], make_float2(v.z, v.w), d1);
    }
    return (d0.x + d0.y) + (d1.x + d1.y);
}
// acc[d] += w * row[d]
__device__ __forceinline__ void axpy32(float2* acc2, float w, const float4* row) {
    const float2 ww = make_float2(w, w);
#pragma unroll
    for (int i = 0; i < kD / 4; ++i) {
        const float4 v = row[i];
        acc2[2 * i] = __ffma2_rn(ww, make_float2(v.x, v.y), acc2[2 * i]);
        acc2[2 * i + 1] = __ffma2_rn(ww, make_float2(v.z, v.w), acc2[2 * i + 1]);
    }
}

__device__ __forceinline__ void load_row32(const float* p, float* r) {
#pragma unroll
    for (int i = 0; i < kD / 4; ++i) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
        r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
    }
}

// Rows of the "other side" (keys for the query-owning kernels, queries for the key-owning kernel) are staged per warp in
// shared memory eight at a time and read back as warp-uniform 128-bit loads. The staging is a per-warp double buffer filled
// with cp.async one block ahead of the arithmetic (as in csrc/sasa.cu, where it took the inference core from 161 to 50 us):
// with synchronous staging every block paid a global-load latency with nothing to overlap it at 8-16 warps per SM.
constexpr int kBlock = 8;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

// keys j0 .. j0 + 7 of (batch row_base, head h): k (32 floats) | v (32 floats) per key -> st[buf][key][64]
__device__ __forceinline__ void prefetch_kv(const Args& a, size_t row_base, int h, int j0, int j_end, uint32_t st_s, int buf, int lane) {
    const int E3 = 3 * a.E;
#pragma unroll
    for (int i = 0; i < kBlock * 16 / 32; ++i) {             // 16 float4 per key: 8 of k, 8 of v
        const int c = lane + 32 * i, key = c >> 4, within = c & 15;
        const int j = min(j0 + key, j_end - 1);
        const float* src = a.qkv + (row_base + j) * E3 + (within < 8 ? a.E + h * kD + within * 4 : 2 * a.E + h * kD + (within - 8) * 4);
        cp_async16(st_s + (uint32_t)((((buf * kBlock + key) * 2 * kD) + within * 4) * 4), src);
    }
    cp_async_commit();
}

// ---------------------------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(kWarps * 32, 2) sasa_train_fwd_kernel(const Args a) {
    extern __shared__ __align__(16) float smem[];
    float* centres = smem;                                   // [Q][2]
    float* part = smem + 2 * ((a.Q + 1) & ~1);               // [kWarps][kD + 2][32]
    float* stage = part + kWarps * (kD + 2) * 32;            // [kWarps][2][kBlock][2 * kD]   (k | v), double-buffered
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 32;
    const size_t row_base = (size_t)b * a.Q;
    compute_centres(a, row_base, centres);
    __syncthreads();
    const int qi = min(q0 + lane, a.Q - 1);
    const int E3 = 3 * a.E;
    float2 q[kD / 2], acc[kD / 2];
    load_row32_f2(a.qkv + (row_base + qi) * E3 + h * kD, q);
#pragma unroll
    for (int d = 0; d < kD / 2; ++d) { q[d].x *= a.scale; q[d].y *= a.scale; acc[d] = make_float2(0.f, 0.f); }
    const float cx = centres[2 * qi], cy = centres[2 * qi + 1];
    const float neg_tau = -__ldg(a.tau + (row_base + qi) * a.H + h);
    float m = -INFINITY, l = 0.f;
    const int per = (a.Q + kWarps - 1) / kWarps;
    const int j_begin = warp * per, j_end = min(a.Q, j_begin + per);
    float* st2 = stage + (size_t)warp * 2 * kBlock * 2 * kD;
    const uint32_t st_s = (uint32_t)__cvta_generic_to_shared(st2);
    if (j_begin < j_end) prefetch_kv(a, row_base, h, j_begin, j_end, st_s, 0, lane);
    int buf = 0;
    for (int j0 = j_begin; j0 < j_end; j0 += kBlock, buf ^= 1) {
        // the block's eight mask bytes are requested up front (independent loads, in flight under the staging wait) instead
        // of one dependent global load per key inside the arithmetic
        unsigned blocked_bits = 0;
        if (a.blocked_t != nullptr) {
#pragma unroll
            for (int jj = 0; jj < kBlock; ++jj)
                blocked_bits |= (a.blocked_t[(size_t)min(j0 + jj, j_end - 1) * a.Q + qi] != 0 ? 1u : 0u) << jj;
        }
        __syncwarp();                                        // every lane is done with the buffer the next prefetch overwrites
        if (j0 + kBlock < j_end) {
            prefetch_kv(a, row_base, h, j0 + kBlock, j_end, st_s, buf ^ 1, lane);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const float* st = st2 + (size_t)buf * kBlock * 2 * kD;
        float s[kBlock];
#pragma unroll
        for (int jj = 0; jj < kBlock; ++jj) {
            const int j = j0 + jj;
            const float dot = dot32(q, reinterpret_cast<const float4*>(st + jj * 2 * kD));
            const bool ok = j < j_end && ((blocked_bits >> jj) & 1u) == 0;
            s[jj] = ok ? __fadd_rn(dot, __fmul_rn(dist_of(cx, cy, centres, min(j, a.Q - 1)), neg_tau)) : -INFINITY;
        }
        float mb = s[0];
#pragma unroll
        for (int jj = 1; jj < kBlock; ++jj) mb = fmaxf(mb, s[jj]);
        const float m_new = fmaxf(m, mb);
        if (m_new == -INFINITY) continue;                    // everything seen so far is blocked (warp-divergent but harmless)
        const float corr = expf(m - m_new);                  // exp(-inf) = 0 when nothing had been accumulated
        m = m_new;
        l *= corr;
#pragma unroll
        for (int d = 0; d < kD / 2; ++d) { acc[d].x *= corr; acc[d].y *= corr; }
#pragma unroll
        for (int jj = 0; jj < kBlock; ++jj) {
            const float p = expf(s[jj] - m);
            l += p;
            const float pa = p * keep_factor(a, b, h, qi, j0 + jj);
            axpy32(acc, pa, reinterpret_cast<const float4*>(st + jj * 2 * kD + kD));
        }
    }
    float* mine = part + (size_t)warp * (kD + 2) * 32 + lane;
    mine[0] = m;
    mine[32] = l;
#pragma unroll
    for (int d = 0; d < kD / 2; ++d) { mine[(2 + 2 * d) * 32] = acc[d].x; mine[(3 + 2 * d) * 32] = acc[d].y; }
    __syncthreads();
    float mx = -INFINITY;
#pragma unroll
    for (int s2 = 0; s2 < kWarps; ++s2) mx = fmaxf(mx, part[(size_t)s2 * (kD + 2) * 32 + lane]);
    float w[kWarps], den = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < kWarps; ++s2) {
        const float* p2 = part + (size_t)s2 * (kD + 2) * 32 + lane;
        w[s2] = (p2[0] == -INFINITY) ? 0.f : expf(p2[0] - mx);
        den = fmaf(p2[32], w[s2], den);
    }
    constexpr int kPer = kD / kWarps;
    float o[kPer];
#pragma unroll
    for (int e = 0; e < kPer; ++e) {
        float v = 0.f;
#pragma unroll
        for (int s2 = 0; s2 < kWarps; ++s2) v = fmaf(part[((size_t)s2 * (kD + 2) + 2 + warp * kPer + e) * 32 + lane], w[s2], v);
        o[e] = v / den;
    }
    if (q0 + lane < a.Q) {
        *reinterpret_cast<float4*>(a.out + (row_base + q0 + lane) * a.E + h * kD + warp * kPer) = make_float4(o[0], o[1], o[2], o[3]);
        if (warp == 0) a.lse[((size_t)b * a.H + h) * a.Q + q0 + lane] = mx + logf(den);
    }
}

// ------------------------------------------------------------------------------------------- backward: dq, dtau (and D)
__global__ void __launch_bounds__(kWarps * 32, 1) sasa_train_bwd_q_kernel(const Args a) {
    extern __shared__ __align__(16) float smem[];
    float* centres = smem;
    float* part = smem + 2 * ((a.Q + 1) & ~1);               // [kWarps][kD + 1][32]: dq, dtau partials
    float* stage = part + kWarps * (kD + 1) * 32;            // [kWarps][2][kBlock][2 * kD]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 32;
    const size_t row_base = (size_t)b * a.Q;
    compute_centres(a, row_base, centres);
    __syncthreads();
    const int qi = min(q0 + lane, a.Q - 1);
    const int E3 = 3 * a.E;
    float2 q[kD / 2], go[kD / 2], dq[kD / 2];
    load_row32_f2(a.qkv + (row_base + qi) * E3 + h * kD, q);
    load_row32_f2(a.grad_out + (row_base + qi) * a.E + h * kD, go);
    float D = 0.f;
    {
        float2 o[kD / 2];
        load_row32_f2(a.out + (row_base + qi) * a.E + h * kD, o);
#pragma unroll
        for (int d = 0; d < kD / 2; ++d) {
            D = fmaf(go[d].x, o[d].x, D);
            D = fmaf(go[d].y, o[d].y, D);
            q[d].x *= a.scale; q[d].y *= a.scale;
            dq[d] = make_float2(0.f, 0.f);
        }
    }
    const float lse = a.lse[((size_t)b * a.H + h) * a.Q + qi];
    const float cx = centres[2 * qi], cy = centres[2 * qi + 1];
    const float neg_tau = -__ldg(a.tau + (row_base + qi) * a.H + h);
    float dtau = 0.f;
    const int per = (a.Q + kWarps - 1) / kWarps;
    const int j_begin = warp * per, j_end = min(a.Q, j_begin + per);
    float* st2 = stage + (size_t)warp * 2 * kBlock * 2 * kD;
    const uint32_t st_s = (uint32_t)__cvta_generic_to_shared(st2);
    if (j_begin < j_end) prefetch_kv(a, row_base, h, j_begin, j_end, st_s, 0, lane);
    int buf = 0;
    for (int j0 = j_begin; j0 < j_end; j0 += kBlock, buf ^= 1) {
        unsigned blocked_bits = 0;
        if (a.blocked_t != nullptr) {
#pragma unroll
            for (int jj = 0; jj < kBlock; ++jj)
                blocked_bits |= (a.blocked_t[(size_t)min(j0 + jj, j_end - 1) * a.Q + qi] != 0 ? 1u : 0u) << jj;
        }
        __syncwarp();
        if (j0 + kBlock < j_end) {
            prefetch_kv(a, row_base, h, j0 + kBlock, j_end, st_s, buf ^ 1, lane);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const float* st = st2 + (size_t)buf * kBlock * 2 * kD;
#pragma unroll 2
        for (int jj = 0; jj < kBlock; ++jj) {
            const int j = j0 + jj;
            if (j >= j_end) break;                           // warp-uniform
            const float4* kp = reinterpret_cast<const float4*>(st + jj * 2 * kD);
            const float dot = dot32(q, kp), da = dot32(go, kp + kD / 4);
            const bool ok = ((blocked_bits >> jj) & 1u) == 0;
            const float dist = dist_of(cx, cy, centres, j);
            const float s = __fadd_rn(dot, __fmul_rn(dist, neg_tau));
            const float p = ok ? expf(s - lse) : 0.f;
            const float ds = p * (keep_factor(a, b, h, qi, j) * da - D);
            dtau = fmaf(-ds, dist, dtau);
            axpy32(dq, ds, kp);
        }
    }
    float* mine = part + (size_t)warp * (kD + 1) * 32 + lane;
#pragma unroll
    for (int d = 0; d < kD / 2; ++d) { mine[(2 * d) * 32] = dq[d].x; mine[(2 * d + 1) * 32] = dq[d].y; }
    mine[kD * 32] = dtau;
    __syncthreads();
    if (q0 + lane < a.Q) {
        constexpr int kPer = kD / kWarps;
        float o[kPer];
#pragma unroll
        for (int e = 0; e < kPer; ++e) {
            float v = 0.f;
#pragma unroll
            for (int s2 = 0; s2 < kWarps; ++s2) v += part[((size_t)s2 * (kD + 1) + warp * kPer + e) * 32 + lane];
            o[e] = v * a.scale;
        }
        *reinterpret_cast<float4*>(a.grad_qkv + (row_base + q0 + lane) * E3 + h * kD + warp * kPer) = make_float4(o[0], o[1], o[2], o[3]);
        if (warp == 0) {
            float t = 0.f;
#pragma unroll
            for (int s2 = 0; s2 < kWarps; ++s2) t += part[((size_t)s2 * (kD + 1) + kD) * 32 + lane];
            a.grad_tau[(row_base + q0 + lane) * a.H + h] = t;
            a.dsum[((size_t)b * a.H + h) * a.Q + q0 + lane] = D;
        }
    }
}

// ------------------------------------------------------------------------------------------------- backward: dk, dv
// lane = key; staged per query: q (scaled) 32 | do 32 | lse, D, -tau, cx, cy (5) -> 72 floats (16-byte aligned rows)
constexpr int kQRow = 72;

__global__ void __launch_bounds__(kWarps * 32, 1) sasa_train_bwd_kv_kernel(const Args a) {
    extern __shared__ __align__(16) float smem[];
    float* centres = smem;
    float* part = smem + 2 * ((a.Q + 1) & ~1);               // [kWarps][2 * kD][32]: dk, dv partials
    float* stage = part + kWarps * 2 * kD * 32;              // [kWarps][2][kBlock][kQRow]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * 32;
    const size_t row_base = (size_t)b * a.Q;
    compute_centres(a, row_base, centres);
    __syncthreads();
    const int kj = min(k0 + lane, a.Q - 1);
    const int E3 = 3 * a.E;
    float2 k[kD / 2], v[kD / 2], dk[kD / 2], dv[kD / 2];
    load_row32_f2(a.qkv + (row_base + kj) * E3 + a.E + h * kD, k);
    load_row32_f2(a.qkv + (row_base + kj) * E3 + 2 * a.E + h * kD, v);
#pragma unroll
    for (int d = 0; d < kD / 2; ++d) { dk[d] = make_float2(0.f, 0.f); dv[d] = make_float2(0.f, 0.f); }
    const float kx = centres[2 * kj], ky = centres[2 * kj + 1];
    const int per = (a.Q + kWarps - 1) / kWarps;
    const int i_begin = warp * per, i_end = min(a.Q, i_begin + per);
    float* st2 = stage + (size_t)warp * 2 * kBlock * kQRow;
    const uint32_t st_s = (uint32_t)__cvta_generic_to_shared(st2);
    auto prefetch_q = [&](int i0, int bufn) {                 // 8 float4 of q, 8 of do per query + lse, D, tau (cp.async), centre
#pragma unroll
        for (int t = 0; t < kBlock * 16 / 32; ++t) {
            const int c = lane + 32 * t, qq = c >> 4, within = c & 15;
            const int i = min(i0 + qq, i_end - 1);
            const float* src = within < 8 ? a.qkv + (row_base + i) * E3 + h * kD + within * 4
                                          : a.grad_out + (row_base + i) * a.E + h * kD + (within - 8) * 4;
            cp_async16(st_s + (uint32_t)(((bufn * kBlock + qq) * kQRow + within * 4) * 4), src);
        }
        if (lane < kBlock) {
            const int i = min(i0 + lane, i_end - 1);
            const uint32_t r = st_s + (uint32_t)(((bufn * kBlock + lane) * kQRow + 2 * kD) * 4);
            cp_async4(r, a.lse + ((size_t)b * a.H + h) * a.Q + i);
            cp_async4(r + 4, a.dsum + ((size_t)b * a.H + h) * a.Q + i);
            cp_async4(r + 8, a.tau + (row_base + i) * a.H + h);               // negated when it is read
            float* rc = st2 + ((size_t)bufn * kBlock + lane) * kQRow + 2 * kD;
            rc[3] = centres[2 * i];
            rc[4] = centres[2 * i + 1];
        }
        cp_async_commit();
    };
    if (i_begin < i_end) prefetch_q(i_begin, 0);
    int buf = 0;
    for (int i0 = i_begin; i0 < i_end; i0 += kBlock, buf ^= 1) {
        unsigned blocked_bits = 0;                           // this key's row of the transposed mask: 8 adjacent bytes
        if (a.blocked_t != nullptr) {
#pragma unroll
            for (int qq = 0; qq < kBlock; ++qq)
                blocked_bits |= (a.blocked_t[(size_t)kj * a.Q + min(i0 + qq, i_end - 1)] != 0 ? 1u : 0u) << qq;
        }
        __syncwarp();
        if (i0 + kBlock < i_end) {
            prefetch_q(i0 + kBlock, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const float* st = st2 + (size_t)buf * kBlock * kQRow;
#pragma unroll 2
        for (int qq = 0; qq < kBlock; ++qq) {
            const int i = i0 + qq;
            if (i >= i_end) break;                           // warp-uniform
            const float* r = st + qq * kQRow;
            const float4* qp = reinterpret_cast<const float4*>(r);
            const float dot = dot32(k, qp), da = dot32(v, qp + kD / 4);
            const float lse = r[2 * kD], D = r[2 * kD + 1], neg_tau = -r[2 * kD + 2];
            const float dx = r[2 * kD + 3] - kx, dy = r[2 * kD + 4] - ky;
            const float dist = sqrtf(fmaf(dy, dy, __fmul_rn(dx, dx)));
            const bool ok = ((blocked_bits >> qq) & 1u) == 0;
            const float s = __fadd_rn(dot * a.scale, __fmul_rn(dist, neg_tau));
            const float p = ok ? expf(s - lse) : 0.f;
            const float keep = keep_factor(a, b, h, i, kj);
            const float pa = p * keep;
            const float ds = p * (keep * da - D) * a.scale;
            axpy32(dk, ds, qp);
            axpy32(dv, pa, qp + kD / 4);
        }
    }
    float* mine = part + (size_t)warp * 2 * kD * 32 + lane;
#pragma unroll
    for (int d = 0; d < kD / 2; ++d) {
        mine[(2 * d) * 32] = dk[d].x; mine[(2 * d + 1) * 32] = dk[d].y;
        mine[(kD + 2 * d) * 32] = dv[d].x; mine[(kD + 2 * d + 1) * 32] = dv[d].y;
    }
    __syncthreads();
    if (k0 + lane < a.Q) {
        constexpr int kPer = 2 * kD / kWarps;                // 8 of the 64 (dk | dv) values per warp
        float o[kPer];
#pragma unroll
        for (int e = 0; e < kPer; ++e) {
            float t = 0.f;
#pragma unroll
            for (int s2 = 0; s2 < kWarps; ++s2) t += part[((size_t)s2 * 2 * kD + warp * kPer + e) * 32 + lane];
            o[e] = t;
        }
        const int col = warp * kPer;                         // 0..31 -> dk, 32..63 -> dv
        float* dst = a.grad_qkv + (row_base + k0 + lane) * E3 + (col < kD ? a.E + h * kD + col : 2 * a.E + h * kD + col - kD);
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
}

static int fill(Args& a, const float* qkv, const float* tau, const float* query_ray, const uint8_t* blocked_t, const double* pc_range,
                int batch, int num_query, int num_heads, int head_dim, int code_size, float drop_p, unsigned seed) {
    if (!qkv || !tau || !query_ray || !pc_range) return RACF_ERR_NULL_POINTER;
    if (batch <= 0 || num_query <= 0 || num_heads <= 0 || code_size < 2 || !(drop_p >= 0.f && drop_p < 1.f)) return RACF_ERR_BAD_SHAPE;
    if (head_dim != kD || batch > 65535 || num_heads > 65535) return RACF_ERR_UNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(qkv) & 15u) return RACF_ERR_UNSUPPORTED;
    a.qkv = qkv; a.tau = tau; a.query_ray = query_ray; a.blocked_t = blocked_t;
    a.Q = num_query; a.H = num_heads; a.E = num_heads * head_dim; a.code = code_size;
    a.scale = (float)(1.0 / sqrt((double)head_dim));
    a.drop_p = drop_p; a.keep_scale = 1.0f / (1.0f - drop_p); a.seed = seed;
    a.x0 = (float)pc_range[0]; a.xs = (float)(pc_range[3] - pc_range[0]);
    a.y0 = (float)pc_range[1]; a.ys = (float)(pc_range[4] - pc_range[1]);
    return RACF_OK;
}

}  // namespace sasat
}  // namespace racf

extern "C" int racf_sasa_attention_train_forward(const float* qkv, const float* tau, const float* query_ray,
                                                 const uint8_t* blocked_t, const double* pc_range, int batch, int num_query,
                                                 int num_heads, int head_dim, int code_size, float drop_p, unsigned seed,
                                                 float* out, float* lse, racf_stream_t stream) {
    using namespace racf::sasat;
    Args a = {};
    const int rc = fill(a, qkv, tau, query_ray, blocked_t, pc_range, batch, num_query, num_heads, head_dim, code_size, drop_p, seed);
    if (rc != RACF_OK) return rc;
    if (!out || !lse) return RACF_ERR_NULL_POINTER;
    if (reinterpret_cast<uintptr_t>(out) & 15u) return RACF_ERR_UNSUPPORTED;
    a.out = out; a.lse = lse;
    const size_t smem = (size_t)(2 * ((num_query + 1) & ~1) + kWarps * (kD + 2) * 32 + kWarps * 2 * kBlock * 2 * kD) * sizeof(float);
    if (smem > 100u * 1024u) return RACF_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(sasa_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const dim3 grid((unsigned)((num_query + 31) / 32), (unsigned)num_heads, (unsigned)batch);
    sasa_train_fwd_kernel<<<grid, kWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(a);
    return (int)cudaGetLastError();
}

extern "C" int racf_sasa_attention_train_backward(const float* qkv, const float* tau, const float* query_ray,
                                                  const uint8_t* blocked_t, const double* pc_range, int batch, int num_query,
                                                  int num_heads, int head_dim, int code_size, float drop_p, unsigned seed,
                                                  const float* out, const float* lse, const float* grad_out, float* dsum,
                                                  float* grad_qkv, float* grad_tau, racf_stream_t stream) {
    using namespace racf::sasat;
    Args a = {};
    const int rc = fill(a, qkv, tau, query_ray, blocked_t, pc_range, batch, num_query, num_heads, head_dim, code_size, drop_p, seed);
    if (rc != RACF_OK) return rc;
    if (!out || !lse || !grad_out || !dsum || !grad_qkv || !grad_tau) return RACF_ERR_NULL_POINTER;
    if ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(grad_qkv)) & 15u)
        return RACF_ERR_UNSUPPORTED;
    a.out = const_cast<float*>(out); a.lse = const_cast<float*>(lse); a.grad_out = grad_out; a.dsum = dsum;
    a.grad_qkv = grad_qkv; a.grad_tau = grad_tau;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const dim3 grid((unsigned)((num_query + 31) / 32), (unsigned)num_heads, (unsigned)batch);
    const size_t cen = (size_t)2 * ((num_query + 1) & ~1);
    const size_t smem_q = (cen + kWarps * (kD + 1) * 32 + kWarps * 2 * kBlock * 2 * kD) * sizeof(float);
    const size_t smem_kv = (cen + kWarps * 2 * kD * 32 + kWarps * 2 * kBlock * kQRow) * sizeof(float);
    if (smem_q > 200u * 1024u || smem_kv > 200u * 1024u) return RACF_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(sasa_train_bwd_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(sasa_train_bwd_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv);
    if (e != cudaSuccess) return (int)e;
    sasa_train_bwd_q_kernel<<<grid, kWarps * 32, smem_q, st>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    sasa_train_bwd_kv_kernel<<<grid, kWarps * 32, smem_kv, st>>>(a);
    return (int)cudaGetLastError();
}
