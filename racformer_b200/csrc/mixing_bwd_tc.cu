// Backward of the fused AdaptiveMixing core on the 5th-generation tensor cores (training) -- SURVEY.md 8f-4.
//
// Same contract as csrc/mixing_bwd.cu (reference: AdaptiveMixing.inner_forward, models/racformer_transformer.py:592-604,
// differentiated by autograd there): per (query, group) item, with x [P_in,64], M [64,64], S [128,P_in],
//     o1 = x M ;  n1 = LN(o1) ;  t = relu(n1) ;  o2 = S t ;  n2 = LN(o2) ;  y = relu(n2)
//     g_n2 = gy [n2 > 0]      g_o2 = rstd2 (g_n2 - mean(g_n2) - n2 mean(g_n2 n2))
//     g_S  = g_o2 t^T         g_t  = S^T g_o2
//     g_n1 = g_t [n1 > 0]     g_o1 = rstd1 (g_n1 - mean(g_n1) - n1 mean(g_n1 n1))
//     g_M  = x^T g_o1         g_x  = g_o1 M^T
// All six products (two recomputed, four gradients; 7.1 MFLOP per item) run as tcgen05.mma on bf16 pieces produced on the
// fly: every fp32 operand is split exactly into three bf16 pieces, the six largest piece products are accumulated in fp32
// in tensor memory (a0*b0 and the cross terms in separate accumulators, see csrc/linear.cu), so every product is
// fp32-grade; the layer norms and their backward run in fp32 on the CUDA cores straight from tensor memory.
//
// One persistent CTA (512 threads) per SM walks its items phase by phase. Operand tiles (bf16 x 3 pieces, shared memory);
// every tile serves two products, once K-major and once MN-major (the descriptor's major-ness bit), so nothing is transposed:
//   X3  [p][c]   128-byte rows, 128-byte swizzle   A of o1 = x M (K-major, K = c)       A of g_M = x^T g_o1 (MN-major, K = p)
//   M3  [c'][c]  = M^T                             B of o1 (K-major, K = c)             B of g_x = g_o1 M^T (MN-major, K = c')
//   S3  [o][p]   32-wide atoms, 64-byte swizzle    A of o2 = S t (K-major, K = p)       A of g_t = S^T g_o2 (MN-major, K = o)
//   T3  [p][c']  t, 128-byte swizzle               B of o2 (MN-major, K = p)            B of g_S = g_o2 t^T (K-major, N = P_in)
//   G2  [o][c']  g_o2                              A of g_S (K-major, K = c')           B of g_t (MN-major, K = o)
//   G1  [p][c']  g_o1, overwrites T3               A of g_x (K-major, K = c')           B of g_M (MN-major, K = p)
// MN-major A (instruction-descriptor bit 15): LBO = distance between the swizzle atoms along M, SBO = distance between the
// 8-row groups along K (measured: the swapped assignment gives wrong results). Rows that are never written (p >= P_in, and
// the upper half of g_M's M = 128) only feed accumulator rows that are never read. Overlap: g_S leaves tensor memory while
// g_t is being computed; the next item's S is split while g_M / g_x are computed.
// 216 KB of shared memory, 448 of the 512 columns of tensor memory.
// Implemented for C == 64, P_out == 128, P_in % 16 == 0, 16 <= P_in <= 96 (csrc/mixing_bwd.cu covers P_in <= 128).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#ifdef RACF_MIXBT_PROFILE
#include <cstdio>
#endif

#include "racformer_ops.h"

namespace racf {
namespace mixbt {

constexpr int kC = 64, kPout = 128, kMaxPin = 96, kThreads = 512;
constexpr int kXPiece = kMaxPin * 128;      // 12 KB  [96 p][64 c]
constexpr int kMPiece = kC * 128;           //  8 KB  [64 c'][64 c]
constexpr int kSPiece = 3 * 8192;           // 24 KB  3 atoms of [128 o][32 p]
constexpr int kTPiece = kMaxPin * 128;      // 12 KB  [96 p][64 c']
constexpr int kGPiece = kPout * 128;        // 16 KB  [128 o][64 c']
constexpr int kX3 = 0;
constexpr int kM3 = kX3 + 3 * kXPiece;
constexpr int kS3 = kM3 + 3 * kMPiece;
constexpr int kT3 = kS3 + 3 * kSPiece;
constexpr int kG3 = kT3 + 3 * kTPiece;
constexpr int kSmemBytes = kG3 + 3 * kGPiece;   // 216 KB
static_assert(kM3 % 1024 == 0 && kS3 % 1024 == 0 && kT3 % 1024 == 0 && kG3 % 1024 == 0, "swizzle atoms are 1 KB aligned");
constexpr int kTmemCols = 512;
// tensor-memory columns (main, cross): D1 (0, 64) o1 | D2 (128, 192) o2 | D3 (128, 224) g_S, 96 wide, over D2 | D4 (320, 384) g_t |
// D5 (128, 192) g_M, over D2 / D3 | D6 (320, 384) g_x, over D4. D1 is free again when g_M / g_x are read out: the next item's
// o1 runs meanwhile.
constexpr uint32_t kD1 = 0, kD2 = 128, kD3 = 128, kD3x = 224, kD4 = 320, kD5 = 128, kD6 = 320;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {   // bounded: a bug must not hang the GPU
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > 2000000000LL) __trap();
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// Shared-memory matrix descriptors (sm_100 UMMA, version 1), the forms used by csrc/mixing_ws.cu.
template <int kRowBytes>
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) {     // K-major, rows one swizzle span wide, 8-row groups SBO apart
    constexpr uint64_t layout = kRowBytes == 128 ? 2 : 4;
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((8 * kRowBytes) >> 4) << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ uint64_t desc_mnmajor_sw128(uint32_t addr) {   // [k rows][64 n] bf16, 128-byte rows, 8-row groups 1 KB apart
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr, uint32_t lbo, uint32_t sbo, uint64_t layout) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | (layout << 61);
}

// exact three-way split of two floats: x = p0 + p1 + p2, bf16 pieces, round to nearest at each step
__device__ __forceinline__ void split3x2(float a, float b, uint32_t (&q)[3]) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        q[k] = *reinterpret_cast<const uint32_t*>(&h);
        if (k < 2) {
            a -= __uint_as_float(q[k] << 16);
            b -= __uint_as_float(q[k] & 0xffff0000u);
        }
    }
}
// eight consecutive fp32 -> one 16-byte chunk per piece at byte offset `off` of each piece tile
__device__ __forceinline__ void split_store8(const float* f, uint8_t* tile, int piece_bytes, int off) {
    uint32_t p[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t q[3];
        split3x2(f[2 * j], f[2 * j + 1], q);
        p[0][j] = q[0]; p[1][j] = q[1]; p[2][j] = q[2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        *reinterpret_cast<uint4*>(tile + k * piece_bytes + off) = make_uint4(p[k][0], p[k][1], p[k][2], p[k][3]);
}

// two block-wide sums at once (fixed order -> deterministic)
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = make_float2(a, b);
    __syncthreads();
    float2 t = make_float2(0.f, 0.f);
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) {
        const float2 v = red[w];
        t.x += v.x;
        t.y += v.y;
    }
    return t;
}

// x [P_in][64] -> X3 (row p, 16-byte chunk of 8 c)
__device__ __forceinline__ void load_split_x(const float* __restrict__ xg, uint8_t* sm, int p_in, int tid) {
    constexpr int kIt = kMaxPin * 8 / kThreads + ((kMaxPin * 8) % kThreads != 0);
    float4 xa[kIt][2];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
        const int i = tid + it * kThreads;
        if (i < p_in * 8) {
            xa[it][0] = __ldg(reinterpret_cast<const float4*>(xg + i * 8));
            xa[it][1] = __ldg(reinterpret_cast<const float4*>(xg + i * 8 + 4));
        }
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
        const int i = tid + it * kThreads;
        if (i < p_in * 8) {
            const int p = i >> 3, ch = i & 7;
            const float f[8] = {xa[it][0].x, xa[it][0].y, xa[it][0].z, xa[it][0].w, xa[it][1].x, xa[it][1].y, xa[it][1].z, xa[it][1].w};
            split_store8(f, sm + kX3, kXPiece, p * 128 + ((ch ^ (p & 7)) << 4));
        }
    }
}
// M [64 c][64 c'] -> M3 = M^T (row c', chunk of 8 c; lanes along c' -> coalesced loads)
__device__ __forceinline__ void load_split_m(const float* __restrict__ mg, uint8_t* sm, int tid) {
    const int cp = tid & 63, ch = tid >> 6;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __ldg(mg + (ch * 8 + j) * kC + cp);
    split_store8(f, sm + kM3, kMPiece, cp * 128 + ((ch ^ (cp & 7)) << 4));
}
// S [128 o][P_in] -> S3: 32-wide K atoms of [128 rows][64 B], 64-byte swizzle
__device__ __forceinline__ void load_split_s(const float* __restrict__ sg, uint8_t* sm, int p_in, int tid) {
    const int s_chunks = p_in >> 3;
    constexpr int kIt = kPout * (kMaxPin / 8) / kThreads;      // 3
    float4 sa[kIt][2];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
        const int i = tid + it * kThreads;
        if (i < kPout * s_chunks) {
            sa[it][0] = __ldg(reinterpret_cast<const float4*>(sg + i * 8));
            sa[it][1] = __ldg(reinterpret_cast<const float4*>(sg + i * 8 + 4));
        }
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
        const int i = tid + it * kThreads;
        if (i < kPout * s_chunks) {
            const int o = i / s_chunks, j = i - o * s_chunks;
            const float f[8] = {sa[it][0].x, sa[it][0].y, sa[it][0].z, sa[it][0].w, sa[it][1].x, sa[it][1].y, sa[it][1].z, sa[it][1].w};
            split_store8(f, sm + kS3, kSPiece, (j >> 2) * 8192 + o * 64 + (((j & 3) ^ ((o >> 1) & 3)) << 4));
        }
    }
}
// One product: D = A B with the six largest piece products; a0*b0 into d_main, the five cross terms into d_cross.
// a_off / b_off: byte offset of K step ks (16 elements of K) inside a piece tile.
template <typename AOff, typename BOff>
__device__ __forceinline__ void issue_product(uint32_t d_main, uint32_t d_cross, uint64_t da, uint64_t db, int a_piece, int b_piece,
                                              int ksteps, uint32_t idesc, AOff a_off, BOff b_off) {
    uint32_t acc_cross = 0;
    for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t ao = (uint64_t)(a_off(ks) >> 4), bo = (uint64_t)(b_off(ks) >> 4);
#pragma unroll
        for (int order = 2; order >= 1; --order)
#pragma unroll
            for (int pa = 0; pa <= order; ++pa) {
                umma_bf16(d_cross, da + ao + (uint64_t)((pa * a_piece) >> 4), db + bo + (uint64_t)(((order - pa) * b_piece) >> 4), idesc,
                          acc_cross);
                acc_cross = 1;
            }
        umma_bf16(d_main, da + ao, db + bo, idesc, ks > 0);
    }
}

// 16 accumulator columns of this thread's row: main + cross terms
__device__ __forceinline__ void load_acc16(uint32_t t_main, uint32_t t_cross, float (&f)[16]) {
    uint32_t v[16], u[16];
    tmem_ld16(t_main, v);
    tmem_ld16(t_cross, u);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + __uint_as_float(u[j]);
}

// 16 consecutive fp32 of row `row` (columns col0 ..) -> two 16-byte chunks per piece of a [rows][64] 128-byte-swizzled tile
__device__ __forceinline__ void split_store_row16(const float (&f)[16], uint8_t* tile, int piece_bytes, int row, int col0) {
    const int ch = col0 >> 3;
    split_store8(f, tile, piece_bytes, row * 128 + ((ch ^ (row & 7)) << 4));
    split_store8(f + 8, tile, piece_bytes, row * 128 + (((ch + 1) ^ (row & 7)) << 4));
}

#ifdef RACF_MIXBT_PROFILE      // tuning aid: per-phase cycle counts of CTA 0 (build with RACF_NVCC_DEFINES=-DRACF_MIXBT_PROFILE)
#define MIXBT_TICK(slot) do { if (tid == 0) { const long long now_ = clock64(); prof[slot] += now_ - tprev; tprev = now_; } } while (0)
#else
#define MIXBT_TICK(slot) do { } while (0)
#endif

__global__ void __launch_bounds__(kThreads, 1)
adaptive_mixing_bwd_tc_kernel(const float* __restrict__ x, const float* __restrict__ params, const float* __restrict__ gy,
                              float* __restrict__ grad_x, float* __restrict__ grad_params, int num_items, int p_in, float eps) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float2 red[kThreads / 32];

    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_addr = smem_u32(sm);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_a = smem_u32(&bars[0]), bar_b = smem_u32(&bars[1]);   // bar_b: g_S alone (its read-out overlaps g_t)

    if (tid == 0) {
        mbar_init(bar_a, 1);
        mbar_init(bar_b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    // instruction descriptors: fp32 accumulate, bf16 x bf16, M = 128; bit 15 / 16 = A / B is MN-major
    constexpr uint32_t kIdescBase = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc_kk = kIdescBase | ((uint32_t)(kC >> 3) << 17);                 // N = 64, both K-major
    constexpr uint32_t idesc_kn = idesc_kk | (1u << 16);                                    // N = 64, B MN-major
    constexpr uint32_t idesc_nn = idesc_kn | (1u << 15);                                    // N = 64, A and B MN-major
    const uint32_t idesc_gs = kIdescBase | ((uint32_t)(p_in >> 3) << 17);                   // N = P_in, both K-major
    const int ksteps_p = p_in >> 4;

    const uint64_t d_x3_k = desc_kmajor<128>(sm_addr + kX3), d_x3_n = desc_mnmajor_sw128(sm_addr + kX3);
    const uint64_t d_m3_k = desc_kmajor<128>(sm_addr + kM3), d_m3_n = desc_mnmajor_sw128(sm_addr + kM3);
    const uint64_t d_s3_k = desc_kmajor<64>(sm_addr + kS3), d_s3_n = desc_mnmajor(sm_addr + kS3, 8192, 512, 4);
    const uint64_t d_t3_n = desc_mnmajor_sw128(sm_addr + kT3), d_t3_k = desc_kmajor<128>(sm_addr + kT3);   // also G1
    const uint64_t d_g2_k = desc_kmajor<128>(sm_addr + kG3), d_g2_n = desc_mnmajor_sw128(sm_addr + kG3);

    // this thread's slice of an accumulator tile: TMEM lane quarter (warp % 4), 16 of 64 columns (warp / 4)
    const int row = (warp & 3) * 32 + lane, cg = warp >> 2, col0 = cg * 16;
    const uint32_t tm_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const long long per_item = kC * kC + (long long)kPout * p_in;
    const bool live = row < p_in;
    const float n1_cnt = (float)(p_in * kC), n2_cnt = (float)(kPout * kC);
    uint32_t phase_a = 0, phase_b = 0;
#ifdef RACF_MIXBT_PROFILE
    long long prof[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif

    // o1 = x M (recompute) of an item is issued as soon as its X3 / M3 are in place: for the first item here, for the others
    // at the end of the previous item's loop body, under the read-out of that item's g_M / g_x
    auto issue_o1 = [&]() {
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            if (elect_one()) {
                tc_fence_after();
                issue_product(tmem + kD1, tmem + kD1 + 64, d_x3_k, d_m3_k, kXPiece, kMPiece, 4, idesc_kk,
                              [](int ks) { return ks * 32; }, [](int ks) { return ks * 32; });
                umma_commit(bar_a);
            }
            __syncwarp();
        }
    };
    auto prefetch_item = [&](int it) {          // pull an item's inputs into L2 well before they are read (128-byte lines)
        const char* px = reinterpret_cast<const char*>(x + (long long)it * p_in * kC);
        const char* pp = reinterpret_cast<const char*>(params + (long long)it * per_item);
        const char* pgy = reinterpret_cast<const char*>(gy + (long long)it * (kPout * kC));
        for (int i = tid * 128; i < p_in * kC * 4; i += kThreads * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(px + i));
        for (int i = tid * 128; i < (int)per_item * 4; i += kThreads * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + i));
        for (int i = tid * 128; i < kPout * kC * 4; i += kThreads * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pgy + i));
    };

    if ((int)blockIdx.x < num_items) {
        const float* pg0 = params + (long long)blockIdx.x * per_item;
        load_split_x(x + (long long)blockIdx.x * p_in * kC, sm, p_in, tid);
        load_split_m(pg0, sm, tid);
        issue_o1();
        load_split_s(pg0 + kC * kC, sm, p_in, tid);
    }

    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int nxt = item + gridDim.x;
        float* gpg = grad_params + (long long)item * per_item;
        MIXBT_TICK(0);

        // ---- o1 = x M (recompute) is in flight -------------------------------------------------------------------------------
        if (nxt < num_items) prefetch_item(nxt);
        float4 gyv[4];                                       // gy[o = row][col0 .. col0 + 15]: needed after o2, loaded now
        {
            const float4* gp = reinterpret_cast<const float4*>(gy + (long long)item * (kPout * kC) + row * kC + col0);
#pragma unroll
            for (int c = 0; c < 4; ++c) gyv[c] = __ldg(gp + c);
        }
        mbar_wait(bar_a, phase_a & 1u);
        ++phase_a;
        tc_fence_after();
        MIXBT_TICK(1);

        // ---- LN1 + ReLU -> T3 -----------------------------------------------------------------------------------------------
        float mean1, rstd1;
        {
            float f[16];
            load_acc16(tm_lane + kD1 + col0, tm_lane + kD1 + 64 + col0, f);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) s += live ? f[j] : 0.f;
            mean1 = block_sum2(s, 0.f, red).x / n1_cnt;
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float d = f[j] - mean1;
                q += live ? d * d : 0.f;
            }
            rstd1 = rsqrtf(block_sum2(q, 0.f, red).x / n1_cnt + eps);
            if (live) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = fmaxf((f[j] - mean1) * rstd1, 0.f);
                split_store_row16(f, sm + kT3, kTPiece, row, col0);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        MIXBT_TICK(2);

        // ---- o2 = S t (recompute) -------------------------------------------------------------------------------------------
        if (warp == 0) {
            if (elect_one()) {
                tc_fence_after();
                issue_product(tmem + kD2, tmem + kD2 + 64, d_s3_k, d_t3_n, kSPiece, kTPiece, ksteps_p, idesc_kn,
                              [](int ks) { return (ks >> 1) * 8192 + (ks & 1) * 32; }, [](int ks) { return ks * 2048; });
                umma_commit(bar_a);
            }
            __syncwarp();
        }
        mbar_wait(bar_a, phase_a & 1u);
        ++phase_a;
        tc_fence_after();
        MIXBT_TICK(3);

        // ---- LN2, ReLU mask, LN2 backward -> G2 = g_o2 -------------------------------------------------------------------------
        {
            float f[16];
            load_acc16(tm_lane + kD2 + col0, tm_lane + kD2 + 64 + col0, f);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) s += f[j];
            const float mean2 = block_sum2(s, 0.f, red).x / n2_cnt;
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float d = f[j] - mean2;
                q += d * d;
            }
            const float rstd2 = rsqrtf(block_sum2(q, 0.f, red).x / n2_cnt + eps);
            const float g[16] = {gyv[0].x, gyv[0].y, gyv[0].z, gyv[0].w, gyv[1].x, gyv[1].y, gyv[1].z, gyv[1].w,
                                 gyv[2].x, gyv[2].y, gyv[2].z, gyv[2].w, gyv[3].x, gyv[3].y, gyv[3].z, gyv[3].w};
            float gn[16];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                f[j] = (f[j] - mean2) * rstd2;               // n2
                gn[j] = f[j] > 0.f ? g[j] : 0.f;
                s1 += gn[j];
                s2 += gn[j] * f[j];
            }
            const float2 t = block_sum2(s1, s2, red);
            const float m1 = t.x / n2_cnt, m2 = t.y / n2_cnt;
#pragma unroll
            for (int j = 0; j < 16; ++j) gn[j] = rstd2 * (gn[j] - m1 - f[j] * m2);
            split_store_row16(gn, sm + kG3, kGPiece, row, col0);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        MIXBT_TICK(4);

        // ---- g_S = g_o2 t^T, then g_t = S^T g_o2; g_S leaves tensor memory while g_t runs ---------------------------------------
        if (warp == 0) {
            if (elect_one()) {
                tc_fence_after();
                issue_product(tmem + kD3, tmem + kD3x, d_g2_k, d_t3_k, kGPiece, kTPiece, 4, idesc_gs,
                              [](int ks) { return ks * 32; }, [](int ks) { return ks * 32; });
                umma_commit(bar_b);
                issue_product(tmem + kD4, tmem + kD4 + 64, d_s3_n, d_g2_n, kSPiece, kGPiece, 8, idesc_nn,
                              [](int ks) { return ks * 1024; }, [](int ks) { return ks * 2048; });
                umma_commit(bar_a);
            }
            __syncwarp();
        }
        mbar_wait(bar_b, phase_b & 1u);
        ++phase_b;
        tc_fence_after();
        MIXBT_TICK(5);
        {
            float* gs = gpg + kC * kC + (long long)row * p_in;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c0 = h * 64 + col0;
                if (c0 < p_in) {                             // warp-uniform
                    float f[16];
                    load_acc16(tm_lane + kD3 + c0, tm_lane + kD3x + c0, f);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<float4*>(gs + c0 + 4 * c) = make_float4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
                }
            }
        }
        MIXBT_TICK(6);
        mbar_wait(bar_a, phase_a & 1u);
        ++phase_a;
        tc_fence_after();
        MIXBT_TICK(7);

        // ---- g_t, ReLU mask, LN1 backward -> G1 = g_o1 (over T3: g_S has completed) ---------------------------------------------
        {
            float n1[16], gn[16];
            load_acc16(tm_lane + kD1 + col0, tm_lane + kD1 + 64 + col0, n1);
            load_acc16(tm_lane + kD4 + col0, tm_lane + kD4 + 64 + col0, gn);
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                n1[j] = (n1[j] - mean1) * rstd1;
                gn[j] = (live && n1[j] > 0.f) ? gn[j] : 0.f;
                n1[j] = live ? n1[j] : 0.f;
                s1 += gn[j];
                s2 += gn[j] * n1[j];
            }
            const float2 t = block_sum2(s1, s2, red);
            const float m1 = t.x / n1_cnt, m2 = t.y / n1_cnt;
            if (live) {
#pragma unroll
                for (int j = 0; j < 16; ++j) gn[j] = rstd1 * (gn[j] - m1 - n1[j] * m2);
                split_store_row16(gn, sm + kT3, kTPiece, row, col0);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        MIXBT_TICK(8);

        // ---- g_M = x^T g_o1 ; g_x = g_o1 M^T ; the next item's S is split meanwhile (S3 is dead: g_t has completed) ---------------
        if (warp == 0) {
            if (elect_one()) {
                tc_fence_after();
                issue_product(tmem + kD5, tmem + kD5 + 64, d_x3_n, d_t3_n, kXPiece, kTPiece, ksteps_p, idesc_nn,
                              [](int ks) { return ks * 2048; }, [](int ks) { return ks * 2048; });
                issue_product(tmem + kD6, tmem + kD6 + 64, d_t3_k, d_m3_n, kTPiece, kMPiece, 4, idesc_kn,
                              [](int ks) { return ks * 32; }, [](int ks) { return ks * 2048; });
                umma_commit(bar_a);
            }
            __syncwarp();
        }
        if (nxt < num_items) load_split_s(params + (long long)nxt * per_item + kC * kC, sm, p_in, tid);
        MIXBT_TICK(9);
        mbar_wait(bar_a, phase_a & 1u);
        ++phase_a;
        tc_fence_after();
        MIXBT_TICK(10);
        if (nxt < num_items) {                               // X3 and M3 are dead: g_M and g_x have completed
            load_split_x(x + (long long)nxt * p_in * kC, sm, p_in, tid);
            load_split_m(params + (long long)nxt * per_item, sm, tid);
            issue_o1();                                      // (block-uniform branch: the barrier inside is safe)
        }
        {
            float f[16];
            if ((warp & 3) < 2) {                            // g_M rows c < 64
                load_acc16(tm_lane + kD5 + col0, tm_lane + kD5 + 64 + col0, f);
                float* gm = gpg + row * kC + col0;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<float4*>(gm + 4 * c) = make_float4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
            }
            load_acc16(tm_lane + kD6 + col0, tm_lane + kD6 + 64 + col0, f);
            if (live) {
                float* gx = grad_x + (long long)item * p_in * kC + row * kC + col0;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<float4*>(gx + 4 * c) = make_float4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
            }
        }
        MIXBT_TICK(11);
        // every thread's tensor-memory reads of this item are complete (tcgen05.wait::ld) before it reaches the next block
        // barrier, and the products that overwrite D2 .. D6 are issued behind one
    }
#ifdef RACF_MIXBT_PROFILE
    if (tid == 0 && blockIdx.x == 0)
        printf("mixbt cycles (CTA 0): top-sync %lld | mma1 %lld | ln1 %lld | mma2 %lld | ln2+bwd %lld | mma3 wait %lld | gS out %lld | mma4 wait %lld | "
               "ln1bwd %lld | S-split(next) %lld | mma56 wait %lld | gM,gx out + x,M-split %lld\n", prof[0], prof[1], prof[2], prof[3],
               prof[4], prof[5], prof[6], prof[7], prof[8], prof[9], prof[10], prof[11]);
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
}

}  // namespace mixbt
}  // namespace racf

// Launcher used by racf_adaptive_mixing_backward_variant (csrc/mixing_bwd.cu); arguments already validated, in_points <= 96.
int racf_mixbwd_tc_launch(const float* x, const float* params, const float* grad_out, int num_query_groups, int in_points, float eps,
                          float* grad_x, float* grad_params, int sms, cudaStream_t st) {
    using namespace racf::mixbt;
    const int smem = kSmemBytes + 1024;
    cudaError_t e = cudaFuncSetAttribute(adaptive_mixing_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    const unsigned grid = (unsigned)(num_query_groups < sms ? num_query_groups : sms);
    adaptive_mixing_bwd_tc_kernel<<<grid, kThreads, smem, st>>>(x, params, grad_out, grad_x, grad_params, num_query_groups, in_points, eps);
    return (int)cudaGetLastError();
}
