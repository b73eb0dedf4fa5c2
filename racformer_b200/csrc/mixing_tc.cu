// AdaptiveMixing core on the 5th-generation tensor cores -- SURVEY.md 8f-4 ("fused grouped GEMM + LayerNorm").
//
// Same contract as csrc/mixing.cu (reference: AdaptiveMixing.inner_forward, models/racformer_transformer.py:592-604), per
// (query, group):   t = relu(LN(x @ M));   out = relu(LN(S @ t))
// x [P_in, 64], M [64, 64], S [128, P_in] fp32. Here both products run as tcgen05.mma on bf16 pieces produced ON THE FLY:
// every fp32 value is split exactly into three bf16 pieces while it is copied from global to shared memory, the six
// largest piece products are accumulated in fp32 in tensor memory (large term and cross terms in separate accumulators,
// see csrc/linear.cu for why), so the products are fp32-grade; the layer norms stay in fp32 on the CUDA cores, fed
// straight from tensor memory. One persistent CTA per SM walks over the (query, group) items:
//
//   P0  global fp32 -> split -> shared, in the layouts the MMA wants (manual 128-/64-byte swizzle):
//         X3 [128 rows p][64 c]  A of product 1 (rows >= P_in stay zero)       M3 [64 rows c'][64 c]  B of product 1 (= M^T)
//         S3 [128 rows o][P_in]  A of product 2, 32-wide K atoms
//   P1  one thread issues product 1: D1[p, c'] (M = 128, N = 64, K = 64; 24 MMAs)
//   P2  D1 -> registers (tcgen05.ld), LayerNorm over P_in x 64, ReLU, split -> T3 [64 rows c'][P_in] (= t^T, B of product 2)
//   P3  product 2: D2[o, c'] (M = 128, N = 64, K = P_in; 36 MMAs at P_in = 96)
//   P4  D2 -> registers, LayerNorm over 128 x 64, ReLU -> fp32 out, or bf16 pieces in the tiled format of out_proj's A operand
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#ifdef RACF_MIXTC_PROFILE
#include <cstdio>
#endif

#include "linear_tiled.cuh"
#include "racformer_ops.h"

namespace racf {
namespace mixtc {

constexpr int kC = 64, kPout = 128, kThreads = 512;   // 16 warps: 4 per TMEM lane quarter, 16 accumulator columns each
constexpr int kX3 = 0;                          // 3 x 16 KB: [128][64] bf16, 128-byte rows, 128-byte swizzle
constexpr int kM3 = kX3 + 3 * 16384;            // 3 x  8 KB: [64][64]
constexpr int kS3 = kM3 + 3 * 8192;             // 3 x 4 atoms x 8 KB: [128][32] per atom, 64-byte rows, 64-byte swizzle
constexpr int kT3 = kS3 + 3 * 4 * 8192;         // 3 x 4 atoms x 4 KB: [64][32] per atom
constexpr int kSmemBytes = kT3 + 3 * 4 * 4096;  // 216 KB for P_in up to 128
constexpr int kTmemCols = 256;                  // D1 main / cross: [0,64) [64,128); D2 main / cross: [128,192) [192,256)

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
        if (clock64() - t0 > 4000000000LL) __trap();
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
// K-major bf16 tile, rows one swizzle span wide (128 or 64 bytes); 8-row groups SBO = 8 rows apart
template <int kRowBytes>
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr) {
    constexpr uint64_t layout = kRowBytes == 128 ? 2 : 4;
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((8 * kRowBytes) >> 4) << 32) | (1ull << 46) | (layout << 61);
}
// Exact three-way split of two floats at once: x = p0 + p1 + p2 with bf16 pieces, round to nearest at each step (the
// residuals are exact in fp32). q[k] holds piece k of (a, b) as a bf16x2 (a in the low half). The packed conversion
// (F2FP) is used on purpose: single cvt.rn.bf16.f32 compiles to F2F, which runs at a quarter of the rate.
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
__device__ __forceinline__ void split_store8(const float (&f)[8], uint8_t* tile, int piece_bytes, int off) {
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
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += red[w];
    return t;
}

// x [P_in][64] -> X3 (A of product 1), M [64 c][64 c'] -> M3 = M^T (B of product 1). All global loads are issued before
// the first use, so their latency is paid once (fixed trip counts for P_in <= 128, predicated).
__device__ __forceinline__ void load_split_xm(const float* __restrict__ xg, const float* __restrict__ mg, uint8_t* sm, int p_in, int tid) {
    constexpr int kXIt = 128 * 8 / kThreads, kMIt = kC * 8 / kThreads;
    float4 xa[kXIt][2];
    float mv[kMIt][8];
#pragma unroll
    for (int it = 0; it < kXIt; ++it) {                         // x: row p, 16-byte chunk ch (8 channels)
        const int i = tid + it * kThreads;
        if (i < p_in * 8) {
            xa[it][0] = __ldg(reinterpret_cast<const float4*>(xg + i * 8));
            xa[it][1] = __ldg(reinterpret_cast<const float4*>(xg + i * 8 + 4));
        }
    }
#pragma unroll
    for (int it = 0; it < kMIt; ++it) {                         // M^T: row c' (lanes -> coalesced), chunk of 8 c
        const int i = tid + it * kThreads, cp = i & 63, ch = i >> 6;
#pragma unroll
        for (int j = 0; j < 8; ++j) mv[it][j] = __ldg(mg + (ch * 8 + j) * kC + cp);
    }
#pragma unroll
    for (int it = 0; it < kXIt; ++it) {
        const int i = tid + it * kThreads;
        if (i < p_in * 8) {
            const int p = i >> 3, ch = i & 7;
            const float f[8] = {xa[it][0].x, xa[it][0].y, xa[it][0].z, xa[it][0].w, xa[it][1].x, xa[it][1].y, xa[it][1].z, xa[it][1].w};
            split_store8(f, sm + kX3, 16384, p * 128 + ((ch ^ (p & 7)) << 4));
        }
    }
#pragma unroll
    for (int it = 0; it < kMIt; ++it) {
        const int i = tid + it * kThreads, cp = i & 63, ch = i >> 6;
        split_store8(mv[it], sm + kM3, 8192, cp * 128 + ((ch ^ (cp & 7)) << 4));
    }
}

// S [128 o][P_in] -> S3 (A of product 2): 32-wide K atoms of [128 rows][64 B], 64-byte swizzle
__device__ __forceinline__ void load_split_s(const float* __restrict__ sg, uint8_t* sm, int p_in, int s3_piece, int tid) {
    const int s_chunks = p_in >> 3;
    constexpr int kSIt = kPout * 16 / kThreads;
    float4 sa[kSIt][2];
#pragma unroll
    for (int it = 0; it < kSIt; ++it) {
        const int i = tid + it * kThreads;
        if (i < kPout * s_chunks) {
            sa[it][0] = __ldg(reinterpret_cast<const float4*>(sg + i * 8));
            sa[it][1] = __ldg(reinterpret_cast<const float4*>(sg + i * 8 + 4));
        }
    }
#pragma unroll
    for (int it = 0; it < kSIt; ++it) {
        const int i = tid + it * kThreads;
        if (i < kPout * s_chunks) {
            const int o = i / s_chunks, j = i - o * s_chunks;
            const float f[8] = {sa[it][0].x, sa[it][0].y, sa[it][0].z, sa[it][0].w, sa[it][1].x, sa[it][1].y, sa[it][1].z, sa[it][1].w};
            split_store8(f, sm + kS3, s3_piece, (j >> 2) * 8192 + o * 64 + (((j & 3) ^ ((o >> 1) & 3)) << 4));
        }
    }
}

#ifdef RACF_MIXTC_PROFILE      // tuning aid: per-phase cycle counts of CTA 0 (build with RACF_NVCC_DEFINES=-DRACF_MIXTC_PROFILE)
#define MIXTC_TICK(slot) do { if (tid == 0) { const long long now_ = clock64(); prof[slot] += now_ - tprev; tprev = now_; } } while (0)
#else
#define MIXTC_TICK(slot) do { } while (0)
#endif

template <bool kSplitOut>
__global__ void __launch_bounds__(kThreads, 1)
adaptive_mixing_tc_kernel(const float* __restrict__ x, const float* __restrict__ params, float* __restrict__ out,
                          __nv_bfloat16* __restrict__ out3, int tiled_groups, int num_items, int p_in, float eps) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ float red[kThreads / 32];

    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_addr = smem_u32(sm);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int natoms = (p_in + 31) >> 5;                     // 32-wide K atoms of product 2
    const int s3_piece = natoms * 8192, t3_piece = natoms * 4096;
    const uint32_t bar_addr = smem_u32(&bar);

    // zero everything once: rows / K tails that are never written must read as 0
    for (int i = tid; i < kSmemBytes / 16; i += kThreads) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(bar_addr, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    // this thread's slice of an accumulator tile: TMEM lane quarter (warp % 4), 16 of the 64 columns (warp / 4)
    const int row = (warp & 3) * 32 + lane, col0 = (warp >> 2) * 16;
    const uint32_t tm_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16) + col0;
    const int m_elems = kC * kC, s_elems = kPout * p_in;
    const long long per_item = m_elems + s_elems;

#ifdef RACF_MIXTC_PROFILE
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
    // Software pipeline over this CTA's items: x / M of item i+1 are split while product 2 of item i runs (X3 / M3 are free
    // once product 1 has completed), S of item i while product 1 of item i runs (S3 is free once product 2 of item i-1 has).
    if ((int)blockIdx.x < num_items)
        load_split_xm(x + (long long)blockIdx.x * p_in * kC, params + (long long)blockIdx.x * per_item, sm, p_in, tid);

    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int nxt = item + gridDim.x;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        MIXTC_TICK(0);

        // ---- product 1: D1[p, c'] = x @ M -----------------------------------------------------------------------
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t acc_cross = 0;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                for (int order = 2; order >= 1; --order)
                    for (int pa = 0; pa <= order; ++pa) {
                        umma_bf16(tmem + 64, umma_desc<128>(sm_addr + kX3 + pa * 16384 + ks * 32),
                                  umma_desc<128>(sm_addr + kM3 + (order - pa) * 8192 + ks * 32), idesc, acc_cross);
                        acc_cross = 1;
                    }
                umma_bf16(tmem, umma_desc<128>(sm_addr + kX3 + ks * 32), umma_desc<128>(sm_addr + kM3 + ks * 32), idesc, ks > 0);
            }
            umma_commit(bar_addr);
        }
        if (nxt < num_items) {      // pull the next item's inputs into L2 well before they are needed
            const char* nx = reinterpret_cast<const char*>(x + (long long)nxt * p_in * kC);
            const char* np = reinterpret_cast<const char*>(params + (long long)nxt * per_item);
            for (int i = tid * 128; i < p_in * kC * 4; i += kThreads * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + i));
            for (int i = tid * 128; i < (int)per_item * 4; i += kThreads * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(np + i));
        }
        load_split_s(params + (long long)item * per_item + m_elems, sm, p_in, s3_piece, tid);   // overlaps product 1
        mbar_wait(bar_addr, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        MIXTC_TICK(1);

        // ---- LayerNorm + ReLU of D1 (rows < P_in), split, store t^T as the B operand of product 2 -----------------
        {
            uint32_t v[16], u[16];
            tmem_ld16(tm_lane, v);
            tmem_ld16(tm_lane + 64, u);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float f[16];
            float s = 0.f;
            const bool live = row < p_in;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                f[j] = __uint_as_float(v[j]) + __uint_as_float(u[j]);
                s += live ? f[j] : 0.f;
            }
            const float n = (float)(p_in * kC);
            const float mean = block_sum(s, red) / n;
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float d = f[j] - mean;
                q += live ? d * d : 0.f;
            }
            const float rstd = rsqrtf(block_sum(q, red) / n + eps);
            if (live) {
                // element (c', p) of t^T: atom p / 32, row c', 16-byte chunk (p % 32) / 8, 2-byte slot p % 8
                const int atom = row >> 5, kk = row & 31;
                uint8_t* base = sm + kT3 + atom * 4096 + (kk & 7) * 2;
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    uint32_t pc[3];
                    split3x2(fmaxf((f[j] - mean) * rstd, 0.f), fmaxf((f[j + 1] - mean) * rstd, 0.f), pc);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int cp = col0 + j + h;
                        uint8_t* e = base + cp * 64 + (((kk >> 3) ^ ((cp >> 1) & 3)) << 4);
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            *reinterpret_cast<uint16_t*>(e + k * t3_piece) = (uint16_t)(h ? (pc[k] >> 16) : (pc[k] & 0xffffu));
                    }
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        MIXTC_TICK(2);

        // ---- product 2: D2[o, c'] = S @ t -----------------------------------------------------------------------
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t acc_cross = 0;
            const int ksteps = p_in >> 4;
            for (int ks = 0; ks < ksteps; ++ks) {
                const uint32_t a_off = (ks >> 1) * 8192 + (ks & 1) * 32, b_off = (ks >> 1) * 4096 + (ks & 1) * 32;
                for (int order = 2; order >= 1; --order)
                    for (int pa = 0; pa <= order; ++pa) {
                        umma_bf16(tmem + 192, umma_desc<64>(sm_addr + kS3 + pa * s3_piece + a_off),
                                  umma_desc<64>(sm_addr + kT3 + (order - pa) * t3_piece + b_off), idesc, acc_cross);
                        acc_cross = 1;
                    }
                umma_bf16(tmem + 128, umma_desc<64>(sm_addr + kS3 + a_off), umma_desc<64>(sm_addr + kT3 + b_off), idesc, ks > 0);
            }
            umma_commit(bar_addr);
        }
        if (nxt < num_items)        // overlaps product 2: X3 / M3 are no longer read
            load_split_xm(x + (long long)nxt * p_in * kC, params + (long long)nxt * per_item, sm, p_in, tid);
        mbar_wait(bar_addr, 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        MIXTC_TICK(3);

        // ---- LayerNorm + ReLU of D2 -> global ----------------------------------------------------------------------
        {
            uint32_t v[16], u[16];
            tmem_ld16(tm_lane + 128, v);
            tmem_ld16(tm_lane + 192, u);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float f[16];
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                f[j] = __uint_as_float(v[j]) + __uint_as_float(u[j]);
                s += f[j];
            }
            const float n = (float)(kPout * kC);
            const float mean = block_sum(s, red) / n;
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float d = f[j] - mean;
                q += d * d;
            }
            const float rstd = rsqrtf(block_sum(q, red) / n + eps);
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf((f[j] - mean) * rstd, 0.f);
            // Stage the tile in the S3 region (dead once product 2 has completed) and write it out in full rows: a thread
            // owns half a row of the accumulator, but neighbouring bytes of the destination belong to other threads.
            uint8_t* stg = sm + kS3;
            if constexpr (kSplitOut) {
                // A operand of out_proj, tiled format: row = query, k = group * 8192 + o * 64 + c'. This thread's 16 values
                // are half of one 64-byte row of a piece tile; chunks go to their swizzled position, which is the same for the
                // whole item (it depends on the query only). Staging rows: [piece][o][half] 64 B, o-stride 144 B (bank spread).
                const long long qi = item / tiled_groups;
                const int sw = (int)((qi & 127) >> 1) & 3;
                const int c_base = (col0 & 31) >> 3;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t p[3][4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint32_t pc[3];
                        split3x2(f[c * 8 + 2 * j], f[c * 8 + 2 * j + 1], pc);
                        p[0][j] = pc[0]; p[1][j] = pc[1]; p[2][j] = pc[2];
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        *reinterpret_cast<uint4*>(stg + k * (kPout * 144) + row * 144 + (col0 >> 5) * 64 + (((c_base + c) ^ sw) << 4)) =
                            make_uint4(p[k][0], p[k][1], p[k][2], p[k][3]);
                }
                __syncthreads();
                const int g = (int)(item - qi * tiled_groups);
                const int kblocks = tiled_groups * (kPout * kC / kTileK);
                // 3 pieces x 128 rows x 2 halves x 4 chunks = 3072 16-byte chunks; 4 consecutive lanes write one 64-byte row
                for (int i = tid; i < 3 * kPout * 8; i += kThreads) {
                    const int pos = i & 3, h = (i >> 2) & 1, o = (i >> 3) & 127, k = i >> 10;
                    const uint4 v4 = *reinterpret_cast<const uint4*>(stg + k * (kPout * 144) + o * 144 + h * 64 + (pos << 4));
                    const long long kb = (long long)g * (kPout * kC / kTileK) + o * 2 + h;      // 32-wide K block of this row
                    __nv_bfloat16* dst = out3 + (((qi >> 7) * kblocks + kb) * 3 + k) * (long long)kTilePieceElems
                                       + (qi & 127) * kTileK + pos * 8;
                    *reinterpret_cast<uint4*>(dst) = v4;
                }
            } else {
                // fp32 rows of 256 B, staged with a 272-byte stride
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    *reinterpret_cast<float4*>(stg + row * 272 + col0 * 4 + c * 16) =
                        make_float4(f[c * 4], f[c * 4 + 1], f[c * 4 + 2], f[c * 4 + 3]);
                __syncthreads();
                float* og = out + (long long)item * (kPout * kC);
                for (int i = tid; i < kPout * 16; i += kThreads) {
                    const int o = i >> 4, c = i & 15;
                    *reinterpret_cast<float4*>(og + o * kC + c * 4) = *reinterpret_cast<const float4*>(stg + o * 272 + c * 16);
                }
            }
            if ((p_in & 31) != 0) {      // the staging area overlaps K tails of S3 that must read as zero: restore them
                __syncthreads();
                for (int i = tid; i < (kT3 - kS3) / 16; i += kThreads) reinterpret_cast<uint4*>(stg)[i] = make_uint4(0, 0, 0, 0);
            }
        }
        MIXTC_TICK(4);
        // the loop-top barrier orders these reads of the staging area / TMEM before the next item's writes
    }
#ifdef RACF_MIXTC_PROFILE
    if (tid == 0 && blockIdx.x == 0)
        printf("mixtc cycles (CTA 0): sync %lld mma1|S-split %lld ln1 %lld mma2|xM-split %lld ln2+out %lld\n", prof[0], prof[1],
               prof[2], prof[3], prof[4]);
#endif
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
}

}  // namespace mixtc
}  // namespace racf

// Which kernel variant 0 selects for in_points <= 96 (tools/mixing_bench.py measures both).
constexpr bool kDefaultWarpSpecialised = true;     // 170 / 176 us vs 179 / 198 us (fp32 / tiled output) at the f8 shapes

int racf_mixws_launch(const float* x, const float* params, int num_query_groups, int in_points, float eps, float* out,
                      void* out3, int tiled_groups, int sms, cudaStream_t st);   // csrc/mixing_ws.cu

// out (fp32 [QG, 128, 64]) or out3 (tiled pieces, tiled_groups = n_groups) -- exactly one of them non-NULL.
// variant: 0 = the warp-specialised kernel (csrc/mixing_ws.cu) where it exists (in_points <= 96), else the phase-serial
// kernel of this file; 1 = phase-serial; 2 = warp-specialised (RACF_ERR_UNSUPPORTED if in_points > 96).
extern "C" int racf_adaptive_mixing_tc_forward_variant(const float* x, const float* params, int num_query_groups, int in_points,
                                                       int out_points, int channels, float eps, float* out, void* out3,
                                                       int tiled_groups, int variant, racf_stream_t stream) {
    using namespace racf::mixtc;
    if (!x || !params || (!out && !out3) || (out && out3)) return RACF_ERR_NULL_POINTER;
    if (num_query_groups <= 0) return RACF_ERR_BAD_SHAPE;
    if (channels != kC || out_points != kPout || in_points < 16 || in_points > 128 || (in_points & 15) != 0)
        return RACF_ERR_UNSUPPORTED;
    if (variant < 0 || variant > 2 || (variant == 2 && in_points > 96)) return RACF_ERR_UNSUPPORTED;
    if (out3 && (tiled_groups <= 0 || num_query_groups % tiled_groups != 0)) return RACF_ERR_BAD_SHAPE;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(out) |
         reinterpret_cast<uintptr_t>(out3)) & 15u)
        return RACF_ERR_UNSUPPORTED;
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (variant == 2 || (variant == 0 && kDefaultWarpSpecialised && in_points <= 96))
        return racf_mixws_launch(x, params, num_query_groups, in_points, eps, out, out3, tiled_groups, sms, st);
    const int smem = kSmemBytes + 1024;
    const unsigned grid = (unsigned)(num_query_groups < sms ? num_query_groups : sms);
    if (out3) {
        e = cudaFuncSetAttribute(adaptive_mixing_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        adaptive_mixing_tc_kernel<true><<<grid, kThreads, smem, st>>>(x, params, nullptr, static_cast<__nv_bfloat16*>(out3),
                                                                      tiled_groups, num_query_groups, in_points, eps);
    } else {
        e = cudaFuncSetAttribute(adaptive_mixing_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        adaptive_mixing_tc_kernel<false><<<grid, kThreads, smem, st>>>(x, params, out, nullptr, 0, num_query_groups, in_points, eps);
    }
    return (int)cudaGetLastError();
}

extern "C" int racf_adaptive_mixing_tc_forward(const float* x, const float* params, int num_query_groups, int in_points,
                                               int out_points, int channels, float eps, float* out, void* out3,
                                               int tiled_groups, racf_stream_t stream) {
    return racf_adaptive_mixing_tc_forward_variant(x, params, num_query_groups, in_points, out_points, channels, eps, out, out3,
                                                   tiled_groups, 0, stream);
}
