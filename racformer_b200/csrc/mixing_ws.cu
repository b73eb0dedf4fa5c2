// AdaptiveMixing core on the tensor cores, warp-specialised -- SURVEY.md 8f-4, second version of csrc/mixing_tc.cu.
//
// Same contract (reference: AdaptiveMixing.inner_forward, models/racformer_transformer.py:592-604), per (query, group):
//     t = relu(LN(x @ M));   out = relu(LN(S @ t))          x [P_in, 64], M [64, 64], S [128, P_in] fp32
// and the same arithmetic as mixing_tc.cu: every fp32 operand is split exactly into three bf16 pieces on the fly, the six
// largest piece products are accumulated in fp32 in tensor memory (a0*b0 and the cross terms in separate accumulators),
// the layer norms run in fp32 on the CUDA cores straight from tensor memory.
//
// mixing_tc.cu walks the phases of one item with all 512 threads behind CTA-wide barriers (load+split, MMA, LayerNorm, ...),
// so the global-load latency, the tensor pipe and the CUDA cores never overlap (profiles/r01c_mixing_tc_ncu_summary.json:
// issue slots 31 %, tensor pipe 7 %). Here one persistent CTA per SM is four roles connected by mbarriers, with three
// items in flight:
//
//   warps 1-14   splitters  the raw fp32 item (x 24 KB, M 16 KB, S 48 KB at P_in = 96) streams through a ring of 4 KB shared-
//                           memory slots (cp.async.bulk); slot g belongs to warp g % 14, which fetches it, splits it exactly into
//                           bf16x3 and writes X3 / M3 / S3 in the swizzled UMMA operand layouts (M3 = M as stored: it is the
//                           MN-major B operand of product 1, so nothing is transposed)
//   warp 0       MMA        one lane issues product 1 of item n+1 (D1 = x @ M) and product 2 of item n (D2 = S @ t);
//                           tcgen05.commit releases the operand buffers and publishes the accumulators
//   warps 8-15   LayerNorm1 D1 -> registers -> LN + ReLU -> split -> T3 (B operand of product 2)
//   warps 16-23  LayerNorm2 D2 -> registers -> LN + ReLU -> global (fp32 rows, or bf16 pieces in out_proj's tiled format)
//
// D1 and D2 are double-buffered in tensor memory (8 x 64 = 512 columns), the operand tiles are single-buffered: product 1
// of item n+1 and LN2 of item n-1 run in the shadow of LN1(n) / product 2(n). Rows >= P_in of X3 and K tails are
// never initialised: an A row only feeds its own accumulator row, and those rows are masked out of LN1.
// t is stored N-major ([p][c'], 128-byte rows, the MN-major UMMA layout), so a LayerNorm thread -- which owns one row p of
// the accumulator -- writes 16-byte chunks instead of the 2-byte transposing stores of mixing_tc.cu.
// A ring slot is released only after its values have been USED: mbarrier.arrive does not wait for outstanding LDS, and
// the bulk copy that refills the slot is not ordered behind them (measured: corrupted items with an early release).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <type_traits>

#include "linear_tiled.cuh"
#include "racformer_ops.h"

namespace racf {
namespace mixws {

constexpr int kC = 64, kPout = 128, kMaxPin = 96;
#ifndef RACF_MIXWS_SPLIT_WARPS
#define RACF_MIXWS_SPLIT_WARPS 14
#endif
#ifndef RACF_MIXWS_MERGED_LN
#define RACF_MIXWS_MERGED_LN 1
#endif
constexpr bool kMergedLn = RACF_MIXWS_MERGED_LN != 0;         // one LayerNorm group does LN1(n) then LN2(n-1) instead of two groups
constexpr int kSplitWarp0 = 1, kSplitWarps = RACF_MIXWS_SPLIT_WARPS, kLnWarps = 8;   // kLnWarps per LN group
constexpr int kLn1Warp0 = kSplitWarp0 + kSplitWarps, kLn2Warp0 = kLn1Warp0 + (kMergedLn ? 0 : kLnWarps);
constexpr int kThreads = 32 * (kLn2Warp0 + kLnWarps);         // 736 (14 splitter warps, one LN group)
#ifndef RACF_MIXWS_SLOT_BYTES
#define RACF_MIXWS_SLOT_BYTES 4096
#endif
constexpr int kSlotBytes = RACF_MIXWS_SLOT_BYTES, kSlotChunks = kSlotBytes / 32;   // a chunk = 8 consecutive fp32 = one 16-byte bf16 chunk
constexpr int kCpt = kSlotChunks / 32;                        // chunks per thread and slot (a warp owns a whole slot)
static_assert(kSlotBytes == 2048 || kSlotBytes == 4096, "slots hold whole 8-row groups of M");
#ifndef RACF_MIXWS_RING
#define RACF_MIXWS_RING (57344 / RACF_MIXWS_SLOT_BYTES)
#endif
constexpr int kRing = RACF_MIXWS_RING;
// A ring slot must always be consumed by the same splitter warp: a warp that runs ahead of its neighbours would otherwise
// wait on a slot's full barrier two phases early, which the parity test cannot tell from "complete" (measured: deadlock).
static_assert(kRing % kSplitWarps == 0, "ring slots must map to a fixed splitter warp");
constexpr int kXPiece = kMaxPin * 128;                        // [96 rows p][64 c] bf16, 128-byte swizzle
constexpr int kMPiece = kC * 128;                             // [64 rows c'][64 c]
constexpr int kSPiece = 3 * 8192;                             // 3 K atoms of [128 rows o][32 p], 64-byte swizzle
constexpr int kTPiece = kMaxPin * 128;                        // [96 rows p][64 c'] (MN-major) or 3 atoms of [64 c'][32 p]
constexpr int kX3 = 0;
constexpr int kM3 = kX3 + 3 * kXPiece;
constexpr int kS3 = kM3 + 3 * kMPiece;
constexpr int kT3 = kS3 + 3 * kSPiece;
constexpr int kRaw = kT3 + 3 * kTPiece;
constexpr int kSmemBytes = kRaw + kRing * kSlotBytes;
constexpr int kTmemCols = 512;                                // D1[b]: b*128 (+64 cross); D2[b]: 256 + b*128 (+64 cross)
// Tuning aid (RACF_NVCC_DEFINES=-DRACF_MIXWS_EXP=n, results are then wrong): 1 = producer + splitters only (no MMA / LN, no
// waits on the operand buffers), 2 = as 1 without the split arithmetic and stores, 3 = as 1 without the stores only,
// 4 = full kernel without the proxy fences, 5 = full kernel issuing only the a0*b0 piece products.
#ifndef RACF_MIXWS_EXP
#define RACF_MIXWS_EXP 0
#endif
#define RACF_MIXWS_SPLIT_ONLY (RACF_MIXWS_EXP >= 1 && RACF_MIXWS_EXP <= 3)
constexpr bool kTMajorMN = true;                              // false: t^T K-major with 2-byte stores, as in mixing_tc.cu

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// No suspend-time hint on purpose: with a hint the wait compiles to a NANOSLEEP loop that sleeps the whole interval even
// when the barrier completes earlier (measured: 60 % of the warp samples in the sleep, kernel 1.5x slower).
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
#ifdef RACF_MIXWS_DEBUG            // debugging aid: who waits where when a barrier wait times out
__shared__ int s_progress[32];
__device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity, int tag) {
    const bool p0 = mbar_try_wait(bar, 0), p1 = mbar_try_wait(bar, 1);
    printf("mixws: barrier wait timed out (block %d, warp %d, wait site %d, smem 0x%x, parity %u, test0 %d test1 %d) progress:",
           (int)blockIdx.x, (int)(threadIdx.x >> 5), tag, bar, parity, (int)p0, (int)p1);
    for (int i = 0; i < 25; ++i) printf(" %d", s_progress[i]);
    printf("\n");
    __trap();
}
#define PROGRESS(v) do { if ((threadIdx.x & 31) == 0) s_progress[threadIdx.x >> 5] = (v); } while (0)
#else
__device__ __forceinline__ void mbar_timeout(uint32_t, uint32_t, int) { __trap(); }
#define PROGRESS(v) do { } while (0)
#endif
// Bounded (a bug must not hang the GPU). kSleepNs > 0: the warp sleeps between polls -- every failed poll is a shared-
// memory transaction, and the LayerNorm warps spend most of their time waiting; the ring hand-offs poll back to back.
template <int kSleepNs>
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, int tag) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) mbar_timeout(bar, parity, tag);
        if (kSleepNs > 0) __nanosleep(kSleepNs);
    }
}
#ifndef RACF_MIXWS_BACKOFF_NS
#define RACF_MIXWS_BACKOFF_NS 64
#endif
// wait sites 1-5: MMA issuer; 6 (raw_full), 7 (xm_free), 8 (s_free): splitters; 9-11: LayerNorm warps waiting for MMA results
#ifndef RACF_MIXWS_NOSLEEP_MAX_TAG
#define RACF_MIXWS_NOSLEEP_MAX_TAG 6
#endif
#define mbar_wait(bar, parity, tag) mbar_wait_t<((tag) <= RACF_MIXWS_NOSLEEP_MAX_TAG) ? 0 : RACF_MIXWS_BACKOFF_NS>(bar, parity, tag)
#ifdef RACF_MIXWS_PROFILE      // tuning aid: cycles each role of CTA 0 spends waiting on each barrier (RACF_NVCC_DEFINES=-DRACF_MIXWS_PROFILE)
#define MBAR_WAIT(bar, parity, slot) do { const long long t_ = clock64(); mbar_wait(bar, parity, slot); prof[slot] += clock64() - t_; } while (0)
#else
#define MBAR_WAIT(bar, parity, slot) mbar_wait(bar, parity, slot)
#endif
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// 32-byte global store (STG.256): half the LSU wavefronts of two 16-byte stores when every lane writes its own line
__device__ __forceinline__ void st_global_256(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
                 "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
// two float4 of one 32-byte chunk, bank-conflict free: lanes with bit 2 set read the upper half first (a plain 32-byte
// stride puts lanes l and l + 4 of a quarter-warp on the same banks)
__device__ __forceinline__ void lds_chunk(const float* chunk, int lane, float4& a, float4& b) {
    const int flip = (lane >> 2) & 1;
    const float4 u = *reinterpret_cast<const float4*>(chunk + 4 * flip);
    const float4 v = *reinterpret_cast<const float4*>(chunk + 4 * (flip ^ 1));
    a = flip ? v : u;
    b = flip ? u : v;
}
__device__ __forceinline__ void fence_async_smem() {
#if RACF_MIXWS_EXP != 4          // experiment 4: full kernel without the proxy fences (timing only, results unreliable)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptors (sm_100 UMMA, version 1). K-major tile whose rows are one swizzle span (128 / 64 bytes)
// wide: 8-row groups are SBO = 8 rows apart. MN-major 128-byte-swizzle tile of N = 64 bf16 (one span): k rows are 128 bytes
// apart, 8-row groups SBO = 1024 bytes apart, LBO (stride between 64-wide N blocks) is not used at N = 64.
template <int kRowBytes>
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) {
    constexpr uint64_t layout = kRowBytes == 128 ? 2 : 4;
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)((8 * kRowBytes) >> 4) << 32) | (1ull << 46) | (layout << 61);
}
__device__ __forceinline__ uint64_t desc_mnmajor_sw128(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

// exact three-way split of two floats: x = p0 + p1 + p2, bf16 pieces, round to nearest at each step (see mixing_tc.cu)
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
// eight consecutive fp32 -> one 16-byte chunk per piece at byte offset `off` of each piece tile (shared memory)
__device__ __forceinline__ void split_store8(const float (&f)[8], uint8_t* tile, int piece_bytes, int off) {
#if RACF_MIXWS_EXP == 2
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc ^= __float_as_uint(f[j]);
    if (acc == 0xdeadbeefu) *reinterpret_cast<uint32_t*>(tile + off) = acc;
#else
    uint32_t p[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t q[3];
        split3x2(f[2 * j], f[2 * j + 1], q);
        p[0][j] = q[0]; p[1][j] = q[1]; p[2][j] = q[2];
    }
#if RACF_MIXWS_EXP == 3
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) acc ^= p[k][0] ^ p[k][1] ^ p[k][2] ^ p[k][3];
    if (acc == 0xdeadbeefu) *reinterpret_cast<uint32_t*>(tile + off) = acc;
#else
#pragma unroll
    for (int k = 0; k < 3; ++k)
        *reinterpret_cast<uint4*>(tile + k * piece_bytes + off) = make_uint4(p[k][0], p[k][1], p[k][2], p[k][3]);
#endif
#endif
}

__device__ __forceinline__ bool elect_one() {      // one lane of the (converged) warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

struct Bars {                       // all mbarriers of the CTA
    uint64_t raw_full[kRing];
    uint64_t xm_ready, s_ready, t_ready, t_free;
    // "operand tile free" signals, one barrier per item modulo 4: a splitter warp does not own a slot of every kind in
    // every item, so it may wait for a completion that is not the next one -- a single barrier's parity cannot express that
    uint64_t xm_free[4], s_free[4];
    uint64_t d1_full[2], d1_free[2], d2_full[2], d2_free[2];
};

// sum over the eight warps of one LayerNorm group (fixed order -> deterministic): shuffles, one named barrier, 8 partials
__device__ __forceinline__ float ln_group_sum(int bar_id, float v, float* red, int lw, int lane, int& flip) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    float* r = red + flip * kLnWarps;
    if (lane == 0) r[lw] = v;
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(kLnWarps * 32) : "memory");
    const float4 a = *reinterpret_cast<const float4*>(r), b = *reinterpret_cast<const float4*>(r + 4);
    flip ^= 1;
    return ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
}

// accumulator slice of this thread (32 columns of one row): main + cross terms -> f, 16 columns at a time (register budget)
__device__ __forceinline__ void load_acc32(uint32_t taddr, float (&f)[32]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint32_t v[16], u[16];
        tmem_ld16(taddr + h * 16, v);
        tmem_ld16(taddr + 64 + h * 16, u);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) f[h * 16 + j] = __uint_as_float(v[j]) + __uint_as_float(u[j]);
    }
}
__device__ __forceinline__ float sum32(const float (&f)[32]) {       // four independent chains
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 4) { s0 += f[j]; s1 += f[j + 1]; s2 += f[j + 2]; s3 += f[j + 3]; }
    return (s0 + s1) + (s2 + s3);
}
__device__ __forceinline__ float sqdev32(const float (&f)[32], float mean) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float d0 = f[j] - mean, d1 = f[j + 1] - mean, d2 = f[j + 2] - mean, d3 = f[j + 3] - mean;
        s0 += d0 * d0; s1 += d1 * d1; s2 += d2 * d2; s3 += d3 * d3;
    }
    return (s0 + s1) + (s2 + s3);
}
// relu((f - mean) * rstd) of 8 consecutive columns -> one 16-byte chunk per piece
__device__ __forceinline__ void norm_split8(const float* f, float mean, float rstd, uint4 (&out)[3]) {
    uint32_t p[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t pc[3];
        split3x2(fmaxf((f[2 * j] - mean) * rstd, 0.f), fmaxf((f[2 * j + 1] - mean) * rstd, 0.f), pc);
        p[0][j] = pc[0]; p[1][j] = pc[1]; p[2][j] = pc[2];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = make_uint4(p[k][0], p[k][1], p[k][2], p[k][3]);
}

template <bool kSplitOut>
__global__ void __launch_bounds__(kThreads, 1)
adaptive_mixing_ws_kernel(const float* __restrict__ x, const float* __restrict__ params, float* __restrict__ out,
                          __nv_bfloat16* __restrict__ out3, int tiled_groups, int num_items, int p_in, float eps) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ Bars bars;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float red[2][2 * kLnWarps];

    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sm_addr = smem_u32(sm);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_local = (num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // items of this CTA (>= 1)
    const int x_bytes = p_in * kC * 4, s_bytes = kPout * p_in * 4;
    const long long per_item = kC * kC + kPout * p_in;
    // ring slots of one item, in stream order: x, M, S
    const int nx = (x_bytes + kSlotBytes - 1) / kSlotBytes, nm = kC * kC * 4 / kSlotBytes, ns = (s_bytes + kSlotBytes - 1) / kSlotBytes;
    const int spi = nx + nm + ns;

    if (tid == 0) {
        for (int s = 0; s < kRing; ++s) mbar_init(smem_u32(&bars.raw_full[s]), 1);
        mbar_init(smem_u32(&bars.xm_ready), nx + nm);
        mbar_init(smem_u32(&bars.s_ready), ns);
        for (int b = 0; b < 4; ++b) { mbar_init(smem_u32(&bars.xm_free[b]), 1); mbar_init(smem_u32(&bars.s_free[b]), 1); }
        mbar_init(smem_u32(&bars.t_ready), kLnWarps); mbar_init(smem_u32(&bars.t_free), 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bars.d1_full[b]), 1); mbar_init(smem_u32(&bars.d1_free[b]), kLnWarps);
            mbar_init(smem_u32(&bars.d2_full[b]), 1); mbar_init(smem_u32(&bars.d2_free[b]), kLnWarps);
        }
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
#ifdef RACF_MIXWS_PROFILE
    long long prof[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_start = clock64();
#endif

    if (warp == 0) {
        // ---------------------------------------------------------------- MMA issuer
        // The whole warp runs the (warp-uniform) event loop and one elected lane issues the MMAs: from a single divergent
        // lane the compiler wraps every tcgen05.mma in a convergence loop and rebuilds both descriptors in the uniform
        // datapath (~190 cycles per MMA, measured in csrc/linear_wide.cu; 60 MMAs per item). Descriptors are a base
        // descriptor plus compile-time constants (address field in 16-byte units).
        if (!RACF_MIXWS_SPLIT_ONLY) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            constexpr uint32_t idesc2 = idesc | (kTMajorMN ? (1u << 16) : 0u);
            constexpr uint32_t idesc1 = idesc | (1u << 16);        // product 1: B = M as stored (MN-major)
            constexpr int kOrder = RACF_MIXWS_EXP == 5 ? 0 : 2;
            const int ksteps2 = p_in >> 4;
            const uint64_t dx = desc_kmajor<128>(sm_addr + kX3), dm = desc_mnmajor_sw128(sm_addr + kM3);
            const uint64_t ds = desc_kmajor<64>(sm_addr + kS3);
            const uint64_t dt = kTMajorMN ? desc_mnmajor_sw128(sm_addr + kT3) : desc_kmajor<64>(sm_addr + kT3);
            auto product1 = [&](int n) {       // D1[b] = X3 @ M3^T   (one elected lane)
                const int b = n & 1;
                tc_fence_after();
                const uint32_t d_main = tmem + b * 128, d_cross = d_main + 64;
                uint32_t acc_cross = 0;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
                    for (int order = kOrder; order >= 1; --order)
#pragma unroll
                        for (int pa = 0; pa <= order; ++pa) {
                            umma_bf16(d_cross, dx + (uint64_t)((pa * kXPiece + ks * 32) >> 4),
                                      dm + (uint64_t)(((order - pa) * kMPiece + ks * 2048) >> 4), idesc1, acc_cross);
                            acc_cross = 1;
                        }
                    umma_bf16(d_main, dx + (uint64_t)((ks * 32) >> 4), dm + (uint64_t)((ks * 2048) >> 4), idesc1, ks > 0);
                }
                umma_commit(smem_u32(&bars.xm_free[n & 3]));
                umma_commit(smem_u32(&bars.d1_full[b]));
            };
            auto product2 = [&](int n) {       // D2[b] = S3 @ t
                const int b = n & 1;
                tc_fence_after();
                const uint32_t d_main = tmem + 256 + b * 128, d_cross = d_main + 64;
                uint32_t acc_cross = 0;
                for (int ks = 0; ks < ksteps2; ++ks) {
                    const uint64_t a_off = (uint64_t)(((ks >> 1) * 8192 + (ks & 1) * 32) >> 4);
                    const uint64_t b_off = (uint64_t)((kTMajorMN ? ks * 2048 : (ks >> 1) * 4096 + (ks & 1) * 32) >> 4);
#pragma unroll
                    for (int order = kOrder; order >= 1; --order)
#pragma unroll
                        for (int pa = 0; pa <= order; ++pa) {
                            umma_bf16(d_cross, ds + a_off + (uint64_t)((pa * kSPiece) >> 4), dt + b_off + (uint64_t)(((order - pa) * kTPiece) >> 4),
                                      idesc2, acc_cross);
                            acc_cross = 1;
                        }
                    umma_bf16(d_main, ds + a_off, dt + b_off, idesc2, ks > 0);
                }
                umma_commit(smem_u32(&bars.s_free[n & 3]));
                umma_commit(smem_u32(&bars.t_free));
                umma_commit(smem_u32(&bars.d2_full[b]));
            };
            // Issue whichever product is ready, product 2 first (it is on the critical path: S3 and T3 are single-buffered
            // and free only when it has completed). A fixed order -- product 1 of item n+1 before product 2 of item n --
            // parks product 2 behind the splitters' x / M slots of the NEXT item (measured: 176 -> 137 us).
            auto ready1 = [&](int n) {
                return mbar_try_wait(smem_u32(&bars.xm_ready), n & 1) && mbar_try_wait(smem_u32(&bars.d1_free[n & 1]), ((n >> 1) & 1) ^ 1);
            };
            auto ready2 = [&](int n) {
                return mbar_try_wait(smem_u32(&bars.s_ready), n & 1) && mbar_try_wait(smem_u32(&bars.t_ready), n & 1) &&
                       mbar_try_wait(smem_u32(&bars.d2_free[n & 1]), ((n >> 1) & 1) ^ 1);
            };
            int n1 = 0, n2 = 0;                     // next product 1 / product 2 to issue
            uint32_t idle = 0;
            while (n2 < n_local) {
                // the readiness tests are made warp-uniform: a barrier may complete between two lanes' polls
                const bool r2 = n2 < n1 && __all_sync(0xffffffffu, ready2(n2));
                const bool r1 = !r2 && n1 < n_local && n1 < n2 + 2 && __all_sync(0xffffffffu, ready1(n1));
                if (r2) {
                    PROGRESS(n2 * 2 + 1);
                    if (elect_one()) product2(n2);
                    __syncwarp();
                    ++n2; idle = 0;
                } else if (r1) {
                    PROGRESS(n1 * 2);
                    if (elect_one()) product1(n1);
                    __syncwarp();
                    ++n1; idle = 0;
                } else if (++idle > (1u << 24)) {
                    mbar_timeout(smem_u32(&bars.xm_ready), (uint32_t)(n1 & 1), 1);
                }
            }
        }
    } else if (warp < kLn1Warp0) {
        // ---------------------------------------------------------------- splitters: raw slot -> bf16x3 operand tiles
        // Slot g of the item stream (x, M, S slots of item 0, then item 1, ...) belongs to warp g % 7, which also FETCHES it:
        // a warp owns two 4 KB ring slots and refills one with its slot g + 14 (cp.async.bulk, lane 0) as soon as it has
        // used the values of slot g -- no producer warp, no "empty" barriers. A thread owns chunks lane + 32 i (i < 4).
        const int sw = warp - kSplitWarp0;
        const int cpr = p_in >> 3;                                // 8-float chunks per row of S
        const uint32_t cpr_magic = 65536u / (uint32_t)cpr + 1u;   // gi / cpr == (gi * magic) >> 16 for gi < 2048
        const int x_chunks = p_in * 8, s_chunks = kPout * cpr;
        const int total = n_local * spi;
        auto fetch = [&](int nf, int wf, uint32_t rs) {            // lane 0: slot wf of item nf -> ring slot rs
            const long long item = (long long)blockIdx.x + (long long)nf * gridDim.x;
            const char* src;
            int nb;
            if (wf < nx) {
                src = reinterpret_cast<const char*>(x + item * (long long)p_in * kC) + wf * kSlotBytes;
                nb = min(kSlotBytes, x_bytes - wf * kSlotBytes);
            } else if (wf < nx + nm) {
                src = reinterpret_cast<const char*>(params + item * per_item) + (wf - nx) * kSlotBytes;
                nb = kSlotBytes;
            } else {
                src = reinterpret_cast<const char*>(params + item * per_item) + kC * kC * 4 + (wf - nx - nm) * kSlotBytes;
                nb = min(kSlotBytes, s_bytes - (wf - nx - nm) * kSlotBytes);
            }
            mbar_arrive_expect_tx(smem_u32(&bars.raw_full[rs]), (uint32_t)nb);
            bulk_load(sm_addr + kRaw + rs * kSlotBytes, src, (uint32_t)nb, smem_u32(&bars.raw_full[rs]));
        };
        int n = 0, w = sw;                                         // slot g = n * spi + w: being consumed
        int nf = 0, wf = sw;                                       // the slot being fetched (kRing ahead once the ring is primed)
        while (wf >= spi) { wf -= spi; ++nf; }
        if (lane == 0 && sw < total) {
            fetch(nf, wf, (uint32_t)sw);
            for (int j = 1; j < kRing / kSplitWarps; ++j) {
                wf += kSplitWarps;
                while (wf >= spi) { wf -= spi; ++nf; }
                if (sw + j * kSplitWarps < total) fetch(nf, wf, (uint32_t)(sw + j * kSplitWarps));
            }
        }
        for (int g = sw; g < total; g += kSplitWarps, w += kSplitWarps) {
            while (w >= spi) { w -= spi; ++n; }
            const uint32_t rs = (uint32_t)g % kRing, rk = (uint32_t)g / kRing;
            PROGRESS(g);
            const float* raw = reinterpret_cast<const float*>(sm + kRaw + rs * kSlotBytes);
            uint64_t* ready;
#ifdef RACF_MIXWS_PROFILE
            long long t_slot;
#endif
            if (w < nx + nm) {
                if (!RACF_MIXWS_SPLIT_ONLY && n > 0) MBAR_WAIT(smem_u32(&bars.xm_free[(n - 1) & 3]), ((n - 1) >> 2) & 1, 7);          // product 1 of item n-1 has read X3 / M3
                MBAR_WAIT(smem_u32(&bars.raw_full[rs]), rk & 1, 6);
#ifdef RACF_MIXWS_PROFILE
                t_slot = clock64();
#endif
                ready = &bars.xm_ready;
                if (w < nx) {                                                  // 16 rows of x
#pragma unroll 1
                    for (int i0 = 0; i0 < kCpt; i0 += 2) {                     // two chunks at a time (register budget)
                        float4 a[2], b[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int c = lane + 32 * (i0 + i);
                            if (w * kSlotChunks + c < x_chunks) lds_chunk(raw + c * 8, lane, a[i], b[i]);
                        }
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int c = lane + 32 * (i0 + i);
                            if (w * kSlotChunks + c < x_chunks) {
                                const int p = w * (kSlotChunks / 8) + (c >> 3), ch = c & 7;
                                const float f[8] = {a[i].x, a[i].y, a[i].z, a[i].w, b[i].x, b[i].y, b[i].z, b[i].w};
                                split_store8(f, sm + kX3, kXPiece, p * 128 + ((ch ^ (p & 7)) << 4));
                            }
                        }
                    }
                } else {                                                       // 16 rows c of M, stored as they are: M3 [c][c'] is the
                    const int h = w - nx;                                      // MN-major B operand of product 1 (K = c, N = c')
#pragma unroll 1
                    for (int i0 = 0; i0 < kCpt; i0 += 2) {
                        float4 a[2], b[2];
#pragma unroll
                        for (int i = 0; i < 2; ++i) lds_chunk(raw + (lane + 32 * (i0 + i)) * 8, lane, a[i], b[i]);
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int c = lane + 32 * (i0 + i);
                            const int r = h * (kSlotChunks / 8) + (c >> 3), ch = c & 7;
                            const float f[8] = {a[i].x, a[i].y, a[i].z, a[i].w, b[i].x, b[i].y, b[i].z, b[i].w};
                            split_store8(f, sm + kM3, kMPiece, r * 128 + ((ch ^ (r & 7)) << 4));
                        }
                    }
                }
            } else {
                if (!RACF_MIXWS_SPLIT_ONLY && n > 0) MBAR_WAIT(smem_u32(&bars.s_free[(n - 1) & 3]), ((n - 1) >> 2) & 1, 8);           // product 2 of item n-1 has read S3
                MBAR_WAIT(smem_u32(&bars.raw_full[rs]), rk & 1, 6);
#ifdef RACF_MIXWS_PROFILE
                t_slot = clock64();
#endif
                ready = &bars.s_ready;
                const int g0 = (w - nx - nm) * kSlotChunks;
#pragma unroll 1
                for (int i0 = 0; i0 < kCpt; i0 += 2) {
                    float4 a[2], b[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int c = lane + 32 * (i0 + i);
                        if (g0 + c < s_chunks) lds_chunk(raw + c * 8, lane, a[i], b[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int gi = g0 + lane + 32 * (i0 + i);
                        if (gi < s_chunks) {
                            const int o = (int)(((uint32_t)gi * cpr_magic) >> 16), j = gi - o * cpr;
                            const float f[8] = {a[i].x, a[i].y, a[i].z, a[i].w, b[i].x, b[i].y, b[i].z, b[i].w};
                            split_store8(f, sm + kS3, kSPiece, (j >> 2) * 8192 + o * 64 + (((j & 3) ^ ((o >> 1) & 3)) << 4));
                        }
                    }
                }
            }
#ifdef RACF_MIXWS_PROFILE
            const long long t_split = clock64();
#endif
            fence_async_smem();                                    // generic-proxy writes -> visible to the MMA (async proxy)
            __syncwarp();
#ifdef RACF_MIXWS_PROFILE
            const long long t_fence = clock64();
            prof[0] += t_split - t_slot;     // LDS + split + STS issue
            prof[1] += t_fence - t_split;    // proxy fence + warp sync
#endif
            if (lane == 0) {
                mbar_arrive(smem_u32(ready));
                if (g + kRing < total) {                           // the slot's values have been used (see header): refill it
                    wf += kSplitWarps;
                    while (wf >= spi) { wf -= spi; ++nf; }
                    fetch(nf, wf, rs);
                }
            }
#ifdef RACF_MIXWS_PROFILE
            prof[2] += clock64() - t_fence;  // arrive + refill issue
#endif
        }
    } else if (RACF_MIXWS_SPLIT_ONLY) {
        // experiment: no LayerNorm roles
    } else {
        // ---------------------------------------------------------------- LayerNorm warps
        // LN1: D1 -> t (B operand of product 2); LN2: D2 -> output. Two groups of eight warps, or (kMergedLn) one group
        // that does LN1(n) and then LN2(n-1).
        const int lw = (warp - kLn1Warp0) & (kLnWarps - 1), quarter = warp & 3, half = lw >> 2;   // TMEM lanes of a warp: 32 * (warp % 4)
        const int row = quarter * 32 + lane, col0 = half * 32;
        const uint32_t tm_row = tmem + ((uint32_t)(quarter * 32) << 16) + col0;
        const bool live = row < p_in;
        const int bar2 = kMergedLn ? 1 : 2;
        float* red2 = kMergedLn ? red[0] : red[1];
        int flip = 0;
        auto ln1_item = [&](int n) {
            const uint32_t tm_lane = tm_row;
            const float cnt = (float)(p_in * kC);
            const int b = n & 1;
            float f[32];
            PROGRESS(n * 2);
            MBAR_WAIT(smem_u32(&bars.d1_full[b]), (n >> 1) & 1, 9);
            tc_fence_after();
            load_acc32(tm_lane + b * 128, f);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars.d1_free[b]));
            const float s = sum32(f);
            const float mean = ln_group_sum(1, live ? s : 0.f, red[0], lw, lane, flip) / cnt;
            const float q = sqdev32(f, mean);
            const float rstd = rsqrtf(ln_group_sum(1, live ? q : 0.f, red[0], lw, lane, flip) / cnt + eps);
            PROGRESS(n * 2 + 1);
            MBAR_WAIT(smem_u32(&bars.t_free), (n & 1) ^ 1, 10);    // product 2 of item n-1 has read T3
            if (live) {
                if constexpr (kTMajorMN) {
                    // t [p][c']: row p = 128 bytes per piece, this thread's 32 columns = chunks 4*half .. 4*half+3
                    uint8_t* base = sm + kT3 + row * 128;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 o3[3];
                        norm_split8(f + c * 8, mean, rstd, o3);
#pragma unroll
                        for (int k = 0; k < 3; ++k)
                            *reinterpret_cast<uint4*>(base + k * kTPiece + (((half * 4 + c) ^ (row & 7)) << 4)) = o3[k];
                    }
                } else {
                    // element (c', p) of t^T: atom p / 32, row c' (64 bytes), 16-byte chunk (p % 32) / 8, 2-byte slot p % 8
                    const int atom = row >> 5, kk = row & 31;
                    uint8_t* base = sm + kT3 + atom * 4096 + (kk & 7) * 2;
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        uint32_t pc[3];
                        split3x2(fmaxf((f[j] - mean) * rstd, 0.f), fmaxf((f[j + 1] - mean) * rstd, 0.f), pc);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int cp = half * 32 + j + h;
                            uint8_t* e = base + cp * 64 + (((kk >> 3) ^ ((cp >> 1) & 3)) << 4);
#pragma unroll
                            for (int k = 0; k < 3; ++k)
                                *reinterpret_cast<uint16_t*>(e + k * kTPiece) = (uint16_t)(h ? (pc[k] >> 16) : (pc[k] & 0xffffu));
                        }
                    }
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars.t_ready));
        };
        auto ln2_item = [&](int n) {
            const uint32_t tm_lane = tm_row + 256;
            const float cnt = (float)(kPout * kC);
            const int b = n & 1;
            float f[32];
            PROGRESS(n);
            MBAR_WAIT(smem_u32(&bars.d2_full[b]), (n >> 1) & 1, 11);
            tc_fence_after();
            load_acc32(tm_lane + b * 128, f);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars.d2_free[b]));
            const float mean = ln_group_sum(bar2, sum32(f), red2, lw, lane, flip) / cnt;
            const float rstd = rsqrtf(ln_group_sum(bar2, sqdev32(f, mean), red2, lw, lane, flip) / cnt + eps);
            const long long item = (long long)blockIdx.x + (long long)n * gridDim.x;
            if constexpr (kSplitOut) {
                // A operand of out_proj, tiled format (linear_tiled.cuh): matrix row = query, k = group * 8192 + o * 64 + c'.
                // This thread's 32 values are one whole 64-byte row (32-wide K block o * 2 + half) of each piece tile.
                const long long qi = item / tiled_groups;
                const int g = (int)(item - qi * tiled_groups);
                const int sw = (int)((qi & 127) >> 1) & 3;
                const long long kblocks = (long long)tiled_groups * (kPout * kC / kTileK);
                const long long kb = (long long)g * (kPout * kC / kTileK) + row * 2 + half;
                __nv_bfloat16* dst = out3 + ((qi >> 7) * kblocks + kb) * 3 * (long long)kTilePieceElems + (qi & 127) * kTileK;
                // the row's four chunks sit at positions c ^ sw (sw is the same for the whole item): two 32-byte stores per piece
                auto emit = [&](auto swc) {
                    constexpr int kSw = decltype(swc)::value;
#pragma unroll
                    for (int j = 0; j < 4; j += 2) {
                        uint4 lo[3], hi[3];
                        norm_split8(f + ((j ^ kSw) * 8), mean, rstd, lo);
                        norm_split8(f + (((j + 1) ^ kSw) * 8), mean, rstd, hi);
#pragma unroll
                        for (int k = 0; k < 3; ++k) st_global_256(dst + k * (long long)kTilePieceElems + j * 8, lo[k], hi[k]);
                    }
                };
                switch (sw) {
                    case 0: emit(std::integral_constant<int, 0>{}); break;
                    case 1: emit(std::integral_constant<int, 1>{}); break;
                    case 2: emit(std::integral_constant<int, 2>{}); break;
                    default: emit(std::integral_constant<int, 3>{}); break;
                }
            } else {
                float* og = out + item * (long long)(kPout * kC) + row * kC + col0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float r[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[j] = fmaxf((f[c * 8 + j] - mean) * rstd, 0.f);
                    st_global_256(og + c * 8, make_uint4(__float_as_uint(r[0]), __float_as_uint(r[1]), __float_as_uint(r[2]), __float_as_uint(r[3])),
                                  make_uint4(__float_as_uint(r[4]), __float_as_uint(r[5]), __float_as_uint(r[6]), __float_as_uint(r[7])));
                }
            }
        };
        if (kMergedLn) {
            for (int n = 0; n < n_local; ++n) {
                ln1_item(n);
                if (n > 0) ln2_item(n - 1);
            }
            ln2_item(n_local - 1);
        } else if (warp < kLn2Warp0) {
            for (int n = 0; n < n_local; ++n) ln1_item(n);
        } else {
            for (int n = 0; n < n_local; ++n) ln2_item(n);
        }
    }
#ifdef RACF_MIXWS_PROFILE
    if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == kSplitWarp0 || warp == kLn1Warp0 + 2 || warp == kLn2Warp0 + 3))
        printf("mixws CTA0 warp %2d total %lld | [splitter: work fence refill | mma: - xm_ready d1_free] %lld %lld %lld s_ready %lld t_ready %lld d2_free %lld | "
               "raw_full %lld xm_free %lld s_free %lld | d1_full %lld t_free %lld d2_full %lld\n", warp, clock64() - t_start, prof[0],
               prof[1], prof[2], prof[3], prof[4], prof[5], prof[6], prof[7], prof[8], prof[9], prof[10], prof[11]);
#endif
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
}

}  // namespace mixws
}  // namespace racf

// Launcher used by racf_adaptive_mixing_tc_forward (csrc/mixing_tc.cu) for in_points <= 96; arguments already validated.
int racf_mixws_launch(const float* x, const float* params, int num_query_groups, int in_points, float eps, float* out,
                      void* out3, int tiled_groups, int sms, cudaStream_t st) {
    using namespace racf::mixws;
    const int smem = kSmemBytes + 1024;
    const unsigned grid = (unsigned)(num_query_groups < sms ? num_query_groups : sms);
    cudaError_t e;
    if (out3) {
        e = cudaFuncSetAttribute(adaptive_mixing_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        adaptive_mixing_ws_kernel<true><<<grid, kThreads, smem, st>>>(x, params, nullptr, static_cast<__nv_bfloat16*>(out3),
                                                                      tiled_groups, num_query_groups, in_points, eps);
    } else {
        e = cudaFuncSetAttribute(adaptive_mixing_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        adaptive_mixing_ws_kernel<false><<<grid, kThreads, smem, st>>>(x, params, out, nullptr, 0, num_query_groups, in_points, eps);
    }
    return (int)cudaGetLastError();
}
