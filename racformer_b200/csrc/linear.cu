// fp32-grade Linear layers on the 5th-generation tensor cores (tcgen05 + TMEM + TMA) -- SURVEY.md 8f-4.
//
// AdaptiveMixing's two large Linear layers (models/racformer_transformer.py:560-566: parameter_generator
// [Q,256] x [65536,256]^T and out_proj [Q,32768] x [256,32768]^T, 45 GFLOP per decoder iteration) are fp32 SGEMMs in the
// reference; on B200 they run on the CUDA cores at ~48 TFLOP/s and are 40 % of the decoder step. Here they run on the
// tensor cores WITHOUT giving up fp32 accuracy:
//
//   * every fp32 operand is split exactly into three bf16 pieces, x = x0 + x1 + x2 (8 + 8 + 8 significand bits,
//     racf_split_bf16x3), so a product a*w = sum_ij a_i*w_j and every a_i*w_j is exact in the tensor core's fp32
//     datapath (8 x 8 bit significands);
//   * terms with i + j <= max_order are accumulated (max_order 4: all 9, nothing dropped; 2: 6 terms, dropped part
//     <= 3 * 2^-24 relative per product, i.e. below fp32's own rounding);
//   * the tensor cores truncate (round toward zero) once per MMA into the fp32 accumulator, a bias of ~0.5 ulp per
//     K = 16 step that grows linearly with the number of MMAs that hit ONE accumulator. So the large term a0*w0
//     gets its own TMEM accumulator (K/16 steps) and the 2^-8-times smaller cross terms share a second one; the two
//     are added in fp32 (round to nearest) in the epilogue, and the host keeps K per accumulator <= 512 by
//     splitting K across CTAs whose partial sums are reduced in fp32 by a second kernel (deterministic, no atomics).
//
// Kernel: one 128 x 128 output tile per CTA, 128 threads, two CTAs per SM (so one CTA's epilogue overlaps the other's
// MMAs). Warp 0 lane 0 = TMA producer (one 3-D box of {BK, 128 rows, 3 pieces} per operand and stage, 64/128-byte
// swizzle), warp 1 lane 0 = MMA issuer (tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 128, K = 16, operands from
// shared memory through UMMA descriptors, D in TMEM), mbarrier full/empty ring; then all four warps read their 32 TMEM
// lanes with tcgen05.ld, stage the tile in (now idle) pipeline shared memory and write it out in full 512-byte rows.
#include <cuda.h>            // CUtensorMap types (no libcuda link: the encoder is fetched with cudaGetDriverEntryPoint)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "linear_tiled.cuh"
#include "racformer_ops.h"

namespace racf {

constexpr int kLinBM = 128;
constexpr int kLinBN = 128;
constexpr int kLinThreads = 128;
constexpr int kLinTmemCols = 256;      // accumulator of a0*w0 in columns [0,128), of the cross terms in [128,256)
constexpr int kLinStgStride = 132;     // floats per staged output row (128 + 4: conflict-free 16-byte row-major stores)

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
// 1-D bulk copy (TMA engine, no tensor map): `bytes` contiguous bytes, 16-byte aligned on both sides
__device__ __forceinline__ void tma_load_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32; issued by ONE thread for the whole CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when all MMAs issued so far by this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread (thread = lane, register = column)
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major [128 rows][BK] bf16 tile whose rows are one swizzle span (64 or 128
// bytes) wide, as TMA writes it: 8-row groups are SBO = 8 * row bytes apart; LBO is unused for swizzled K-major.
template <int kRowBytes>
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    constexpr uint64_t layout = kRowBytes == 128 ? 2 : (kRowBytes == 64 ? 4 : 6);   // SWIZZLE_128B / 64B / 32B
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4)
         | ((uint64_t)((8 * kRowBytes) >> 4) << 32)
         | (1ull << 46)                                                               // descriptor version (sm_100)
         | (layout << 61);
}

// Instruction descriptor: D fp32, A and B bf16, both K-major, M = 128, N = 128, dense, no negate.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ bool elect_one_lane() {      // one lane of the (converged) warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive_plain(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// All MMAs of one K block of kBK: piece products a_i * w_j with i + j <= kMaxOrder, smallest terms first; the cross terms
// accumulate in TMEM columns [128,256), a0 * w0 in [0,128). a_desc / w_desc: descriptors of piece 0, K step 0 of the
// stage; the other tiles are constant offsets (address field in 16-byte units). first == 0: overwrite.
template <int kBK, int kPieceBytes, int kMaxOrder>
__device__ __forceinline__ void issue_kblock(uint32_t tmem, uint64_t a_desc, uint64_t w_desc, uint32_t idesc, uint32_t first) {
    uint32_t acc_cross = first;
#pragma unroll
    for (int ks = 0; ks < kBK / 16; ++ks) {
#pragma unroll
        for (int order = kMaxOrder; order >= 1; --order) {
#pragma unroll
            for (int pa = 0; pa <= 2; ++pa) {
                const int pw = order - pa;
                if (pw < 0 || pw > 2) continue;
                umma_bf16(tmem + kLinBN, a_desc + (uint64_t)((pa * kPieceBytes + ks * 32) >> 4),
                          w_desc + (uint64_t)((pw * kPieceBytes + ks * 32) >> 4), idesc, acc_cross);
                acc_cross = 1;
            }
        }
        umma_bf16(tmem, a_desc + (uint64_t)((ks * 32) >> 4), w_desc + (uint64_t)((ks * 32) >> 4), idesc, (ks > 0) ? 1u : first);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// GEMM kernel
// ---------------------------------------------------------------------------------------------------------------------
struct LinArgs {
    const float* bias;      // [N] or nullptr; added only when the kernel writes the final result (num_splits == 1)
    float* out;             // [M, N] (num_splits == 1) or workspace [num_splits, M, N]
    int M, N;
    int m_tiles, n_tiles;
    int num_kblocks;        // ceil(K / BK)
    int kblocks_per_split;
    int max_order;          // accumulate a_i * w_j for i + j <= max_order (4 = all nine terms)
    const uint8_t* a_tiled; // operands in the pre-tiled format of linear_tiled.cuh (kTiled kernels), else nullptr
    const uint8_t* w_tiled;
    // Several Linear layers that share the input (racf_linear_bf16x3_multi_forward): their weights are stacked along N,
    // each padded with zero rows to a multiple of 128, and every layer has its own dense output and bias.
    int num_segments;       // 0 = one layer (the fields above)
    int seg_tile_start[RACF_LINEAR_MAX_SEGMENTS + 1];
    int seg_n[RACF_LINEAR_MAX_SEGMENTS];
    float* seg_out[RACF_LINEAR_MAX_SEGMENTS];
    const float* seg_bias[RACF_LINEAR_MAX_SEGMENTS];
};

template <int kBK, int kStages, bool kTiled>
__global__ void __launch_bounds__(kLinThreads, 2)
linear_bf16x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const LinArgs args) {
    constexpr int kRowBytes = kBK * 2;
    constexpr int kPieceBytes = 128 * kRowBytes;          // one [128][BK] bf16 tile
    constexpr int kStageBytes = 6 * kPieceBytes;          // a0 a1 a2 w0 w1 w2
    static_assert(kStages * kStageBytes >= kLinBM * kLinStgStride * 4, "pipeline smem is reused to stage the output tile");

    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kStages + 1];
    __shared__ uint32_t tmem_base_slot;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // swizzled tiles need 1024-byte alignment
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int tile = blockIdx.x;
    const int m_tile = tile % args.m_tiles;
    const int n_tile = (tile / args.m_tiles) % args.n_tiles;
    const int split = tile / (args.m_tiles * args.n_tiles);
    const int kb_begin = split * args.kblocks_per_split;
    const int kb_end = min(kb_begin + args.kblocks_per_split, args.num_kblocks);
    const int num_kb = kb_end - kb_begin;
    const int m0 = m_tile * kLinBM, n0 = n_tile * kLinBN;

    auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
    auto empty_bar = [&](int s) { return smem_u32(&bars[kStages + s]); };
    const uint32_t tmem_full_bar = smem_u32(&bars[2 * kStages]);

    if (!kTiled && warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // one warp allocates (and later frees) the CTA's TMEM columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_slot)), "n"(kLinTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        for (int i = 0; i < num_kb; ++i) {
            const int s = i % kStages;
            mbar_wait(empty_bar(s), ((i / kStages) & 1) ^ 1);
            mbar_arrive_expect_tx(full_bar(s), kStageBytes);
            const uint32_t dst = smem_base + s * kStageBytes;
            if constexpr (kTiled) {     // one contiguous 24 KB block per operand and stage
                static_assert(!kTiled || (kBK == kTileK && 3 * kPieceBytes == kTileStageBytes), "tiled format is BK = 32");
                const long long kb = kb_begin + i;
                tma_load_bulk(dst, args.a_tiled + ((long long)m_tile * args.num_kblocks + kb) * kTileStageBytes,
                              kTileStageBytes, full_bar(s));
                tma_load_bulk(dst + 3 * kPieceBytes, args.w_tiled + ((long long)n_tile * args.num_kblocks + kb) * kTileStageBytes,
                              kTileStageBytes, full_bar(s));
            } else {
                const int k0 = (kb_begin + i) * kBK;
                tma_load_3d(dst, &map_a, k0, m0, 0, full_bar(s));
                tma_load_3d(dst + 3 * kPieceBytes, &map_w, k0, n0, 0, full_bar(s));
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // The whole warp walks the (warp-uniform) loop and one elected lane issues; the term loops are unrolled at compile
        // time for the usual orders and a tile's descriptor is the stage's base descriptor plus a constant. Issued from a
        // single divergent lane with run-time term loops, every tcgen05.mma cost ~190 cycles of descriptor arithmetic in the
        // uniform datapath inside a compiler-generated convergence loop (csrc/linear_wide.cu measured it): three MMAs' worth.
        constexpr uint32_t idesc = umma_idesc_bf16(kLinBM, kLinBN);
        const uint64_t desc0 = umma_desc<kRowBytes>(smem_base);
        for (int i = 0; i < num_kb; ++i) {
            const int s = i % kStages;
            mbar_wait(full_bar(s), (i / kStages) & 1);
            tcgen05_fence_after();
            if (elect_one_lane()) {
                const uint64_t a_desc = desc0 + (uint64_t)((s * kStageBytes) >> 4), w_desc = a_desc + (uint64_t)((3 * kPieceBytes) >> 4);
                const uint32_t first = i == 0 ? 0u : 1u;
                switch (args.max_order) {
                    case 2: issue_kblock<kBK, kPieceBytes, 2>(tmem_base, a_desc, w_desc, idesc, first); break;
                    case 4: issue_kblock<kBK, kPieceBytes, 4>(tmem_base, a_desc, w_desc, idesc, first); break;
                    case 0: issue_kblock<kBK, kPieceBytes, 0>(tmem_base, a_desc, w_desc, idesc, first); break;
                    case 1: issue_kblock<kBK, kPieceBytes, 1>(tmem_base, a_desc, w_desc, idesc, first); break;
                    default: issue_kblock<kBK, kPieceBytes, 3>(tmem_base, a_desc, w_desc, idesc, first); break;
                }
                umma_commit(empty_bar(s));     // the stage may be refilled once these MMAs have read it
                if (i == num_kb - 1) umma_commit(tmem_full_bar);        // accumulators complete
            }
            __syncwarp();
        }
        if (num_kb == 0 && lane == 0) mbar_arrive_plain(tmem_full_bar);   // nothing to wait for (cannot happen: no empty splits)
    }
    __syncwarp();

    // ===== epilogue: all four warps, warp w owns TMEM lanes / output rows [32w, 32w + 32) =====
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();

    float* stg = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)));
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    const bool has_cross = args.max_order >= 1 && num_kb > 0;
#pragma unroll 1
    for (int c = 0; c < kLinBN / 32; ++c) {
        uint32_t v[32], u[32];
        if (num_kb > 0) {
            tmem_ld32(lane_addr + c * 32, v);
            if (has_cross) tmem_ld32(lane_addr + kLinBN + c * 32, u);
            tmem_ld_wait();
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 o;
            if (num_kb > 0) {
                o.x = __uint_as_float(v[4 * j + 0]); o.y = __uint_as_float(v[4 * j + 1]);
                o.z = __uint_as_float(v[4 * j + 2]); o.w = __uint_as_float(v[4 * j + 3]);
                if (has_cross) {
                    o.x += __uint_as_float(u[4 * j + 0]); o.y += __uint_as_float(u[4 * j + 1]);
                    o.z += __uint_as_float(u[4 * j + 2]); o.w += __uint_as_float(u[4 * j + 3]);
                }
            } else {
                o = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            *reinterpret_cast<float4*>(stg + row * kLinStgStride + c * 32 + j * 4) = o;
        }
    }
    __syncwarp();     // each warp reads back only the 32 rows it staged itself

    const int num_splits = gridDim.x / (args.m_tiles * args.n_tiles);
    float* outp = args.out + (long long)split * args.M * args.N;
    const float* biasp = num_splits == 1 ? args.bias : nullptr;
    int N = args.N, gn = n0 + lane * 4;
    if (args.num_segments > 0) {     // which stacked layer does this column tile belong to?
        int sg = 0;
        while (sg + 1 < args.num_segments && n_tile >= args.seg_tile_start[sg + 1]) ++sg;
        outp = args.seg_out[sg];
        biasp = args.seg_bias[sg];
        N = args.seg_n[sg];
        gn = (n_tile - args.seg_tile_start[sg]) * kLinBN + lane * 4;
    }
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (biasp != nullptr) {
        if (gn + 0 < N) b4.x = __ldg(biasp + gn + 0);
        if (gn + 1 < N) b4.y = __ldg(biasp + gn + 1);
        if (gn + 2 < N) b4.z = __ldg(biasp + gn + 2);
        if (gn + 3 < N) b4.w = __ldg(biasp + gn + 3);
    }
    const bool vec_ok = (N & 3) == 0 && gn + 3 < N;
    for (int r = 0; r < 32; ++r) {
        const int gm = m0 + warp * 32 + r;
        if (gm >= args.M || gn >= N) break;
        float4 o = *reinterpret_cast<const float4*>(stg + (warp * 32 + r) * kLinStgStride + lane * 4);
        o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
        float* dst = outp + (long long)gm * N + gn;
        if (vec_ok) {
            *reinterpret_cast<float4*>(dst) = o;
        } else {
            if (gn + 0 < N) dst[0] = o.x;
            if (gn + 1 < N) dst[1] = o.y;
            if (gn + 2 < N) dst[2] = o.z;
            if (gn + 3 < N) dst[3] = o.w;
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kLinTmemCols) : "memory");
    }
}

// out[m, n] = bias[n] + sum_s ws[s, m, n]   (fp32, round to nearest, fixed order)
__global__ void __launch_bounds__(256)
linear_splitk_reduce_kernel(const float* __restrict__ ws, const float* __restrict__ bias, float* __restrict__ out,
                            long long mn, int n, int num_splits) {
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= mn) return;
    if (i4 + 3 < mn && (n & 3) == 0) {
        float4 acc = *reinterpret_cast<const float4*>(ws + i4);
        for (int s = 1; s < num_splits; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(ws + (long long)s * mn + i4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        if (bias != nullptr) {
            const int c = (int)(i4 % n);
            acc.x += __ldg(bias + c); acc.y += __ldg(bias + c + 1); acc.z += __ldg(bias + c + 2); acc.w += __ldg(bias + c + 3);
        }
        *reinterpret_cast<float4*>(out + i4) = acc;
    } else {
        for (long long i = i4; i < mn && i < i4 + 4; ++i) {
            float acc = ws[i];
            for (int s = 1; s < num_splits; ++s) acc += ws[(long long)s * mn + i];
            if (bias != nullptr) acc += __ldg(bias + (int)(i % n));
            out[i] = acc;
        }
    }
}

// x = p0 + p1 + p2 exactly (round-to-nearest bf16 at each step; the residuals are exact in fp32)
__device__ __forceinline__ void split3(float x, __nv_bfloat16& p0, __nv_bfloat16& p1, __nv_bfloat16& p2) {
    p0 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(p0);
    p1 = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(p1);
    p2 = __float2bfloat16_rn(r2);
}

__global__ void __launch_bounds__(256)
split_bf16x3_kernel(const float* __restrict__ x, long long n, __nv_bfloat16* __restrict__ out) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    if (i + 3 < n) {
        const float4 v = *reinterpret_cast<const float4*>(x + i);
        const float f[4] = {v.x, v.y, v.z, v.w};
        __align__(8) __nv_bfloat16 p[3][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) split3(f[j], p[0][j], p[1][j], p[2][j]);
#pragma unroll
        for (int k = 0; k < 3; ++k) *reinterpret_cast<uint2*>(out + k * n + i) = *reinterpret_cast<const uint2*>(p[k]);
    } else {
        for (long long j = i; j < n; ++j) split3(x[j], out[j], out[n + j], out[2 * n + j]);
    }
}

// x [rows][K] fp32 -> the pre-tiled pieces of linear_tiled.cuh; one thread per 16-byte chunk (8 k) of the padded domain
__global__ void __launch_bounds__(256)
split_bf16x3_tiled_kernel(const float* __restrict__ x, const float* __restrict__ addend, int addend_rows, int rows, int K,
                          int num_kblocks, long long num_chunks, __nv_bfloat16* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_chunks) return;
    const int chunks_per_row = num_kblocks * 4;
    const long long row = i / chunks_per_row;
    const int k0 = (int)(i - row * chunks_per_row) * 8;
    __align__(16) __nv_bfloat16 p[3][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float v = (row < rows && k0 + j < K) ? x[row * K + k0 + j] : 0.f;
        if (addend != nullptr && row < rows && k0 + j < K) v += __ldg(addend + (row % addend_rows) * K + k0 + j);
        split3(v, p[0][j], p[1][j], p[2][j]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        *reinterpret_cast<uint4*>(out + tiled_offset(row, k0, num_kblocks, k)) = *reinterpret_cast<const uint4*>(p[k]);
}

// (in[bt][c][s] + pos[c][s]) -> bf16 pieces out3[piece][bt * S + s][c]: the channel-first BEV maps become the K-major A
// operand of value_proj (models/bev_self_attention.py:162-174 does this with an add, a permute + copy and a GEMM whose
// bias is added in a fourth pass). 64 channels x 32 positions per CTA through a padded shared-memory tile.
__global__ void __launch_bounds__(256)
split_bf16x3_chw_to_hwc_kernel(const float* __restrict__ in, const float* __restrict__ pos, int C, int S,
                               long long piece_stride, int tiled_kblocks, __nv_bfloat16* __restrict__ out) {
    __shared__ float tile[64][33];
    const int s0 = blockIdx.x * 32, c0 = blockIdx.y * 64, bt = blockIdx.z;
    const int t = threadIdx.x;
    {
        const int sl = t & 31, s = s0 + sl;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int cl = (t >> 5) + 8 * i, c = c0 + cl;
            float v = 0.f;
            if (s < S && c < C) {
                v = in[((long long)bt * C + c) * S + s];
                if (pos != nullptr) v += __ldg(pos + (long long)c * S + s);
            }
            tile[cl][sl] = v;
        }
    }
    __syncthreads();
    const int sl = t >> 3, c8 = (t & 7) * 8, s = s0 + sl;
    const int c_end = tiled_kblocks > 0 ? tiled_kblocks * kTileK : C;      // the tiled format zero-fills its K tail
    if (s >= S || c0 + c8 >= c_end) return;                                // C % 8 == 0 is checked by the host
    __align__(16) __nv_bfloat16 p[3][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split3(tile[c8 + j][sl], p[0][j], p[1][j], p[2][j]);
    const long long row = (long long)bt * S + s;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        __nv_bfloat16* o = tiled_kblocks > 0 ? out + tiled_offset(row, c0 + c8, tiled_kblocks, k)
                                             : out + row * C + c0 + c8 + k * piece_stride;
        *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(p[k]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// [3 pieces][rows][K] bf16, box {BK, 128 rows, 3 pieces}; out-of-range rows / K are zero-filled by TMA
static int make_operand_map(CUtensorMap* map, const void* base, int rows, int K, int bk) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (enc == nullptr) return RACF_ERR_UNSUPPORTED;
    const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, 3};
    const cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)rows * (cuuint64_t)K * 2};
    const cuuint32_t box[3] = {(cuuint32_t)bk, 128, 3};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapSwizzle swz = bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : RACF_ERR_BAD_SHAPE;
}

template <int kBK, int kStages, bool kTiled>
static int launch_linear(const CUtensorMap& ma, const CUtensorMap& mw, const LinArgs& args, int num_splits, cudaStream_t st) {
    constexpr int smem = kStages * 6 * 128 * kBK * 2 + 1024;
    cudaError_t e = cudaFuncSetAttribute(linear_bf16x3_kernel<kBK, kStages, kTiled>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    const unsigned grid = (unsigned)(args.m_tiles * args.n_tiles * num_splits);
    linear_bf16x3_kernel<kBK, kStages, kTiled><<<grid, kLinThreads, smem, st>>>(ma, mw, args);
    return (int)cudaGetLastError();
}

}  // namespace racf

extern "C" int racf_split_bf16x3(const float* x, long long count, void* out3, racf_stream_t stream) {
    using namespace racf;
    if (!x || !out3) return RACF_ERR_NULL_POINTER;
    if (count <= 0) return RACF_ERR_BAD_SHAPE;
    if ((reinterpret_cast<uintptr_t>(x) & 15u) || (reinterpret_cast<uintptr_t>(out3) & 7u) || (count & 3)) return RACF_ERR_UNSUPPORTED;
    const long long threads = (count + 3) / 4;
    split_bf16x3_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, count, static_cast<__nv_bfloat16*>(out3));
    return (int)cudaGetLastError();
}

extern "C" long long racf_linear_tiled_bytes(long long rows, int K) {
    if (rows <= 0 || K <= 0) return 0;
    return ((rows + racf::kTileRows - 1) / racf::kTileRows) * (long long)((K + racf::kTileK - 1) / racf::kTileK) * racf::kTileStageBytes;
}

extern "C" int racf_split_bf16x3_tiled(const float* x, long long rows, int K, void* out, racf_stream_t stream) {
    return racf_split_bf16x3_tiled_add(x, rows, K, nullptr, 1, out, stream);
}

extern "C" int racf_split_bf16x3_tiled_add(const float* x, long long rows, int K, const float* addend, long long addend_rows,
                                           void* out, racf_stream_t stream) {
    using namespace racf;
    if (!x || !out) return RACF_ERR_NULL_POINTER;
    if (rows <= 0 || K <= 0 || addend_rows <= 0 || addend_rows > 0x7fffffffLL) return RACF_ERR_BAD_SHAPE;
    if (reinterpret_cast<uintptr_t>(out) & 15u) return RACF_ERR_UNSUPPORTED;
    const int num_kblocks = (K + kTileK - 1) / kTileK;
    const long long rows_pad = (rows + kTileRows - 1) / kTileRows * kTileRows;
    const long long chunks = rows_pad * num_kblocks * 4;
    if ((chunks + 255) / 256 > 0x7fffffffLL) return RACF_ERR_BAD_SHAPE;
    split_bf16x3_tiled_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, addend, (int)addend_rows, (int)rows, K, num_kblocks, chunks, static_cast<__nv_bfloat16*>(out));
    return (int)cudaGetLastError();
}

extern "C" int racf_split_bf16x3_chw_to_hwc(const float* in, const float* pos, int batch, int channels, int positions,
                                            int tiled, void* out3, racf_stream_t stream) {
    using namespace racf;
    if (!in || !out3) return RACF_ERR_NULL_POINTER;
    if (batch <= 0 || channels <= 0 || positions <= 0 || batch > 65535) return RACF_ERR_BAD_SHAPE;
    if ((channels & 7) != 0 || (reinterpret_cast<uintptr_t>(out3) & 15u)) return RACF_ERR_UNSUPPORTED;
    const dim3 grid((unsigned)((positions + 31) / 32), (unsigned)((channels + 63) / 64), (unsigned)batch);
    const int kblocks = tiled ? (channels + kTileK - 1) / kTileK : 0;
    const dim3 grid_t((unsigned)((positions + 31) / 32), (unsigned)((kblocks * kTileK + 63) / 64), (unsigned)batch);
    split_bf16x3_chw_to_hwc_kernel<<<tiled ? grid_t : grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, pos, channels, positions, (long long)batch * positions * channels, kblocks, static_cast<__nv_bfloat16*>(out3));
    return (int)cudaGetLastError();
}

extern "C" int racf_linear_bf16x3_plan(int M, int N, int K, int* split_k, long long* workspace_bytes) {
    if (!split_k || !workspace_bytes) return RACF_ERR_NULL_POINTER;
    if (M <= 0 || N <= 0 || K <= 0) return RACF_ERR_BAD_SHAPE;
    // at most 512 of K (32 MMA steps) per TMEM accumulator: bounds the tensor cores' truncation bias (see the header)
    const int s = (K + 511) / 512;
    *split_k = s;
    *workspace_bytes = s > 1 ? (long long)s * M * N * 4 : 0;
    return RACF_OK;
}

int racf_linear_wide_launch(const void* a_tiled, const void* w_tiled, const float* bias, float* out, int M, int N, int num_kblocks,
                            int kblocks_per_pass, int kblocks_per_split, int num_splits, int max_order, cudaStream_t st);   // csrc/linear_wide.cu

extern "C" int racf_linear_bf16x3_forward(const void* a3, const void* w3, const float* bias, int M, int N, int K,
                                          int max_order, int split_k, int variant, float* workspace, float* out,
                                          racf_stream_t stream) {
    using namespace racf;
    if (!a3 || !w3 || !out) return RACF_ERR_NULL_POINTER;
    if (M <= 0 || N <= 0 || K <= 0 || split_k <= 0) return RACF_ERR_BAD_SHAPE;
    if (variant < 0 || variant > 3 || max_order < 0 || max_order > 4) return RACF_ERR_UNSUPPORTED;
    if (variant < 2 && (K & 7) != 0) return RACF_ERR_UNSUPPORTED;                      // tensor maps: 16-byte global strides
    if ((reinterpret_cast<uintptr_t>(a3) | reinterpret_cast<uintptr_t>(w3) | reinterpret_cast<uintptr_t>(out)) & 15u)
        return RACF_ERR_UNSUPPORTED;
    if (split_k > 1 && (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 15u))) return RACF_ERR_NULL_POINTER;
    const int bk = variant == 1 ? 64 : 32;
    const int num_kblocks = (K + bk - 1) / bk;
    if (split_k > num_kblocks) split_k = num_kblocks;
    const int kbps = (num_kblocks + split_k - 1) / split_k;
    split_k = (num_kblocks + kbps - 1) / kbps;           // no empty splits

    CUtensorMap ma, mw;
    int rc = 0;
    if (variant == 3) {                  // 128 x 256 tiles, persistent CTAs (csrc/linear_wide.cu); tiled operands
        cudaStream_t st3 = static_cast<cudaStream_t>(stream);
        // split_k passes of kbps K blocks (<= 512 of K per tensor-memory accumulator). A work item takes several consecutive
        // passes and adds them in registers: as few work items as still fill the SMs once, and a workspace / reduction
        // that is as many times smaller (the caller's workspace is sized for split_k slabs; the first `splits` are used).
        int dev = 0, sms = 0;
        cudaError_t e3 = cudaGetDevice(&dev);
        if (e3 == cudaSuccess) e3 = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e3 != cudaSuccess) return (int)e3;
        const long long base_tiles = (long long)((M + 127) / 128) * (((N + 127) / 128 + 1) / 2);
        long long ppi = ((long long)split_k * base_tiles + sms - 1) / sms;          // passes per work item
        if (ppi < 1) ppi = 1;
        if (ppi > split_k) ppi = split_k;
        const int splits = (int)((split_k + ppi - 1) / ppi);
        rc = racf_linear_wide_launch(a3, w3, bias, splits > 1 ? workspace : out, M, N, num_kblocks, kbps, (int)ppi * kbps, splits,
                                     max_order, st3);
        if (rc != 0) return rc;
        if (splits > 1) {
            const long long mn = (long long)M * N;
            const long long threads = (mn + 3) / 4;
            linear_splitk_reduce_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st3>>>(workspace, bias, out, mn, N, splits);
            return (int)cudaGetLastError();
        }
        return RACF_OK;
    }
    if (variant < 2) {
        rc = make_operand_map(&ma, a3, M, K, bk);
        if (rc != 0) return rc;
        rc = make_operand_map(&mw, w3, N, K, bk);
        if (rc != 0) return rc;
    }

    LinArgs args;
    args.a_tiled = static_cast<const uint8_t*>(a3);
    args.w_tiled = static_cast<const uint8_t*>(w3);
    args.num_segments = 0;
    args.bias = bias;
    args.out = split_k > 1 ? workspace : out;
    args.M = M; args.N = N;
    args.m_tiles = (M + kLinBM - 1) / kLinBM;
    args.n_tiles = (N + kLinBN - 1) / kLinBN;
    args.num_kblocks = num_kblocks;
    args.kblocks_per_split = kbps;
    args.max_order = max_order;
    if ((long long)args.m_tiles * args.n_tiles * split_k > 0x7fffffffLL) return RACF_ERR_BAD_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = variant == 2 ? launch_linear<32, 2, true>(ma, mw, args, split_k, st)
       : variant == 1 ? launch_linear<64, 1, false>(ma, mw, args, split_k, st)
                      : launch_linear<32, 2, false>(ma, mw, args, split_k, st);
    if (rc != 0) return rc;
    if (split_k > 1) {
        const long long mn = (long long)M * N;
        const long long threads = (mn + 3) / 4;
        linear_splitk_reduce_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(workspace, bias, out, mn, N, split_k);
        return (int)cudaGetLastError();
    }
    return RACF_OK;
}

extern "C" int racf_linear_bf16x3_multi_forward(const void* a3, const void* w3, int M, int K, int num_segments,
                                                const int* seg_n, const float* const* seg_bias, float* const* seg_out,
                                                int max_order, int tiled, racf_stream_t stream) {
    using namespace racf;
    if (!a3 || !w3 || !seg_n || !seg_out) return RACF_ERR_NULL_POINTER;
    if (M <= 0 || K <= 0 || num_segments <= 0) return RACF_ERR_BAD_SHAPE;
    if (num_segments > RACF_LINEAR_MAX_SEGMENTS || (!tiled && (K & 7) != 0) || K > 512 || max_order < 0 || max_order > 4)
        return RACF_ERR_UNSUPPORTED;    // K <= 512: one accumulator pass, no K split
    if ((reinterpret_cast<uintptr_t>(a3) | reinterpret_cast<uintptr_t>(w3)) & 15u) return RACF_ERR_UNSUPPORTED;
    LinArgs args;
    args.a_tiled = static_cast<const uint8_t*>(a3);
    args.w_tiled = static_cast<const uint8_t*>(w3);
    args.bias = nullptr;
    args.out = nullptr;
    args.M = M;
    args.num_segments = num_segments;
    int tiles = 0;
    for (int i = 0; i < num_segments; ++i) {
        if (seg_n[i] <= 0) return RACF_ERR_BAD_SHAPE;
        if (!seg_out[i]) return RACF_ERR_NULL_POINTER;
        if ((reinterpret_cast<uintptr_t>(seg_out[i]) & 15u) && (seg_n[i] & 3) == 0) return RACF_ERR_UNSUPPORTED;
        args.seg_tile_start[i] = tiles;
        args.seg_n[i] = seg_n[i];
        args.seg_out[i] = seg_out[i];
        args.seg_bias[i] = seg_bias ? seg_bias[i] : nullptr;
        tiles += (seg_n[i] + kLinBN - 1) / kLinBN;
    }
    args.seg_tile_start[num_segments] = tiles;
    for (int i = num_segments; i < RACF_LINEAR_MAX_SEGMENTS; ++i) {
        args.seg_tile_start[i + 1] = tiles; args.seg_n[i] = 0; args.seg_out[i] = nullptr; args.seg_bias[i] = nullptr;
    }
    args.N = tiles * kLinBN;                  // rows of the stacked, padded weight
    args.m_tiles = (M + kLinBM - 1) / kLinBM;
    args.n_tiles = tiles;
    args.num_kblocks = (K + 31) / 32;
    args.kblocks_per_split = args.num_kblocks;
    args.max_order = max_order;
    CUtensorMap ma, mw;
    if (tiled) return launch_linear<32, 2, true>(ma, mw, args, 1, static_cast<cudaStream_t>(stream));
    int rc = make_operand_map(&ma, a3, M, K, 32);
    if (rc != 0) return rc;
    rc = make_operand_map(&mw, w3, args.N, K, 32);
    if (rc != 0) return rc;
    return launch_linear<32, 2, false>(ma, mw, args, 1, static_cast<cudaStream_t>(stream));
}
