// fp32-grade Linear on the tensor cores, wide-tile persistent version -- SURVEY.md 8f-4, second kernel of csrc/linear.cu.
//
// Same arithmetic and operand format as linear_bf16x3_kernel<32, 2, true> (csrc/linear.cu: bf16x3 operand split, six piece
// products, a0*w0 and the cross terms in separate TMEM accumulators, operands pre-tiled as in linear_tiled.cuh). What
// changes is the tile and the schedule. linear_bf16x3_kernel is bound by the L2 -> SM operand traffic, not by the tensor
// pipe (profiles/r01c_linear_ncu_summary.json: 1.61 GB of bulk-copy loads per parameter_generator launch at 7.7 TB/s, tensor
// pipe 26 %; with one piece product instead of six it still takes 141 of 185 us): a 128 x 128 tile reads 48 KB of operands
// per 32-wide K block for 6.3 MFLOP (131 flop/byte). Here
//   * a tile is 128 x 256: the A block is read once for two 128-row blocks of W -- 72 KB per K block for 12.6 MFLOP
//     (175 flop/byte, 25 % fewer operand bytes per launch). Each piece product is ONE N = 256 MMA (the producer lays the
//     two 128-row blocks of a W piece next to each other), accumulating into TMEM columns [0,256) (a0*w0) and [256,512)
//     (cross terms): all 512 columns, so one CTA per SM. The MMAs read 12 KB of shared memory per 256 columns instead of
//     16 KB -- the tensor pipe of these small-K-block MMAs is fed at the shared-memory bandwidth;
//   * the CTA is persistent (tile = blockIdx.x + i * gridDim.x, m fastest so that concurrent CTAs share W blocks in L2)
//     and warp-specialised: warp 0 = bulk-copy producer, warp 1 = MMA issuer, warps 2-17 = epilogue. The producer runs
//     ahead into the NEXT tile while the epilogue (warps 2-17) drains tensor memory, so the operand stream -- the binding resource --
//     does not stop at tile boundaries (linear_bf16x3_kernel relies on a second CTA per SM for that);
//   * the epilogue first drains tensor memory into registers (128 values per thread) and hands it back to the MMA warp,
//     then writes the tile through a small staging block per warp (8 rows x 32 columns) in 128-byte row segments.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>

#include "linear_tiled.cuh"
#include "racformer_ops.h"

namespace racf {
namespace linwide {

constexpr int kBM = 128, kStages = 2;                            // tile 128 x 256 (two 128-row blocks of W)
constexpr int kPieceBytes = kTilePieceElems * 2;                 // 8 KB: [128 rows][32 k] bf16, 64-byte rows, 64-byte swizzle
constexpr int kABytes = 3 * kPieceBytes, kWHalfBytes = 3 * kPieceBytes;
constexpr int kStageBytes = kABytes + 2 * kWHalfBytes;           // 72 KB
constexpr int kEpiWarps = 16, kThreads = 32 * (2 + kEpiWarps);   // 576: an epilogue thread owns 64 columns of one row
constexpr int kStgRows = 32, kStgStride = 36;                    // epilogue staging: 32 rows x 32 columns per warp (+4: conflict-free)
constexpr int kStgBytes = kStgRows * kStgStride * 4;
constexpr int kSmemBytes = kStages * kStageBytes + kEpiWarps * kStgBytes;
constexpr int kTmemCols = 512;

struct Args {
    const uint8_t* a_tiled;   // [m_tiles][num_kblocks][3][128 x 32] bf16
    const uint8_t* w_tiled;   // [n_tiles128][num_kblocks][3][128 x 32] bf16
    const float* bias;        // [N] or nullptr; added only when num_splits == 1
    float* out;               // [M, N] (num_splits == 1) or workspace [num_splits, M, N]
    int M, N;
    int m_tiles, n_tiles128, n_pairs;
    int num_kblocks;
    int kblocks_per_pass;     // K blocks accumulated in tensor memory before the epilogue takes the partial sum (<= 16: K <= 512
                              // per accumulator, see csrc/linear.cu on the truncating accumulation)
    int kblocks_per_split;    // K blocks of one work item = passes_per_split * kblocks_per_pass; its passes are added in registers
    int num_splits;
    int max_order;
    int num_tiles;            // m_tiles * n_pairs * num_splits
};

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
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {   // bounded: a bug must not hang the GPU
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 26)) __trap();
}
#ifdef RACF_LINW_PROFILE          // tuning aid: cycles each role of CTA 0 waits on each barrier
#define MBAR_WAIT(bar, parity, slot) do { const long long t_ = clock64(); mbar_wait(bar, parity); prof[slot] += clock64() - t_; } while (0)
#else
#define MBAR_WAIT(bar, parity, slot) mbar_wait(bar, parity)
#endif
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major [128 rows][32 k] bf16 tile, 64-byte rows, 64-byte swizzle: 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t umma_desc64(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}

__device__ __forceinline__ bool elect_one() {      // one lane of the (converged) warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// All MMAs of one 32-wide K block: piece products a_i * w_j with i + j <= kMaxOrder, smallest terms first, each as one
// N = 128 MMA per 128-row half of W. a_desc / w_desc: descriptors of piece 0, K step 0 of the stage's A / W block; the other
// tiles are constant offsets (address field in 16-byte units). first == 0: the tile's first K block (overwrite).
template <int kMaxOrder>
__device__ __forceinline__ void issue_kblock(uint32_t tmem, uint64_t a_desc, uint64_t w_desc, uint32_t idesc, uint32_t first) {
    uint32_t acc_cross = first;
#pragma unroll
    for (int ks = 0; ks < kTileK / 16; ++ks) {
#pragma unroll
        for (int order = kMaxOrder; order >= 1; --order) {
#pragma unroll
            for (int pa = 0; pa <= 2; ++pa) {
                const int pw = order - pa;
                if (pw < 0 || pw > 2) continue;
                umma_bf16(tmem + 256, a_desc + (uint64_t)((pa * kPieceBytes + ks * 32) >> 4),
                          w_desc + (uint64_t)((pw * 2 * kPieceBytes + ks * 32) >> 4), idesc, acc_cross);
                acc_cross = 1;
            }
        }
        umma_bf16(tmem, a_desc + (uint64_t)((ks * 32) >> 4), w_desc + (uint64_t)((ks * 32) >> 4), idesc, (ks > 0) ? 1u : first);
    }
}

__global__ void __launch_bounds__(kThreads, 1) linear_bf16x3_wide_kernel(const Args args) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * kStages + 2];
    __shared__ uint32_t tmem_slot;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // swizzled tiles need 1024-byte alignment
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
    auto empty_bar = [&](int s) { return smem_u32(&bars[kStages + s]); };
    const uint32_t tmem_full = smem_u32(&bars[2 * kStages]), tmem_empty = smem_u32(&bars[2 * kStages + 1]);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, kEpiWarps);
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
#ifdef RACF_LINW_PROFILE
    long long prof[4] = {0, 0, 0, 0};
    const long long t_start = clock64();
#endif

    // tile -> (m_tile, n_pair, split): m fastest
#define LINW_DECODE(tile)                                                                                   \
    const int m_tile = (tile) % args.m_tiles, rest_ = (tile) / args.m_tiles;                                \
    const int n_pair = rest_ % args.n_pairs, split = rest_ / args.n_pairs;                                  \
    const int kb_begin = split * args.kblocks_per_split;                                                    \
    const int num_kb = min(kb_begin + args.kblocks_per_split, args.num_kblocks) - kb_begin;

    if (warp == 0) {
        // ===== bulk-copy producer: runs ahead of the MMA warp across tile boundaries =====
        if (lane == 0) {
            uint32_t it = 0;                       // K blocks issued so far (ring position)
            for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
                LINW_DECODE(tile)
                const bool two = 2 * n_pair + 1 < args.n_tiles128;
                const uint8_t* a_src = args.a_tiled + ((long long)m_tile * args.num_kblocks + kb_begin) * kTileStageBytes;
                const uint8_t* w_src = args.w_tiled + ((long long)(2 * n_pair) * args.num_kblocks + kb_begin) * kTileStageBytes;
                const long long w_half = (long long)args.num_kblocks * kTileStageBytes;
                for (int i = 0; i < num_kb; ++i, ++it) {
                    const int s = it % kStages;
                    MBAR_WAIT(empty_bar(s), ((it / kStages) & 1) ^ 1, 0);
                    mbar_arrive_expect_tx(full_bar(s), kABytes + (two ? 2 : 1) * kWHalfBytes);
                    const uint32_t dst = smem_base + s * kStageBytes;
                    bulk_load(dst, a_src + (long long)i * kTileStageBytes, kABytes, full_bar(s));
                    // W block in shared memory: [piece][256 rows] -- the two 128-row blocks of a piece are adjacent, so that one
                    // N = 256 MMA reads them through one descriptor (six 8 KB copies instead of two 24 KB ones)
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        bulk_load(dst + kABytes + p * 2 * kPieceBytes, w_src + (long long)i * kTileStageBytes + p * kPieceBytes, kPieceBytes, full_bar(s));
                        if (two)
                            bulk_load(dst + kABytes + p * 2 * kPieceBytes + kPieceBytes,
                                      w_src + w_half + (long long)i * kTileStageBytes + p * kPieceBytes, kPieceBytes, full_bar(s));
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // The whole warp runs the (warp-uniform) loops and one elected lane issues: with a single divergent lane the
        // compiler wraps every tcgen05.mma in a convergence loop and rebuilds both descriptors in the uniform datapath
        // (~30 dependent instructions, measured ~190 cycles per MMA -- three times the MMA itself; the issuing thread, not
        // the operand stream, was the limit). Here the term loops are unrolled at compile time and a descriptor is the
        // stage's base descriptor plus a constant.
        constexpr uint32_t idesc128 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
        constexpr uint32_t idesc256 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
        const uint64_t desc0 = umma_desc64(smem_base);
        uint32_t it = 0, pass_iter = 0;
        for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
            LINW_DECODE(tile)
            const bool two = 2 * n_pair + 1 < args.n_tiles128;
            for (int p0 = 0; p0 < num_kb; p0 += args.kblocks_per_pass, ++pass_iter) {
                const int pass_kb = min(args.kblocks_per_pass, num_kb - p0);
                MBAR_WAIT(tmem_empty, (pass_iter & 1) ^ 1, 1);     // the epilogue has read the previous pass's accumulators
                tc_fence_after();
                for (int i = 0; i < pass_kb; ++i, ++it) {
                    const int s = it % kStages;
                    MBAR_WAIT(full_bar(s), (it / kStages) & 1, 2);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t a_desc = desc0 + (uint64_t)((s * kStageBytes) >> 4), w_desc = a_desc + (kABytes >> 4);
                        const uint32_t first = i == 0 ? 0u : 1u;
                        const uint32_t idesc = two ? idesc256 : idesc128;     // N = 256: both 128-row blocks of W in one MMA
                        if (args.max_order == 2) issue_kblock<2>(tmem, a_desc, w_desc, idesc, first);
                        else if (args.max_order == 4) issue_kblock<4>(tmem, a_desc, w_desc, idesc, first);
                        else if (args.max_order == 0) issue_kblock<0>(tmem, a_desc, w_desc, idesc, first);
                        else if (args.max_order == 1) issue_kblock<1>(tmem, a_desc, w_desc, idesc, first);
                        else issue_kblock<3>(tmem, a_desc, w_desc, idesc, first);
                        umma_commit(empty_bar(s));     // the stage may be refilled once these MMAs have read it
                        if (i == pass_kb - 1) umma_commit(tmem_full);   // accumulators of this pass complete
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===== epilogue: warp e owns TMEM lanes 32 * (warp % 4) and 64 of the tile's 256 columns =====
        const int e = warp - 2, quarter = warp & 3, cq = e >> 2, half = cq >> 1;     // cq: column quarter, half: 128-row block of W
        float* stg = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + kStages * kStageBytes + e * kStgBytes);
        const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16) + cq * 64;
        const bool has_cross = args.max_order >= 1;
        uint32_t pass_iter = 0;
        for (int tile = blockIdx.x; tile < args.num_tiles; tile += gridDim.x) {
            LINW_DECODE(tile)
            const int n_tile = 2 * n_pair + half;
            const bool active = n_tile < args.n_tiles128;
            float* outp = args.out + (long long)split * args.M * args.N;
            const float* biasp = args.num_splits == 1 ? args.bias : nullptr;
            const int m0 = m_tile * kBM + quarter * 32;
            // Drain tensor memory into registers (64 values per thread) and hand it back to the MMA warp BEFORE the stores:
            // the accumulators are single-buffered (all 512 columns), so whatever the epilogue does while it still holds
            // them is serial with the next MMAs. The passes of a work item (K > 512 per split) are added here in fp32,
            // in order -- fewer, longer work items and a K-split workspace that is passes_per_split times smaller.
            float f[2][32];
            for (int p0 = 0; p0 < num_kb; p0 += args.kblocks_per_pass, ++pass_iter) {
                MBAR_WAIT(tmem_full, pass_iter & 1, 3);
                tc_fence_after();
#ifdef RACF_LINW_PROFILE
                const long long t_ready = clock64();
#endif
                if (active) {
                    const bool first_pass = p0 == 0;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {                       // 8 columns at a time (register budget: 64 live sums)
                        uint32_t v[8], u[8];
                        tmem_ld8(lane_addr + c * 8, v);
                        if (has_cross) tmem_ld8(lane_addr + 256 + c * 8, u);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float x = has_cross ? __uint_as_float(v[j]) + __uint_as_float(u[j]) : __uint_as_float(v[j]);
                            float& acc = f[c >> 2][(c & 3) * 8 + j];
                            acc = first_pass ? x : acc + x;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty);
#ifdef RACF_LINW_PROFILE
                prof[0] += clock64() - t_ready;      // epilogue: tensor memory -> registers
#endif
            }
#ifdef RACF_LINW_PROFILE
            const long long t_drained = clock64();
#endif
            // Stores: a 32 x 32 chunk at a time goes through this warp's staging block, so that a store instruction writes
            // four 128-byte row segments (straight from registers every lane would write its own row,
            // 256 KB apart at N = 65536: measured 1300 cycles per store instruction, the epilogue became the bottleneck).
            if (active) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int gn = n_tile * 128 + (cq & 1) * 64 + c * 32 + (lane & 7) * 4;
                    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (biasp != nullptr) {
                        if (gn + 0 < args.N) b4.x = __ldg(biasp + gn + 0);
                        if (gn + 1 < args.N) b4.y = __ldg(biasp + gn + 1);
                        if (gn + 2 < args.N) b4.z = __ldg(biasp + gn + 2);
                        if (gn + 3 < args.N) b4.w = __ldg(biasp + gn + 3);
                    }
                    const bool vec_ok = (args.N & 3) == 0 && gn + 3 < args.N;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(stg + lane * kStgStride + j * 4) =
                            make_float4(f[c][4 * j], f[c][4 * j + 1], f[c][4 * j + 2], f[c][4 * j + 3]);
                    __syncwarp();
#pragma unroll 2
                    for (int r0 = 0; r0 < 32; r0 += 4) {             // 8 lanes x 16 bytes = one 128-byte row segment, 4 rows per store
                        const int r = r0 + (lane >> 3), gm = m0 + r;
                        if (gm < args.M && gn < args.N) {
                            float4 o = *reinterpret_cast<const float4*>(stg + r * kStgStride + (lane & 7) * 4);
                            o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
                            float* dst = outp + (long long)gm * args.N + gn;
                            if (vec_ok) {
                                *reinterpret_cast<float4*>(dst) = o;
                            } else {
                                if (gn + 0 < args.N) dst[0] = o.x;
                                if (gn + 1 < args.N) dst[1] = o.y;
                                if (gn + 2 < args.N) dst[2] = o.z;
                                if (gn + 3 < args.N) dst[3] = o.w;
                            }
                        }
                    }
                    __syncwarp();
                }
            }
#ifdef RACF_LINW_PROFILE
            prof[1] += clock64() - t_drained;    // epilogue: stores
#endif
        }
    }
#ifdef RACF_LINW_PROFILE
    if (blockIdx.x == 0 && lane == 0 && warp <= 2)
        printf("linw CTA0 warp %d total %lld | producer: empty (epilogue: drain) %lld | mma: tmem_empty (epilogue: stores) %lld full %lld | epilogue: tmem_full %lld\n", warp,
               clock64() - t_start, prof[0], prof[1], prof[2], prof[3]);
#endif
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
}

}  // namespace linwide
}  // namespace racf

// Launcher used by racf_linear_bf16x3_forward (variant 3; csrc/linear.cu validates the arguments and runs the K-split
// reduction). The operands are TiledOperands; partial tiles of a K split go to `out` = workspace [num_splits, M, N].
int racf_linear_wide_launch(const void* a_tiled, const void* w_tiled, const float* bias, float* out, int M, int N, int num_kblocks,
                            int kblocks_per_pass, int kblocks_per_split, int num_splits, int max_order, cudaStream_t st) {
    using namespace racf::linwide;
    Args args;
    args.a_tiled = static_cast<const uint8_t*>(a_tiled);
    args.w_tiled = static_cast<const uint8_t*>(w_tiled);
    args.bias = bias;
    args.out = out;
    args.M = M; args.N = N;
    args.m_tiles = (M + kBM - 1) / kBM;
    args.n_tiles128 = (N + 127) / 128;
    args.n_pairs = (args.n_tiles128 + 1) / 2;
    args.num_kblocks = num_kblocks;
    args.kblocks_per_pass = kblocks_per_pass;
    args.kblocks_per_split = kblocks_per_split;
    args.num_splits = num_splits;
    args.max_order = max_order;
    const long long tiles = (long long)args.m_tiles * args.n_pairs * num_splits;
    if (tiles > 0x7fffffffLL) return RACF_ERR_BAD_SHAPE;
    args.num_tiles = (int)tiles;
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return (int)e;
    const int smem = kSmemBytes + 1024;
    e = cudaFuncSetAttribute(linear_bf16x3_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    linear_bf16x3_wide_kernel<<<grid, kThreads, smem, st>>>(args);
    return (int)cudaGetLastError();
}
