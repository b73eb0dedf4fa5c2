// Row programs: the per-query "glue" of a decoder iteration in ONE launch (sm_100a).
//
// Between the sampling ops, the decoder layer (models/racformer_transformer.py:239-262) is a chain of row-wise
// operators on the [B*Q, 256] query matrix: Linear, LayerNorm, ReLU, residual adds, concatenation
// (position_encoder :204-211, norm1..3, fusion :231, FFN, cls/reg branches :214-226; the softmax queue fusion and
// output_proj of models/bev_self_attention.py:206-225). PyTorch runs them as ~150 launches of 2-12 us per iteration.
// Rows never interact, so a CTA can carry a tile of kRows rows through the WHOLE chain in shared memory: the host
// describes the chain as a short program of racf_row_op_t records (passed by value in the kernel parameters, so a
// CUDA graph captures it), every CTA interprets the same program on its own rows.
//
// Arithmetic is fp32 FMA on the CUDA cores; a dot product is two ascending-k partial sums (the 16-wide k halves of
// every 32-wide tile) added at the end, the bias last (the order of a GEMM epilogue). Weights are kept by the host as
// chunked transposes W^T [N/256][K][256]; a [32 k][256 column] tile is one contiguous 32 KB block that ONE thread moves
// into shared memory with cp.async.bulk (TMA engine) behind full / empty mbarriers -- a dedicated producer warp runs
// the whole program's tile stream through a four-deep ring, across operator boundaries, while 8 consumer warps compute. Every CTA reads the same <= 0.8 MB per Linear from L2: L2 -> SM bandwidth and the FMA pipe bound the
// kernel, not HBM (a first version with per-thread LDG weight loads was latency-bound at 40 us per 256x256 layer).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "racformer_ops.h"

namespace racf {

constexpr int kRowThreads = 256;                          // consumer threads (8 warps); one more warp only moves weights
constexpr int kRowWarps = kRowThreads / 32;
constexpr int kLaunchThreads = kRowThreads + 32;
constexpr int kChunkCols = RACF_ROW_CHUNK_COLS;           // 256 output columns per weight chunk
constexpr int kTileK = 32;                                // k rows per weight tile
constexpr int kTileFloats = kTileK * kChunkCols;          // 32 KB
constexpr int kStages = 4;                                // weight tile ring (6 stages measure the same)
constexpr int kSplitK = kRowThreads / 64;                 // a tile's 32 k are split over 4 thread groups of 64

struct RowProgram {
    racf_row_op_t ops[RACF_ROW_MAX_OPS];
    int num_ops;
    int rows;        // total rows
    int width;       // floats per buffer row (multiple of 4)
    int num_bufs;
};

// ---- PTX wrappers (mbarrier + 1-D bulk copy on the TMA engine)
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
        if (clock64() - t0 > 2000000000LL) __trap();
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// barrier of the consumer warps only (the producer warp never joins it)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kRowThreads) : "memory"); }
__device__ __forceinline__ void tma_load_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// The weight-tile stream: every LINEAR operator of the program, in program order, contributes its tiles
// [chunk][k block] (each a contiguous <= 32 KB block of the chunked transposed weight). One thread walks this
// cursor and keeps kStages tiles in flight across operator boundaries, so the next Linear's weights are already
// arriving while the CTA is still in a LayerNorm.
struct TileCursor {
    int op, chunk, kb;
};

__device__ __forceinline__ bool next_tile(const RowProgram& prog, TileCursor& cur, const float*& src, uint32_t& bytes) {
    while (cur.op < prog.num_ops) {
        const racf_row_op_t& op = prog.ops[cur.op];
        if (op.kind == RACF_ROWOP_LINEAR) {
            const int kblocks = (op.k + kTileK - 1) / kTileK, chunks = (op.n + kChunkCols - 1) / kChunkCols;
            if (cur.chunk < chunks) {
                const int k0 = cur.kb * kTileK;
                const int kn = min(kTileK, op.k - k0);
                src = op.p0 + ((size_t)cur.chunk * op.k + k0) * kChunkCols;
                bytes = (uint32_t)kn * kChunkCols * 4u;
                if (++cur.kb == kblocks) { cur.kb = 0; ++cur.chunk; }
                return true;
            }
        } else if (op.kind == RACF_ROWOP_LINEAR_NARROW && cur.chunk == 0) {   // the whole [n][k] weight is one tile
            src = op.p0;
            bytes = (uint32_t)(op.n * op.k) * 4u;
            cur.chunk = 1;
            return true;
        }
        ++cur.op; cur.chunk = 0; cur.kb = 0;
    }
    return false;
}

// dst[r][j] = act(bias[j] + sum_k src[r][k] * W[j][k]); the weight arrives as tiles of W^T through shared memory.
// Thread = (column quad, k group): it owns 4 columns x kRows rows (4 * kRows independent accumulators) and 8 of the 32 k
// of every tile, so each weight element is read from shared memory once (LDS.128) and each input once per warp
// (LDS.128 broadcast): 4 + kRows shared-memory instructions per 16 * kRows FMAs. A warp releases a stage by arriving on
// its "empty" barrier -- the consumer warps never wait for each other inside a Linear. The four k groups are added at
// the end of a chunk through a scratch area, in a fixed order, behind one barrier.
template <int kRows>
__device__ __forceinline__ void op_linear(const racf_row_op_t& op, float* bufs, float* scratch, const float* wtiles,
                                          uint32_t full_bar, uint32_t empty_bar, int width, int tid, unsigned& consumed) {
    static_assert(kSplitK >= 2, "the epilogue sums kSplitK - 1 parked partials and the owner's accumulators");
    constexpr int kGroupK = kTileK / kSplitK;   // 8
    const int jq = tid & 63, h = tid >> 6, lane = tid & 31;
    const float* src = bufs + (size_t)op.src * kRows * width + op.src_col;
    float* dst = bufs + (size_t)op.dst * kRows * width + op.dst_col;
    const int N = op.n, K = op.k;
    const bool relu = (op.flags & RACF_ROWOP_RELU) != 0, accum = (op.flags & RACF_ROWOP_ACCUM) != 0;
    const int kblocks = (K + kTileK - 1) / kTileK, chunks = (N + kChunkCols - 1) / kChunkCols;
    for (int c = 0; c < chunks; ++c) {
        float acc[kRows][4];
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
        for (int kb = 0; kb < kblocks; ++kb) {
            const int stage = consumed % kStages;
            mbar_wait(full_bar + 8u * stage, (consumed / kStages) & 1u);
            const float* ws = wtiles + (size_t)stage * kTileFloats + 4 * jq;
            const float* xs = src + kb * kTileK;
            const int kn = min(kTileK, K - kb * kTileK);
            if (kn == kTileK) {
                const float* wh = ws + h * kGroupK * kChunkCols;
                const float* xh = xs + h * kGroupK;
#pragma unroll
                for (int kk = 0; kk < kGroupK; kk += 4) {
                    float4 w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) w[i] = *reinterpret_cast<const float4*>(wh + (kk + i) * kChunkCols);
                    // packed FMAs (fma.rn.f32x2): two columns per instruction -- a 3-register FFMA issues every second
                    // cycle per scheduler on sm_100, the packed form does two of them in the same slot. k outer, rows
                    // inner: 16 independent instructions between two updates of one accumulator pair.
                    float4 x4[kRows];
#pragma unroll
                    for (int r = 0; r < kRows; ++r) x4[r] = *reinterpret_cast<const float4*>(xh + r * width + kk);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 wlo = make_float2(w[i].x, w[i].y), whi = make_float2(w[i].z, w[i].w);
#pragma unroll
                        for (int r = 0; r < kRows; ++r) {
                            const float xs1 = i == 0 ? x4[r].x : i == 1 ? x4[r].y : i == 2 ? x4[r].z : x4[r].w;
                            const float2 xx = make_float2(xs1, xs1);
                            const float2 lo = __ffma2_rn(xx, wlo, make_float2(acc[r][0], acc[r][1]));
                            const float2 hi = __ffma2_rn(xx, whi, make_float2(acc[r][2], acc[r][3]));
                            acc[r][0] = lo.x; acc[r][1] = lo.y; acc[r][2] = hi.x; acc[r][3] = hi.y;
                        }
                    }
                }
            } else {                               // K tail (or a tiny K): the groups take every fourth k
                for (int kk = h; kk < kn; kk += kSplitK) {
                    const float4 w = *reinterpret_cast<const float4*>(ws + kk * kChunkCols);
#pragma unroll
                    for (int r = 0; r < kRows; ++r) {
                        const float x = xs[r * width + kk];
                        acc[r][0] = fmaf(x, w.x, acc[r][0]); acc[r][1] = fmaf(x, w.y, acc[r][1]);
                        acc[r][2] = fmaf(x, w.z, acc[r][2]); acc[r][3] = fmaf(x, w.w, acc[r][3]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar + 8u * stage);   // this warp is done with the stage
            ++consumed;
        }
        // dst = ((p3 + p2) + p1) + p0, then bias and activation. Groups 1..3 park their partial sums in a scratch area
        // ([3][kRows][256] floats behind the row buffers), one barrier, group 0 finishes the chunk.
        const int j = c * kChunkCols + 4 * jq;
        if (h != 0) {
#pragma unroll
            for (int r = 0; r < kRows; ++r)
                *reinterpret_cast<float4*>(scratch + ((h - 1) * kRows + r) * kChunkCols + 4 * jq) =
                    make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        }
        consumer_sync();
        if (h == 0 && j < N) {
            float b[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) b[e] = (op.p1 != nullptr && j + e < N) ? __ldg(op.p1 + j + e) : 0.f;
            const bool vec = (j + 3 < N) && ((op.dst_col & 3) == 0);
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                float4 v = *reinterpret_cast<const float4*>(scratch + ((kSplitK - 2) * kRows + r) * kChunkCols + 4 * jq);
#pragma unroll
                for (int g = kSplitK - 3; g >= 0; --g) {
                    const float4 p = *reinterpret_cast<const float4*>(scratch + (g * kRows + r) * kChunkCols + 4 * jq);
                    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
                }
                v.x = (v.x + acc[r][0]) + b[0]; v.y = (v.y + acc[r][1]) + b[1];
                v.z = (v.z + acc[r][2]) + b[2]; v.w = (v.w + acc[r][3]) + b[3];
                if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                if (vec) {
                    if (accum) {
                        const float4 o = *reinterpret_cast<const float4*>(dst + r * width + j);
                        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                    }
                    *reinterpret_cast<float4*>(dst + r * width + j) = v;
                } else {
                    const float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (j + e < N) dst[r * width + j + e] = accum ? dst[r * width + j + e] + o[e] : o[e];
                }
            }
        }
        if (c + 1 < chunks) consumer_sync();       // the next chunk reuses the scratch area
    }
}

// Narrow outputs (a few columns, n * k <= 8192): the weight W [n][k] (nn.Linear's own layout) arrives as ONE tile of the
// stream; one warp per row, lanes split k, warp-shuffle reduction per output column.
template <int kRows>
__device__ __forceinline__ void op_linear_narrow(const racf_row_op_t& op, float* bufs, const float* wtiles, uint32_t full_bar,
                                                 uint32_t empty_bar, int width, int tid, unsigned& consumed) {
    const int warp = tid >> 5, lane = tid & 31;
    const int N = op.n, K = op.k;
    const bool relu = (op.flags & RACF_ROWOP_RELU) != 0;
    const int stage = consumed % kStages;
    mbar_wait(full_bar + 8u * stage, (consumed / kStages) & 1u);
    const float* w = wtiles + (size_t)stage * kTileFloats;
    for (int r = warp; r < kRows; r += kRowWarps) {
        const float* src = bufs + ((size_t)op.src * kRows + r) * width + op.src_col;
        float* dst = bufs + ((size_t)op.dst * kRows + r) * width + op.dst_col;
        for (int j0 = 0; j0 < N; j0 += 4) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k = lane; k < K; k += 32) {
                const float x = src[k];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (j0 + e < N) acc[e] = fmaf(x, w[(j0 + e) * K + k], acc[e]);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float s = warp_sum(acc[e]);
                if (lane == 0 && j0 + e < N) {
                    float v = s + (op.p1 != nullptr ? __ldg(op.p1 + j0 + e) : 0.f);
                    if (relu) v = fmaxf(v, 0.f);
                    dst[j0 + e] = v;
                }
            }
        }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar + 8u * stage);
    ++consumed;
}

// In-place LayerNorm over n columns (two-pass mean / variance, biased variance, rsqrtf as PyTorch), optional ReLU.
// Vector path (n, dst_col multiples of 4, n <= 256, 16-byte aligned gamma / beta): the row lives in registers and the
// gamma / beta loads are issued before the reductions, so their latency is hidden.
template <int kRows>
__device__ __forceinline__ void op_layernorm(const racf_row_op_t& op, float* bufs, int width, int warp, int lane) {
    constexpr int kV = 2;                     // float4 per lane on the vector path
    const int n = op.n;
    const bool relu = (op.flags & RACF_ROWOP_RELU) != 0;
    const bool vec = ((n | op.dst_col) & 3) == 0 && n <= 128 * kV &&
                     ((reinterpret_cast<uintptr_t>(op.p0) | reinterpret_cast<uintptr_t>(op.p1)) & 15u) == 0;
    for (int r = warp; r < kRows; r += kRowWarps) {
        float* x = bufs + ((size_t)op.dst * kRows + r) * width + op.dst_col;
        if (vec) {
            const int n4 = n >> 2;
            float4 v[kV], g[kV], b[kV];
#pragma unroll
            for (int i = 0; i < kV; ++i) {
                const int c4 = lane + 32 * i;
                if (c4 < n4) {
                    g[i] = op.p0 != nullptr ? __ldg(reinterpret_cast<const float4*>(op.p0) + c4) : make_float4(1.f, 1.f, 1.f, 1.f);
                    b[i] = op.p1 != nullptr ? __ldg(reinterpret_cast<const float4*>(op.p1) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[i] = *reinterpret_cast<const float4*>(x + 4 * c4);
                }
            }
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < kV; ++i)
                if (lane + 32 * i < n4) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            const float mean = warp_sum(s) / (float)n;
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < kV; ++i)
                if (lane + 32 * i < n4) {
                    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
                    q = fmaf(v[i].x, v[i].x, q); q = fmaf(v[i].y, v[i].y, q);
                    q = fmaf(v[i].z, v[i].z, q); q = fmaf(v[i].w, v[i].w, q);
                }
            const float rstd = rsqrtf(warp_sum(q) / (float)n + op.eps);
#pragma unroll
            for (int i = 0; i < kV; ++i) {
                const int c4 = lane + 32 * i;
                if (c4 < n4) {
                    float4 o;
                    o.x = v[i].x * rstd * g[i].x + b[i].x; o.y = v[i].y * rstd * g[i].y + b[i].y;
                    o.z = v[i].z * rstd * g[i].z + b[i].z; o.w = v[i].w * rstd * g[i].w + b[i].w;
                    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    *reinterpret_cast<float4*>(x + 4 * c4) = o;
                }
            }
            continue;
        }
        float s = 0.f;
        for (int c = lane; c < n; c += 32) s += x[c];
        const float mean = warp_sum(s) / (float)n;
        float q = 0.f;
        for (int c = lane; c < n; c += 32) {
            const float d = x[c] - mean;
            q = fmaf(d, d, q);
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)n + op.eps);
        for (int c = lane; c < n; c += 32) {
            float v = (x[c] - mean) * rstd;
            if (op.p0 != nullptr) v = v * __ldg(op.p0 + c);
            if (op.p1 != nullptr) v = v + __ldg(op.p1 + c);
            if (relu) v = fmaxf(v, 0.f);
            x[c] = v;
        }
    }
}

// Counter-based keep mask of the DROPOUT operator: a murmur-style finaliser of (seed, row, column). The same record in the
// forward and the backward program regenerates the same mask.
__device__ __forceinline__ float dropout_scale(unsigned seed, long long row, int col, float p) {
    unsigned h = seed ^ (unsigned)(row * 0x9E3779B1ull) ^ ((unsigned)col * 0x85EBCA77u);
    h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
    const float u = (float)(h >> 8) * (1.0f / 16777216.0f);
    return u >= p ? 1.0f / (1.0f - p) : 0.f;
}

// Backward of op_layernorm on the gradient held in buf[dst]: one warp per row recomputes mean / rstd / xhat from the saved
// input row p2, applies the ReLU mask of the forward (y = xhat * gamma + beta > 0), parks g and g * xhat in the scratch area
// for the gamma / beta gradients and rewrites the gradient in place: rstd * (g' - mean(g') - xhat * mean(g' * xhat)), g' = g * gamma.
template <int kRows>
__device__ __forceinline__ void op_layernorm_bwd(const racf_row_op_t& op, float* bufs, float* scratch, int width, long long row0,
                                                 int valid, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    const int n = op.n;
    const bool relu = (op.flags & RACF_ROWOP_RELU) != 0;
    float* part_b = scratch;                      // [kRows][n]: masked g        -> d beta
    float* part_g = scratch + kRows * n;          // [kRows][n]: masked g * xhat -> d gamma
    for (int r = warp; r < kRows; r += kRowWarps) {
        float* g = bufs + ((size_t)op.dst * kRows + r) * width + op.dst_col;
        if (r >= valid) {
            for (int c = lane; c < n; c += 32) { part_b[r * n + c] = 0.f; part_g[r * n + c] = 0.f; g[c] = 0.f; }
            continue;
        }
        const float* x = op.p2 + (row0 + r) * op.ld;
        float s = 0.f;
        for (int c = lane; c < n; c += 32) s += __ldg(x + c);
        const float mean = warp_sum(s) / (float)n;
        float q = 0.f;
        for (int c = lane; c < n; c += 32) {
            const float d = __ldg(x + c) - mean;
            q = fmaf(d, d, q);
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)n + op.eps);
        float m1 = 0.f, m2 = 0.f;
        for (int c = lane; c < n; c += 32) {
            const float xh = (__ldg(x + c) - mean) * rstd;
            const float gam = op.p0 != nullptr ? __ldg(op.p0 + c) : 1.f;
            float gv = g[c];
            if (relu) {
                const float y = xh * gam + (op.p1 != nullptr ? __ldg(op.p1 + c) : 0.f);
                if (!(y > 0.f)) gv = 0.f;
            }
            part_b[r * n + c] = gv;
            part_g[r * n + c] = gv * xh;
            const float gp = gv * gam;
            g[c] = gp;                            // g' parked in place for the second pass
            m1 += gp;
            m2 = fmaf(gp, xh, m2);
        }
        m1 = warp_sum(m1) / (float)n;
        m2 = warp_sum(m2) / (float)n;
        for (int c = lane; c < n; c += 32) {
            const float xh = (__ldg(x + c) - mean) * rstd;
            g[c] = rstd * (g[c] - m1 - xh * m2);
        }
    }
    consumer_sync();
    for (int c = tid; c < n; c += kRowThreads) {
        float sb = 0.f, sg = 0.f;
#pragma unroll
        for (int r = 0; r < kRows; ++r) { sb += part_b[r * n + c]; sg += part_g[r * n + c]; }
        if (op.out != nullptr) atomicAdd(op.out + c, sg);
        if (op.out2 != nullptr) atomicAdd(op.out2 + c, sb);
    }
}

template <int kRows>
__global__ void __launch_bounds__(kLaunchThreads, 1)
row_program_kernel(const __grid_constant__ RowProgram prog) {
    extern __shared__ __align__(128) float smem[];
    __shared__ float queue_w[kRows][RACF_ROW_MAX_QUEUE];
    __shared__ float queue_d[kRows][RACF_ROW_MAX_QUEUE];
    __shared__ __align__(8) unsigned long long bars[2 * kStages];
    float* wtiles = smem;                                  // [kStages][32][256]
    float* bufs = smem + (size_t)kStages * kTileFloats;    // [num_bufs][kRows][width]
    float* scratch = bufs + (size_t)prog.num_bufs * kRows * prog.width;   // [kSplitK - 1][kRows][256]: Linear k-group partials
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int width = prog.width;
    const long long row0 = (long long)blockIdx.x * kRows;
    const int valid = (int)min((long long)kRows, (long long)prog.rows - row0);
    const uint32_t full_bar = smem_u32(bars), empty_bar = full_bar + 8u * kStages;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar + 8u * s, 1);
            mbar_init(empty_bar + 8u * s, kRowWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == kRowWarps) {
        // Producer warp: one thread walks the weight-tile stream of the whole program and keeps the ring full, waiting
        // only for the consumers to release a stage.
        if (lane == 0) {
            TileCursor cursor = {0, 0, 0};
            const float* g;
            uint32_t bytes;
            for (unsigned issued = 0; next_tile(prog, cursor, g, bytes); ++issued) {
                const unsigned s = issued % kStages;
                if (issued >= kStages) mbar_wait(empty_bar + 8u * s, ((issued / kStages) - 1u) & 1u);
                mbar_arrive_expect_tx(full_bar + 8u * s, bytes);
                tma_load_bulk(smem_u32(wtiles + (size_t)s * kTileFloats), g, bytes, full_bar + 8u * s);
            }
        }
        return;
    }

    unsigned consumed = 0;
    for (int i = 0; i < prog.num_ops; ++i) {
        const racf_row_op_t& op = prog.ops[i];
        switch (op.kind) {
        case RACF_ROWOP_LOAD: {      // dst[r][c] = p0[row * ld + c]; rows past the end read as 0
            float* dst = bufs + (size_t)op.dst * kRows * width + op.dst_col;
            const bool accum = (op.flags & RACF_ROWOP_ACCUM) != 0;
            if (((op.n | op.ld | op.dst_col) & 3) == 0 && (reinterpret_cast<uintptr_t>(op.p0) & 15u) == 0) {
                const int n4 = op.n >> 2;
                for (int e = tid; e < kRows * n4; e += kRowThreads) {
                    const int r = e / n4, c = (e - r * n4) * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < valid) v = __ldg(reinterpret_cast<const float4*>(op.p0 + (row0 + r) * op.ld + c));
                    if (accum) {
                        const float4 o = *reinterpret_cast<const float4*>(dst + r * width + c);
                        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                    }
                    *reinterpret_cast<float4*>(dst + r * width + c) = v;
                }
            } else {
                for (int e = tid; e < kRows * op.n; e += kRowThreads) {
                    const int r = e / op.n, c = e - r * op.n;
                    const float v = r < valid ? __ldg(op.p0 + (row0 + r) * op.ld + c) : 0.f;
                    dst[r * width + c] = accum ? dst[r * width + c] + v : v;
                }
            }
            break;
        }
        case RACF_ROWOP_ZERO: {
            float* dst = bufs + (size_t)op.dst * kRows * width + op.dst_col;
            for (int e = tid; e < kRows * op.n; e += kRowThreads) {
                const int r = e / op.n, c = e - r * op.n;
                dst[r * width + c] = 0.f;
            }
            break;
        }
        case RACF_ROWOP_RELU_MASK: {
            float* dst = bufs + (size_t)op.dst * kRows * width + op.dst_col;
            for (int e = tid; e < kRows * op.n; e += kRowThreads) {
                const int r = e / op.n, c = e - r * op.n;
                if (r >= valid || !(__ldg(op.p0 + (row0 + r) * op.ld + c) > 0.f)) dst[r * width + c] = 0.f;
            }
            break;
        }
        case RACF_ROWOP_DROPOUT: {
            float* dst = bufs + (size_t)op.dst * kRows * width + op.dst_col;
            for (int e = tid; e < kRows * op.n; e += kRowThreads) {
                const int r = e / op.n, c = e - r * op.n;
                dst[r * width + c] *= dropout_scale((unsigned)op.aux, row0 + r, op.k + c, op.eps);
            }
            break;
        }
        case RACF_ROWOP_LAYERNORM_BWD:
            op_layernorm_bwd<kRows>(op, bufs, scratch, width, row0, valid, tid);
            break;
        case RACF_ROWOP_STORE_COLSUM: {
            const float* src = bufs + (size_t)op.src * kRows * width + op.src_col;
            if (op.out != nullptr)
                for (int e = tid; e < valid * op.n; e += kRowThreads) {
                    const int r = e / op.n, c = e - r * op.n;
                    op.out[(row0 + r) * op.ld + c] = src[r * width + c];
                }
            if (op.out2 != nullptr)
                for (int c = tid; c < op.n; c += kRowThreads) {
                    float sum = 0.f;
                    for (int r = 0; r < valid; ++r) sum += src[r * width + c];
                    atomicAdd(op.out2 + c, sum);
                }
            break;
        }
        case RACF_ROWOP_QUEUE_BWD: {   // one warp per row: d_t = <g, value_t>, grad value_t = w_t g, grad logit_t = w_t (d_t - sum_s w_s d_s)
            const int T = op.aux, Q = op.k;
            for (int r = warp; r < valid; r += kRowWarps) {
                const float* g = bufs + ((size_t)op.src * kRows + r) * width + op.src_col;
                const long long row = row0 + r;
                const long long b = row / Q, q = row - b * Q;
                float* w = queue_w[r];            // this warp's row: softmax weights, then the dots
                float* d = queue_d[r];
                if (lane == 0) {
                    if (op.p1 != nullptr) {
                        const float* lg = op.p1 + row * T;
                        float m = -INFINITY;
                        for (int t = 0; t < T; ++t) m = fmaxf(m, __ldg(lg + t));
                        float den = 0.f;
                        for (int t = 0; t < T; ++t) den += expf(__ldg(lg + t) - m);
                        for (int t = 0; t < T; ++t) w[t] = expf(__ldg(lg + t) - m) / den;
                    } else {
                        for (int t = 0; t < T; ++t) w[t] = 1.f / (float)T;
                    }
                }
                __syncwarp();
                float dot_all = 0.f;
                for (int t = 0; t < T; ++t) {
                    const long long off = ((b * T + t) * Q + q) * op.ld;
                    const float wt = w[t];
                    float acc = 0.f;
                    for (int c = lane; c < op.n; c += 32) {
                        const float gv = g[c];
                        acc = fmaf(gv, __ldg(op.p0 + off + c), acc);
                        if (op.out != nullptr) op.out[off + c] = wt * gv;
                    }
                    acc = warp_sum(acc);
                    if (lane == 0) d[t] = acc;
                    dot_all = fmaf(wt, acc, dot_all);
                }
                __syncwarp();
                if (op.out2 != nullptr && op.p1 != nullptr)
                    for (int t = lane; t < T; t += 32) op.out2[row * T + t] = w[t] * (d[t] - dot_all);
            }
            break;
        }
        case RACF_ROWOP_LOAD_QUEUE: {   // w = softmax over the T queue logits p1[row][T] (NULL: 1/T); dst = sum_t w_t * p0[b*T + t][q][c]
            float* dst = bufs + (size_t)op.dst * kRows * width + op.dst_col;
            const int T = op.aux, Q = op.k;
            if (tid < kRows) {
                const long long row = row0 + tid;
                if (tid < valid && op.p1 != nullptr) {
                    const float* lg = op.p1 + row * T;
                    float m = -INFINITY;
                    for (int t = 0; t < T; ++t) m = fmaxf(m, __ldg(lg + t));
                    float den = 0.f;
                    for (int t = 0; t < T; ++t) den += expf(__ldg(lg + t) - m);
                    for (int t = 0; t < T; ++t) queue_w[tid][t] = expf(__ldg(lg + t) - m) / den;
                } else {
                    for (int t = 0; t < T; ++t) queue_w[tid][t] = 1.f / (float)T;
                }
            }
            consumer_sync();
            const int n4 = op.n >> 2;      // validated: n, ld, dst_col multiples of 4, p0 16-byte aligned
            for (int e = tid; e < kRows * n4; e += kRowThreads) {
                const int r = e / n4, c = (e - r * n4) * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < valid) {
                    const long long row = row0 + r;
                    const long long b = row / Q, q = row - b * Q;
                    const float* base = op.p0 + ((b * T) * Q + q) * op.ld + c;
                    const long long tstride = (long long)Q * op.ld;
                    for (int t0 = 0; t0 < T; t0 += 8) {
                        float4 x[8];
#pragma unroll
                        for (int t = 0; t < 8; ++t)
                            if (t0 + t < T) x[t] = __ldg(reinterpret_cast<const float4*>(base + (t0 + t) * tstride));
#pragma unroll
                        for (int t = 0; t < 8; ++t)
                            if (t0 + t < T) {
                                const float wq = queue_w[r][t0 + t];
                                v.x = fmaf(x[t].x, wq, v.x); v.y = fmaf(x[t].y, wq, v.y);
                                v.z = fmaf(x[t].z, wq, v.z); v.w = fmaf(x[t].w, wq, v.w);
                            }
                    }
                }
                *reinterpret_cast<float4*>(dst + r * width + c) = v;
            }
            break;
        }
        case RACF_ROWOP_STORE: {
            const float* src = bufs + (size_t)op.src * kRows * width + op.src_col;
            for (int e = tid; e < valid * op.n; e += kRowThreads) {
                const int r = e / op.n, c = e - r * op.n;
                op.out[(row0 + r) * op.ld + c] = src[r * width + c];
            }
            break;
        }
        case RACF_ROWOP_ADD: {
            const float* src = bufs + (size_t)op.src * kRows * width + op.src_col;
            float* dst = bufs + (size_t)op.dst * kRows * width + op.dst_col;
            for (int e = tid; e < kRows * op.n; e += kRowThreads) {
                const int r = e / op.n, c = e - r * op.n;
                dst[r * width + c] += src[r * width + c];
            }
            break;
        }
        case RACF_ROWOP_LINEAR:
            op_linear<kRows>(op, bufs, scratch, wtiles, full_bar, empty_bar, width, tid, consumed);
            break;
        case RACF_ROWOP_LINEAR_NARROW:
            op_linear_narrow<kRows>(op, bufs, wtiles, full_bar, empty_bar, width, tid, consumed);
            break;
        case RACF_ROWOP_LAYERNORM: op_layernorm<kRows>(op, bufs, width, warp, lane); break;
        default: break;
        }
        consumer_sync();
    }
}

static int validate(const racf_row_op_t& op, int width, int num_bufs) {
    const bool reads_buf = op.kind == RACF_ROWOP_STORE || op.kind == RACF_ROWOP_ADD || op.kind == RACF_ROWOP_LINEAR ||
                           op.kind == RACF_ROWOP_LINEAR_NARROW || op.kind == RACF_ROWOP_STORE_COLSUM ||
                           op.kind == RACF_ROWOP_QUEUE_BWD;
    const bool writes_buf = op.kind != RACF_ROWOP_STORE && op.kind != RACF_ROWOP_STORE_COLSUM && op.kind != RACF_ROWOP_QUEUE_BWD;
    if (op.kind < RACF_ROWOP_LOAD || op.kind > RACF_ROWOP_QUEUE_BWD) return RACF_ERR_UNSUPPORTED;
    if (op.n <= 0) return RACF_ERR_BAD_SHAPE;
    const bool is_linear = op.kind == RACF_ROWOP_LINEAR || op.kind == RACF_ROWOP_LINEAR_NARROW;
    const int in_w = is_linear ? op.k : op.n;
    if (is_linear && op.k <= 0) return RACF_ERR_BAD_SHAPE;
    if (reads_buf && (op.src < 0 || op.src >= num_bufs || op.src_col < 0 || op.src_col + in_w > width)) return RACF_ERR_BAD_SHAPE;
    if (writes_buf && (op.dst < 0 || op.dst >= num_bufs || op.dst_col < 0 || op.dst_col + op.n > width)) return RACF_ERR_BAD_SHAPE;
    switch (op.kind) {
    case RACF_ROWOP_LOAD:
        if (!op.p0) return RACF_ERR_NULL_POINTER;
        if (op.ld < op.n) return RACF_ERR_BAD_SHAPE;
        break;
    case RACF_ROWOP_LOAD_QUEUE:
        if (!op.p0) return RACF_ERR_NULL_POINTER;
        if (op.ld < op.n || op.aux <= 0 || op.aux > RACF_ROW_MAX_QUEUE || op.k <= 0) return RACF_ERR_BAD_SHAPE;
        if (((op.n | op.ld | op.dst_col) & 3) != 0 || (reinterpret_cast<uintptr_t>(op.p0) & 15u)) return RACF_ERR_UNSUPPORTED;
        break;
    case RACF_ROWOP_STORE:
        if (!op.out) return RACF_ERR_NULL_POINTER;
        if (op.ld < op.n) return RACF_ERR_BAD_SHAPE;
        break;
    case RACF_ROWOP_STORE_COLSUM:
        if (!op.out && !op.out2) return RACF_ERR_NULL_POINTER;
        if (op.out && op.ld < op.n) return RACF_ERR_BAD_SHAPE;
        break;
    case RACF_ROWOP_RELU_MASK:
        if (!op.p0) return RACF_ERR_NULL_POINTER;
        if (op.ld < op.n) return RACF_ERR_BAD_SHAPE;
        break;
    case RACF_ROWOP_DROPOUT:
        if (!(op.eps >= 0.f && op.eps < 1.f)) return RACF_ERR_BAD_SHAPE;
        break;
    case RACF_ROWOP_LAYERNORM_BWD:
        if (!op.p2) return RACF_ERR_NULL_POINTER;
        if (op.ld < op.n) return RACF_ERR_BAD_SHAPE;
        if (op.n > 384) return RACF_ERR_UNSUPPORTED;                // two [rows_per_cta][n] partials in the scratch area
        break;
    case RACF_ROWOP_QUEUE_BWD:
        if (!op.p0) return RACF_ERR_NULL_POINTER;
        if (op.ld < op.n || op.aux <= 0 || op.aux > RACF_ROW_MAX_QUEUE || op.k <= 0) return RACF_ERR_BAD_SHAPE;
        break;
    case RACF_ROWOP_LINEAR:
        if (!op.p0) return RACF_ERR_NULL_POINTER;
        if (op.src == op.dst) return RACF_ERR_UNSUPPORTED;          // not in place
        if ((op.src_col & 3) != 0) return RACF_ERR_UNSUPPORTED;     // 128-bit shared-memory reads
        if (reinterpret_cast<uintptr_t>(op.p0) & 15u) return RACF_ERR_UNSUPPORTED;   // bulk copies of the weight tiles
        break;
    case RACF_ROWOP_LINEAR_NARROW:
        if (!op.p0) return RACF_ERR_NULL_POINTER;
        if (op.src == op.dst) return RACF_ERR_UNSUPPORTED;
        if ((long long)op.n * op.k > 8192 || ((op.n * op.k) & 3) != 0 || (reinterpret_cast<uintptr_t>(op.p0) & 15u))
            return RACF_ERR_UNSUPPORTED;                                // one weight tile, moved by one bulk copy
        break;
    default: break;
    }
    return RACF_OK;
}

template <int kRows>
static int launch(const RowProgram& prog, cudaStream_t st) {
    const size_t smem = ((size_t)kStages * kTileFloats + (size_t)prog.num_bufs * kRows * prog.width +
                         (size_t)(kSplitK - 1) * kRows * kChunkCols) * sizeof(float);
    if (smem > 227u * 1024u - 1024u) return RACF_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(row_program_kernel<kRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const long long grid = ((long long)prog.rows + kRows - 1) / kRows;
    if (grid > 0x7fffffffLL) return RACF_ERR_BAD_SHAPE;
    row_program_kernel<kRows><<<(unsigned)grid, kLaunchThreads, smem, st>>>(prog);
    return (int)cudaGetLastError();
}

}  // namespace racf

extern "C" int racf_row_program_forward(const racf_row_op_t* ops, int num_ops, int rows, int rows_per_cta, int num_bufs,
                                        int width, racf_stream_t stream) {
    using namespace racf;
    if (!ops) return RACF_ERR_NULL_POINTER;
    if (num_ops <= 0 || num_ops > RACF_ROW_MAX_OPS || rows <= 0 || num_bufs <= 0 || width <= 0) return RACF_ERR_BAD_SHAPE;
    if ((width & 3) != 0 || rows_per_cta < 4 || rows_per_cta > 8) return RACF_ERR_UNSUPPORTED;
    RowProgram prog;
    for (int i = 0; i < num_ops; ++i) {
        const int rc = validate(ops[i], width, num_bufs);
        if (rc != RACF_OK) return rc;
        prog.ops[i] = ops[i];
    }
    prog.num_ops = num_ops;
    prog.rows = rows;
    prog.width = width;
    prog.num_bufs = num_bufs;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (rows_per_cta) {
        case 4: return launch<4>(prog, st);
        case 5: return launch<5>(prog, st);
        case 6: return launch<6>(prog, st);
        case 7: return launch<7>(prog, st);
        default: return launch<8>(prog, st);
    }
}
