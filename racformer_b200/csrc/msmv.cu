// Multi-scale multi-view (MSMV) sampling for B200 (sm_100a): forward, backward, mask dump.
//
// Operator semantics: models/csrc/msmv_sampling/msmv_sampling_forward.cu:27-164 and
// msmv_sampling_backward.cu:29-224 of the reference (restated in SURVEY.md App. A.1/A.2).
//
// Work decomposition (fast path, C == 64):
//   * one warp owns one query (b, q): all P sample points, all L levels, all 64 channels. Because the
//     warp owns every contribution to grad_loc[b,q,:,:] and grad_weights[b,q,:,:], those gradients are
//     reduced with warp shuffles and written with plain stores -- no atomics, no pre-zeroing.
//   * a pixel is 64 fp32 = 256 B = 16 lanes x 128 bit. In channel-last layout the two x-neighbours of
//     a bilinear cell are adjacent in memory, so ONE warp-wide 128-bit load fetches 512 contiguous
//     bytes = both corners of a row: lanes 0-15 hold (y, x0), lanes 16-31 hold (y, x0+1). A tap is two
//     such loads (top row, bottom row); the two half-warps are combined with a single shfl_xor(16).
//   * per-tap geometry (corner offset, validity bits, bilinear and level weights) is computed once by
//     one lane per tap and staged in shared memory; every lane then reads it back as a broadcast.
//   * backward scatters the feature gradient with red.global.add.v4.f32: one 512-byte vector
//     reduction per row per warp instead of 128 scalar atomics.
#include <cstdlib>

#include "racf_common.cuh"
#include "racformer_ops.h"

namespace racf {

constexpr int kMsmvChunk = 4;   // sample points handled per inner round (also the float4 width along P)
#ifndef RACF_MSMV_WARPS
#define RACF_MSMV_WARPS 8
#endif
constexpr int kMsmvWarps = RACF_MSMV_WARPS;   // warps (queries) per CTA on the fast path
constexpr int kLanesPerPixel = 16;  // C == 64 -> 16 float4 lanes

template <int L>
struct MsmvArgs {
    const float* feat[L];
    float* grad_feat[L];
    int H[L];
    int W[L];
    const float* loc;      // [B,Q,P,3]
    const float* wts;      // [B,Q,P,L]
    const float* grad_out; // [B,Q,C,P]
    float* out;            // [B,Q,C,P]
    float* grad_loc;       // [B,Q,P,3]
    float* grad_wts;       // [B,Q,P,L]
    int B, N, Q, P, C;
    int out_T, out_G;   // > 0: forward writes the un-packed layout [B/(T*G), Q, G, T*P, C] instead of [B,Q,C,P]
};

template <int L>
__device__ __forceinline__ void level_dims(const MsmvArgs<L>& a, int l, int& H, int& W) {
    H = a.H[0];
    W = a.W[0];
#pragma unroll
    for (int k = 1; k < L; ++k)
        if (k == l) { H = a.H[k]; W = a.W[k]; }
}

// ------------------------------------------------------------------------------------------------
// Fast path, forward. grid = ceil(B*Q / kMsmvWarps), block = kMsmvWarps warps.
//
// Latency, not bandwidth, limited the first version (ncu r01: 75 % long-scoreboard stalls, DRAM 50 %), so the
// loads are software-pipelined: the geometry of a whole batch of points (up to 64 taps) is staged once, and the
// 2L 128-bit loads of point p+1 are issued into a second register buffer before the FMAs of point p run.
// ------------------------------------------------------------------------------------------------
template <int L>
struct PointRegs {
    float4 top[L];
    float4 bot[L];
};

#ifndef RACF_FWD_MIN_BLOCKS
#define RACF_FWD_MIN_BLOCKS 2   // 128 registers: room for two point buffers (16 x 128-bit loads in flight per lane)
#endif
template <int L>
__global__ void __launch_bounds__(kMsmvWarps * 32, RACF_FWD_MIN_BLOCKS) msmv_fwd_c64_kernel(const MsmvArgs<L> a) {
    constexpr int G = kLanesPerPixel;
    constexpr int PB = (L <= 4) ? 16 : 12;      // points per staged batch (multiple of kMsmvChunk)
    constexpr int NT = PB * L;                  // taps per batch (<= 64)
    constexpr int NTP = NT + L;                 // + one all-invalid pad point so the pipeline can run one ahead
    static_assert(NT <= 64 && PB % kMsmvChunk == 0, "batch geometry");
    __shared__ float2 s_w[kMsmvWarps][NTP][2];  // [tap][x-slot] = {w_top, w_bot} * level weight
    __shared__ int2 s_om[kMsmvWarps][NTP];      // {float4 offset of the top-left pixel, corner mask}

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long bq = (long long)blockIdx.x * kMsmvWarps + warp;
    if (bq >= (long long)a.B * a.Q) return;  // whole warp exits together; only __syncwarp below
    const int b = (int)(bq / a.Q);
    const int slot = lane >> 4, j = lane & 15;

    const float4* base[L];
    int row[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        base[l] = reinterpret_cast<const float4*>(a.feat[l] + (size_t)b * a.N * a.H[l] * a.W[l] * 64) + slot * G + j;
        row[l] = a.W[l] * G;
    }
    const float* loc_q = a.loc + bq * a.P * 3;
    const float* wts_q = a.wts + bq * a.P * L;
    float* out_q = a.out + bq * 64 * a.P;
    const bool vec_store = (a.P % 4) == 0;
    if (lane < L) {
        s_w[warp][NT + lane][0] = make_float2(0.f, 0.f);
        s_w[warp][NT + lane][1] = make_float2(0.f, 0.f);
        s_om[warp][NT + lane] = make_int2(0, 0);
    }

    auto issue = [&](PointRegs<L>& r, int pt) {
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const int2 om = s_om[warp][pt * L + l];
            const unsigned m = (unsigned)om.y >> slot;  // bit0 = top corner of my column, bit2 = bottom
            const float4* p = base[l] + om.x;
            r.top[l] = make_float4(0.f, 0.f, 0.f, 0.f);
            r.bot[l] = r.top[l];
            if (m & 1u) r.top[l] = ldg128(p);
            if (m & 4u) r.bot[l] = ldg128(p + row[l]);
        }
    };
    auto consume = [&](const PointRegs<L>& r, int pt, float4& acc) {
        acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const float2 w = s_w[warp][pt * L + l][slot];
            acc.x = fmaf(w.y, r.bot[l].x, fmaf(w.x, r.top[l].x, acc.x));
            acc.y = fmaf(w.y, r.bot[l].y, fmaf(w.x, r.top[l].y, acc.y));
            acc.z = fmaf(w.y, r.bot[l].z, fmaf(w.x, r.top[l].z, acc.z));
            acc.w = fmaf(w.y, r.bot[l].w, fmaf(w.x, r.top[l].w, acc.w));
        }
    };

    for (int pb = 0; pb < a.P; pb += PB) {
        __syncwarp();
        for (int t = lane; t < NT; t += 32) {
            const int pp = t / L, l = t % L, p = pb + pp;
            float2 w0 = make_float2(0.f, 0.f), w1 = w0;
            int2 om = make_int2(0, 0);
            if (p < a.P) {
                const float lx = loc_q[p * 3], ly = loc_q[p * 3 + 1], lz = loc_q[p * 3 + 2];
                const float s = wts_q[p * L + l];
                int H, W;
                level_dims<L>(a, l, H, W);
                const int v = msmv_view(lz, a.N);
                const TapGeom g = tap_geometry(msmv_pixel(ly, H), msmv_pixel(lx, W), H, W);
                if (g.mask != 0u && v >= 0 && v < a.N) {
                    const float hh = 1.f - g.lh, hw = 1.f - g.lw;
                    w0 = make_float2(hh * hw * s, g.lh * hw * s);
                    w1 = make_float2(hh * g.lw * s, g.lh * g.lw * s);
                    om.x = ((v * H + g.h_low) * W + g.w_low) * G;
                    om.y = (int)g.mask;
                }
            }
            s_w[warp][t][0] = w0;
            s_w[warp][t][1] = w1;
            s_om[warp][t] = om;
        }
        __syncwarp();

        const int npts = min(PB, a.P - pb);
        PointRegs<L> ra, rb;
        issue(ra, 0);
        for (int c = 0; c < npts; c += kMsmvChunk) {
            float4 acc[kMsmvChunk];
            issue(rb, c + 1);
            consume(ra, c + 0, acc[0]);
            issue(ra, c + 2);
            consume(rb, c + 1, acc[1]);
            issue(rb, c + 3);
            consume(ra, c + 2, acc[2]);
            issue(ra, c + 4);  // first point of the next chunk (or the all-invalid pad point)
            consume(rb, c + 3, acc[3]);
            // combine the x0 / x1 half-warps
#pragma unroll
            for (int pp = 0; pp < kMsmvChunk; ++pp) {
                acc[pp].x += __shfl_xor_sync(0xffffffffu, acc[pp].x, 16);
                acc[pp].y += __shfl_xor_sync(0xffffffffu, acc[pp].y, 16);
                acc[pp].z += __shfl_xor_sync(0xffffffffu, acc[pp].z, 16);
                acc[pp].w += __shfl_xor_sync(0xffffffffu, acc[pp].w, 16);
            }
            const int p0 = pb + c;
            if (a.out_T > 0) {
                // un-packed layout of sampling_4d's tail (sparsebev_sampling.py:128-131): out[bb,q,g,t*P+p,:], one
                // pixel-sized 256 B row per point; half-warp 0 stores points 0,1 and half-warp 1 points 2,3
                const int g = b % a.out_G, t = (b / a.out_G) % a.out_T, bb = b / (a.out_G * a.out_T);
                const int q = (int)(bq - (long long)b * a.Q);
                float* rowp = a.out + ((((size_t)bb * a.Q + q) * a.out_G + g) * ((size_t)a.out_T * a.P) + (size_t)t * a.P) * 64;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int pp = 2 * slot + h;
                    const float4 val = slot ? (h ? acc[3] : acc[2]) : (h ? acc[1] : acc[0]);
                    if (p0 + pp < a.P) *reinterpret_cast<float4*>(rowp + (size_t)(p0 + pp) * 64 + 4 * j) = val;
                }
                continue;
            }
            // out[b,q,c,p]: lane (slot, j) stores channels 4j + 2*slot + {0,1}, points p0..p0+3
            const int c0 = 4 * j + 2 * slot;
            const float4 e0 = slot ? make_float4(acc[0].z, acc[1].z, acc[2].z, acc[3].z)
                                   : make_float4(acc[0].x, acc[1].x, acc[2].x, acc[3].x);
            const float4 e1 = slot ? make_float4(acc[0].w, acc[1].w, acc[2].w, acc[3].w)
                                   : make_float4(acc[0].y, acc[1].y, acc[2].y, acc[3].y);
            if (vec_store) {
                *reinterpret_cast<float4*>(out_q + (size_t)c0 * a.P + p0) = e0;
                *reinterpret_cast<float4*>(out_q + (size_t)(c0 + 1) * a.P + p0) = e1;
            } else {
                const float v0[4] = {e0.x, e0.y, e0.z, e0.w}, v1[4] = {e1.x, e1.y, e1.z, e1.w};
#pragma unroll
                for (int pp = 0; pp < kMsmvChunk; ++pp)
                    if (p0 + pp < a.P) {
                        out_q[(size_t)c0 * a.P + p0 + pp] = v0[pp];
                        out_q[(size_t)(c0 + 1) * a.P + p0 + pp] = v1[pp];
                    }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fast path, forward, persistent variant with software L2 prefetch (used when a query fits one staged batch).
//
// The gather is limited by bytes in flight per SM (ncu r01b; every occupancy/register trade lands on the same
// ~130 us plateau, profiles/r01_msmv_fwd_variants.txt). Each warp walks its queries in a 3-stage pipeline, so the
// loc/weight loads and the tap-record computation of the following queries hide behind the gather of the current
// one (133 -> 125 us). The optional prefetch.global.L2 of the next query's rows was measured too and is OFF: it
// costs L2 request bandwidth, which is the binding resource (145-168 us with 1-4 levels prefetched).
//   iteration i:  (a) turn the loc/weights loaded during iteration i-1 into the tap records of query i+1 and issue
//                     prefetch.global.L2 for their cell rows on the DRAM-sized levels,
//                 (b) load loc/weights of query i+2 into registers,
//                 (c) gather + blend query i, whose rows were prefetched one iteration ago.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int L>
__global__ void __launch_bounds__(kMsmvWarps * 32, 2) msmv_fwd_c64_pf_kernel(const MsmvArgs<L> a, int prefetch_levels) {
    constexpr int G = kLanesPerPixel;
    constexpr int PB = (L <= 4) ? 16 : 12;
    constexpr int NT = PB * L;
    constexpr int NTP = NT + L;
    constexpr int TPL = (NT + 31) / 32;          // taps owned by a lane (<= 2)
    __shared__ float2 s_w[kMsmvWarps][2][NTP][2];
    __shared__ int2 s_om[kMsmvWarps][2][NTP];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = lane >> 4, j = lane & 15;
    const long long nq = (long long)a.B * a.Q;
    const long long stride = (long long)gridDim.x * kMsmvWarps;
    const long long q0 = (long long)blockIdx.x * kMsmvWarps + warp;
    if (q0 >= nq) return;

    const float4* lbase[L];
    int row[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        lbase[l] = reinterpret_cast<const float4*>(a.feat[l]) + slot * G + j;
        row[l] = a.W[l] * G;
    }
    const bool vec_store = (a.P % 4) == 0;
    if (lane < L) {
#pragma unroll
        for (int bsel = 0; bsel < 2; ++bsel) {
            s_w[warp][bsel][NT + lane][0] = make_float2(0.f, 0.f);
            s_w[warp][bsel][NT + lane][1] = make_float2(0.f, 0.f);
            s_om[warp][bsel][NT + lane] = make_int2(0, 0);
        }
    }

    float lx[TPL], ly[TPL], lz[TPL], sw[TPL];
    auto load_locw = [&](long long q) {
#pragma unroll
        for (int k = 0; k < TPL; ++k) {
            const int t = lane + 32 * k, pp = t / L, l = t % L;
            lx[k] = ly[k] = lz[k] = sw[k] = 0.f;
            if (q < nq && t < NT && pp < a.P) {
                const float* lp = a.loc + (q * a.P + pp) * 3;
                lx[k] = __ldg(lp); ly[k] = __ldg(lp + 1); lz[k] = __ldg(lp + 2);
                sw[k] = __ldg(a.wts + (q * a.P + pp) * L + l);
            }
        }
    };
    auto stage = [&](long long q, int buf) {
        const int b = (int)(q / a.Q);
#pragma unroll
        for (int k = 0; k < TPL; ++k) {
            const int t = lane + 32 * k;
            if (t >= NT) continue;
            const int pp = t / L, l = t % L;
            float2 w0 = make_float2(0.f, 0.f), w1 = w0;
            int2 om = make_int2(0, 0);
            if (pp < a.P) {
                int H, W;
                level_dims<L>(a, l, H, W);
                const int v = msmv_view(lz[k], a.N);
                const TapGeom g = tap_geometry(msmv_pixel(ly[k], H), msmv_pixel(lx[k], W), H, W);
                if (g.mask != 0u && v >= 0 && v < a.N) {
                    const float hh = 1.f - g.lh, hw = 1.f - g.lw, s = sw[k];
                    w0 = make_float2(hh * hw * s, g.lh * hw * s);
                    w1 = make_float2(hh * g.lw * s, g.lh * g.lw * s);
                    om.x = ((((b * a.N + v) * H) + g.h_low) * W + g.w_low) * G;
                    om.y = (int)g.mask;
                    if (l < prefetch_levels) {
                        const float* f = a.feat[0];
#pragma unroll
                        for (int kk = 1; kk < L; ++kk)
                            if (kk == l) f = a.feat[kk];
                        const char* top = reinterpret_cast<const char*>(f) + (long long)om.x * 16;
                        const char* bot = top + (long long)W * G * 16;
                        if (g.mask & kTL) { prefetch_l2(top); prefetch_l2(top + 128); }
                        if (g.mask & kTR) { prefetch_l2(top + 256); prefetch_l2(top + 384); }
                        if (g.mask & kBL) { prefetch_l2(bot); prefetch_l2(bot + 128); }
                        if (g.mask & kBR) { prefetch_l2(bot + 256); prefetch_l2(bot + 384); }
                    }
                }
            }
            s_w[warp][buf][t][0] = w0;
            s_w[warp][buf][t][1] = w1;
            s_om[warp][buf][t] = om;
        }
    };

    // prologue: query 0 staged directly, loc/weights of query 1 in registers
    load_locw(q0);
    stage(q0, 0);
    load_locw(q0 + stride);
    int cur = 0;
    for (long long q = q0; q < nq; q += stride, cur ^= 1) {
        if (q + stride < nq) stage(q + stride, cur ^ 1);
        load_locw(q + 2 * stride);
        __syncwarp();

        auto issue = [&](PointRegs<L>& r, int pt) {
#pragma unroll
            for (int l = 0; l < L; ++l) {
                const int2 om = s_om[warp][cur][pt * L + l];
                const unsigned m = (unsigned)om.y >> slot;
                const float4* p = lbase[l] + om.x;
                r.top[l] = make_float4(0.f, 0.f, 0.f, 0.f);
                r.bot[l] = r.top[l];
                if (m & 1u) r.top[l] = ldg128(p);
                if (m & 4u) r.bot[l] = ldg128(p + row[l]);
            }
        };
        auto consume = [&](const PointRegs<L>& r, int pt, float4& acc) {
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int l = 0; l < L; ++l) {
                const float2 w = s_w[warp][cur][pt * L + l][slot];
                acc.x = fmaf(w.y, r.bot[l].x, fmaf(w.x, r.top[l].x, acc.x));
                acc.y = fmaf(w.y, r.bot[l].y, fmaf(w.x, r.top[l].y, acc.y));
                acc.z = fmaf(w.y, r.bot[l].z, fmaf(w.x, r.top[l].z, acc.z));
                acc.w = fmaf(w.y, r.bot[l].w, fmaf(w.x, r.top[l].w, acc.w));
            }
        };

        const int b = (int)(q / a.Q);
        float* out_q = a.out + q * 64 * a.P;
        PointRegs<L> ra, rb;
        issue(ra, 0);
        for (int c = 0; c < a.P; c += kMsmvChunk) {
            float4 acc[kMsmvChunk];
            issue(rb, c + 1);
            consume(ra, c + 0, acc[0]);
            issue(ra, c + 2);
            consume(rb, c + 1, acc[1]);
            issue(rb, c + 3);
            consume(ra, c + 2, acc[2]);
            issue(ra, c + 4);
            consume(rb, c + 3, acc[3]);
#pragma unroll
            for (int pp = 0; pp < kMsmvChunk; ++pp) {
                acc[pp].x += __shfl_xor_sync(0xffffffffu, acc[pp].x, 16);
                acc[pp].y += __shfl_xor_sync(0xffffffffu, acc[pp].y, 16);
                acc[pp].z += __shfl_xor_sync(0xffffffffu, acc[pp].z, 16);
                acc[pp].w += __shfl_xor_sync(0xffffffffu, acc[pp].w, 16);
            }
            const int p0 = c;
            if (a.out_T > 0) {
                const int g = b % a.out_G, t = (b / a.out_G) % a.out_T, bb = b / (a.out_G * a.out_T);
                const int qi = (int)(q - (long long)b * a.Q);
                float* rowp = a.out + ((((size_t)bb * a.Q + qi) * a.out_G + g) * ((size_t)a.out_T * a.P) + (size_t)t * a.P) * 64;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int pp = 2 * slot + h;
                    const float4 val = slot ? (h ? acc[3] : acc[2]) : (h ? acc[1] : acc[0]);
                    if (p0 + pp < a.P) *reinterpret_cast<float4*>(rowp + (size_t)(p0 + pp) * 64 + 4 * j) = val;
                }
                continue;
            }
            const int c0 = 4 * j + 2 * slot;
            const float4 e0 = slot ? make_float4(acc[0].z, acc[1].z, acc[2].z, acc[3].z)
                                   : make_float4(acc[0].x, acc[1].x, acc[2].x, acc[3].x);
            const float4 e1 = slot ? make_float4(acc[0].w, acc[1].w, acc[2].w, acc[3].w)
                                   : make_float4(acc[0].y, acc[1].y, acc[2].y, acc[3].y);
            if (vec_store) {
                *reinterpret_cast<float4*>(out_q + (size_t)c0 * a.P + p0) = e0;
                *reinterpret_cast<float4*>(out_q + (size_t)(c0 + 1) * a.P + p0) = e1;
            } else {
                const float v0[4] = {e0.x, e0.y, e0.z, e0.w}, v1[4] = {e1.x, e1.y, e1.z, e1.w};
#pragma unroll
                for (int pp = 0; pp < kMsmvChunk; ++pp)
                    if (p0 + pp < a.P) {
                        out_q[(size_t)c0 * a.P + p0 + pp] = v0[pp];
                        out_q[(size_t)(c0 + 1) * a.P + p0 + pp] = v1[pp];
                    }
            }
        }
        __syncwarp();   // every lane is done with buffer `cur` before the next iteration restages it
    }
}

// ------------------------------------------------------------------------------------------------
// Fast path, backward. Same ownership as forward.
// ------------------------------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(kMsmvWarps * 32) msmv_bwd_c64_kernel(const MsmvArgs<L> a) {
    constexpr int G = kLanesPerPixel;
    constexpr int TAPS = kMsmvChunk * L;
    static_assert(TAPS <= 32, "one lane per tap");
    __shared__ float4 s_rec[kMsmvWarps][TAPS];           // {lh, lw, level weight, bits(offset)}
    __shared__ unsigned s_mask[kMsmvWarps][TAPS];
    __shared__ __align__(16) float s_g[kMsmvWarps][kMsmvChunk][64];  // grad_out chunk, [point][channel]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long bq = (long long)blockIdx.x * kMsmvWarps + warp;
    if (bq >= (long long)a.B * a.Q) return;
    const int b = (int)(bq / a.Q);
    const int slot = lane >> 4, j = lane & 15;

    const float4* base[L];
    float* gbase[L];
    int row[L];
    float sx_scale[L], sy_scale[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
        const size_t boff = (size_t)b * a.N * a.H[l] * a.W[l] * 64;
        base[l] = reinterpret_cast<const float4*>(a.feat[l] + boff) + slot * G + j;
        gbase[l] = a.grad_feat[l] + boff + (size_t)(slot * G + j) * 4;
        row[l] = a.W[l] * G;
        sx_scale[l] = (float)(a.W[l] - 1);
        sy_scale[l] = (float)(a.H[l] - 1);
    }
    const float* loc_q = a.loc + bq * a.P * 3;
    const float* wts_q = a.wts + bq * a.P * L;
    const float* go_q = a.grad_out + bq * 64 * a.P;
    const bool grouped = a.out_T > 0;       // grad_out arrives in the forward_grouped layout [B/(T*G), Q, G, T*P, C]
    if (grouped) {
        const int q = (int)(bq % a.Q);
        const int g = b % a.out_G, t = (b / a.out_G) % a.out_T, bb = b / (a.out_G * a.out_T);
        go_q = a.grad_out + ((((size_t)bb * a.Q + q) * a.out_G + g) * ((size_t)a.out_T * a.P) + (size_t)t * a.P) * 64;
    }
    float* gl_q = a.grad_loc + bq * a.P * 3;
    float* gw_q = a.grad_wts + bq * a.P * L;
    const bool vec_load = (a.P % 4) == 0;
    const float sgn = slot ? 1.f : -1.f;

    for (int p0 = 0; p0 < a.P; p0 += kMsmvChunk) {
        __syncwarp();
        // stage grad_out[b,q,:,p0:p0+4] transposed to [point][channel] (the grouped layout is already channel-contiguous)
#pragma unroll
        for (int h = 0; h < 2 && !grouped; ++h) {
            const int c = lane + 32 * h;
            float g4[4] = {0.f, 0.f, 0.f, 0.f};
            if (vec_load) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(go_q + (size_t)c * a.P + p0));
                g4[0] = t.x; g4[1] = t.y; g4[2] = t.z; g4[3] = t.w;
            } else {
#pragma unroll
                for (int pp = 0; pp < kMsmvChunk; ++pp)
                    if (p0 + pp < a.P) g4[pp] = __ldg(go_q + (size_t)c * a.P + p0 + pp);
            }
#pragma unroll
            for (int pp = 0; pp < kMsmvChunk; ++pp) s_g[warp][pp][c] = g4[pp];
        }
        if (lane < TAPS) {
            const int pp = lane / L, l = lane % L, p = p0 + pp;
            float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
            unsigned mask = 0u;
            if (p < a.P) {
                const float lx = loc_q[p * 3], ly = loc_q[p * 3 + 1], lz = loc_q[p * 3 + 2];
                const float s = wts_q[p * L + l];
                int H, W;
                level_dims<L>(a, l, H, W);
                const int v = msmv_view(lz, a.N);
                const TapGeom g = tap_geometry(msmv_pixel(ly, H), msmv_pixel(lx, W), H, W);
                if (g.mask != 0u && v >= 0 && v < a.N) {
                    rec = make_float4(g.lh, g.lw, s, __int_as_float(((v * H + g.h_low) * W + g.w_low) * G));
                    mask = g.mask;
                }
            }
            s_rec[warp][lane] = rec;
            s_mask[warp][lane] = mask;
        }
        __syncwarp();

#pragma unroll
        for (int pp = 0; pp < kMsmvChunk; ++pp) {
            float4 g;
            if (grouped)
                g = (p0 + pp < a.P) ? __ldg(reinterpret_cast<const float4*>(go_q + (size_t)(p0 + pp) * 64) + j)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
            else
                g = *reinterpret_cast<const float4*>(&s_g[warp][pp][4 * j]);
            float glx = 0.f, gly = 0.f;
            float gw[L];
#pragma unroll
            for (int l = 0; l < L; ++l) {
                const int t = pp * L + l;
                gw[l] = 0.f;
                const unsigned mask = s_mask[warp][t];
                if (mask == 0u) continue;  // warp-uniform
                const float4 rec = s_rec[warp][t];
                const float lh = rec.x, lw = rec.y, s = rec.z;
                const int off = __float_as_int(rec.w);
                const float hh = 1.f - lh, hw = 1.f - lw;
                const float wx = slot ? lw : hw;
                const float w_top = hh * wx, w_bot = lh * wx;
                const unsigned m = mask >> slot;
                const float4* p = base[l] + off;
                float* gp = gbase[l] + (size_t)off * 4;
                float4 top = make_float4(0.f, 0.f, 0.f, 0.f), bot = top;
                if (m & 1u) top = ldg128(p);
                if (m & 4u) bot = ldg128(p + row[l]);
                if (m & 1u) {
                    const float k = w_top * s;
                    red_add_v4(gp, k * g.x, k * g.y, k * g.z, k * g.w);
                }
                if (m & 4u) {
                    const float k = w_bot * s;
                    red_add_v4(gp + (size_t)row[l] * 4, k * g.x, k * g.y, k * g.z, k * g.w);
                }
                const float A = dot4(g, top), Bv = dot4(g, bot);
                gw[l] = fmaf(w_top, A, w_bot * Bv);                       // -> grad_weights
                gly = fmaf(sy_scale[l] * s, wx * (Bv - A), gly);           // d/dy
                glx = fmaf(sx_scale[l] * s, sgn * fmaf(hh, A, lh * Bv), glx);  // d/dx
            }
            const int p = p0 + pp;
            if (p < a.P) {  // warp-uniform
                glx = warp_sum(glx);
                gly = warp_sum(gly);
#pragma unroll
                for (int l = 0; l < L; ++l) gw[l] = warp_sum(gw[l]);
                if (lane == 0) {
                    gl_q[p * 3 + 0] = glx;
                    gl_q[p * 3 + 1] = gly;
                    gl_q[p * 3 + 2] = 0.f;
#pragma unroll
                    for (int l = 0; l < L; ++l) gw_q[p * L + l] = gw[l];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Generic path (any C, any L <= RACF_MAX_LEVELS): one thread per (b,q,c) / (b,q,p,c). Slow but complete.
// ------------------------------------------------------------------------------------------------
struct MsmvArgsDyn {
    const float* feat[RACF_MAX_LEVELS];
    float* grad_feat[RACF_MAX_LEVELS];
    int H[RACF_MAX_LEVELS];
    int W[RACF_MAX_LEVELS];
    const float* loc;
    const float* wts;
    const float* grad_out;
    float* out;
    float* grad_loc;
    float* grad_wts;
    int B, N, Q, P, C, L;
};

__global__ void __launch_bounds__(256) msmv_fwd_generic_kernel(const MsmvArgsDyn a) {
    __shared__ int sH[RACF_MAX_LEVELS], sW[RACF_MAX_LEVELS];
    if (threadIdx.x < a.L) { sH[threadIdx.x] = a.H[threadIdx.x]; sW[threadIdx.x] = a.W[threadIdx.x]; }
    __syncthreads();
    const long long total = (long long)a.B * a.Q * a.C;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % a.C);
        const long long bq = idx / a.C;
        const int b = (int)(bq / a.Q);
        for (int p = 0; p < a.P; ++p) {
            const float* lp = a.loc + (bq * a.P + p) * 3;
            const float lx = lp[0], ly = lp[1];
            const int v = msmv_view(lp[2], a.N);
            float res = 0.f;
            if (v >= 0 && v < a.N) {
                for (int l = 0; l < a.L; ++l) {
                    const int H = sH[l], W = sW[l];
                    const TapGeom g = tap_geometry(msmv_pixel(ly, H), msmv_pixel(lx, W), H, W);
                    if (g.mask == 0u) continue;
                    const float* f = nullptr;
#pragma unroll
                    for (int k = 0; k < RACF_MAX_LEVELS; ++k)
                        if (k == l) f = a.feat[k];
                    f += (((size_t)b * a.N + v) * H * W) * a.C + c;
                    const size_t o = ((size_t)g.h_low * W + g.w_low) * a.C;  // may wrap for -1: only used when valid
                    const size_t rs = (size_t)W * a.C;
                    const float hh = 1.f - g.lh, hw = 1.f - g.lw;
                    const float v1 = (g.mask & kTL) ? __ldg(f + o) : 0.f;
                    const float v2 = (g.mask & kTR) ? __ldg(f + o + a.C) : 0.f;
                    const float v3 = (g.mask & kBL) ? __ldg(f + o + rs) : 0.f;
                    const float v4 = (g.mask & kBR) ? __ldg(f + o + rs + a.C) : 0.f;
                    const float val = hh * hw * v1 + hh * g.lw * v2 + g.lh * hw * v3 + g.lh * g.lw * v4;
                    res = fmaf(val, a.wts[(bq * a.P + p) * a.L + l], res);
                }
            }
            a.out[idx * a.P + p] = res;
        }
    }
}

// grad_loc / grad_wts must be zero on entry (the launcher memsets them).
__global__ void __launch_bounds__(256) msmv_bwd_generic_kernel(const MsmvArgsDyn a) {
    __shared__ int sH[RACF_MAX_LEVELS], sW[RACF_MAX_LEVELS];
    if (threadIdx.x < a.L) { sH[threadIdx.x] = a.H[threadIdx.x]; sW[threadIdx.x] = a.W[threadIdx.x]; }
    __syncthreads();
    const long long total = (long long)a.B * a.Q * a.P * a.C;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % a.C);
        const long long bqp = idx / a.C;
        const int p = (int)(bqp % a.P);
        const long long bq = bqp / a.P;
        const int b = (int)(bq / a.Q);
        const float go = a.grad_out[(bq * a.C + c) * a.P + p];
        const float* lp = a.loc + bqp * 3;
        const float lx = lp[0], ly = lp[1];
        const int v = msmv_view(lp[2], a.N);
        if (v < 0 || v >= a.N) continue;
        float glx = 0.f, gly = 0.f;
        for (int l = 0; l < a.L; ++l) {
            const int H = sH[l], W = sW[l];
            const TapGeom g = tap_geometry(msmv_pixel(ly, H), msmv_pixel(lx, W), H, W);
            if (g.mask == 0u) continue;
            const float* f = nullptr;
            float* gf = nullptr;
#pragma unroll
            for (int k = 0; k < RACF_MAX_LEVELS; ++k)
                if (k == l) { f = a.feat[k]; gf = a.grad_feat[k]; }
            const size_t vb = (((size_t)b * a.N + v) * H * W) * a.C + c;
            f += vb;
            gf += vb;
            const size_t o = ((size_t)g.h_low * W + g.w_low) * a.C;
            const size_t rs = (size_t)W * a.C;
            const float s = a.wts[bqp * a.L + l];
            const float hh = 1.f - g.lh, hw = 1.f - g.lw;
            const float tv = go * s;
            float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
            if (g.mask & kTL) { v1 = __ldg(f + o); atomicAdd(gf + o, hh * hw * tv); }
            if (g.mask & kTR) { v2 = __ldg(f + o + a.C); atomicAdd(gf + o + a.C, hh * g.lw * tv); }
            if (g.mask & kBL) { v3 = __ldg(f + o + rs); atomicAdd(gf + o + rs, g.lh * hw * tv); }
            if (g.mask & kBR) { v4 = __ldg(f + o + rs + a.C); atomicAdd(gf + o + rs + a.C, g.lh * g.lw * tv); }
            const float val = hh * hw * v1 + hh * g.lw * v2 + g.lh * hw * v3 + g.lh * g.lw * v4;
            const float gh = -hw * v1 - g.lw * v2 + hw * v3 + g.lw * v4;
            const float gwd = -hh * v1 + hh * v2 - g.lh * v3 + g.lh * v4;
            atomicAdd(a.grad_wts + bqp * a.L + l, go * val);
            glx = fmaf((float)(W - 1) * gwd, tv, glx);
            gly = fmaf((float)(H - 1) * gh, tv, gly);
        }
        atomicAdd(a.grad_loc + bqp * 3 + 0, glx);
        atomicAdd(a.grad_loc + bqp * 3 + 1, gly);
    }
}

struct MaskArgs {
    int H[RACF_MAX_LEVELS];
    int W[RACF_MAX_LEVELS];
    const float* loc;
    int32_t* view;
    uint8_t* mask;
    long long npts;
    int N, L;
};

__global__ void __launch_bounds__(256) msmv_mask_kernel(const MaskArgs a) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.npts;
         i += (long long)gridDim.x * blockDim.x) {
        const float lx = a.loc[i * 3], ly = a.loc[i * 3 + 1];
        if (a.view) a.view[i] = msmv_view(a.loc[i * 3 + 2], a.N);
        if (a.mask) {
#pragma unroll
            for (int l = 0; l < RACF_MAX_LEVELS; ++l) {
                if (l < a.L) {
                    const TapGeom g = tap_geometry(msmv_pixel(ly, a.H[l]), msmv_pixel(lx, a.W[l]), a.H[l], a.W[l]);
                    a.mask[i * a.L + l] = (uint8_t)((g.in_range ? 1u : 0u) | (g.mask << 1));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Host launchers
// ------------------------------------------------------------------------------------------------
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int check_msmv_common(const float* const* feats, const int* hw, int L, const float* loc, const float* wts,
                             int B, int C, int N, int Q, int P) {
    if (!feats || !hw || !loc || !wts) return RACF_ERR_NULL_POINTER;
    if (L < 1 || L > RACF_MAX_LEVELS) return RACF_ERR_BAD_LEVELS;
    if (B <= 0 || C <= 0 || N <= 0 || Q <= 0 || P <= 0) return RACF_ERR_BAD_SHAPE;
    if (P > RACF_MSMV_MAX_POINT) return RACF_ERR_TOO_MANY_PTS;
    for (int l = 0; l < L; ++l) {
        if (!feats[l]) return RACF_ERR_NULL_POINTER;
        if (hw[2 * l] <= 0 || hw[2 * l + 1] <= 0) return RACF_ERR_BAD_SHAPE;
        // per-batch-element view stack is indexed with int32 float4 offsets on the fast path
        if ((long long)N * hw[2 * l] * hw[2 * l + 1] * C >= (1LL << 31)) return RACF_ERR_BAD_SHAPE;
    }
    return RACF_OK;
}

// Forward variants (explicit argument of racf_msmv_forward_variant; the plain entry points use kMsmvFwdDefault):
// 0 one-query-per-warp kernel, -1 persistent kernel without prefetch (measured best on B200,
// profiles/r01_msmv_fwd_variants.txt), k > 0 persistent kernel prefetching the first k levels into L2.
constexpr int kMsmvFwdDefault = -1;

template <int L>
static int launch_fast(bool backward, const float* grad_out, const float* const* feats, float* const* grad_feats,
                       const int* hw, const float* loc, const float* wts, int B, int C, int N, int Q, int P,
                       float* out, float* grad_loc, float* grad_wts, cudaStream_t st, int out_T = 0, int out_G = 0,
                       int mode = kMsmvFwdDefault) {
    MsmvArgs<L> a;
    a.out_T = out_T;
    a.out_G = out_G;
    for (int l = 0; l < L; ++l) {
        a.feat[l] = feats[l];
        a.grad_feat[l] = grad_feats ? grad_feats[l] : nullptr;
        a.H[l] = hw[2 * l];
        a.W[l] = hw[2 * l + 1];
    }
    a.loc = loc; a.wts = wts; a.grad_out = grad_out; a.out = out; a.grad_loc = grad_loc; a.grad_wts = grad_wts;
    a.B = B; a.N = N; a.Q = Q; a.P = P; a.C = C;
    const long long nq = (long long)B * Q;
    const unsigned grid = (unsigned)((nq + kMsmvWarps - 1) / kMsmvWarps);
    if (backward) {
        msmv_bwd_c64_kernel<L><<<grid, kMsmvWarps * 32, 0, st>>>(a);
    } else {
        constexpr int PB = (L <= 4) ? 16 : 12;
        bool fits32 = true;   // the persistent kernel indexes whole levels (all batch elements) with 32-bit float4 offsets
        for (int l = 0; l < L; ++l)
            if ((long long)B * N * hw[2 * l] * hw[2 * l + 1] * 16 >= (1LL << 31)) fits32 = false;
        if (mode != 0 && P <= PB && fits32) {
            const unsigned cap = 2u * (unsigned)sm_count();               // persistent: 2 CTAs per SM
            const unsigned pgrid = grid < cap ? grid : cap;
            msmv_fwd_c64_pf_kernel<L><<<pgrid, kMsmvWarps * 32, 0, st>>>(a, mode < 0 ? 0 : (mode > L ? L : mode));
        } else {
            msmv_fwd_c64_kernel<L><<<grid, kMsmvWarps * 32, 0, st>>>(a);
        }
    }
    return (int)cudaGetLastError();
}

static int launch_generic(bool backward, const float* grad_out, const float* const* feats, float* const* grad_feats,
                          const int* hw, int L, const float* loc, const float* wts, int B, int C, int N, int Q,
                          int P, float* out, float* grad_loc, float* grad_wts, cudaStream_t st) {
    MsmvArgsDyn a;
    for (int l = 0; l < RACF_MAX_LEVELS; ++l) {
        a.feat[l] = l < L ? feats[l] : nullptr;
        a.grad_feat[l] = (l < L && grad_feats) ? grad_feats[l] : nullptr;
        a.H[l] = l < L ? hw[2 * l] : 1;
        a.W[l] = l < L ? hw[2 * l + 1] : 1;
    }
    a.loc = loc; a.wts = wts; a.grad_out = grad_out; a.out = out; a.grad_loc = grad_loc; a.grad_wts = grad_wts;
    a.B = B; a.N = N; a.Q = Q; a.P = P; a.C = C; a.L = L;
    const long long total = (long long)B * Q * C * (backward ? P : 1);
    const long long gcap = 64LL * sm_count();
    const unsigned grid = (unsigned)((total + 255) / 256 > gcap ? gcap : (total + 255) / 256);
    if (backward) {
        cudaError_t e = cudaMemsetAsync(grad_loc, 0, sizeof(float) * (size_t)B * Q * P * 3, st);
        if (e != cudaSuccess) return (int)e;
        e = cudaMemsetAsync(grad_wts, 0, sizeof(float) * (size_t)B * Q * P * L, st);
        if (e != cudaSuccess) return (int)e;
        msmv_bwd_generic_kernel<<<grid, 256, 0, st>>>(a);
    } else {
        msmv_fwd_generic_kernel<<<grid, 256, 0, st>>>(a);
    }
    return (int)cudaGetLastError();
}

static bool fast_ok(const float* const* feats, float* const* grad_feats, int L, int C, const float* io) {
    if (C != 64 || !(L == 2 || L == 4 || L == 5) || !aligned16(io)) return false;
    for (int l = 0; l < L; ++l) {
        if (!aligned16(feats[l])) return false;
        if (grad_feats && !aligned16(grad_feats[l])) return false;
    }
    return true;
}

}  // namespace racf

using namespace racf;

extern "C" int racf_msmv_forward_variant(const float* const* feats, const int* hw, int num_levels, const float* loc,
                                         const float* weights, int batch, int channels, int num_views, int num_query,
                                         int num_point, int variant, float* out, racf_stream_t stream) {
    int rc = check_msmv_common(feats, hw, num_levels, loc, weights, batch, channels, num_views, num_query, num_point);
    if (rc != RACF_OK) return rc;
    if (!out) return RACF_ERR_NULL_POINTER;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (fast_ok(feats, nullptr, num_levels, channels, out)) {
        switch (num_levels) {
            case 2: return launch_fast<2>(false, nullptr, feats, nullptr, hw, loc, weights, batch, channels, num_views, num_query, num_point, out, nullptr, nullptr, st, 0, 0, variant);
            case 4: return launch_fast<4>(false, nullptr, feats, nullptr, hw, loc, weights, batch, channels, num_views, num_query, num_point, out, nullptr, nullptr, st, 0, 0, variant);
            case 5: return launch_fast<5>(false, nullptr, feats, nullptr, hw, loc, weights, batch, channels, num_views, num_query, num_point, out, nullptr, nullptr, st, 0, 0, variant);
        }
    }
    return launch_generic(false, nullptr, feats, nullptr, hw, num_levels, loc, weights, batch, channels, num_views,
                          num_query, num_point, out, nullptr, nullptr, st);
}

extern "C" int racf_msmv_forward(const float* const* feats, const int* hw, int num_levels, const float* loc,
                                 const float* weights, int batch, int channels, int num_views, int num_query,
                                 int num_point, float* out, racf_stream_t stream) {
    return racf_msmv_forward_variant(feats, hw, num_levels, loc, weights, batch, channels, num_views, num_query,
                                     num_point, kMsmvFwdDefault, out, stream);
}

extern "C" int racf_msmv_forward_grouped(const float* const* feats, const int* hw, int num_levels, const float* loc,
                                         const float* weights, int batch, int channels, int num_views, int num_query,
                                         int num_point, int num_frames, int num_groups, float* out,
                                         racf_stream_t stream) {
    int rc = check_msmv_common(feats, hw, num_levels, loc, weights, batch, channels, num_views, num_query, num_point);
    if (rc != RACF_OK) return rc;
    if (!out) return RACF_ERR_NULL_POINTER;
    if (num_frames <= 0 || num_groups <= 0 || batch % (num_frames * num_groups) != 0) return RACF_ERR_BAD_SHAPE;
    if (!fast_ok(feats, nullptr, num_levels, channels, out)) return RACF_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (num_levels) {
        case 2: return launch_fast<2>(false, nullptr, feats, nullptr, hw, loc, weights, batch, channels, num_views, num_query, num_point, out, nullptr, nullptr, st, num_frames, num_groups);
        case 4: return launch_fast<4>(false, nullptr, feats, nullptr, hw, loc, weights, batch, channels, num_views, num_query, num_point, out, nullptr, nullptr, st, num_frames, num_groups);
        case 5: return launch_fast<5>(false, nullptr, feats, nullptr, hw, loc, weights, batch, channels, num_views, num_query, num_point, out, nullptr, nullptr, st, num_frames, num_groups);
    }
    return RACF_ERR_UNSUPPORTED;
}

extern "C" int racf_msmv_backward(const float* grad_out, const float* const* feats, const int* hw, int num_levels,
                                  const float* loc, const float* weights, int batch, int channels, int num_views,
                                  int num_query, int num_point, float* const* grad_feats, float* grad_loc,
                                  float* grad_weights, int zero_grad_feats, racf_stream_t stream) {
    int rc = check_msmv_common(feats, hw, num_levels, loc, weights, batch, channels, num_views, num_query, num_point);
    if (rc != RACF_OK) return rc;
    if (!grad_out || !grad_feats || !grad_loc || !grad_weights) return RACF_ERR_NULL_POINTER;
    for (int l = 0; l < num_levels; ++l)
        if (!grad_feats[l]) return RACF_ERR_NULL_POINTER;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (zero_grad_feats) {
        for (int l = 0; l < num_levels; ++l) {
            const size_t n = (size_t)batch * num_views * hw[2 * l] * hw[2 * l + 1] * channels;
            cudaError_t e = cudaMemsetAsync(grad_feats[l], 0, n * sizeof(float), st);
            if (e != cudaSuccess) return (int)e;
        }
    }
    if (fast_ok(feats, grad_feats, num_levels, channels, grad_out)) {
        switch (num_levels) {
            case 2: return launch_fast<2>(true, grad_out, feats, grad_feats, hw, loc, weights, batch, channels, num_views, num_query, num_point, nullptr, grad_loc, grad_weights, st);
            case 4: return launch_fast<4>(true, grad_out, feats, grad_feats, hw, loc, weights, batch, channels, num_views, num_query, num_point, nullptr, grad_loc, grad_weights, st);
            case 5: return launch_fast<5>(true, grad_out, feats, grad_feats, hw, loc, weights, batch, channels, num_views, num_query, num_point, nullptr, grad_loc, grad_weights, st);
        }
    }
    return launch_generic(true, grad_out, feats, grad_feats, hw, num_levels, loc, weights, batch, channels, num_views,
                          num_query, num_point, nullptr, grad_loc, grad_weights, st);
}

extern "C" int racf_msmv_backward_grouped(const float* grad_out, const float* const* feats, const int* hw, int num_levels,
                                          const float* loc, const float* weights, int batch, int channels, int num_views,
                                          int num_query, int num_point, int num_frames, int num_groups,
                                          float* const* grad_feats, float* grad_loc, float* grad_weights,
                                          int zero_grad_feats, racf_stream_t stream) {
    int rc = check_msmv_common(feats, hw, num_levels, loc, weights, batch, channels, num_views, num_query, num_point);
    if (rc != RACF_OK) return rc;
    if (!grad_out || !grad_feats || !grad_loc || !grad_weights) return RACF_ERR_NULL_POINTER;
    for (int l = 0; l < num_levels; ++l)
        if (!grad_feats[l]) return RACF_ERR_NULL_POINTER;
    if (num_frames <= 0 || num_groups <= 0 || batch % (num_frames * num_groups) != 0) return RACF_ERR_BAD_SHAPE;
    if (!fast_ok(feats, grad_feats, num_levels, channels, grad_out)) return RACF_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (zero_grad_feats) {
        for (int l = 0; l < num_levels; ++l) {
            const size_t n = (size_t)batch * num_views * hw[2 * l] * hw[2 * l + 1] * channels;
            cudaError_t e = cudaMemsetAsync(grad_feats[l], 0, n * sizeof(float), st);
            if (e != cudaSuccess) return (int)e;
        }
    }
    switch (num_levels) {
        case 2: return launch_fast<2>(true, grad_out, feats, grad_feats, hw, loc, weights, batch, channels, num_views, num_query, num_point, nullptr, grad_loc, grad_weights, st, num_frames, num_groups);
        case 4: return launch_fast<4>(true, grad_out, feats, grad_feats, hw, loc, weights, batch, channels, num_views, num_query, num_point, nullptr, grad_loc, grad_weights, st, num_frames, num_groups);
        case 5: return launch_fast<5>(true, grad_out, feats, grad_feats, hw, loc, weights, batch, channels, num_views, num_query, num_point, nullptr, grad_loc, grad_weights, st, num_frames, num_groups);
    }
    return RACF_ERR_UNSUPPORTED;
}

extern "C" int racf_msmv_tap_masks(const int* hw, int num_levels, const float* loc, int batch, int num_views,
                                   int num_query, int num_point, int32_t* view_index, uint8_t* tap_mask,
                                   racf_stream_t stream) {
    if (!hw || !loc) return RACF_ERR_NULL_POINTER;
    if (num_levels < 1 || num_levels > RACF_MAX_LEVELS) return RACF_ERR_BAD_LEVELS;
    if (batch <= 0 || num_views <= 0 || num_query <= 0 || num_point <= 0) return RACF_ERR_BAD_SHAPE;
    MaskArgs a;
    for (int l = 0; l < RACF_MAX_LEVELS; ++l) {
        a.H[l] = l < num_levels ? hw[2 * l] : 1;
        a.W[l] = l < num_levels ? hw[2 * l + 1] : 1;
    }
    a.loc = loc; a.view = view_index; a.mask = tap_mask;
    a.npts = (long long)batch * num_query * num_point;
    a.N = num_views; a.L = num_levels;
    const long long gcap = 32LL * sm_count();
    const unsigned grid = (unsigned)((a.npts + 255) / 256 > gcap ? gcap : (a.npts + 255) / 256);
    msmv_mask_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return (int)cudaGetLastError();
}
