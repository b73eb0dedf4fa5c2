// BEVPoolv2 (Lift-Splat-Shoot voxel pooling) for B200 (sm_100a) -- the reference's other in-tree native op
// (SURVEY.md 8f-4). Semantics: models/csrc/bev_pool_v2/src/bev_pool_cuda.cu:21-48 (forward), :67-121 (backward).
//
//   forward : for every interval (run of points that fall into the same BEV cell, points pre-sorted by ranks_bev)
//             out[ranks_bev[start], :] = sum_i feat[ranks_feat[start+i], :] * depth[ranks_depth[start+i]]
//   backward: intervals are runs of equal ranks_feat (the Python side re-sorts, bev_pool.py:50-60)
//             depth_grad[ranks_depth[p]]   = < out_grad[ranks_bev[p], :], feat[ranks_feat[p], :] >      per point
//             feat_grad[ranks_feat[start]] = sum_i out_grad[ranks_bev[start+i], :] * depth[ranks_depth[start+i]]
//
// Reference mapping: forward one thread per (interval, channel) -> every channel thread re-reads the three rank
// arrays and the depth; backward ONE thread per interval looping over all points and all C channels serially.
// Here one warp owns an interval: lanes hold 4*NV consecutive channels each as float4 (a C = 256 feature row is two
// coalesced 512-byte loads per point), rank/depth values are loaded once per point by one lane group and broadcast,
// the per-channel sums run over the points in the reference's order with explicit fmaf (bit-identical to the
// reference's contracted `psum += feat * depth`), and the backward depth-gradient dot product is a warp shuffle tree.
#include <cuda_runtime.h>
#include <stdint.h>

#include "racf_common.cuh"
#include "racformer_ops.h"

namespace racf {

constexpr int kPoolWarps = 8;

struct PoolArgs {
    const float* depth;
    const float* feat;
    const float* out_grad;
    const int* ranks_depth;
    const int* ranks_feat;
    const int* ranks_bev;
    const int* starts;
    const int* lengths;
    float* out;
    float* depth_grad;
    float* feat_grad;
    int n_intervals, c;
};

// NV float4 per lane: covers C <= 128 * NV channels, C % 4 == 0
template <int NV>
__global__ void __launch_bounds__(kPoolWarps * 32) bev_pool_fwd_kernel(const PoolArgs a) {
    const int lane = threadIdx.x & 31;
    const int iv = blockIdx.x * kPoolWarps + (threadIdx.x >> 5);
    if (iv >= a.n_intervals) return;
    const int start = a.starts[iv], len = a.lengths[iv];
    const int c4 = a.c >> 2;
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i0 = 0; i0 < len; i0 += 32) {
        // one coalesced read of up to 32 points' ranks + depths, broadcast point by point
        const int n = min(32, len - i0);
        int rf = 0;
        float d = 0.f;
        if (lane < n) {
            rf = __ldg(a.ranks_feat + start + i0 + lane);
            d = __ldg(a.depth + __ldg(a.ranks_depth + start + i0 + lane));
        }
        for (int i = 0; i < n; ++i) {
            const int rfi = __shfl_sync(0xffffffffu, rf, i);
            const float di = __shfl_sync(0xffffffffu, d, i);
            const float4* f = reinterpret_cast<const float4*>(a.feat + (size_t)rfi * a.c);
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int k = lane + 32 * v;
                if (k < c4) {
                    const float4 x = __ldg(f + k);
                    acc[v].x = fmaf(x.x, di, acc[v].x);
                    acc[v].y = fmaf(x.y, di, acc[v].y);
                    acc[v].z = fmaf(x.z, di, acc[v].z);
                    acc[v].w = fmaf(x.w, di, acc[v].w);
                }
            }
        }
    }
    float4* o = reinterpret_cast<float4*>(a.out + (size_t)__ldg(a.ranks_bev + start) * a.c);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int k = lane + 32 * v;
        if (k < c4) o[k] = acc[v];
    }
}

template <int NV>
__global__ void __launch_bounds__(kPoolWarps * 32) bev_pool_bwd_kernel(const PoolArgs a) {
    const int lane = threadIdx.x & 31;
    const int iv = blockIdx.x * kPoolWarps + (threadIdx.x >> 5);
    if (iv >= a.n_intervals) return;
    const int start = a.starts[iv], len = a.lengths[iv];
    const int c4 = a.c >> 2;
    const int rf0 = __ldg(a.ranks_feat + start);
    const float4* f = reinterpret_cast<const float4*>(a.feat + (size_t)rf0 * a.c);
    float4 fv[NV], acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int k = lane + 32 * v;
        fv[v] = k < c4 ? __ldg(f + k) : make_float4(0.f, 0.f, 0.f, 0.f);   // all points of the run share this pixel
        acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i0 = 0; i0 < len; i0 += 32) {
        const int n = min(32, len - i0);
        int rb = 0, rd = 0;
        float d = 0.f, my_dgrad = 0.f;
        if (lane < n) {
            rb = __ldg(a.ranks_bev + start + i0 + lane);
            rd = __ldg(a.ranks_depth + start + i0 + lane);
            d = __ldg(a.depth + rd);
        }
        for (int i = 0; i < n; ++i) {
            const int rbi = __shfl_sync(0xffffffffu, rb, i);
            const float di = __shfl_sync(0xffffffffu, d, i);
            const float4* g = reinterpret_cast<const float4*>(a.out_grad + (size_t)rbi * a.c);
            float dot = 0.f;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int k = lane + 32 * v;
                if (k < c4) {
                    const float4 x = __ldg(g + k);
                    dot = fmaf(x.x, fv[v].x, fmaf(x.y, fv[v].y, fmaf(x.z, fv[v].z, fmaf(x.w, fv[v].w, dot))));
                    acc[v].x = fmaf(x.x, di, acc[v].x);
                    acc[v].y = fmaf(x.y, di, acc[v].y);
                    acc[v].z = fmaf(x.z, di, acc[v].z);
                    acc[v].w = fmaf(x.w, di, acc[v].w);
                }
            }
            dot = warp_sum(dot);
            if (lane == i) my_dgrad = dot;
        }
        if (lane < n) a.depth_grad[rd] = my_dgrad;   // one coalesced-ish store per 32 points
    }
    float4* o = reinterpret_cast<float4*>(a.feat_grad + (size_t)rf0 * a.c);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int k = lane + 32 * v;
        if (k < c4) o[k] = acc[v];
    }
}

// Any channel count: one warp per interval, lanes stride over channels with scalar loads.
__global__ void __launch_bounds__(kPoolWarps * 32) bev_pool_fwd_generic_kernel(const PoolArgs a) {
    const int lane = threadIdx.x & 31;
    const int iv = blockIdx.x * kPoolWarps + (threadIdx.x >> 5);
    if (iv >= a.n_intervals) return;
    const int start = a.starts[iv], len = a.lengths[iv];
    float* o = a.out + (size_t)a.ranks_bev[start] * a.c;
    for (int c = lane; c < a.c; c += 32) {
        float psum = 0.f;
        for (int i = 0; i < len; ++i)
            psum = fmaf(a.feat[(size_t)a.ranks_feat[start + i] * a.c + c], a.depth[a.ranks_depth[start + i]], psum);
        o[c] = psum;
    }
}

__global__ void __launch_bounds__(kPoolWarps * 32) bev_pool_bwd_generic_kernel(const PoolArgs a) {
    const int lane = threadIdx.x & 31;
    const int iv = blockIdx.x * kPoolWarps + (threadIdx.x >> 5);
    if (iv >= a.n_intervals) return;
    const int start = a.starts[iv], len = a.lengths[iv];
    const float* f = a.feat + (size_t)a.ranks_feat[start] * a.c;
    for (int i = 0; i < len; ++i) {
        const float* g = a.out_grad + (size_t)a.ranks_bev[start + i] * a.c;
        float dot = 0.f;
        for (int c = lane; c < a.c; c += 32) dot = fmaf(g[c], f[c], dot);
        dot = warp_sum(dot);
        if (lane == 0) a.depth_grad[a.ranks_depth[start + i]] = dot;
    }
    float* o = a.feat_grad + (size_t)a.ranks_feat[start] * a.c;
    for (int c = lane; c < a.c; c += 32) {
        float s = 0.f;
        for (int i = 0; i < len; ++i)
            s = fmaf(a.out_grad[(size_t)a.ranks_bev[start + i] * a.c + c], a.depth[a.ranks_depth[start + i]], s);
        o[c] = s;
    }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <bool BWD>
static int launch_pool(const PoolArgs& a, bool vec_ok, cudaStream_t st) {
    const unsigned grid = (unsigned)((a.n_intervals + kPoolWarps - 1) / kPoolWarps);
    const int nv = (a.c + 127) / 128;
    if (vec_ok && nv <= 8) {
#define RACF_POOL_CASE(NV)                                                         \
    if (nv <= NV) {                                                                \
        if (BWD) bev_pool_bwd_kernel<NV><<<grid, kPoolWarps * 32, 0, st>>>(a);     \
        else bev_pool_fwd_kernel<NV><<<grid, kPoolWarps * 32, 0, st>>>(a);         \
        return (int)cudaGetLastError();                                            \
    }
        RACF_POOL_CASE(1) RACF_POOL_CASE(2) RACF_POOL_CASE(4) RACF_POOL_CASE(8)
#undef RACF_POOL_CASE
    }
    if (BWD) bev_pool_bwd_generic_kernel<<<grid, kPoolWarps * 32, 0, st>>>(a);
    else bev_pool_fwd_generic_kernel<<<grid, kPoolWarps * 32, 0, st>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace racf

using namespace racf;

extern "C" int racf_bev_pool_v2_forward(const float* depth, const float* feat, const int* ranks_depth,
                                        const int* ranks_feat, const int* ranks_bev, const int* interval_starts,
                                        const int* interval_lengths, int n_intervals, int channels, float* out,
                                        racf_stream_t stream) {
    if (!depth || !feat || !ranks_depth || !ranks_feat || !ranks_bev || !interval_starts || !interval_lengths || !out)
        return RACF_ERR_NULL_POINTER;
    if (n_intervals < 0 || channels <= 0) return RACF_ERR_BAD_SHAPE;
    if (n_intervals == 0) return RACF_OK;
    PoolArgs a{};
    a.depth = depth; a.feat = feat; a.ranks_depth = ranks_depth; a.ranks_feat = ranks_feat; a.ranks_bev = ranks_bev;
    a.starts = interval_starts; a.lengths = interval_lengths; a.out = out; a.n_intervals = n_intervals; a.c = channels;
    return launch_pool<false>(a, channels % 4 == 0 && al16(feat) && al16(out), static_cast<cudaStream_t>(stream));
}

extern "C" int racf_bev_pool_v2_backward(const float* out_grad, const float* depth, const float* feat,
                                         const int* ranks_depth, const int* ranks_feat, const int* ranks_bev,
                                         const int* interval_starts, const int* interval_lengths, int n_intervals,
                                         int channels, float* depth_grad, float* feat_grad, racf_stream_t stream) {
    if (!out_grad || !depth || !feat || !ranks_depth || !ranks_feat || !ranks_bev || !interval_starts ||
        !interval_lengths || !depth_grad || !feat_grad)
        return RACF_ERR_NULL_POINTER;
    if (n_intervals < 0 || channels <= 0) return RACF_ERR_BAD_SHAPE;
    if (n_intervals == 0) return RACF_OK;
    PoolArgs a{};
    a.out_grad = out_grad; a.depth = depth; a.feat = feat; a.ranks_depth = ranks_depth; a.ranks_feat = ranks_feat;
    a.ranks_bev = ranks_bev; a.starts = interval_starts; a.lengths = interval_lengths; a.depth_grad = depth_grad;
    a.feat_grad = feat_grad; a.n_intervals = n_intervals; a.c = channels;
    return launch_pool<true>(a, channels % 4 == 0 && al16(feat) && al16(out_grad) && al16(feat_grad),
                             static_cast<cudaStream_t>(stream));
}
