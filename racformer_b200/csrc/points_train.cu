// Backward kernels of the fused sampling-point chain and of the box refinement (training; SURVEY.md 8f-2 / a6 under
// autograd). The forward kernels are the ones of csrc/points.cu; these turn the location / weight gradients the sampling
// ops return (msmv_sampling backward: grad_loc [B*T*G,Q,P,3], grad_weights [B*G*T,Q,P,L]; MSDA backward: grad_loc
// [T*B,Q,M,1,P,2], grad_attn [T*B,Q,M,1,P]) into gradients of the four Linear heads' outputs and of the query ray --
// what PyTorch does with ~250 autograd nodes per branch and iteration (models/racformer_transformer.py:361-408, 493-529,
// models/sparsebev_sampling.py:8-25, 45-120).
//
// One CTA per query (b, q): a thread owns a box-relative point (g, p) and walks its T frames, so the offset gradient is a
// register sum written once; the ray-logit gradients (shared by the points of a depth bin) and the query-frame gradients
// (shared by all points) are reduced in shared memory. The per-point rules live in points_bwd.cuh (also compiled for the
// host by the CPU test). The discrete view choice is read back from loc[...,2].
#include "points_bwd.cuh"
#include "racf_common.cuh"
#include "racformer_ops.h"

namespace racf {
namespace ptbwd {

struct ImgBwdArgs {
    const float *ray, *offset, *ray_logit, *time_diff, *lidar2img, *depth_base;
    const float *loc, *weights;            // forward outputs (view index, softmax values)
    const float *grad_loc, *grad_weights;
    float *g_ray, *g_offset, *g_logit, *g_scale;
    Consts k;
    float image_w, image_h, eps;
    int B, Q, T, G, Pn, N, L;
};

__device__ __forceinline__ void reduce_query_grad(const QueryGrad& acc, float* sm_q) {
    float v[8] = {acc.cx, acc.cy, acc.cz, acc.w, acc.l, acc.h, acc.s, acc.c};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sm_q[i], x);
    }
}

__global__ void __launch_bounds__(128) msmv_points_bwd_kernel(const ImgBwdArgs a) {
    extern __shared__ float sm[];          // [D] logit gradients, [8] query-frame gradients
    const int D = a.k.D, P = a.Pn * D, GP = a.G * P;
    float* sm_q = sm + D;
    for (int i = threadIdx.x; i < D + 8; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const long long bq = blockIdx.x;
    const int b = (int)(bq / a.Q), q = (int)(bq % a.Q);
    const QueryFrame f = decode(a.ray + bq * 10, a.k);
    QueryGrad acc;
    zero(acc);
    for (int gp = threadIdx.x; gp < GP; gp += blockDim.x) {
        const int g = gp / P, p = gp % P, d = p % D;
        const float* off = a.offset + (bq * GP + gp) * 3;
        const float ox = off[0], oy = off[1], oz = off[2];
        const float logit = a.ray_logit[bq * D + d], base = a.depth_base[d];
        float s_ox = 0.f, s_oy = 0.f, s_oz = 0.f, s_logit = 0.f;
        for (int t = 0; t < a.T; ++t) {
            const float td = a.time_diff[b * a.T + t];
            const PointFwd pf = point_forward(f, ox, oy, oz, logit, base, td, a.k);
            const long long row = ((((long long)b * a.T + t) * a.G + g) * a.Q + q) * P + p;
            const float gu = a.grad_loc[row * 3], gv = a.grad_loc[row * 3 + 1];
            const int view = __float2int_rn(a.loc[row * 3 + 2] * (float)(a.N - 1));
            const float* m = a.lidar2img + (((long long)b * a.T + t) * a.N + view) * 16;
            float g_x2, g_y2, g_Z, g_ox, g_oy, g_oz, g_lg;
            project_backward(pf, m, gu, gv, a.image_w, a.image_h, a.eps, a.k, g_x2, g_y2, g_Z);
            point_backward(f, pf, ox, oy, oz, g_x2, g_y2, g_Z, a.k, g_ox, g_oy, g_oz, g_lg, acc);
            s_ox += g_ox; s_oy += g_oy; s_oz += g_oz; s_logit += g_lg;
            // softmax over levels: weights[(b*G+g)*T+t, q, p, :] came from scale_raw[b,q,g,t,p,:]   (G-major quirk kept)
            const long long wrow = (((((long long)b * a.G + g) * a.T + t) * a.Q + q) * P + p) * a.L;
            float dot = 0.f;
            for (int l = 0; l < a.L; ++l) dot += a.grad_weights[wrow + l] * a.weights[wrow + l];
            float* gs = a.g_scale + ((((bq * a.G + g) * a.T + t) * P) + p) * a.L;
            for (int l = 0; l < a.L; ++l) gs[l] = a.weights[wrow + l] * (a.grad_weights[wrow + l] - dot);
        }
        float* go = a.g_offset + (bq * GP + gp) * 3;
        go[0] = s_ox; go[1] = s_oy; go[2] = s_oz;
        atomicAdd(&sm[d], s_logit);
    }
    reduce_query_grad(acc, sm_q);
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) a.g_logit[bq * D + i] = sm[i];
    if (threadIdx.x == 0 && a.g_ray != nullptr) {
        QueryGrad tot;
        tot.cx = sm_q[0]; tot.cy = sm_q[1]; tot.cz = sm_q[2]; tot.w = sm_q[3]; tot.l = sm_q[4]; tot.h = sm_q[5];
        tot.s = sm_q[6]; tot.c = sm_q[7];
        query_backward(f, tot, a.k, a.g_ray + bq * 10);
    }
}

struct BevBwdArgs {
    const float *ray, *offset, *ray_logit, *time_diff, *depth_base;
    const float *attn;                     // forward output [T*B,Q,M,1,P]; the t = 0 slice holds softmax(attn_raw[b])
    const float *grad_loc, *grad_attn;
    float *g_ray, *g_offset, *g_logit, *g_attn_raw;
    Consts k;
    int B, Q, T, M, Pn;
};

__global__ void __launch_bounds__(128) bev_points_bwd_kernel(const BevBwdArgs a) {
    extern __shared__ float sm[];          // [D] logit gradients, [8] query-frame gradients, [M*P] summed weight gradients
    const int D = a.k.D, P = a.Pn * D, MP = a.M * P;
    float* sm_q = sm + D;
    float* sm_ga = sm + D + 8;
    for (int i = threadIdx.x; i < D + 8; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const long long bq = blockIdx.x;
    const int b = (int)(bq / a.Q), q = (int)(bq % a.Q);
    const QueryFrame f = decode(a.ray + bq * 10, a.k);
    QueryGrad acc;
    zero(acc);
    const long long frame_stride = (long long)a.B * a.Q * MP;       // outputs are queue-major: (t*B + b, q, m, p)
    for (int mp = threadIdx.x; mp < MP; mp += blockDim.x) {
        const int d = (mp % P) % D;
        const float* off = a.offset + (bq * MP + mp) * 2;
        const float ox = off[0], oy = off[1];
        const float logit = a.ray_logit[bq * D + d], base = a.depth_base[d];
        float s_ox = 0.f, s_oy = 0.f, s_logit = 0.f, s_ga = 0.f;
        for (int t = 0; t < a.T; ++t) {
            const float td = a.time_diff[b * a.T + t];
            const PointFwd pf = point_forward(f, ox, oy, 0.f, logit, base, td, a.k);
            const long long o = (long long)t * frame_stride + bq * MP + mp;
            const float2 gl = *reinterpret_cast<const float2*>(a.grad_loc + o * 2);
            float g_ox, g_oy, g_oz, g_lg;
            point_backward(f, pf, ox, oy, 0.f, gl.x, gl.y, 0.f, a.k, g_ox, g_oy, g_oz, g_lg, acc);
            s_ox += g_ox; s_oy += g_oy; s_logit += g_lg;
            s_ga += a.grad_attn[o];
        }
        *reinterpret_cast<float2*>(a.g_offset + (bq * MP + mp) * 2) = make_float2(s_ox, s_oy);
        atomicAdd(&sm[d], s_logit);
        sm_ga[mp] = s_ga;
    }
    reduce_query_grad(acc, sm_q);
    __syncthreads();
    // softmax over the P points of a head, shared by the T frames
    const float* w = a.attn + bq * MP;
    for (int mp = threadIdx.x; mp < MP; mp += blockDim.x) {
        const int m = mp / P;
        float dot = 0.f;
        for (int kk = 0; kk < P; ++kk) dot += sm_ga[m * P + kk] * w[m * P + kk];
        a.g_attn_raw[bq * MP + mp] = w[mp] * (sm_ga[mp] - dot);
    }
    for (int i = threadIdx.x; i < D; i += blockDim.x) a.g_logit[bq * D + i] = sm[i];
    if (threadIdx.x == 0 && a.g_ray != nullptr) {
        QueryGrad tot;
        tot.cx = sm_q[0]; tot.cy = sm_q[1]; tot.cz = sm_q[2]; tot.w = sm_q[3]; tot.l = sm_q[4]; tot.h = sm_q[5];
        tot.s = sm_q[6]; tot.c = sm_q[7];
        query_backward(f, tot, a.k, a.g_ray + bq * 10);
    }
}

// Box refinement (models/racformer_transformer.py:255-279) + theta_d2xy_coods of the prediction: gradient of pred_xy
// w.r.t. the reg-branch output (and, when asked for, the proposal). One thread per query row.
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float inverse_sigmoid_(float x, float eps, float& dydx) {
    const bool inside = x >= 0.f && x <= 1.f;
    x = fminf(fmaxf(x, 0.f), 1.f);
    const float num = fmaxf(x, eps), den = fmaxf(1.f - x, eps);
    dydx = inside ? ((x >= eps ? 1.f / num : 0.f) + (1.f - x >= eps ? 1.f / den : 0.f)) : 0.f;
    return logf(num / den);
}

__global__ void __launch_bounds__(128)
refine_bbox_bwd_kernel(const float* __restrict__ proposal, const float* __restrict__ delta, const float* __restrict__ time_diff,
                       const float* __restrict__ grad_xy, int rows, int num_query, int num_frames, int code, float num_ray,
                       float* __restrict__ g_delta, float* __restrict__ g_proposal) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const float* p = proposal + (size_t)row * code;
    const float* d = delta + (size_t)row * code;
    const float* g = grad_xy + (size_t)row * code;
    float* gd = g_delta + (size_t)row * code;
    const float s0 = sigmoidf_(d[0]);
    const float theta = p[0] + (s0 * 2.f - 1.f) / num_ray;
    float di1, di2;
    const float dist = sigmoidf_(d[1] + inverse_sigmoid_(p[1], 1e-5f, di1));
    const float z = sigmoidf_(d[2] + inverse_sigmoid_(p[2], 1e-5f, di2));
    const float ang = theta * kTwoPi, r = dist * kRayR, ca = cosf(ang), sa = sinf(ang);
    const float x = (kCentre + r * ca) / kMapSize, y = (kCentre + r * sa) / kMapSize;
    const float gx = (x >= 0.f && x <= 1.f) ? g[0] : 0.f, gy = (y >= 0.f && y <= 1.f) ? g[1] : 0.f;
    const float g_r = (gx * ca + gy * sa) / kMapSize, g_a = (gy * ca - gx * sa) * r / kMapSize;
    const float g_theta = g_a * kTwoPi, g_dist = g_r * kRayR;
    const float g_l1 = g_dist * dist * (1.f - dist), g_l2 = g[2] * z * (1.f - z);
    gd[0] = g_theta * 2.f * s0 * (1.f - s0) / num_ray;
    gd[1] = g_l1;
    gd[2] = g_l2;
    float td = 1.f;
    if (time_diff != nullptr && num_frames > 1) {
        td = time_diff[(size_t)(row / num_query) * num_frames + 1];
        if (td < 1e-5f) td = 1.f;
    }
    for (int c = 3; c < code; ++c) gd[c] = (c >= 8 && time_diff != nullptr && num_frames > 1) ? g[c] / td : g[c];
    if (g_proposal != nullptr) {
        float* gp = g_proposal + (size_t)row * code;
        gp[0] = g_theta;
        gp[1] = g_l1 * di1;
        gp[2] = g_l2 * di2;
        for (int c = 3; c < code; ++c) gp[c] = 0.f;
    }
}

static void fill_consts(Consts& k, const double* pc_range, float d_region, int depth_num) {
    for (int i = 0; i < 6; ++i) k.pc[i] = (float)pc_range[i];
    for (int i = 0; i < 3; ++i) k.span[i] = (float)(pc_range[3 + i] - pc_range[i]);
    k.d_region = d_region;
    k.D = depth_num;
}

static int block_for(int points) {
    int t = (points + 31) / 32 * 32;
    return t < 32 ? 32 : (t > 128 ? 128 : t);
}

}  // namespace ptbwd
}  // namespace racf

using namespace racf;
using namespace racf::ptbwd;

extern "C" int racf_msmv_points_backward(const float* query_ray, const float* offset, const float* ray_logit,
                                         const float* time_diff, const float* lidar2img, const float* depth_base,
                                         const double* pc_range, float d_region, float image_w, float image_h, float eps,
                                         int batch, int num_query, int num_frames, int num_groups, int num_points,
                                         int depth_num, int num_views, int num_levels, const float* loc,
                                         const float* weights, const float* grad_loc, const float* grad_weights,
                                         float* grad_ray, float* grad_offset, float* grad_ray_logit, float* grad_scale_raw,
                                         racf_stream_t stream) {
    if (!query_ray || !offset || !ray_logit || !time_diff || !lidar2img || !depth_base || !pc_range || !loc || !weights ||
        !grad_loc || !grad_weights || !grad_offset || !grad_ray_logit || !grad_scale_raw)
        return RACF_ERR_NULL_POINTER;
    if (batch <= 0 || num_query <= 0 || num_frames <= 0 || num_groups <= 0 || num_points <= 0 || depth_num <= 0 ||
        num_views < 2 || (long long)batch * num_query > 0x7fffffffLL)
        return RACF_ERR_BAD_SHAPE;
    if (num_levels < 1 || num_levels > RACF_MAX_LEVELS) return RACF_ERR_BAD_LEVELS;
    ImgBwdArgs a;
    a.ray = query_ray; a.offset = offset; a.ray_logit = ray_logit; a.time_diff = time_diff; a.lidar2img = lidar2img;
    a.depth_base = depth_base; a.loc = loc; a.weights = weights; a.grad_loc = grad_loc; a.grad_weights = grad_weights;
    a.g_ray = grad_ray; a.g_offset = grad_offset; a.g_logit = grad_ray_logit; a.g_scale = grad_scale_raw;
    fill_consts(a.k, pc_range, d_region, depth_num);
    a.image_w = image_w; a.image_h = image_h; a.eps = eps;
    a.B = batch; a.Q = num_query; a.T = num_frames; a.G = num_groups; a.Pn = num_points; a.N = num_views; a.L = num_levels;
    const size_t smem = (size_t)(depth_num + 8) * sizeof(float);
    if (smem > 48 * 1024) return RACF_ERR_UNSUPPORTED;
    msmv_points_bwd_kernel<<<(unsigned)(batch * num_query), block_for(num_groups * num_points * depth_num), smem,
                             static_cast<cudaStream_t>(stream)>>>(a);
    return (int)cudaGetLastError();
}

extern "C" int racf_bev_points_backward(const float* query_ray, const float* offset, const float* ray_logit,
                                        const float* time_diff, const float* depth_base, const double* pc_range,
                                        float d_region, int batch, int num_query, int num_frames, int num_heads,
                                        int num_points, int depth_num, const float* attn, const float* grad_loc,
                                        const float* grad_attn, float* grad_ray, float* grad_offset, float* grad_ray_logit,
                                        float* grad_attn_raw, racf_stream_t stream) {
    if (!query_ray || !offset || !ray_logit || !time_diff || !depth_base || !pc_range || !attn || !grad_loc || !grad_attn ||
        !grad_offset || !grad_ray_logit || !grad_attn_raw)
        return RACF_ERR_NULL_POINTER;
    if (batch <= 0 || num_query <= 0 || num_frames <= 0 || num_heads <= 0 || num_points <= 0 || depth_num <= 0 ||
        (long long)batch * num_query > 0x7fffffffLL)
        return RACF_ERR_BAD_SHAPE;
    BevBwdArgs a;
    a.ray = query_ray; a.offset = offset; a.ray_logit = ray_logit; a.time_diff = time_diff; a.depth_base = depth_base;
    a.attn = attn; a.grad_loc = grad_loc; a.grad_attn = grad_attn;
    a.g_ray = grad_ray; a.g_offset = grad_offset; a.g_logit = grad_ray_logit; a.g_attn_raw = grad_attn_raw;
    fill_consts(a.k, pc_range, d_region, depth_num);
    a.B = batch; a.Q = num_query; a.T = num_frames; a.M = num_heads; a.Pn = num_points;
    const size_t smem = (size_t)(depth_num + 8 + num_heads * num_points * depth_num) * sizeof(float);
    if (smem > 48 * 1024) return RACF_ERR_UNSUPPORTED;
    bev_points_bwd_kernel<<<(unsigned)(batch * num_query), block_for(num_heads * num_points * depth_num), smem,
                            static_cast<cudaStream_t>(stream)>>>(a);
    return (int)cudaGetLastError();
}

extern "C" int racf_refine_bbox_backward(const float* proposal, const float* delta, const float* time_diff,
                                         const float* grad_pred_xy, int batch, int num_query, int num_frames, int code_size,
                                         float num_ray, float* grad_delta, float* grad_proposal, racf_stream_t stream) {
    if (!proposal || !delta || !grad_pred_xy || !grad_delta) return RACF_ERR_NULL_POINTER;
    if (batch <= 0 || num_query <= 0 || num_frames <= 0 || code_size < 3 || !(num_ray > 0.f)) return RACF_ERR_BAD_SHAPE;
    const long long rows = (long long)batch * num_query;
    if (rows > 0x7fffffffLL) return RACF_ERR_BAD_SHAPE;
    refine_bbox_bwd_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        proposal, delta, time_diff, grad_pred_xy, (int)rows, num_query, num_frames, code_size, num_ray, grad_delta,
        grad_proposal);
    return (int)cudaGetLastError();
}
