// Fused sampling-point generation for the two decoder branches (inference / forward only) -- SURVEY.md 8f-2.
//
// What the reference does with ~150 small PyTorch kernels per decoder iteration, in one launch per branch:
//   image branch : RaCFormerSampling.inner_forward (models/racformer_transformer.py:361-408) followed by the front
//                  half of sampling_4d (models/sparsebev_sampling.py:45-120): box decode, box-relative offsets,
//                  velocity warp over T frames, polar depth jitter, projection into the N cameras of the frame,
//                  first-valid-view selection, packing of `loc` as [B*T*G,Q,P,3] and of softmax(scale weights) as
//                  [B*G*T,Q,P,L] (the reference's mismatched packing is preserved).
//   BEV branch   : BEVSampling.inner_forward (models/racformer_transformer.py:493-529) + the queue-major packing of
//                  BEVSelfAttention.forward (models/bev_self_attention.py:176-188): loc [T*B,Q,M,1,P,2],
//                  softmax(attention weights) [T*B,Q,M,1,P].
//
// Arithmetic mirrors the PyTorch op sequence in fp32 (separate multiply/add where PyTorch issues separate kernels,
// FMA chains in k order where it calls a GEMM) so that the discrete view selection agrees with the eager path up to
// last-ulp effects at image borders; tests/test_points.py states the tolerance.
#include "racf_common.cuh"
#include "racformer_ops.h"

namespace racf {

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

constexpr float kMapSize = 102.4f, kCentre = 51.2f, kRayR = 65.0f;
constexpr float kTwoPi = 6.283185307179586f;

// theta_d2xy_coods (models/bbox/utils.py:82-90): polar (turns, units of r) -> normalised cartesian, clamped to [0,1]
__device__ __forceinline__ void polar_to_xy(float theta, float d, float& x, float& y) {
    const float ang = fmul(theta, kTwoPi);
    const float rad = fmul(d, kRayR);
    x = fdiv(fadd(kCentre, fmul(rad, cosf(ang))), kMapSize);
    y = fdiv(fadd(kCentre, fmul(rad, sinf(ang))), kMapSize);
    x = fminf(fmaxf(x, 0.f), 1.f);
    y = fminf(fmaxf(y, 0.f), 1.f);
}

// xy2theta_d_coods (models/bbox/utils.py:93-106, norm=True)
__device__ __forceinline__ void xy_to_polar(float x, float y, float& theta, float& d) {
    const float dx = fsub(fmul(x, kMapSize), kCentre);
    const float dy = fsub(fmul(y, kMapSize), kCentre);
    d = fdiv(sqrtf(fadd(fmul(dx, dx), fmul(dy, dy))), kRayR);
    float t = fadd(atan2f(dy, dx), kTwoPi);
    t = fmodf(t, kTwoPi);
    if (t < 0.f) t = fadd(t, kTwoPi);
    theta = fdiv(t, kTwoPi);
}

struct BoxFrame {   // decode_bbox(theta_d2xy_coods(query_ray), pc_range) (models/bbox/utils.py:66-80)
    float cx, cy, cz, w, l, h, s, c, vx, vy;
};

__device__ __forceinline__ BoxFrame decode_query(const float* __restrict__ ray, const float* pc, const float* span) {
    BoxFrame f;
    float qx, qy;
    polar_to_xy(ray[0], ray[1], qx, qy);
    f.cx = fadd(fmul(qx, span[0]), pc[0]);
    f.cy = fadd(fmul(qy, span[1]), pc[1]);
    f.cz = fadd(fmul(ray[2], span[2]), pc[2]);
    f.w = expf(ray[3]);
    f.l = expf(ray[4]);
    f.h = expf(ray[5]);
    const float ang = atan2f(ray[6], ray[7]);
    f.s = sinf(ang);
    f.c = cosf(ang);
    f.vx = ray[8];
    f.vy = ray[9];
    return f;
}

// make_sample_points (models/sparsebev_sampling.py:8-25): box-relative offset -> lidar frame. The rotation is a
// [P,3]x[3,3] bmm in the reference (models/utils.py:78); products are accumulated in k order.
__device__ __forceinline__ void box_point(const BoxFrame& f, float ox, float oy, float oz, float& x, float& y, float& z) {
    const float dx = fmul(f.w, ox), dy = fmul(f.l, oy), dz = fmul(f.h, oz);
    const float rx = fmaf(dz, 0.f, fmaf(dy, -f.s, fmul(dx, f.c)));
    const float ry = fmaf(dz, 0.f, fmaf(dy, f.c, fmul(dx, f.s)));
    const float rz = fmaf(dz, 1.f, fmaf(dy, 0.f, fmul(dx, 0.f)));
    x = fadd(f.cx, rx);
    y = fadd(f.cy, ry);
    z = fadd(f.cz, rz);
}

__device__ __forceinline__ float sigmoidf_torch(float x) { return fdiv(1.f, fadd(1.f, expf(-x))); }

// depth jitter along the ray: linspace(-d,d,D)[k] + (sigmoid(logit)*2-1)*d_region/D/2 (racformer_transformer.py:395-396)
__device__ __forceinline__ float depth_offset(float base, float logit, float d_region, int D) {
    const float j = fdiv(fdiv(fmul(fsub(fmul(sigmoidf_torch(logit), 2.f), 1.f), d_region), (float)D), 2.f);
    return fadd(base, j);
}

struct ImgPointArgs {
    const float* ray;       // [B,Q,10]
    const float* offset;    // [B,Q,G*Pn*D,3]
    const float* ray_logit; // [B,Q,D]
    const float* scale_raw; // [B,Q,G,T,Pn*D,L]
    const float* time_diff; // [B,T]
    const float* lidar2img; // [B,T*N,4,4]
    const float* depth_base;// [D]
    float* loc;             // [B*T*G,Q,P,3]
    float* weights;         // [B*G*T,Q,P,L]
    float pc[6];     // pc_range as fp32 scalars
    float span[3];   // (pc[3+i] - pc[i]) evaluated in double like the Python expression, then rounded to fp32
    float d_region, image_w, image_h, eps;
    int B, Q, T, G, Pn, D, N, L;
};

__global__ void __launch_bounds__(256) msmv_points_kernel(const ImgPointArgs a) {
    const int P = a.Pn * a.D;
    const long long total = (long long)a.B * a.T * a.G * a.Q * P;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        // idx enumerates the OUTPUT order of loc: ((b*T + t)*G + g, q, p)
        long long r = idx;
        const int p = (int)(r % P); r /= P;
        const int q = (int)(r % a.Q); r /= a.Q;
        const int g = (int)(r % a.G); r /= a.G;
        const int t = (int)(r % a.T);
        const int b = (int)(r / a.T);
        const int d = p % a.D;

        const long long bq = (long long)b * a.Q + q;
        const BoxFrame f = decode_query(a.ray + bq * 10, a.pc, a.span);
        const float* off = a.offset + (bq * (a.G * P) + (long long)g * P + p) * 3;
        float x, y, z;
        box_point(f, off[0], off[1], off[2], x, y, z);
        // ego-motion-free warp by the query velocity, then normalise to [0,1]
        const float td = a.time_diff[b * a.T + t];
        x = fdiv(fsub(fsub(x, fmul(f.vx, td)), a.pc[0]), a.span[0]);
        y = fdiv(fsub(fsub(y, fmul(f.vy, td)), a.pc[1]), a.span[1]);
        float theta, dist;
        xy_to_polar(x, y, theta, dist);
        dist = fadd(dist, depth_offset(a.depth_base[d], a.ray_logit[bq * a.D + d], a.d_region, a.D));
        polar_to_xy(theta, dist, x, y);
        x = fadd(fmul(x, a.span[0]), a.pc[0]);
        y = fadd(fmul(y, a.span[1]), a.pc[1]);

        // project into the N cameras of frame t; keep the first view that sees the point, else view 0
        float su = 0.f, sv = 0.f;
        int view = 0;
        bool found = false;
        for (int n = 0; n < a.N; ++n) {
            const float* m = a.lidar2img + ((long long)b * a.T * a.N + (long long)t * a.N + n) * 16;
            const float cx = fmaf(1.f, m[3], fmaf(z, m[2], fmaf(y, m[1], fmul(x, m[0]))));
            const float cy = fmaf(1.f, m[7], fmaf(z, m[6], fmaf(y, m[5], fmul(x, m[4]))));
            const float cz = fmaf(1.f, m[11], fmaf(z, m[10], fmaf(y, m[9], fmul(x, m[8]))));
            const float hz = fmaxf(cz, a.eps);
            const float u = fdiv(fdiv(cx, hz), a.image_w);
            const float v = fdiv(fdiv(cy, hz), a.image_h);
            const bool ok = (cz > a.eps) && (v > 0.f) && (v < 1.f) && (u > 0.f) && (u < 1.f);
            if (n == 0) { su = u; sv = v; }                       // nobody sees it -> view 0 (zero padding does the rest)
            if (ok && !found) { su = u; sv = v; view = n; found = true; }   // argmax over views = FIRST valid view
        }
        float* lo = a.loc + idx * 3;
        lo[0] = su;
        lo[1] = sv;
        lo[2] = fdiv((float)view, (float)(a.N - 1));

        // softmax over levels of scale_raw[b,q,g,t,p,:] -> weights[(b*G+g)*T+t, q, p, :]   (reference quirk: G-major)
        const float* sr = a.scale_raw + ((((bq * a.G + g) * a.T + t) * P) + p) * a.L;
        float mx = sr[0];
        for (int l = 1; l < a.L; ++l) mx = fmaxf(mx, sr[l]);
        float sum = 0.f;
        for (int l = 0; l < a.L; ++l) sum = fadd(sum, expf(fsub(sr[l], mx)));
        float* wo = a.weights + ((((long long)(b * a.G + g) * a.T + t) * a.Q + q) * P + p) * a.L;
        for (int l = 0; l < a.L; ++l) wo[l] = fdiv(expf(fsub(sr[l], mx)), sum);
    }
}

struct BevPointArgs {
    const float* ray;       // [B,Q,10]
    const float* offset;    // [B,Q,M*Pn*D,2]
    const float* ray_logit; // [B,Q,D]
    const float* attn_raw;  // [B,Q,M,P]      (num_levels == 1)
    const float* time_diff; // [B,T]
    const float* depth_base;// [D]
    float* loc;             // [T*B,Q,M,1,P,2]
    float* attn;            // [T*B,Q,M,1,P]
    float pc[6];
    float span[3];
    float d_region;
    int B, Q, T, M, Pn, D;
};

// One thread per (b, q, head, point): the box decode, the box-relative point and the attention softmax do not depend on
// the frame, so they are done once and only the velocity warp / polar jitter runs per frame (the first version spent
// 22 us per launch redoing them for each of the T frames; same arithmetic per output element).
__global__ void __launch_bounds__(256) bev_points_kernel(const BevPointArgs a) {
    const int P = a.Pn * a.D;
    const long long per_frame = (long long)a.B * a.Q * a.M * P;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < per_frame;
         idx += (long long)gridDim.x * blockDim.x) {
        // idx enumerates (b, q, m, p); outputs are queue-major: (t*B + b, q, m, p)   (bev_self_attention.py:185-188)
        long long r = idx;
        const int p = (int)(r % P); r /= P;
        const int m = (int)(r % a.M); r /= a.M;
        const int q = (int)(r % a.Q);
        const int b = (int)(r / a.Q);
        const int d = p % a.D;

        const long long bq = (long long)b * a.Q + q;
        const BoxFrame f = decode_query(a.ray + bq * 10, a.pc, a.span);
        const float* off = a.offset + (bq * (a.M * P) + (long long)m * P + p) * 2;
        float px, py, pz;
        box_point(f, off[0], off[1], 0.f, px, py, pz);
        const float jitter = depth_offset(a.depth_base[d], a.ray_logit[bq * a.D + d], a.d_region, a.D);

        const float* ar = a.attn_raw + (bq * a.M + m) * P;
        float mx = ar[0];
        for (int k = 1; k < P; ++k) mx = fmaxf(mx, ar[k]);
        float sum = 0.f;
        for (int k = 0; k < P; ++k) sum = fadd(sum, expf(fsub(ar[k], mx)));
        const float weight = fdiv(expf(fsub(ar[p], mx)), sum);

        const long long within = ((long long)q * a.M + m) * P + p;
        for (int t = 0; t < a.T; ++t) {
            const float td = a.time_diff[b * a.T + t];
            float x = fdiv(fsub(fsub(px, fmul(f.vx, td)), a.pc[0]), a.span[0]);
            float y = fdiv(fsub(fsub(py, fmul(f.vy, td)), a.pc[1]), a.span[1]);
            float theta, dist;
            xy_to_polar(x, y, theta, dist);
            dist = fadd(dist, jitter);
            polar_to_xy(theta, dist, x, y);
            const long long o = ((long long)t * a.B + b) * ((long long)a.Q * a.M * P) + within;
            *reinterpret_cast<float2*>(a.loc + o * 2) = make_float2(x, y);
            a.attn[o] = weight;
        }
    }
}

static unsigned grid_for(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = 16LL * sm_count();
    return (unsigned)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace racf

using namespace racf;

extern "C" int racf_msmv_points_forward(const float* query_ray, const float* offset, const float* ray_logit,
                                        const float* scale_raw, const float* time_diff, const float* lidar2img,
                                        const float* depth_base, const double* pc_range, float d_region, float image_w,
                                        float image_h, float eps, int batch, int num_query, int num_frames,
                                        int num_groups, int num_points, int depth_num, int num_views, int num_levels,
                                        float* loc, float* weights, racf_stream_t stream) {
    if (!query_ray || !offset || !ray_logit || !scale_raw || !time_diff || !lidar2img || !depth_base || !pc_range ||
        !loc || !weights)
        return RACF_ERR_NULL_POINTER;
    if (batch <= 0 || num_query <= 0 || num_frames <= 0 || num_groups <= 0 || num_points <= 0 || depth_num <= 0 ||
        num_views < 2)
        return RACF_ERR_BAD_SHAPE;
    if (num_levels < 1 || num_levels > RACF_MAX_LEVELS) return RACF_ERR_BAD_LEVELS;
    ImgPointArgs a;
    a.ray = query_ray; a.offset = offset; a.ray_logit = ray_logit; a.scale_raw = scale_raw; a.time_diff = time_diff;
    a.lidar2img = lidar2img; a.depth_base = depth_base; a.loc = loc; a.weights = weights;
    for (int i = 0; i < 6; ++i) a.pc[i] = (float)pc_range[i];
    for (int i = 0; i < 3; ++i) a.span[i] = (float)(pc_range[3 + i] - pc_range[i]);
    a.d_region = d_region; a.image_w = image_w; a.image_h = image_h; a.eps = eps;
    a.B = batch; a.Q = num_query; a.T = num_frames; a.G = num_groups; a.Pn = num_points; a.D = depth_num;
    a.N = num_views; a.L = num_levels;
    const long long total = (long long)batch * num_frames * num_groups * num_query * num_points * depth_num;
    msmv_points_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return (int)cudaGetLastError();
}

extern "C" int racf_bev_points_forward(const float* query_ray, const float* offset, const float* ray_logit,
                                       const float* attn_raw, const float* time_diff, const float* depth_base,
                                       const double* pc_range, float d_region, int batch, int num_query, int num_frames,
                                       int num_heads, int num_points, int depth_num, float* loc, float* attn,
                                       racf_stream_t stream) {
    if (!query_ray || !offset || !ray_logit || !attn_raw || !time_diff || !depth_base || !pc_range || !loc || !attn)
        return RACF_ERR_NULL_POINTER;
    if (batch <= 0 || num_query <= 0 || num_frames <= 0 || num_heads <= 0 || num_points <= 0 || depth_num <= 0)
        return RACF_ERR_BAD_SHAPE;
    BevPointArgs a;
    a.ray = query_ray; a.offset = offset; a.ray_logit = ray_logit; a.attn_raw = attn_raw; a.time_diff = time_diff;
    a.depth_base = depth_base; a.loc = loc; a.attn = attn;
    for (int i = 0; i < 6; ++i) a.pc[i] = (float)pc_range[i];
    for (int i = 0; i < 3; ++i) a.span[i] = (float)(pc_range[3 + i] - pc_range[i]);
    a.d_region = d_region;
    a.B = batch; a.Q = num_query; a.T = num_frames; a.M = num_heads; a.Pn = num_points; a.D = depth_num;
    const long long total = (long long)batch * num_query * num_heads * num_points * depth_num;   // one thread per (b,q,m,p)
    bev_points_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// Box refinement at the end of a decoder iteration (models/racformer_transformer.py:264-279 refine_bbox + the
// relative -> absolute velocity scaling :255-259 + theta_d2xy_coods of the prediction, models/bbox/utils.py:82-90):
// ~25 PyTorch elementwise launches per iteration as one. One thread per query; separate multiplies / adds where PyTorch
// runs separate kernels.
// ---------------------------------------------------------------------------------------------------------------------
namespace racf {

__device__ __forceinline__ float torch_sigmoid(float x) { return fdiv(1.f, fadd(1.f, expf(-x))); }

__device__ __forceinline__ float torch_inverse_sigmoid(float x, float eps) {
    x = fminf(fmaxf(x, 0.f), 1.f);
    return logf(fdiv(fmaxf(x, eps), fmaxf(fsub(1.f, x), eps)));
}

__global__ void __launch_bounds__(128)
refine_bbox_kernel(const float* __restrict__ proposal, const float* __restrict__ delta, const float* __restrict__ time_diff,
                   int rows, int num_query, int num_frames, int code, float num_ray, float* __restrict__ pred,
                   float* __restrict__ pred_xy) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const float* p = proposal + (size_t)row * code;
    const float* d = delta + (size_t)row * code;
    float* o = pred + (size_t)row * code;
    float* oxy = pred_xy + (size_t)row * code;
    const float theta = fadd(p[0], fdiv(fsub(fmul(torch_sigmoid(d[0]), 2.f), 1.f), num_ray));
    const float dist = torch_sigmoid(fadd(d[1], torch_inverse_sigmoid(p[1], 1e-5f)));
    const float z = torch_sigmoid(fadd(d[2], torch_inverse_sigmoid(p[2], 1e-5f)));
    o[0] = theta; o[1] = dist; o[2] = z;
    float x, y;
    polar_to_xy(theta, dist, x, y);
    oxy[0] = x; oxy[1] = y; oxy[2] = z;
    float td = 1.f;
    if (time_diff != nullptr && num_frames > 1) {      // relative -> absolute velocity
        td = time_diff[(size_t)(row / num_query) * num_frames + 1];
        if (td < 1e-5f) td = 1.f;
    }
    for (int c = 3; c < code; ++c) {
        float v = d[c];
        if (c >= 8 && time_diff != nullptr && num_frames > 1) v = fdiv(v, td);
        o[c] = v;
        oxy[c] = v;
    }
}

}  // namespace racf

extern "C" int racf_refine_bbox_forward(const float* proposal, const float* delta, const float* time_diff, int batch,
                                        int num_query, int num_frames, int code_size, float num_ray, float* pred,
                                        float* pred_xy, racf_stream_t stream) {
    using namespace racf;
    if (!proposal || !delta || !pred || !pred_xy) return RACF_ERR_NULL_POINTER;
    if (batch <= 0 || num_query <= 0 || num_frames <= 0 || code_size < 3 || !(num_ray > 0.f)) return RACF_ERR_BAD_SHAPE;
    const long long rows = (long long)batch * num_query;
    if (rows > 0x7fffffffLL) return RACF_ERR_BAD_SHAPE;
    refine_bbox_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        proposal, delta, time_diff, (int)rows, num_query, num_frames, code_size, num_ray, pred, pred_xy);
    return (int)cudaGetLastError();
}
