// Multi-scale deformable attention (MSDA) for B200 (sm_100a): forward, backward, mask dump.
//
// Operator semantics: mmcv-full 1.6.0 ms_deform_attn (third-party, not in the reference tree), called
// from models/multi_scale_deformable_attn_function.py:118-124,150-160; restated in SURVEY.md App. A.3/A.4.
//
// Work decomposition (fast path, head_dim == 64):
//   * one warp owns one (b, q, head): all L*P taps and the 64 channels of that head, so the output row,
//     grad_attn[b,q,m,:,:] and grad_loc[b,q,m,:,:,:] are produced without atomics.
//   * value is [B,S,M,D]: a (pixel, head) is 256 contiguous bytes = 16 lanes x 128 bit. The x-neighbour
//     is M*D floats further, so lanes 0-15 fetch (y, x0) and lanes 16-31 fetch (y, x0+1) in the same
//     warp-wide 128-bit load; two loads per tap, halves combined by one shfl_xor(16) per output.
//   * tap geometry is computed by one lane per tap (32 taps per round) and staged in shared memory.
//   * backward scatters grad_value with red.global.add.v4.f32.
#include "racf_common.cuh"
#include "racformer_ops.h"

namespace racf {

constexpr int kMsdaWarps = 8;

struct MsdaArgs {
    const float* value;       // [B,S,M,D]
    const int64_t* shapes;    // [L,2] device
    const int64_t* lsi;       // [L] device
    const float* loc;         // [B,Q,M,L,P,2]
    const float* attn;        // [B,Q,M,L,P]
    const float* grad_out;    // [B,Q,M*D]
    float* out;               // [B,Q,M*D]
    float* grad_value;
    float* grad_loc;
    float* grad_attn;
    int B, S, M, D, L, Q, P;
    // second problem of identical geometry served by the same launch (blockIdx.y == 1): racf_msda_forward_pair
    const float* value2;
    const float* loc2;
    const float* attn2;
    float* out2;
};

// ------------------------------------------------------------------------------------------------
// Fast path forward (D == 64)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMsdaWarps * 32) msda_fwd_d64_kernel(const MsdaArgs a) {
    __shared__ float2 s_w[kMsdaWarps][32][2];  // [tap][x-slot] {w_top, w_bot} * attention weight
    __shared__ int4 s_om[kMsdaWarps][32];      // {float4 offset of top-left (pixel, head), mask, row stride, -}

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long bqm = (long long)blockIdx.x * kMsdaWarps + warp;
    if (bqm >= (long long)a.B * a.Q * a.M) return;
    const bool second = blockIdx.y != 0;             // the other BEV branch of the pair (same shapes, other tensors)
    const float* value = second ? a.value2 : a.value;
    const int m = (int)(bqm % a.M);
    const int b = (int)(bqm / ((long long)a.Q * a.M));
    const int slot = lane >> 4, j = lane & 15;
    const int px = a.M * 16;  // float4 stride between x-neighbours
    const float4* base =
        reinterpret_cast<const float4*>(value + ((size_t)b * a.S * a.M + m) * 64) + slot * px + j;
    const int T = a.L * a.P;
    const float* loc_q = (second ? a.loc2 : a.loc) + bqm * T * 2;
    const float* aw_q = (second ? a.attn2 : a.attn) + bqm * T;

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t0 = 0; t0 < T; t0 += 32) {
        __syncwarp();
        {
            const int t = t0 + lane;
            float2 w0 = make_float2(0.f, 0.f), w1 = w0;
            int4 om = make_int4(0, 0, 0, 0);
            if (t < T) {
                const int l = t / a.P;
                const int H = (int)a.shapes[2 * l], W = (int)a.shapes[2 * l + 1];
                const int start = (int)a.lsi[l];
                const float2 xy = __ldg(reinterpret_cast<const float2*>(loc_q) + t);
                const float aw = __ldg(aw_q + t);
                const TapGeom g = tap_geometry(msda_pixel(xy.y, H), msda_pixel(xy.x, W), H, W);
                if (g.mask != 0u) {
                    const float hh = 1.f - g.lh, hw = 1.f - g.lw;
                    w0 = make_float2(hh * hw * aw, g.lh * hw * aw);
                    w1 = make_float2(hh * g.lw * aw, g.lh * g.lw * aw);
                    om.x = (start + g.h_low * W + g.w_low) * px;
                    om.y = (int)g.mask;
                    om.z = W * px;
                }
            }
            s_w[warp][lane][0] = w0;
            s_w[warp][lane][1] = w1;
            s_om[warp][lane] = om;
        }
        __syncwarp();
        const int n = min(32, T - t0);
        for (int tt = 0; tt < n; tt += 4) {  // records past n have mask 0
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int4 om = s_om[warp][tt + u];
                const float2 w = s_w[warp][tt + u][slot];
                const unsigned mk = (unsigned)om.y >> slot;
                const float4* p = base + om.x;
                float4 top = make_float4(0.f, 0.f, 0.f, 0.f), bot = top;
                if (mk & 1u) top = ldg128(p);
                if (mk & 4u) bot = ldg128(p + om.z);
                acc.x = fmaf(w.y, bot.x, fmaf(w.x, top.x, acc.x));
                acc.y = fmaf(w.y, bot.y, fmaf(w.x, top.y, acc.y));
                acc.z = fmaf(w.y, bot.z, fmaf(w.x, top.z, acc.z));
                acc.w = fmaf(w.y, bot.w, fmaf(w.x, top.w, acc.w));
            }
        }
    }
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
    if (slot == 0) reinterpret_cast<float4*>((second ? a.out2 : a.out) + bqm * 64)[j] = acc;
}

// ------------------------------------------------------------------------------------------------
// Fast path backward (D == 64)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMsdaWarps * 32) msda_bwd_d64_kernel(const MsdaArgs a) {
    __shared__ float4 s_rec[kMsdaWarps][32];  // {lh, lw, attention weight, bits(offset)}
    __shared__ int4 s_aux[kMsdaWarps][32];    // {mask, row stride (float4), W, H}

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long bqm = (long long)blockIdx.x * kMsdaWarps + warp;
    if (bqm >= (long long)a.B * a.Q * a.M) return;
    const int m = (int)(bqm % a.M);
    const int b = (int)(bqm / ((long long)a.Q * a.M));
    const int slot = lane >> 4, j = lane & 15;
    const int px = a.M * 16;
    const size_t boff = ((size_t)b * a.S * a.M + m) * 64;
    const float4* base = reinterpret_cast<const float4*>(a.value + boff) + slot * px + j;
    float* gbase = a.grad_value + boff + (size_t)(slot * px + j) * 4;
    const int T = a.L * a.P;
    const float* loc_q = a.loc + bqm * T * 2;
    const float* aw_q = a.attn + bqm * T;
    float* gl_q = a.grad_loc + bqm * T * 2;
    float* ga_q = a.grad_attn + bqm * T;
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.grad_out + bqm * 64) + j);
    const float sgn = slot ? 1.f : -1.f;

    for (int t0 = 0; t0 < T; t0 += 32) {
        __syncwarp();
        {
            const int t = t0 + lane;
            float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
            int4 aux = make_int4(0, 0, 0, 0);
            if (t < T) {
                const int l = t / a.P;
                const int H = (int)a.shapes[2 * l], W = (int)a.shapes[2 * l + 1];
                const int start = (int)a.lsi[l];
                const float2 xy = __ldg(reinterpret_cast<const float2*>(loc_q) + t);
                const float aw = __ldg(aw_q + t);
                const TapGeom tg = tap_geometry(msda_pixel(xy.y, H), msda_pixel(xy.x, W), H, W);
                if (tg.mask != 0u) {
                    rec = make_float4(tg.lh, tg.lw, aw, __int_as_float((start + tg.h_low * W + tg.w_low) * px));
                    aux = make_int4((int)tg.mask, W * px, W, H);
                }
            }
            s_rec[warp][lane] = rec;
            s_aux[warp][lane] = aux;
        }
        __syncwarp();
        const int n = min(32, T - t0);
        for (int tt = 0; tt < n; ++tt) {
            const int4 aux = s_aux[warp][tt];
            float g_aw = 0.f, g_x = 0.f, g_y = 0.f;
            if (aux.x != 0) {  // warp-uniform
                const float4 rec = s_rec[warp][tt];
                const float lh = rec.x, lw = rec.y, aw = rec.z;
                const int off = __float_as_int(rec.w);
                const float hh = 1.f - lh, hw = 1.f - lw;
                const float wx = slot ? lw : hw;
                const float w_top = hh * wx, w_bot = lh * wx;
                const unsigned mk = (unsigned)aux.x >> slot;
                const float4* p = base + off;
                float* gp = gbase + (size_t)off * 4;
                float4 top = make_float4(0.f, 0.f, 0.f, 0.f), bot = top;
                if (mk & 1u) top = ldg128(p);
                if (mk & 4u) bot = ldg128(p + aux.y);
                if (mk & 1u) {
                    const float k = w_top * aw;
                    red_add_v4(gp, k * g.x, k * g.y, k * g.z, k * g.w);
                }
                if (mk & 4u) {
                    const float k = w_bot * aw;
                    red_add_v4(gp + (size_t)aux.y * 4, k * g.x, k * g.y, k * g.z, k * g.w);
                }
                const float A = dot4(g, top), Bv = dot4(g, bot);
                g_aw = warp_sum(fmaf(w_top, A, w_bot * Bv));
                g_x = (float)aux.z * aw * warp_sum(sgn * fmaf(hh, A, lh * Bv));
                g_y = (float)aux.w * aw * warp_sum(wx * (Bv - A));
            }
            if (lane == 0) {
                ga_q[t0 + tt] = g_aw;
                reinterpret_cast<float2*>(gl_q)[t0 + tt] = make_float2(g_x, g_y);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Generic path (any head_dim): one thread per output channel / per (tap, channel).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) msda_fwd_generic_kernel(const MsdaArgs a) {
    const long long total = (long long)a.B * a.Q * a.M * a.D;
    const int T = a.L * a.P;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(idx % a.D);
        const long long bqm = idx / a.D;
        const int m = (int)(bqm % a.M);
        const int b = (int)(bqm / ((long long)a.Q * a.M));
        const float* vb = a.value + ((size_t)b * a.S * a.M + m) * a.D + d;
        const size_t px = (size_t)a.M * a.D;
        float col = 0.f;
        for (int t = 0; t < T; ++t) {
            const int l = t / a.P;
            const int H = (int)a.shapes[2 * l], W = (int)a.shapes[2 * l + 1];
            const float x = a.loc[(bqm * T + t) * 2], y = a.loc[(bqm * T + t) * 2 + 1];
            const TapGeom g = tap_geometry(msda_pixel(y, H), msda_pixel(x, W), H, W);
            if (g.mask == 0u) continue;
            const float* f = vb + ((size_t)a.lsi[l] + (size_t)((long long)g.h_low * W + g.w_low)) * px;
            const size_t rs = (size_t)W * px;
            const float hh = 1.f - g.lh, hw = 1.f - g.lw;
            const float v1 = (g.mask & kTL) ? __ldg(f) : 0.f;
            const float v2 = (g.mask & kTR) ? __ldg(f + px) : 0.f;
            const float v3 = (g.mask & kBL) ? __ldg(f + rs) : 0.f;
            const float v4 = (g.mask & kBR) ? __ldg(f + rs + px) : 0.f;
            const float val = hh * hw * v1 + hh * g.lw * v2 + g.lh * hw * v3 + g.lh * g.lw * v4;
            col = fmaf(val, a.attn[bqm * T + t], col);
        }
        a.out[idx] = col;
    }
}

// grad_loc / grad_attn must be zero on entry (the launcher memsets them).
__global__ void __launch_bounds__(256) msda_bwd_generic_kernel(const MsdaArgs a) {
    const int T = a.L * a.P;
    const long long total = (long long)a.B * a.Q * a.M * T * a.D;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(idx % a.D);
        const long long bqmt = idx / a.D;
        const int t = (int)(bqmt % T);
        const long long bqm = bqmt / T;
        const int m = (int)(bqm % a.M);
        const int b = (int)(bqm / ((long long)a.Q * a.M));
        const int l = t / a.P;
        const int H = (int)a.shapes[2 * l], W = (int)a.shapes[2 * l + 1];
        const float x = a.loc[bqmt * 2], y = a.loc[bqmt * 2 + 1];
        const TapGeom g = tap_geometry(msda_pixel(y, H), msda_pixel(x, W), H, W);
        if (g.mask == 0u) continue;
        const float go = a.grad_out[bqm * a.D + d];
        const float aw = a.attn[bqmt];
        const size_t px = (size_t)a.M * a.D;
        const size_t o = ((size_t)b * a.S * a.M + m) * a.D + d +
                         ((size_t)a.lsi[l] + (size_t)((long long)g.h_low * W + g.w_low)) * px;
        const float* f = a.value + o;
        float* gf = a.grad_value + o;
        const size_t rs = (size_t)W * px;
        const float hh = 1.f - g.lh, hw = 1.f - g.lw;
        const float tv = go * aw;
        float v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
        if (g.mask & kTL) { v1 = __ldg(f); atomicAdd(gf, hh * hw * tv); }
        if (g.mask & kTR) { v2 = __ldg(f + px); atomicAdd(gf + px, hh * g.lw * tv); }
        if (g.mask & kBL) { v3 = __ldg(f + rs); atomicAdd(gf + rs, g.lh * hw * tv); }
        if (g.mask & kBR) { v4 = __ldg(f + rs + px); atomicAdd(gf + rs + px, g.lh * g.lw * tv); }
        const float val = hh * hw * v1 + hh * g.lw * v2 + g.lh * hw * v3 + g.lh * g.lw * v4;
        const float gh = -hw * v1 - g.lw * v2 + hw * v3 + g.lw * v4;
        const float gwd = -hh * v1 + hh * v2 - g.lh * v3 + g.lh * v4;
        atomicAdd(a.grad_attn + bqmt, go * val);
        atomicAdd(a.grad_loc + bqmt * 2 + 0, (float)W * gwd * tv);
        atomicAdd(a.grad_loc + bqmt * 2 + 1, (float)H * gh * tv);
    }
}

struct MsdaMaskArgs {
    const int64_t* shapes;
    const float* loc;
    uint8_t* mask;
    long long ntaps;
    int L, P;
};

__global__ void __launch_bounds__(256) msda_mask_kernel(const MsdaMaskArgs a) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.ntaps;
         i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)((i / a.P) % a.L);
        const int H = (int)a.shapes[2 * l], W = (int)a.shapes[2 * l + 1];
        const TapGeom g = tap_geometry(msda_pixel(a.loc[i * 2 + 1], H), msda_pixel(a.loc[i * 2], W), H, W);
        a.mask[i] = (uint8_t)((g.in_range ? 1u : 0u) | (g.mask << 1));
    }
}

static int check_msda(const float* value, const int64_t* shapes, const int64_t* lsi, const float* loc,
                      const float* attn, int B, int S, int M, int D, int L, int Q, int P, int im2col_step) {
    if (!value || !shapes || !lsi || !loc || !attn) return RACF_ERR_NULL_POINTER;
    if (L < 1) return RACF_ERR_BAD_LEVELS;
    if (B <= 0 || S <= 0 || M <= 0 || D <= 0 || Q <= 0 || P <= 0) return RACF_ERR_BAD_SHAPE;
    if ((long long)S * M * D >= (1LL << 31)) return RACF_ERR_BAD_SHAPE;
    if (im2col_step <= 0) return RACF_ERR_IM2COL_STEP;
    const int step = B < im2col_step ? B : im2col_step;
    if (B % step != 0) return RACF_ERR_IM2COL_STEP;
    return RACF_OK;
}

static unsigned capped_grid(long long total, int threads, long long cap) {
    long long g = (total + threads - 1) / threads;
    return (unsigned)(g > cap ? cap : (g < 1 ? 1 : g));
}

static bool aligned16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace racf

using namespace racf;

extern "C" int racf_msda_forward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                                 const float* loc, const float* attn, int batch, int spatial_size, int num_heads,
                                 int head_dim, int num_levels, int num_query, int num_point, int im2col_step,
                                 float* out, racf_stream_t stream) {
    int rc = check_msda(value, spatial_shapes, level_start_index, loc, attn, batch, spatial_size, num_heads, head_dim,
                        num_levels, num_query, num_point, im2col_step);
    if (rc != RACF_OK) return rc;
    if (!out) return RACF_ERR_NULL_POINTER;
    MsdaArgs a{};
    a.value = value; a.shapes = spatial_shapes; a.lsi = level_start_index; a.loc = loc; a.attn = attn; a.out = out;
    a.B = batch; a.S = spatial_size; a.M = num_heads; a.D = head_dim; a.L = num_levels; a.Q = num_query; a.P = num_point;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (head_dim == 64 && aligned16p(value) && aligned16p(out) && (reinterpret_cast<uintptr_t>(loc) & 7u) == 0) {
        const long long nw = (long long)batch * num_query * num_heads;
        msda_fwd_d64_kernel<<<(unsigned)((nw + kMsdaWarps - 1) / kMsdaWarps), kMsdaWarps * 32, 0, st>>>(a);
    } else {
        const long long total = (long long)batch * num_query * num_heads * head_dim;
        msda_fwd_generic_kernel<<<capped_grid(total, 256, 64LL * sm_count()), 256, 0, st>>>(a);
    }
    return (int)cudaGetLastError();
}

extern "C" int racf_msda_forward_pair(const float* value_a, const float* loc_a, const float* attn_a, float* out_a,
                                      const float* value_b, const float* loc_b, const float* attn_b, float* out_b,
                                      const int64_t* spatial_shapes, const int64_t* level_start_index, int batch,
                                      int spatial_size, int num_heads, int head_dim, int num_levels, int num_query,
                                      int num_point, int im2col_step, racf_stream_t stream) {
    int rc = check_msda(value_a, spatial_shapes, level_start_index, loc_a, attn_a, batch, spatial_size, num_heads, head_dim,
                        num_levels, num_query, num_point, im2col_step);
    if (rc == RACF_OK)
        rc = check_msda(value_b, spatial_shapes, level_start_index, loc_b, attn_b, batch, spatial_size, num_heads, head_dim,
                        num_levels, num_query, num_point, im2col_step);
    if (rc != RACF_OK) return rc;
    if (!out_a || !out_b) return RACF_ERR_NULL_POINTER;
    const bool fast = head_dim == 64 && aligned16p(value_a) && aligned16p(value_b) && aligned16p(out_a) && aligned16p(out_b) &&
                      ((reinterpret_cast<uintptr_t>(loc_a) | reinterpret_cast<uintptr_t>(loc_b)) & 7u) == 0;
    if (!fast) {       // no paired generic kernel: two launches
        rc = racf_msda_forward(value_a, spatial_shapes, level_start_index, loc_a, attn_a, batch, spatial_size, num_heads, head_dim,
                               num_levels, num_query, num_point, im2col_step, out_a, stream);
        if (rc != RACF_OK) return rc;
        return racf_msda_forward(value_b, spatial_shapes, level_start_index, loc_b, attn_b, batch, spatial_size, num_heads,
                                 head_dim, num_levels, num_query, num_point, im2col_step, out_b, stream);
    }
    MsdaArgs a{};
    a.value = value_a; a.loc = loc_a; a.attn = attn_a; a.out = out_a;
    a.value2 = value_b; a.loc2 = loc_b; a.attn2 = attn_b; a.out2 = out_b;
    a.shapes = spatial_shapes; a.lsi = level_start_index;
    a.B = batch; a.S = spatial_size; a.M = num_heads; a.D = head_dim; a.L = num_levels; a.Q = num_query; a.P = num_point;
    const long long nw = (long long)batch * num_query * num_heads;
    const dim3 grid((unsigned)((nw + kMsdaWarps - 1) / kMsdaWarps), 2, 1);
    msda_fwd_d64_kernel<<<grid, kMsdaWarps * 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return (int)cudaGetLastError();
}

extern "C" int racf_msda_backward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                                  const float* loc, const float* attn, const float* grad_out, int batch,
                                  int spatial_size, int num_heads, int head_dim, int num_levels, int num_query,
                                  int num_point, int im2col_step, float* grad_value, float* grad_loc,
                                  float* grad_attn, racf_stream_t stream) {
    int rc = check_msda(value, spatial_shapes, level_start_index, loc, attn, batch, spatial_size, num_heads, head_dim,
                        num_levels, num_query, num_point, im2col_step);
    if (rc != RACF_OK) return rc;
    if (!grad_out || !grad_value || !grad_loc || !grad_attn) return RACF_ERR_NULL_POINTER;
    MsdaArgs a{};
    a.value = value; a.shapes = spatial_shapes; a.lsi = level_start_index; a.loc = loc; a.attn = attn;
    a.grad_out = grad_out; a.grad_value = grad_value; a.grad_loc = grad_loc; a.grad_attn = grad_attn;
    a.B = batch; a.S = spatial_size; a.M = num_heads; a.D = head_dim; a.L = num_levels; a.Q = num_query; a.P = num_point;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (head_dim == 64 && aligned16p(value) && aligned16p(grad_value) && aligned16p(grad_out) &&
        (reinterpret_cast<uintptr_t>(loc) & 7u) == 0 && (reinterpret_cast<uintptr_t>(grad_loc) & 7u) == 0) {
        const long long nw = (long long)batch * num_query * num_heads;
        msda_bwd_d64_kernel<<<(unsigned)((nw + kMsdaWarps - 1) / kMsdaWarps), kMsdaWarps * 32, 0, st>>>(a);
    } else {
        const size_t ntap = (size_t)batch * num_query * num_heads * num_levels * num_point;
        cudaError_t e = cudaMemsetAsync(grad_loc, 0, ntap * 2 * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
        e = cudaMemsetAsync(grad_attn, 0, ntap * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
        const long long total = (long long)ntap * head_dim;
        msda_bwd_generic_kernel<<<capped_grid(total, 256, 64LL * sm_count()), 256, 0, st>>>(a);
    }
    return (int)cudaGetLastError();
}

extern "C" int racf_msda_tap_masks(const int64_t* spatial_shapes, const float* loc, int batch, int num_heads,
                                   int num_levels, int num_query, int num_point, uint8_t* tap_mask,
                                   racf_stream_t stream) {
    if (!spatial_shapes || !loc || !tap_mask) return RACF_ERR_NULL_POINTER;
    if (num_levels < 1) return RACF_ERR_BAD_LEVELS;
    if (batch <= 0 || num_heads <= 0 || num_query <= 0 || num_point <= 0) return RACF_ERR_BAD_SHAPE;
    MsdaMaskArgs a;
    a.shapes = spatial_shapes; a.loc = loc; a.mask = tap_mask;
    a.ntaps = (long long)batch * num_query * num_heads * num_levels * num_point;
    a.L = num_levels; a.P = num_point;
    msda_mask_kernel<<<capped_grid(a.ntaps, 256, 32LL * sm_count()), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return (int)cudaGetLastError();
}

extern "C" int racf_version(void) { return 100; }

extern "C" const char* racf_status_string(int status) {
    switch (status) {
        case RACF_OK: return "ok";
        case RACF_ERR_NULL_POINTER: return "null pointer argument";
        case RACF_ERR_BAD_LEVELS: return "num_levels out of range";
        case RACF_ERR_BAD_SHAPE: return "invalid or too large dimension";
        case RACF_ERR_TOO_MANY_PTS: return "num_point exceed limits";
        case RACF_ERR_IM2COL_STEP: return "batch must be divisible by min(batch, im2col_step)";
        case RACF_ERR_UNSUPPORTED: return "variant not available for these shapes";
        default: break;
    }
    if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
    return "unknown status";
}
