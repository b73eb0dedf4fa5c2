/*
 * racformer_ops.h -- C ABI of libracformer_ops.so (B200 / sm_100a).
 *
 * The drop-in boundary for RaCFormer's sampling hot path. Every entry point
 * takes plain device pointers, integer sizes and a CUDA stream; the library
 * never allocates or frees device memory, never synchronises, keeps no
 * mutable global state, reads no environment variable and returns an int
 * status instead of printing.
 *
 * Reference interfaces replaced (paths relative to the RaCFormer tree):
 *   racf_msmv_forward   <- ms_deformable_im2col_cuda_c{45,2345,23456}
 *                          models/csrc/msmv_sampling/msmv_sampling_forward.cu:336-428
 *                          (declared models/csrc/msmv_sampling/msmv_sampling.cpp:5-60)
 *   racf_msmv_backward  <- ms_deformable_col2im_cuda_c{45,2345,23456}
 *                          models/csrc/msmv_sampling/msmv_sampling_backward.cu:442-563
 *   racf_msda_forward   <- mmcv-full 1.6.0 `_ext.ms_deform_attn_forward`, bound at
 *                          models/multi_scale_deformable_attn_function.py:10-12,118-124
 *   racf_msda_backward  <- mmcv-full 1.6.0 `_ext.ms_deform_attn_backward`, bound at
 *                          models/multi_scale_deformable_attn_function.py:150-160
 *
 * All tensors are fp32, dense and contiguous in the layouts given below.
 */
#ifndef RACFORMER_OPS_H_
#define RACFORMER_OPS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes. 0 = success; >0 = a cudaError_t from the launch; <0 = argument error. */
#define RACF_OK                 0
#define RACF_ERR_NULL_POINTER  (-1)
#define RACF_ERR_BAD_LEVELS    (-2)  /* num_levels outside [1, RACF_MAX_LEVELS] */
#define RACF_ERR_BAD_SHAPE     (-3)  /* a dimension is <= 0 or a per-batch extent overflows int32 */
#define RACF_ERR_TOO_MANY_PTS  (-4)  /* num_point > RACF_MSMV_MAX_POINT (reference: msmv_sampling.cpp:159) */
#define RACF_ERR_IM2COL_STEP   (-5)  /* batch % min(batch, im2col_step) != 0 (mmcv contract) */
#define RACF_ERR_UNSUPPORTED   (-6)  /* the requested variant only exists for the fast path (C == 64, L in {2,4,5}) */

#define RACF_MAX_LEVELS        8
#define RACF_MSMV_MAX_POINT    128

/* Opaque stream handle: a cudaStream_t / CUstream (NULL = legacy default stream). */
typedef void* racf_stream_t;

/* ABI version of this header (major*100 + minor). */
int racf_version(void);

/* Human-readable text for a status code returned by this library (never NULL). */
const char* racf_status_string(int status);

/*
 * Multi-scale multi-view sampling, forward.
 *   feats[l]   : [B, N, H_l, W_l, C]   channel-last feature maps, l = 0..L-1 (host array of device pointers)
 *   hw[2l..]   : (H_l, W_l)             host array
 *   loc        : [B, Q, P, 3]           (x, y in normalised image units, z = view / (N-1))
 *   weights    : [B, Q, P, L]           per-level scale weights
 *   out        : [B, Q, C, P]           fully overwritten (no pre-zeroing needed)
 * out[b,q,c,p] = sum_l weights[b,q,p,l] * bilinear(feats[l][b, round(z*(N-1))], x*(W_l-1), y*(H_l-1))[c]
 * with zero padding and align_corners=True; arithmetic follows msmv_sampling_forward.cu:27-73,95-163.
 */
int racf_msmv_forward(const float* const* feats, const int* hw, int num_levels,
                      const float* loc, const float* weights,
                      int batch, int channels, int num_views, int num_query, int num_point,
                      float* out, racf_stream_t stream);

/*
 * racf_msmv_forward with an explicit kernel variant (tuning / profiling; results are identical):
 *   variant  0 : one query per warp, one launch-time CTA per 8 queries
 *   variant -1 : persistent warps with a 3-stage query pipeline (what racf_msmv_forward uses)
 *   variant k>0: the persistent kernel, prefetching the first k levels of the next query into L2
 */
int racf_msmv_forward_variant(const float* const* feats, const int* hw, int num_levels,
                              const float* loc, const float* weights,
                              int batch, int channels, int num_views, int num_query, int num_point,
                              int variant, float* out, racf_stream_t stream);

/*
 * Forward with the un-packing of sampling_4d's tail fused in (models/sparsebev_sampling.py:128-131): batch is
 * B*T*G (T = num_frames, G = num_groups, G fastest) and out is [B, Q, G, T*P, C] -- what AdaptiveMixing consumes --
 * instead of [B*T*G, Q, C, P]. Same arithmetic as racf_msmv_forward; forward only; fast path only
 * (returns RACF_ERR_UNSUPPORTED otherwise, callers then use racf_msmv_forward + a permute).
 */
int racf_msmv_forward_grouped(const float* const* feats, const int* hw, int num_levels,
                              const float* loc, const float* weights,
                              int batch, int channels, int num_views, int num_query, int num_point,
                              int num_frames, int num_groups, float* out, racf_stream_t stream);

/*
 * Multi-scale multi-view sampling, backward (msmv_sampling_backward.cu:29-105,132-223).
 *   grad_out      : [B, Q, C, P]
 *   grad_feats[l] : [B, N, H_l, W_l, C]  accumulated with atomics. If zero_grad_feats != 0 the library
 *                   zero-fills them on `stream` first (what at::zeros_like did, msmv_sampling.cpp:322-325);
 *                   otherwise the gradients are added to the existing contents.
 *   grad_loc      : [B, Q, P, 3]  fully overwritten; the z (view) component is always 0
 *   grad_weights  : [B, Q, P, L]  fully overwritten
 */
int racf_msmv_backward(const float* grad_out,
                       const float* const* feats, const int* hw, int num_levels,
                       const float* loc, const float* weights,
                       int batch, int channels, int num_views, int num_query, int num_point,
                       float* const* grad_feats, float* grad_loc, float* grad_weights,
                       int zero_grad_feats, racf_stream_t stream);

/*
 * Backward of racf_msmv_forward_grouped: grad_out is [B/(T*G), Q, G, T*P, C] (the layout AdaptiveMixing's gradient has),
 * everything else as racf_msmv_backward. With zero_grad_feats == 0 the feature gradients of several calls accumulate in
 * one set of buffers -- the decoder's six iterations sample the same pyramid, so their gradients are summed by the
 * scatter itself instead of by six 1.5 GB additions. Fast path only (RACF_ERR_UNSUPPORTED otherwise).
 */
int racf_msmv_backward_grouped(const float* grad_out,
                               const float* const* feats, const int* hw, int num_levels,
                               const float* loc, const float* weights,
                               int batch, int channels, int num_views, int num_query, int num_point,
                               int num_frames, int num_groups,
                               float* const* grad_feats, float* grad_loc, float* grad_weights,
                               int zero_grad_feats, racf_stream_t stream);

/*
 * Debug/verification entry: the integer decisions of the MSMV kernels.
 *   view_index : [B, Q, P]     int32  round(z*(N-1))
 *   tap_mask   : [B, Q, P, L]  uint8  bit0 = tap in range (msmv_sampling_forward.cu:126),
 *                                     bit1..4 = corner (h0,w0),(h0,w1),(h1,w0),(h1,w1) is read (:48-67)
 */
int racf_msmv_tap_masks(const int* hw, int num_levels, const float* loc,
                        int batch, int num_views, int num_query, int num_point,
                        int32_t* view_index, uint8_t* tap_mask, racf_stream_t stream);

/*
 * Multi-scale deformable attention, forward (mmcv 1.6.0 ms_deformable_im2col_gpu_kernel).
 *   value            : [B, S, M, D]
 *   spatial_shapes   : [L, 2] int64 DEVICE pointer, (H_l, W_l)
 *   level_start_index: [L]    int64 DEVICE pointer
 *   loc              : [B, Q, M, L, P, 2]  (x, y) in [0,1] units of the level
 *   attn             : [B, Q, M, L, P]
 *   out              : [B, Q, M*D]          fully overwritten
 * Sampling uses align_corners=False: pixel = fma(loc, size, -0.5), zero padding.
 * im2col_step only has mmcv's divisibility contract; it does not change results.
 */
int racf_msda_forward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                      const float* loc, const float* attn,
                      int batch, int spatial_size, int num_heads, int head_dim,
                      int num_levels, int num_query, int num_point, int im2col_step,
                      float* out, racf_stream_t stream);

/*
 * Two MSDA forward problems of identical geometry (the radar and the LSS BEV branch of a decoder iteration,
 * models/racformer_transformer.py:229-236: same spatial_shapes, batch, heads, queries, points; different value maps, sampling
 * locations and attention weights) in ONE launch (grid.y = 2): same arithmetic as two racf_msda_forward calls, bit for bit.
 */
int racf_msda_forward_pair(const float* value_a, const float* loc_a, const float* attn_a, float* out_a,
                           const float* value_b, const float* loc_b, const float* attn_b, float* out_b,
                           const int64_t* spatial_shapes, const int64_t* level_start_index, int batch,
                           int spatial_size, int num_heads, int head_dim, int num_levels, int num_query,
                           int num_point, int im2col_step, racf_stream_t stream);

/*
 * Multi-scale deformable attention, backward.
 *   grad_out   : [B, Q, M*D]
 *   grad_value : [B, S, M, D]          accumulated with atomics into the caller's (pre-zeroed) buffer,
 *                                      as mmcv does (multi_scale_deformable_attn_function.py:146-160)
 *   grad_loc   : [B, Q, M, L, P, 2]    fully overwritten
 *   grad_attn  : [B, Q, M, L, P]       fully overwritten
 */
int racf_msda_backward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                       const float* loc, const float* attn, const float* grad_out,
                       int batch, int spatial_size, int num_heads, int head_dim,
                       int num_levels, int num_query, int num_point, int im2col_step,
                       float* grad_value, float* grad_loc, float* grad_attn, racf_stream_t stream);

/*
 * Debug/verification entry for MSDA: tap_mask[b,q,m,l,p] uint8 with the same bit layout as
 * racf_msmv_tap_masks (bit0 in range, bit1..4 corners read).
 */
int racf_msda_tap_masks(const int64_t* spatial_shapes, const float* loc,
                        int batch, int num_heads, int num_levels, int num_query, int num_point,
                        uint8_t* tap_mask, racf_stream_t stream);

/*
 * ---- "next" row: fused sampling-point generation, forward / inference only ----------
 *
 * racf_msmv_points_forward replaces the PyTorch op chain of RaCFormerSampling.inner_forward
 * (models/racformer_transformer.py:361-408) plus the projection / view-selection / packing half of sampling_4d
 * (models/sparsebev_sampling.py:45-120). Device inputs (fp32, contiguous):
 *   query_ray [B,Q,10]   polar query boxes            offset    [B,Q,G*Pn*D,3]  sampling_offset(query_feat)
 *   ray_logit [B,Q,D]    ray_points_offset(query_feat) scale_raw [B,Q,G,T,Pn*D,L] scale_weights(query_feat), pre-softmax
 *   time_diff [B,T]      lidar2img [B,T*N,4,4]         depth_base [D]            linspace(-d_region, d_region, D)
 * Host inputs: pc_range (6 doubles). Outputs: loc [B*T*G,Q,Pn*D,3] and weights [B*G*T,Q,Pn*D,L] -- exactly the two
 * tensors sampling_4d hands to msmv_sampling (including the reference's B*T*G vs B*G*T packing).
 */
int racf_msmv_points_forward(const float* query_ray, const float* offset, const float* ray_logit,
                             const float* scale_raw, const float* time_diff, const float* lidar2img,
                             const float* depth_base, const double* pc_range, float d_region, float image_w,
                             float image_h, float eps, int batch, int num_query, int num_frames, int num_groups,
                             int num_points, int depth_num, int num_views, int num_levels,
                             float* loc, float* weights, racf_stream_t stream);

/*
 * racf_bev_points_forward replaces BEVSampling.inner_forward's point math (models/racformer_transformer.py:493-529)
 * plus the queue-major packing of BEVSelfAttention.forward (models/bev_self_attention.py:176-188), num_levels == 1:
 *   offset [B,Q,M*Pn*D,2], attn_raw [B,Q,M,Pn*D] (pre-softmax) ->
 *   loc [T*B,Q,M,1,Pn*D,2], attn [T*B,Q,M,1,Pn*D] (softmax over the Pn*D points, repeated over the T frames).
 */
int racf_bev_points_forward(const float* query_ray, const float* offset, const float* ray_logit,
                            const float* attn_raw, const float* time_diff, const float* depth_base,
                            const double* pc_range, float d_region, int batch, int num_query, int num_frames,
                            int num_heads, int num_points, int depth_num, float* loc, float* attn,
                            racf_stream_t stream);

/*
 * Training (autograd) counterparts: backward of the two point kernels. They take the forward's inputs and outputs and the
 * location / weight gradients the sampling ops' backward returns, and write the gradients of the Linear heads' outputs
 * (fully overwritten): grad_offset like offset, grad_ray_logit [B,Q,D], grad_scale_raw / grad_attn_raw like scale_raw /
 * attn_raw, and -- when grad_ray is not NULL -- grad_ray [B,Q,10] (the velocity columns are detached in the reference's warp,
 * models/racformer_transformer.py:381, so they get 0). The camera view of a point is not differentiable; it is read back
 * from loc[...,2]. grad_loc [B*T*G,Q,Pn*D,3] (the third component is ignored) / grad_weights [B*G*T,Q,Pn*D,L] are what
 * racf_msmv_backward produces; grad_loc [T*B,Q,M,1,Pn*D,2] / grad_attn [T*B,Q,M,1,Pn*D] what racf_msda_backward produces.
 */
int racf_msmv_points_backward(const float* query_ray, const float* offset, const float* ray_logit,
                              const float* time_diff, const float* lidar2img, const float* depth_base,
                              const double* pc_range, float d_region, float image_w, float image_h, float eps,
                              int batch, int num_query, int num_frames, int num_groups, int num_points,
                              int depth_num, int num_views, int num_levels, const float* loc,
                              const float* weights, const float* grad_loc, const float* grad_weights,
                              float* grad_ray, float* grad_offset, float* grad_ray_logit, float* grad_scale_raw,
                              racf_stream_t stream);
int racf_bev_points_backward(const float* query_ray, const float* offset, const float* ray_logit,
                             const float* time_diff, const float* depth_base, const double* pc_range,
                             float d_region, int batch, int num_query, int num_frames, int num_heads,
                             int num_points, int depth_num, const float* attn, const float* grad_loc,
                             const float* grad_attn, float* grad_ray, float* grad_offset, float* grad_ray_logit,
                             float* grad_attn_raw, racf_stream_t stream);

/*
 * "next" row: channel-last re-layout of one FPN level,
 * in [B, T*N, G*C, H, W] -> out [B*T*G, N, H, W, C] (models/racformer_transformer.py:112-124), C == 64.
 */
int racf_to_sampling_layout(const float* in, float* out, int batch, int num_frames, int num_views,
                            int num_groups, int channels, int height, int width, racf_stream_t stream);

/*
 * Inverse of racf_to_sampling_layout (its backward): in [B*T*G, N, H, W, C] -> out [B, T*N, G*C, H, W], C == 64.
 */
int racf_from_sampling_layout(const float* in, float* out, int batch, int num_frames, int num_views,
                              int num_groups, int channels, int height, int width, racf_stream_t stream);

/*
 * The same re-layout from fp16 storage, upcast fused into the copy: `in` holds IEEE binary16 values, `out` is the fp32
 * sampling layout the MSMV kernels read. The reference's image branch runs in fp16 and casts its FPN outputs up to fp32
 * (models/racformer.py:106 `auto_fp16(apply_to=('img'), out_fp32=True)`), so the result is bit-identical to the
 * reference's fp32 tensors; what changes is that the 2-byte form is what is transported / re-read.
 *   channels_last == 0 : in [B, T*N, G*C, H, W] halves
 *   channels_last != 0 : in [B*T*N, H, W, G*C] halves (the layout cuDNN's fp16 convolutions emit), 16-byte aligned
 */
int racf_to_sampling_layout_f16(const void* in, float* out, int batch, int num_frames, int num_views,
                                int num_groups, int channels, int height, int width, int channels_last,
                                racf_stream_t stream);

/*
 * "next" row: BEVPoolv2, the reference's other in-tree native op
 * (models/csrc/bev_pool_v2/src/bev_pool.cpp:30-104 -> bev_pool_cuda.cu:125-140). All pointers are device pointers.
 *   depth [b,n,d,h,w] fp32, feat [b,n,h,w,c] fp32, ranks_* int32 [n_points], interval_* int32 [n_intervals].
 * forward : out[ranks_bev[s], :] = sum over the interval of feat[ranks_feat[i], :] * depth[ranks_depth[i]];
 *           cells without an interval are not touched (the caller zero-fills `out`, bev_pool.py:30).
 * backward: intervals are runs of equal ranks_feat (bev_pool.py:50-60); depth_grad / feat_grad entries that belong
 *           to no point are not touched (the caller zero-fills them, bev_pool.py:70-71).
 */
int racf_bev_pool_v2_forward(const float* depth, const float* feat, const int* ranks_depth, const int* ranks_feat,
                             const int* ranks_bev, const int* interval_starts, const int* interval_lengths,
                             int n_intervals, int channels, float* out, racf_stream_t stream);
int racf_bev_pool_v2_backward(const float* out_grad, const float* depth, const float* feat, const int* ranks_depth,
                              const int* ranks_feat, const int* ranks_bev, const int* interval_starts,
                              const int* interval_lengths, int n_intervals, int channels,
                              float* depth_grad, float* feat_grad, racf_stream_t stream);

/*
 * "next" row: fused core of AdaptiveMixing for inference
 * (models/racformer_transformer.py:592-604): per (query, group)
 *     t = relu(layer_norm(x @ M));  out = relu(layer_norm(S @ t))
 * x [QG, in_points, C], params [QG, C*C + out_points*in_points] (M then S, as parameter_generator emits them),
 * out [QG, out_points, C]; layer norms over the whole [points, C] slab, no affine. fp32 FMA on the CUDA cores.
 * Implemented for C == 64, out_points == 128, in_points % 4 == 0, in_points <= 128 (else RACF_ERR_UNSUPPORTED).
 */
int racf_adaptive_mixing_forward(const float* x, const float* params, int num_query_groups, int in_points,
                                 int out_points, int channels, float eps, float* out, racf_stream_t stream);
/* Same, but the result is written as three bf16 pieces out3 [3][QG][out_points][C] with value == p0 + p1 + p2 exactly:
 * the A operand of racf_linear_bf16x3_forward (out_proj), without an fp32 round trip through HBM. */
/* tiled_groups > 0: out3 is instead the pre-tiled A operand (racf_linear_tiled_bytes(QG / tiled_groups, tiled_groups *
 * out_points * C) bytes) of a [QG / tiled_groups, tiled_groups * out_points * C] matrix, tiled_groups = n_groups. */
int racf_adaptive_mixing_forward_split(const float* x, const float* params, int num_query_groups, int in_points,
                                       int out_points, int channels, float eps, void* out3, int tiled_groups,
                                       racf_stream_t stream);

/*
 * The same AdaptiveMixing core on the tcgen05 tensor cores (csrc/mixing_tc.cu): both products as bf16x3 operand-split
 * MMAs with fp32 accumulation in tensor memory (six piece products, large / cross terms in separate accumulators), the
 * layer norms in fp32 straight from tensor memory. Exactly one of `out` (fp32 [QG, out_points, C]) and `out3` (the tiled
 * A operand of out_proj, tiled_groups = n_groups, see racf_adaptive_mixing_forward_split) is non-NULL.
 * Implemented for C == 64, out_points == 128, in_points % 16 == 0, 16 <= in_points <= 128 (else RACF_ERR_UNSUPPORTED).
 */
int racf_adaptive_mixing_tc_forward(const float* x, const float* params, int num_query_groups, int in_points,
                                    int out_points, int channels, float eps, float* out, void* out3,
                                    int tiled_groups, racf_stream_t stream);
/* The same with the kernel chosen explicitly (measurement and parity of both versions): variant 0 = default (the
 * warp-specialised kernel csrc/mixing_ws.cu for in_points <= 96, else the phase-serial kernel csrc/mixing_tc.cu),
 * 1 = phase-serial, 2 = warp-specialised (RACF_ERR_UNSUPPORTED for in_points > 96). Results agree to fp32 rounding of the
 * layer-norm sums (different summation trees), not bit for bit. */
int racf_adaptive_mixing_tc_forward_variant(const float* x, const float* params, int num_query_groups, int in_points,
                                            int out_points, int channels, float eps, float* out, void* out3,
                                            int tiled_groups, int variant, racf_stream_t stream);

/*
 * Backward of the AdaptiveMixing core for training (autograd of models/racformer_transformer.py:592-604): given
 * grad_out = dL/d(out) [QG, out_points, C] it recomputes the forward from x and params (nothing else is saved) and writes
 * grad_x [QG, in_points, C] and grad_params [QG, C*C + out_points*in_points] (dL/dM then dL/dS), both fully overwritten.
 * fp32 FMA on the CUDA cores. C == 64, out_points == 128, in_points % 16 == 0, 16 <= in_points <= 128, 16-byte aligned
 * pointers (RACF_ERR_UNSUPPORTED otherwise).
 */
int racf_adaptive_mixing_backward(const float* x, const float* params, const float* grad_out, int num_query_groups,
                                  int in_points, int out_points, int channels, float eps, float* grad_x,
                                  float* grad_params, racf_stream_t stream);
/* The same with an explicit kernel choice: variant 0 = default, 1 = fp32 FMA on the CUDA cores (csrc/mixing_bwd.cu),
 * 2 = all six products on the tensor cores (tcgen05, exact bf16x3 operand splitting, csrc/mixing_bwd_tc.cu; in_points <= 96,
 * else RACF_ERR_UNSUPPORTED). Results agree to fp32 rounding of the products. */
int racf_adaptive_mixing_backward_variant(const float* x, const float* params, const float* grad_out, int num_query_groups,
                                          int in_points, int out_points, int channels, float eps, float* grad_x,
                                          float* grad_params, int variant, racf_stream_t stream);


/*
 * "next" row: AdaptiveMixing's two large Linear layers (parameter_generator and out_proj,
 * models/racformer_transformer.py:560-566, F.linear in fp32) on the tcgen05 tensor cores at fp32-grade accuracy.
 *
 * racf_split_bf16x3: x [count] fp32 -> out3 [3][count] bf16 with x == out3[0] + out3[1] + out3[2] exactly
 *   (count % 4 == 0, x 16-byte aligned).
 * racf_linear_bf16x3_plan: the K split and workspace size racf_linear_bf16x3_forward wants for a problem
 *   (K per accumulator is kept <= 512 to bound the tensor cores' truncating accumulation).
 * racf_linear_bf16x3_forward: out[M,N] = a[M,K] . w[N,K]^T + bias[N] with a3 = split(a) [3][M][K], w3 = split(w)
 *   [3][N][K] (nn.Linear's weight layout), products a_i * w_j with i + j <= max_order accumulated in fp32 (4: all nine,
 *   2: six). bias may be NULL. K % 8 == 0. workspace: split_k * M * N floats when split_k > 1 (else may be NULL).
 *   variant: 0 = plain pieces through tensor maps, 32-wide K blocks / 64-byte swizzle / 2 stages; 1 = the same with
 *   64-wide K blocks / 128-byte swizzle / 1 stage; 2 = a3 and w3 are in the tiled format (bulk copies, any K);
 *   3 = tiled operands, 128 x 256 output tiles on persistent warp-specialised CTAs (csrc/linear_wide.cu: 25 % less
 *   L2 -> SM operand traffic, the binding resource of variant 2). Variants 2 and 3 issue the same MMAs in the same order
 *   per accumulator and give bit-identical results.
 */
int racf_split_bf16x3(const float* x, long long count, void* out3, racf_stream_t stream);
/* (in [batch][channels][positions] + pos [channels][positions] (may be NULL)) -> out3 [3][batch * positions][channels]:
 * the add, permute and copy in front of BEVSelfAttention.value_proj (models/bev_self_attention.py:162-174) fused with the
 * operand split; channels % 8 == 0. */
int racf_split_bf16x3_chw_to_hwc(const float* in, const float* pos, int batch, int channels, int positions,
                                 int tiled, void* out3, racf_stream_t stream);
/*
 * The kernel's preferred operand format ("tiled", variant 2 / tiled != 0 below): the three bf16 pieces cut into
 * [128 rows][32 k] tiles stored in global memory as the shared-memory image the MMA reads (K-major, 64-byte swizzle),
 * ordered [row tile][k block][piece], so that a pipeline stage is one contiguous 24 KB bulk copy
 * (racformer_b200/csrc/linear_tiled.cuh). racf_linear_tiled_bytes: buffer size for a [rows][K] matrix;
 * racf_split_bf16x3_tiled: fp32 [rows][K] row-major -> tiled pieces (any K; the K tail is zero-filled).
 */
long long racf_linear_tiled_bytes(long long rows, int K);
int racf_split_bf16x3_tiled(const float* x, long long rows, int K, void* out, racf_stream_t stream);
/* The same with a row-periodic fp32 addend: splits x[row][k] + addend[row % addend_rows][k] (addend may be NULL). Used
 * for value_proj when the BEV maps are channel-last: x = the [B*T*H*W, C] pixel matrix, addend = the [H*W, C] positional
 * encoding (models/bev_self_attention.py:162-174 adds it with a separate pass). */
int racf_split_bf16x3_tiled_add(const float* x, long long rows, int K, const float* addend, long long addend_rows,
                                void* out, racf_stream_t stream);
int racf_linear_bf16x3_plan(int M, int N, int K, int* split_k, long long* workspace_bytes);
int racf_linear_bf16x3_forward(const void* a3, const void* w3, const float* bias, int M, int N, int K,
                               int max_order, int split_k, int variant, float* workspace, float* out,
                               racf_stream_t stream);

/*
 * Several Linear layers that share one input (the sampling heads of a decoder iteration all read the same query
 * features: sampling_offset / scale_weights / ray_points_offset of RaCFormerSampling and the two BEVSampling branches,
 * bev_queue_weight; models/racformer_transformer.py:361-366,493-496, models/bev_self_attention.py:193) in ONE launch:
 * out_i[M, seg_n[i]] = a . w_i^T + bias_i. w3 is [3][sum_i pad128(seg_n[i])][K]: the bf16 pieces of the weights stacked
 * along N, each layer padded with zero rows to a multiple of 128. K <= 512, num_segments <= RACF_LINEAR_MAX_SEGMENTS.
 * seg_n / seg_bias / seg_out are HOST arrays (of device pointers); seg_bias or its entries may be NULL.
 */
#define RACF_LINEAR_MAX_SEGMENTS 16
int racf_linear_bf16x3_multi_forward(const void* a3, const void* w3, int M, int K, int num_segments,
                                     const int* seg_n, const float* const* seg_bias, float* const* seg_out,
                                     int max_order, int tiled, racf_stream_t stream);

/*
 * Radar temporal encoder (models/racformer_transformer.py:709-720, ConvGRUCell.forward after the gates convolution), channel-
 * last tensors: gates [pixels, 3 * hidden] = (z | r | cand) pre-activations, h_prev [pixels, hidden] ->
 * h = (1 - sigmoid(z)) * h_prev + sigmoid(z) * tanh(cand + sigmoid(r) * h_prev), in PyTorch's fp32 operation order.
 * hidden % 4 == 0, 16-byte aligned pointers (else RACF_ERR_UNSUPPORTED). Inference only.
 */
int racf_convgru_gates_forward(const float* gates, const float* h_prev, long long pixels, int hidden_channels,
                               float* h, racf_stream_t stream);
/* nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True) of a channel-last tensor (the same encoder,
 * models/racformer_transformer.py:637-640): in [batch, height, width, channels] -> out [batch, 2 height, 2 width, channels].
 * channels % 4 == 0, 16-byte aligned pointers (else RACF_ERR_UNSUPPORTED). Inference only. */
int racf_upsample2x_bilinear_nhwc(const float* in, int batch, int height, int width, int channels, float* out,
                                  racf_stream_t stream);

/*
 * in [batch][channels][positions] -> out[batch][positions][ld] (the first `channels` floats of every ld-long pixel row)
 * and, when out2 != NULL, the same values into out2 with row length ld2: NCHW -> NHWC in front of the radar temporal
 * encoder's convolutions (models/racformer_transformer.py:645-656), a tiled transpose.
 */
int racf_chw_to_hwc(const float* in, int batch, int channels, int positions, float* out, int ld, float* out2,
                    int ld2, racf_stream_t stream);

/*
 * Call-site row: the attention core of ScaleAdaptiveSelfAttention
 * (models/racformer_transformer.py:283-336: pairwise centre distances, dist * tau as the additive mask of mmcv's
 * MultiheadAttention, softmax, weighted sum) as one launch, without materialising the [B, H, Q, Q] mask.
 *   qkv       : [batch * num_query, 3 * H * head_dim]  in_proj output (q | k | v), head h at columns h * head_dim
 *   tau       : [batch * num_query, H]                  gen_tau output
 *   query_ray : [batch * num_query, code_size]          (theta, d, ...): centre = decode_bbox(theta_d2xy_coods(.))[:2]
 *   pc_range  : 6 host doubles
 *   out       : [batch * num_query, H * head_dim]       heads merged, ready for out_proj
 * score = q.k / sqrt(head_dim) - tau[i, h] * |centre_i - centre_j|. head_dim must be 32. Inference only (no extra mask).
 */
int racf_sasa_attention_forward(const float* qkv, const float* tau, const float* query_ray, const double* pc_range,
                                int batch, int num_query, int num_heads, int head_dim, int code_size,
                                float* out, racf_stream_t stream);

/*
 * The same attention core for TRAINING (autograd of models/racformer_transformer.py:296-336 incl. the query-denoising mask
 * of models/racformer_head.py:205-206 and MultiheadAttention's attention dropout): forward with an optional blocked-pair
 * mask and dropout, saving only the log-sum-exp of every row; backward recomputes the probabilities.
 *   blocked_t : uint8 [num_query, num_query] TRANSPOSED (blocked_t[j * Q + i] != 0: query i may not attend to key j) or NULL
 *   drop_p    : attention dropout probability (0: none); the keep mask is a hash of (seed, batch, head, query, key)
 *   lse       : [batch, num_heads, num_query]   dsum: same shape, scratch written by the backward
 *   grad_qkv  : [batch * num_query, 3E] (dq | dk | dv), grad_tau [batch * num_query, num_heads]; both fully overwritten
 * No gradient flows to query_ray (the reference computes the centre distances under no_grad). head_dim == 32.
 */
int racf_sasa_attention_train_forward(const float* qkv, const float* tau, const float* query_ray,
                                      const uint8_t* blocked_t, const double* pc_range, int batch, int num_query,
                                      int num_heads, int head_dim, int code_size, float drop_p, unsigned seed,
                                      float* out, float* lse, racf_stream_t stream);
int racf_sasa_attention_train_backward(const float* qkv, const float* tau, const float* query_ray,
                                       const uint8_t* blocked_t, const double* pc_range, int batch, int num_query,
                                       int num_heads, int head_dim, int code_size, float drop_p, unsigned seed,
                                       const float* out, const float* lse, const float* grad_out, float* dsum,
                                       float* grad_qkv, float* grad_tau, racf_stream_t stream);

/*
 * Call-site row: box refinement at the end of a decoder iteration
 * (models/racformer_transformer.py:255-259,264-279 and theta_d2xy_coods, models/bbox/utils.py:82-90) as one launch.
 *   proposal, delta : [batch * num_query, code_size]  the iteration's input rays and the reg branch output
 *   time_diff       : [batch, num_frames] or NULL; with num_frames > 1 the velocity columns (>= 8) are divided by
 *                     time_diff[b, 1] (1 where it is < 1e-5)
 *   pred            : (theta + (2 sigmoid(delta_0) - 1) / num_ray, sigmoid(delta_1:3 + inverse_sigmoid(proposal_1:3)),
 *                     delta_3:)                                              -- the next iteration's query rays
 *   pred_xy         : pred with (theta, d) mapped to clamped normalised (x, y)  -- the iteration's box output
 */
int racf_refine_bbox_forward(const float* proposal, const float* delta, const float* time_diff, int batch,
                             int num_query, int num_frames, int code_size, float num_ray, float* pred,
                             float* pred_xy, racf_stream_t stream);

/* Backward of racf_refine_bbox_forward w.r.t. its pred_xy output (pred feeds the next iteration detached,
 * models/racformer_transformer.py:132): grad_delta [rows, code_size] and, when not NULL, grad_proposal (both overwritten). */
int racf_refine_bbox_backward(const float* proposal, const float* delta, const float* time_diff,
                              const float* grad_pred_xy, int batch, int num_query, int num_frames, int code_size,
                              float num_ray, float* grad_delta, float* grad_proposal, racf_stream_t stream);

/*
 * Call-site row: the row-wise operators between the sampling ops of one decoder iteration
 * (models/racformer_transformer.py:204-262: position_encoder, norm1..3, fusion, FFN, cls / reg branches; the softmax
 * queue fusion + output_proj of models/bev_self_attention.py:206-225) as ONE launch. Rows (queries) never interact, so
 * each CTA carries rows_per_cta rows through the whole chain in shared memory, interpreting a short program.
 *
 * A program works on `num_bufs` shared-memory buffers of [rows_per_cta][width] floats. Operators (kind):
 *   LOAD           buf[dst][r][dst_col + c] = p0[row * ld + c],                               c < n
 *   LOAD_QUEUE     buf[dst][r][dst_col + c] = sum_t softmax_t(p1[row * aux + t]) * p0[((b * aux + t) * k + q) * ld + c]
 *                  with row = b * k + q (k = rows per batch element, aux = queue length <= RACF_ROW_MAX_QUEUE;
 *                  p1 == NULL: plain mean over the queue)
 *   STORE          out[row * ld + c] = buf[src][r][src_col + c],                              c < n
 *   ADD            buf[dst][.. dst_col + c] += buf[src][.. src_col + c],                      c < n
 *   LINEAR         buf[dst][.. dst_col + j] = act(p1[j] + sum_{i<k} buf[src][.. src_col + i] * p0[i * n + j]),  j < n
 *                  p0 = the nn.Linear weight W [n][k] as its chunked transpose [ceil(n / 256)][k][256]
 *                  (element [c][i][jj] = W[c * 256 + jj][i], zero for c * 256 + jj >= n; 16-byte aligned),
 *                  p1 = bias or NULL; dst != src, src_col % 4 == 0
 *   LINEAR_NARROW  the same for a few output columns with p0 = the weight in nn.Linear's own layout [n][k]
 *                  (n * k <= 8192, a multiple of 4, 16-byte aligned)
 *   LAYERNORM      in place over buf[dst][.. dst_col .. dst_col + n), gamma p0 / beta p1 (either may be NULL), eps
 * flags & RACF_ROWOP_RELU applies max(., 0) to the operator's result (LINEAR*, LAYERNORM).
 * flags & RACF_ROWOP_ACCUM (LOAD, LINEAR): the result is added to buf[dst] instead of replacing it.
 * Operators of the training programs (the backward pass of a chain is itself a row program over gradient buffers, built by
 * the host from the forward program; the forward saves what it needs with STORE operators):
 *   ZERO           buf[dst][.. dst_col + c] = 0
 *   RELU_MASK      buf[dst][.. dst_col + c] = 0 where the saved forward output p0[row * ld + c] <= 0
 *   DROPOUT        buf[dst][.. dst_col + c] *= keep / (1 - eps), keep = [hash(aux, row, k + c) >= eps] (eps = drop probability,
 *                  aux = seed, k = column offset of the stream): the same record replays the same mask in the backward program
 *   LAYERNORM_BWD  in place on the gradient in buf[dst][.. dst_col .. dst_col + n): p2 = the LayerNorm's saved INPUT rows
 *                  [rows][ld] (mean / rstd are recomputed), gamma p0 / beta p1, RELU flag = the forward applied max(., 0);
 *                  adds the column sums of g * xhat to out[c] (d gamma) and of g to out2[c] (d beta) with one atomicAdd per
 *                  CTA and column (either may be NULL). n <= 384.
 *   STORE_COLSUM   out[row * ld + c] = buf[src][.. src_col + c] (out may be NULL) and out2[c] += sum over the CTA's rows
 *                  (atomicAdd; out2 may be NULL): a Linear's output gradient for the weight-gradient GEMM + its bias gradient
 *   QUEUE_BWD      backward of LOAD_QUEUE: g = buf[src][.. src_col .. + n); out[((b * aux + t) * k + q) * ld + c] = w_t * g[c]
 *                  (gradient of the queue values) and out2[row * aux + t] = w_t * (d_t - sum_s w_s d_s), d_t = <g, value_t>
 *                  (gradient of the logits; skipped when p1 == NULL)
 * LOAD_QUEUE needs n, ld, dst_col multiples of 4 and a 16-byte aligned p0.
 * fp32 FMA on the CUDA cores, bias added last. `ops` is a HOST array, copied into the kernel parameters
 * (capture-safe). rows_per_cta is 4 .. 8; 128 KB (weight tiles) + (num_bufs * width + 768) * rows_per_cta * 4 bytes <= 226 KB.
 */
#define RACF_ROW_MAX_OPS        128
#define RACF_ROW_MAX_QUEUE      16
#define RACF_ROW_CHUNK_COLS    256
#define RACF_ROWOP_LOAD          1
#define RACF_ROWOP_LOAD_QUEUE    2
#define RACF_ROWOP_STORE         3
#define RACF_ROWOP_ADD           4
#define RACF_ROWOP_LINEAR        5
#define RACF_ROWOP_LINEAR_NARROW 6
#define RACF_ROWOP_LAYERNORM     7
#define RACF_ROWOP_ZERO          8
#define RACF_ROWOP_RELU_MASK     9
#define RACF_ROWOP_DROPOUT      10
#define RACF_ROWOP_LAYERNORM_BWD 11
#define RACF_ROWOP_STORE_COLSUM 12
#define RACF_ROWOP_QUEUE_BWD    13
#define RACF_ROWOP_RELU          1   /* flags bit */
#define RACF_ROWOP_ACCUM         2   /* flags bit */
typedef struct racf_row_op {
    int kind;
    int dst, dst_col;
    int src, src_col;
    int n, k;
    int flags;
    int ld;
    int aux;
    float eps;
    const float* p0;
    const float* p1;
    float* out;
    const float* p2;
    float* out2;
} racf_row_op_t;
int racf_row_program_forward(const racf_row_op_t* ops, int num_ops, int rows, int rows_per_cta, int num_bufs,
                             int width, racf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RACFORMER_OPS_H_ */
