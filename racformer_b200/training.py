"""Training-mode (autograd) bindings of the fused kernels around the sampling ops.

The reference trains the decoder through `MSMVSamplingC2345` / `MultiScaleDeformableAttnFunction_fp32`
(models/csrc/wrapper.py:78-142, models/multi_scale_deformable_attn_function.py:90-163) plus PyTorch ops for everything
between them. The Functions here keep those two contracts (inputs saved with `save_for_backward` only, fp32, gradients
for features / locations / weights) and add what the B200 path needs to stop re-reading and re-adding 1.5 GB tensors:

* `SamplingLayout`      : channel-last re-layout of an FPN level (racformer_transformer.py:112-124) with the tiled-transpose
                          kernels in both directions instead of PyTorch's strided copies.
* `MSMVGrouped`         : msmv_sampling with sampling_4d's un-packing (sparsebev_sampling.py:128-131) fused into the forward
                          output AND into the backward's grad_out read -- no [B',Q,C,P] <-> [B,Q,G,T*P,C] permute copies.
* `SharedGrad`          : the six decoder iterations read the SAME pyramid / BEV value maps (one shared layer,
                          racformer_transformer.py:84-89). Autograd would materialise six zero-filled 1.5 GB feature gradients
                          and add them pairwise; here every backward scatters into one shared buffer (one zero-fill) and only
                          the last backward to run hands the buffer to autograd (the others return None = zero).
* `MSDAShared`          : MultiScaleDeformableAttnFunction_fp32 with the value gradient accumulated the same way.
* `MSMVPoints` / `BEVPoints` / `RefineBBox` : the point-generation chains in front of the two sampling ops and the box
                          refinement behind the reg branch (racformer_transformer.py:361-408, 493-529, 255-279) as one
                          forward and one backward launch each (csrc/points.cu, csrc/points_train.cu) instead of ~400
                          autograd nodes per branch and iteration.

No CPU fallback; nothing here imports `oracle/`.
"""
import ctypes

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib, points, wrapper
from .multi_scale_deformable_attn_function import ext_module


class SharedGrad:
    """One gradient buffer (or list of buffers) shared by several autograd nodes that consume the same tensors.

    forward of every consumer calls `join()`; backward calls `buffers()` to get the (lazily zero-filled) buffers to
    accumulate into and then `leave()`, which returns the buffers when this was the last outstanding consumer (they are
    then returned to autograd as the gradient) and None otherwise (autograd treats None as zero)."""

    def __init__(self, like):
        self.like = list(like)
        self.users = 0
        self.bufs = None

    def join(self):
        self.users += 1

    def buffers(self):
        if self.bufs is None:
            self.bufs = [torch.zeros_like(t) for t in self.like]
        return self.bufs

    def leave(self):
        self.users -= 1
        if self.users == 0:
            bufs, self.bufs = self.buffers(), None
            return bufs
        return None


class SharedGrads:
    """Per-forward registry: tensors (by identity of the first tensor) -> SharedGrad."""

    def __init__(self):
        self._by_key = {}

    def get(self, tensors):
        key = tuple(t.data_ptr() for t in tensors)
        sg = self._by_key.get(key)
        if sg is None:
            sg = self._by_key[key] = SharedGrad(tensors)
        return sg


class SamplingLayout(Function):
    """[B, T*N, G*C, H, W] -> [B*T*G, N, H, W, C] (C == 64), both directions on the tiled-transpose kernels."""

    @staticmethod
    def forward(ctx, feat, num_cams, num_groups):
        ctx.shape, ctx.num_cams, ctx.num_groups = feat.shape, num_cams, num_groups
        return points.to_sampling_layout(feat.contiguous(), num_cams, num_groups)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad):
        B, TN, GC, H, W = ctx.shape
        N, G = ctx.num_cams, ctx.num_groups
        grad = grad.contiguous()
        out = torch.empty(ctx.shape, dtype=torch.float32, device=grad.device)
        with torch.cuda.device(grad.device):
            rc = _lib.load().racf_from_sampling_layout(grad.data_ptr(), out.data_ptr(), B, TN // N, N, G, GC // G, H, W,
                                                       wrapper._stream(grad.device))
        _lib.check(rc, "racf_from_sampling_layout")
        return out, None, None


class MSMVGrouped(Function):
    """apply(shared, num_frames, num_groups, loc [B*T*G,Q,P,3], weights [B*T*G,Q,P,L], *feats) -> [B, Q, G, T*P, C].
    `shared`: a SharedGrad over `feats` or None (then the feature gradients are fresh zero-filled tensors)."""

    @staticmethod
    def forward(ctx, shared, num_frames, num_groups, loc, weights, *feats):
        ctx.shared, ctx.T, ctx.G = shared, int(num_frames), int(num_groups)
        if shared is not None:
            shared.join()
        ctx.save_for_backward(loc, weights, *feats)
        return wrapper.msmv_forward_grouped(feats, loc, weights, num_frames, num_groups)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        loc, weights, *feats = ctx.saved_tensors
        Bp, N, C, Q, P = wrapper._check_inputs(feats, loc, weights)
        grad_out = grad_out.contiguous()
        shared = ctx.shared
        grad_feats = shared.buffers() if shared is not None else [torch.empty_like(f) for f in feats]
        grad_loc, grad_w = torch.empty_like(loc), torch.empty_like(weights)
        with torch.cuda.device(loc.device):
            rc = _lib.load().racf_msmv_backward_grouped(
                grad_out.data_ptr(), wrapper._ptr_array(feats), wrapper._hw_array(feats), len(feats), loc.data_ptr(),
                weights.data_ptr(), Bp, C, N, Q, P, ctx.T, ctx.G, wrapper._ptr_array(grad_feats), grad_loc.data_ptr(),
                grad_w.data_ptr(), 0 if shared is not None else 1, wrapper._stream(loc.device))
        _lib.check(rc, "racf_msmv_backward_grouped")
        if shared is not None:
            grad_feats = shared.leave() or [None] * len(feats)
        return (None, None, None, grad_loc, grad_w, *grad_feats)


def msmv_grouped_supported(feats, loc, weights):
    f = feats[0]
    return (f.is_cuda and f.dtype == torch.float32 and f.shape[-1] == 64 and len(feats) in (2, 4, 5)
            and all(t.data_ptr() % 16 == 0 for t in feats))


class MSDAShared(Function):
    """MultiScaleDeformableAttnFunction_fp32 whose value gradient accumulates into a SharedGrad (the hoisted value map is
    read by all six iterations). apply(shared, value, shapes, level_start, loc, attn, im2col_step)."""

    @staticmethod
    def forward(ctx, shared, value, shapes, lsi, loc, attn, im2col_step):
        ctx.shared, ctx.im2col_step = shared, im2col_step
        shared.join()
        ctx.save_for_backward(value, shapes, lsi, loc, attn)
        return ext_module.ms_deform_attn_forward(value, shapes, lsi, loc, attn, im2col_step=im2col_step)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        grad_value = ctx.shared.buffers()[0]
        grad_loc, grad_attn = torch.empty_like(loc), torch.empty_like(attn)
        ext_module.ms_deform_attn_backward(value, shapes, lsi, loc, attn, grad_out.contiguous(), grad_value, grad_loc,
                                           grad_attn, im2col_step=ctx.im2col_step)
        done = ctx.shared.leave()
        return None, (done[0] if done is not None else None), None, None, grad_loc, grad_attn, None


class AdaptiveMixingCore(Function):
    """relu(LN(S @ relu(LN(x @ M)))) per (query, group) -- models/racformer_transformer.py:592-604 -- as one forward kernel
    (tcgen05, csrc/mixing_tc.cu) and one backward kernel that recomputes the forward (csrc/mixing_bwd.cu); only x and the
    generated parameters are saved. apply(x [QG,P_in,C], params [QG, C*C + P_out*P_in], out_points) -> [QG, P_out, C]."""

    @staticmethod
    def supported(x, params, out_points):
        QG, P_in, C = x.shape
        return (x.is_cuda and x.dtype == torch.float32 and params.dtype == torch.float32 and C == 64 and out_points == 128
                and P_in % 16 == 0 and 16 <= P_in <= 128)

    @staticmethod
    def forward(ctx, x, params, out_points):
        x, params = x.contiguous(), params.contiguous()
        ctx.out_points = out_points
        ctx.save_for_backward(x, params)
        return points.adaptive_mixing_core(x, params, out_points)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        x, params = ctx.saved_tensors
        gx, gp = points.adaptive_mixing_core_backward(x, params, grad_out.contiguous(), ctx.out_points)
        return gx, gp, None


class SasaAttention(Function):
    """Attention core of ScaleAdaptiveSelfAttention with autograd (csrc/sasa_train.cu): apply(qkv [B,Q,3E], tau [B,Q,H],
    query_ray [B,Q,code], blocked (bool [Q,Q], True = query i may not see key j, or None), pc_range, num_heads, drop_p, seed)
    -> [B,Q,E]. Saves qkv, tau, the output and one log-sum-exp per row; the [B,H,Q,Q] mask / probabilities never exist."""

    @staticmethod
    def supported(qkv, tau, num_heads):
        return (qkv.is_cuda and qkv.dtype == torch.float32 and tau.dtype == torch.float32 and qkv.shape[-1] == 3 * num_heads * 32
                and qkv.shape[1] <= 4096)

    @staticmethod
    def forward(ctx, qkv, tau, query_ray, blocked, pc_range, num_heads, drop_p, seed):
        qkv, tau, query_ray = qkv.contiguous(), tau.contiguous(), query_ray.detach().contiguous()
        B, Q, E3 = qkv.shape
        E = E3 // 3
        blocked_t = None if blocked is None else blocked.t().contiguous().to(torch.uint8)
        out = torch.empty((B, Q, E), dtype=torch.float32, device=qkv.device)
        lse = torch.empty((B, num_heads, Q), dtype=torch.float32, device=qkv.device)
        pc = (ctypes.c_double * 6)(*[float(v) for v in pc_range])
        with torch.cuda.device(qkv.device):
            rc = _lib.load().racf_sasa_attention_train_forward(
                qkv.data_ptr(), tau.data_ptr(), query_ray.data_ptr(), blocked_t.data_ptr() if blocked_t is not None else None, pc,
                B, Q, num_heads, E // num_heads, query_ray.shape[2], float(drop_p), int(seed) & 0xffffffff, out.data_ptr(),
                lse.data_ptr(), wrapper._stream(qkv.device))
        _lib.check(rc, "racf_sasa_attention_train_forward")
        ctx.save_for_backward(qkv, tau, query_ray, out, lse, *([blocked_t] if blocked_t is not None else []))
        ctx.args = (tuple(float(v) for v in pc_range), num_heads, float(drop_p), int(seed) & 0xffffffff)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        qkv, tau, query_ray, out, lse, *rest = ctx.saved_tensors
        blocked_t = rest[0] if rest else None
        pc_range, num_heads, drop_p, seed = ctx.args
        B, Q, E3 = qkv.shape
        grad_out = grad_out.contiguous()
        grad_qkv, grad_tau, dsum = torch.empty_like(qkv), torch.empty_like(tau), torch.empty_like(lse)
        pc = (ctypes.c_double * 6)(*pc_range)
        with torch.cuda.device(qkv.device):
            rc = _lib.load().racf_sasa_attention_train_backward(
                qkv.data_ptr(), tau.data_ptr(), query_ray.data_ptr(), blocked_t.data_ptr() if blocked_t is not None else None, pc,
                B, Q, num_heads, E3 // 3 // num_heads, query_ray.shape[2], drop_p, seed, out.data_ptr(), lse.data_ptr(),
                grad_out.data_ptr(), dsum.data_ptr(), grad_qkv.data_ptr(), grad_tau.data_ptr(), wrapper._stream(qkv.device))
        _lib.check(rc, "racf_sasa_attention_train_backward")
        return grad_qkv, grad_tau, None, None, None, None, None, None


class MSMVPoints(Function):
    """Image-branch point generation with autograd: apply(query_ray [B,Q,10], offset [B,Q,G*Pn*D*3], ray_logit [B,Q,D],
    scale_raw [B,Q,G*T*Pn*D*L], time_diff [B,T], lidar2img [B,T*N,4,4], depth_base [D], geom) -> (loc [B*T*G,Q,P,3],
    weights [B*G*T,Q,P,L]) -- the tensors sampling_4d hands to msmv_sampling. geom = (pc_range, d_region, image_w, image_h,
    T, G, Pn, D, L). The camera view (loc[...,2]) is a discrete choice: no gradient flows through it, as in the reference,
    whose argmax / gather pick it (sparsebev_sampling.py:96-107)."""

    @staticmethod
    def forward(ctx, query_ray, offset, ray_logit, scale_raw, time_diff, lidar2img, depth_base, geom):
        pc_range, d_region, image_w, image_h, T, G, Pn, D, L = geom
        query_ray, offset, ray_logit, scale_raw = (t.contiguous() for t in (query_ray, offset, ray_logit, scale_raw))
        time_diff, lidar2img = time_diff.contiguous(), lidar2img.contiguous()
        loc, weights = points.msmv_points(query_ray, offset, ray_logit, scale_raw, time_diff, lidar2img, depth_base, pc_range,
                                          d_region, image_w, image_h, T, G, Pn, D, L)
        ctx.save_for_backward(query_ray, offset, ray_logit, time_diff, lidar2img, depth_base, loc, weights)
        ctx.geom = geom
        ctx.shapes = (offset.shape, ray_logit.shape, scale_raw.shape)
        return loc, weights

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loc, grad_weights):
        query_ray, offset, ray_logit, time_diff, lidar2img, depth_base, loc, weights = ctx.saved_tensors
        pc_range, d_region, image_w, image_h, T, G, Pn, D, L = ctx.geom
        grad_loc = torch.zeros_like(loc) if grad_loc is None else grad_loc.contiguous()
        grad_weights = torch.zeros_like(weights) if grad_weights is None else grad_weights.contiguous()
        g_ray, g_off, g_logit, g_scale = points.msmv_points_backward(
            query_ray, offset, ray_logit, time_diff, lidar2img, depth_base, pc_range, d_region, image_w, image_h, T, G, Pn, D, L,
            loc, weights, grad_loc, grad_weights, need_ray_grad=ctx.needs_input_grad[0])
        return g_ray, g_off.view(ctx.shapes[0]), g_logit.view(ctx.shapes[1]), g_scale.view(ctx.shapes[2]), None, None, None, None


class BEVPoints(Function):
    """BEV-branch point generation with autograd: apply(query_ray, offset [B,Q,M*Pn*D*2], ray_logit [B,Q,D], attn_raw
    [B,Q,M*Pn*D], time_diff, depth_base, geom) -> (loc [T*B,Q,M,1,P,2], attn [T*B,Q,M,1,P]), queue-major as
    BEVSelfAttention hands them to MSDA. geom = (pc_range, d_region, T, M, Pn, D)."""

    @staticmethod
    def forward(ctx, query_ray, offset, ray_logit, attn_raw, time_diff, depth_base, geom):
        pc_range, d_region, T, M, Pn, D = geom
        query_ray, offset, ray_logit, attn_raw = (t.contiguous() for t in (query_ray, offset, ray_logit, attn_raw))
        time_diff = time_diff.contiguous()
        loc, attn = points.bev_points(query_ray, offset, ray_logit, attn_raw, time_diff, depth_base, pc_range, d_region, T, M,
                                      Pn, D)
        ctx.save_for_backward(query_ray, offset, ray_logit, time_diff, depth_base, attn)
        ctx.geom = geom
        ctx.shapes = (offset.shape, ray_logit.shape, attn_raw.shape, loc.shape)
        return loc, attn

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_loc, grad_attn):
        query_ray, offset, ray_logit, time_diff, depth_base, attn = ctx.saved_tensors
        pc_range, d_region, T, M, Pn, D = ctx.geom
        grad_loc = (torch.zeros(ctx.shapes[3], dtype=torch.float32, device=attn.device) if grad_loc is None
                    else grad_loc.contiguous())
        grad_attn = torch.zeros_like(attn) if grad_attn is None else grad_attn.contiguous()
        g_ray, g_off, g_logit, g_raw = points.bev_points_backward(
            query_ray, offset, ray_logit, time_diff, depth_base, pc_range, d_region, T, M, Pn, D, attn, grad_loc, grad_attn,
            need_ray_grad=ctx.needs_input_grad[0])
        return g_ray, g_off.view(ctx.shapes[0]), g_logit.view(ctx.shapes[1]), g_raw.view(ctx.shapes[2]), None, None, None


class RefineBBox(Function):
    """Box refinement + velocity scaling + polar -> cartesian output transform (racformer_transformer.py:255-279) with
    autograd: apply(proposal [B,Q,code], delta [B,Q,code], time_diff [B,T], num_ray) -> (bbox_pred, bbox_xy). bbox_pred (the
    next iteration's query rays, which the reference detaches, :132) is marked non-differentiable; gradients flow from
    bbox_xy to delta and the proposal."""

    @staticmethod
    def forward(ctx, proposal, delta, time_diff, num_ray):
        proposal, delta, time_diff = proposal.contiguous(), delta.contiguous(), time_diff.contiguous()
        pred, pred_xy = points.refine_bbox(proposal, delta, time_diff, num_ray)
        ctx.save_for_backward(proposal, delta, time_diff)
        ctx.num_ray = num_ray
        ctx.mark_non_differentiable(pred)
        return pred, pred_xy

    @staticmethod
    @once_differentiable
    def backward(ctx, _grad_pred, grad_xy):
        proposal, delta, time_diff = ctx.saved_tensors
        g_delta, g_prop = points.refine_bbox_backward(proposal, delta, time_diff, ctx.num_ray, grad_xy.contiguous(),
                                                      need_proposal_grad=ctx.needs_input_grad[0])
        return g_prop, g_delta, None, None
