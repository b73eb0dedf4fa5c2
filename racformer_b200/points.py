"""Python binding of the fused sampling-point kernels (racformer_b200/csrc/points.cu, csrc/points_train.cu).
The forward functions are what the decoder harness calls when autograd is off; racformer_b200/training.py wraps them and
the *_backward functions below into autograd Functions for training."""
import ctypes

import torch

from . import _lib

_lib.load()


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check(*tensors):
    dev = tensors[0].device
    for t in tensors:
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.device == dev):
            raise RuntimeError("fused point kernels need contiguous fp32 CUDA tensors on one device")


def _pc(pc_range):
    return (ctypes.c_double * 6)(*[float(v) for v in pc_range])


def msmv_points(query_ray, offset, ray_logit, scale_raw, time_diff, lidar2img, depth_base, pc_range, d_region, image_w,
                image_h, num_frames, num_groups, num_points, depth_num, num_levels, eps=1e-5):
    """-> loc [B*T*G,Q,Pn*D,3], weights [B*G*T,Q,Pn*D,L]: the two tensors sampling_4d passes to msmv_sampling."""
    _check(query_ray, offset, ray_logit, scale_raw, time_diff, lidar2img, depth_base)
    B, Q, _ = query_ray.shape
    T, G, Pn, D, L = num_frames, num_groups, num_points, depth_num, num_levels
    N = lidar2img.shape[1] // T
    P = Pn * D
    if offset.numel() != B * Q * G * P * 3 or scale_raw.numel() != B * Q * G * T * P * L or ray_logit.numel() != B * Q * D:
        raise RuntimeError("msmv_points: inconsistent input sizes")
    loc = torch.empty((B * T * G, Q, P, 3), dtype=torch.float32, device=query_ray.device)
    weights = torch.empty((B * G * T, Q, P, L), dtype=torch.float32, device=query_ray.device)
    with torch.cuda.device(query_ray.device):
        rc = _lib.load().racf_msmv_points_forward(
            query_ray.data_ptr(), offset.data_ptr(), ray_logit.data_ptr(), scale_raw.data_ptr(), time_diff.data_ptr(),
            lidar2img.data_ptr(), depth_base.data_ptr(), _pc(pc_range), float(d_region), float(image_w), float(image_h),
            float(eps), B, Q, T, G, Pn, D, N, L, loc.data_ptr(), weights.data_ptr(), _stream(query_ray.device))
    _lib.check(rc, "racf_msmv_points_forward")
    return loc, weights


def bev_points(query_ray, offset, ray_logit, attn_raw, time_diff, depth_base, pc_range, d_region, num_frames, num_heads,
               num_points, depth_num):
    """-> loc [T*B,Q,M,1,Pn*D,2], attn [T*B,Q,M,1,Pn*D] (queue-major, as BEVSelfAttention hands them to MSDA)."""
    _check(query_ray, offset, ray_logit, attn_raw, time_diff, depth_base)
    B, Q, _ = query_ray.shape
    T, M, Pn, D = num_frames, num_heads, num_points, depth_num
    P = Pn * D
    if offset.numel() != B * Q * M * P * 2 or attn_raw.numel() != B * Q * M * P or ray_logit.numel() != B * Q * D:
        raise RuntimeError("bev_points: inconsistent input sizes")
    loc = torch.empty((T * B, Q, M, 1, P, 2), dtype=torch.float32, device=query_ray.device)
    attn = torch.empty((T * B, Q, M, 1, P), dtype=torch.float32, device=query_ray.device)
    with torch.cuda.device(query_ray.device):
        rc = _lib.load().racf_bev_points_forward(
            query_ray.data_ptr(), offset.data_ptr(), ray_logit.data_ptr(), attn_raw.data_ptr(), time_diff.data_ptr(),
            depth_base.data_ptr(), _pc(pc_range), float(d_region), B, Q, T, M, Pn, D, loc.data_ptr(), attn.data_ptr(),
            _stream(query_ray.device))
    _lib.check(rc, "racf_bev_points_forward")
    return loc, attn


def msmv_points_backward(query_ray, offset, ray_logit, time_diff, lidar2img, depth_base, pc_range, d_region, image_w, image_h,
                         num_frames, num_groups, num_points, depth_num, num_levels, loc, weights, grad_loc, grad_weights,
                         need_ray_grad=False, eps=1e-5):
    """Backward of msmv_points: (grad_ray or None, grad_offset, grad_ray_logit, grad_scale_raw [B,Q,G,T,Pn*D,L])."""
    _check(query_ray, offset, ray_logit, time_diff, lidar2img, depth_base, loc, weights, grad_loc, grad_weights)
    B, Q, _ = query_ray.shape
    T, G, Pn, D, L = num_frames, num_groups, num_points, depth_num, num_levels
    N = lidar2img.shape[1] // T
    P = Pn * D
    if loc.shape != (B * T * G, Q, P, 3) or grad_loc.shape != loc.shape or weights.shape != (B * G * T, Q, P, L) \
            or grad_weights.shape != weights.shape or offset.numel() != B * Q * G * P * 3 or ray_logit.numel() != B * Q * D:
        raise RuntimeError("msmv_points_backward: inconsistent input sizes")
    dev = query_ray.device
    g_ray = torch.empty((B, Q, 10), dtype=torch.float32, device=dev) if need_ray_grad else None
    g_off = torch.empty_like(offset)
    g_logit = torch.empty((B, Q, D), dtype=torch.float32, device=dev)
    g_scale = torch.empty((B, Q, G, T, P, L), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().racf_msmv_points_backward(
            query_ray.data_ptr(), offset.data_ptr(), ray_logit.data_ptr(), time_diff.data_ptr(), lidar2img.data_ptr(),
            depth_base.data_ptr(), _pc(pc_range), float(d_region), float(image_w), float(image_h), float(eps), B, Q, T, G, Pn, D,
            N, L, loc.data_ptr(), weights.data_ptr(), grad_loc.data_ptr(), grad_weights.data_ptr(),
            g_ray.data_ptr() if g_ray is not None else None, g_off.data_ptr(), g_logit.data_ptr(), g_scale.data_ptr(),
            _stream(dev))
    _lib.check(rc, "racf_msmv_points_backward")
    return g_ray, g_off, g_logit, g_scale


def bev_points_backward(query_ray, offset, ray_logit, time_diff, depth_base, pc_range, d_region, num_frames, num_heads,
                        num_points, depth_num, attn, grad_loc, grad_attn, need_ray_grad=False):
    """Backward of bev_points: (grad_ray or None, grad_offset, grad_ray_logit, grad_attn_raw [B,Q,M,Pn*D])."""
    _check(query_ray, offset, ray_logit, time_diff, depth_base, attn, grad_loc, grad_attn)
    B, Q, _ = query_ray.shape
    T, M, Pn, D = num_frames, num_heads, num_points, depth_num
    P = Pn * D
    if attn.shape != (T * B, Q, M, 1, P) or grad_attn.shape != attn.shape or grad_loc.shape != (T * B, Q, M, 1, P, 2) \
            or offset.numel() != B * Q * M * P * 2 or ray_logit.numel() != B * Q * D:
        raise RuntimeError("bev_points_backward: inconsistent input sizes")
    dev = query_ray.device
    g_ray = torch.empty((B, Q, 10), dtype=torch.float32, device=dev) if need_ray_grad else None
    g_off = torch.empty_like(offset)
    g_logit = torch.empty((B, Q, D), dtype=torch.float32, device=dev)
    g_attn_raw = torch.empty((B, Q, M, P), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().racf_bev_points_backward(
            query_ray.data_ptr(), offset.data_ptr(), ray_logit.data_ptr(), time_diff.data_ptr(), depth_base.data_ptr(),
            _pc(pc_range), float(d_region), B, Q, T, M, Pn, D, attn.data_ptr(), grad_loc.data_ptr(), grad_attn.data_ptr(),
            g_ray.data_ptr() if g_ray is not None else None, g_off.data_ptr(), g_logit.data_ptr(), g_attn_raw.data_ptr(),
            _stream(dev))
    _lib.check(rc, "racf_bev_points_backward")
    return g_ray, g_off, g_logit, g_attn_raw


def to_sampling_layout(feat, num_cams, num_groups=4):
    """[B, T*N, G*C, H, W] -> [B*T*G, N, H, W, C] with the tiled-transpose kernel (C == 64, fp32 CUDA, no autograd)."""
    _check(feat)
    B, TN, GC, H, W = feat.shape
    T, C = TN // num_cams, GC // num_groups
    out = torch.empty((B * T * num_groups, num_cams, H, W, C), dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        rc = _lib.load().racf_to_sampling_layout(feat.data_ptr(), out.data_ptr(), B, T, num_cams, num_groups, C, H, W,
                                                 _stream(feat.device))
    _lib.check(rc, "racf_to_sampling_layout")
    return out


def to_sampling_layout_f16(feat, num_cams, num_groups=4):
    """fp16 features -> fp32 sampling layout [B*T*G, N, H, W, C], the upcast fused into the re-layout (C == 64).
    feat: [B, T*N, G*C, H, W] halves, either contiguous or channels-last per image (memory [B*T*N, H, W, G*C], what cuDNN's
    fp16 convolutions produce). Bit-identical to `to_sampling_layout(feat.float(), ...)`."""
    if not (feat.is_cuda and feat.dtype == torch.float16 and feat.dim() == 5):
        raise RuntimeError("to_sampling_layout_f16 needs a [B, T*N, G*C, H, W] float16 CUDA tensor")
    B, TN, GC, H, W = feat.shape
    T, C = TN // num_cams, GC // num_groups
    if feat.is_contiguous():
        nhwc = 0
    elif feat.permute(0, 1, 3, 4, 2).is_contiguous():
        nhwc = 1
    else:
        feat, nhwc = feat.contiguous(), 0
    out = torch.empty((B * T * num_groups, num_cams, H, W, C), dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        rc = _lib.load().racf_to_sampling_layout_f16(feat.data_ptr(), out.data_ptr(), B, T, num_cams, num_groups, C, H, W,
                                                     nhwc, _stream(feat.device))
    _lib.check(rc, "racf_to_sampling_layout_f16")
    return out


def to_channels_last(x, concat_channels=0):
    """x [N, C, H, W] contiguous -> (x as a channels_last tensor, and -- concat_channels > 0 -- a channels_last
    [N, C + concat_channels, H, W] buffer whose first C channels are x; the rest is uninitialised): one tiled transpose
    instead of PyTorch's strided copies in front of the radar temporal encoder."""
    _check(x)
    N, C, H, W = x.shape
    dense = torch.empty((N, C, H, W), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    both = None
    if concat_channels > 0:
        both = torch.empty((N, C + concat_channels, H, W), dtype=torch.float32, device=x.device,
                           memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        rc = _lib.load().racf_chw_to_hwc(x.data_ptr(), N, C, H * W, dense.data_ptr(), C,
                                         both.data_ptr() if both is not None else None, C + concat_channels, _stream(x.device))
    _lib.check(rc, "racf_chw_to_hwc")
    return dense, both


def convgru_gates(gates, h_prev):
    """ConvGRU cell after the gates convolution (racformer_transformer.py:709-720) in one launch. gates [N, 3*Hc, H, W] and
    h_prev [N, Hc, H, W], both fp32 CUDA tensors in channels_last memory format -> h (channels_last). None when the
    tensors are not dense channel-last (the caller runs the PyTorch ops)."""
    N, C3, H, W = gates.shape
    hc = C3 // 3
    if not (gates.is_cuda and gates.dtype == torch.float32 and h_prev.dtype == torch.float32 and hc % 4 == 0
            and tuple(h_prev.shape) == (N, hc, H, W) and gates.permute(0, 2, 3, 1).is_contiguous()
            and h_prev.permute(0, 2, 3, 1).is_contiguous()):
        return None
    h = torch.empty_like(h_prev, memory_format=torch.channels_last)
    with torch.cuda.device(gates.device):
        rc = _lib.load().racf_convgru_gates_forward(gates.data_ptr(), h_prev.data_ptr(), N * H * W, hc, h.data_ptr(),
                                                    _stream(gates.device))
    if rc == -6:
        return None
    _lib.check(rc, "racf_convgru_gates_forward")
    return h


def upsample2x_bilinear(x):
    """nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True) of a channels_last fp32 CUDA tensor [N, C, H, W] ->
    [N, C, 2H, 2W] (channels_last) in one launch; None when the tensor is not dense channel-last."""
    N, C, H, W = x.shape
    if not (x.is_cuda and x.dtype == torch.float32 and C % 4 == 0 and x.permute(0, 2, 3, 1).is_contiguous()):
        return None
    out = torch.empty((N, C, 2 * H, 2 * W), dtype=torch.float32, device=x.device, memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        rc = _lib.load().racf_upsample2x_bilinear_nhwc(x.data_ptr(), N, H, W, C, out.data_ptr(), _stream(x.device))
    if rc == -6:
        return None
    _lib.check(rc, "racf_upsample2x_bilinear_nhwc")
    return out


def adaptive_mixing_core(x, params, out_points, eps=1e-5, split=False, tiled_groups=0, tensor_cores=True, variant=0):
    """x [QG, P_in, C], params [QG, C*C + out_points*P_in] -> relu(LN(S @ relu(LN(x @ M)))) [QG, out_points, C], one kernel
    (csrc/mixing.cu). With split=True the result comes back as its three bf16 pieces [3, QG, out_points, C] (their sum is
    the fp32 result exactly), or -- tiled_groups = n_groups -- as the linear.TiledOperand of the [QG / n_groups, n_groups *
    out_points * C] matrix out_proj multiplies. Returns None when the fused kernel does not exist for the shapes (callers
    use the PyTorch chain). tensor_cores: use csrc/mixing_tc.cu (tcgen05, bf16x3 operand splitting, fp32 accumulation)
    where it exists (P_in % 16 == 0) for the fp32 and the tiled result; otherwise the CUDA-core kernel csrc/mixing.cu.
    variant (tensor-core kernels only): 0 = default (warp-specialised csrc/mixing_ws.cu for P_in <= 96), 1 = phase-serial
    csrc/mixing_tc.cu, 2 = warp-specialised."""
    _check(x, params)
    QG, P_in, C = x.shape
    if params.shape != (QG, C * C + out_points * P_in):
        raise RuntimeError("adaptive_mixing_core: params must be [QG, C*C + out_points*P_in]")
    lib = _lib.load()
    if tensor_cores and (not split or tiled_groups) and C == 64 and out_points == 128 and P_in % 16 == 0 and 16 <= P_in <= 128:
        with torch.cuda.device(x.device):
            if split:
                from . import linear
                out = linear.empty_tiled(QG // tiled_groups, tiled_groups * out_points * C, x.device)
                rc = lib.racf_adaptive_mixing_tc_forward_variant(x.data_ptr(), params.data_ptr(), QG, P_in, out_points, C,
                                                                 float(eps), None, out.buf.data_ptr(), int(tiled_groups),
                                                                 int(variant), _stream(x.device))
            else:
                out = torch.empty((QG, out_points, C), dtype=torch.float32, device=x.device)
                rc = lib.racf_adaptive_mixing_tc_forward_variant(x.data_ptr(), params.data_ptr(), QG, P_in, out_points, C,
                                                                 float(eps), out.data_ptr(), None, 0, int(variant),
                                                                 _stream(x.device))
        _lib.check(rc, "racf_adaptive_mixing_tc_forward")
        return out
    with torch.cuda.device(x.device):
        if split and tiled_groups:
            from . import linear
            out = linear.empty_tiled(QG // tiled_groups, tiled_groups * out_points * C, x.device)
            rc = lib.racf_adaptive_mixing_forward_split(x.data_ptr(), params.data_ptr(), QG, P_in, out_points, C,
                                                        float(eps), out.buf.data_ptr(), int(tiled_groups), _stream(x.device))
        elif split:
            out = torch.empty((3, QG, out_points, C), dtype=torch.bfloat16, device=x.device)
            rc = lib.racf_adaptive_mixing_forward_split(x.data_ptr(), params.data_ptr(), QG, P_in, out_points, C,
                                                        float(eps), out.data_ptr(), 0, _stream(x.device))
        else:
            out = torch.empty((QG, out_points, C), dtype=torch.float32, device=x.device)
            rc = lib.racf_adaptive_mixing_forward(x.data_ptr(), params.data_ptr(), QG, P_in, out_points, C, float(eps),
                                                  out.data_ptr(), _stream(x.device))
    if rc == -6:
        return None
    _lib.check(rc, "racf_adaptive_mixing_forward")
    return out


def adaptive_mixing_core_backward(x, params, grad_out, out_points, eps=1e-5, variant=0):
    """Backward of adaptive_mixing_core: (grad_x like x, grad_params like params); None when the kernel does not exist for
    the shapes. Recomputes the forward from x and params. variant: 0 = default, 1 = CUDA cores (csrc/mixing_bwd.cu),
    2 = tensor cores (csrc/mixing_bwd_tc.cu, P_in <= 96)."""
    _check(x, params, grad_out)
    QG, P_in, C = x.shape
    if not (C == 64 and out_points == 128 and P_in % 16 == 0 and 16 <= P_in <= 128):
        return None
    grad_x, grad_params = torch.empty_like(x), torch.empty_like(params)
    with torch.cuda.device(x.device):
        rc = _lib.load().racf_adaptive_mixing_backward_variant(x.data_ptr(), params.data_ptr(), grad_out.data_ptr(), QG, P_in,
                                                               out_points, C, float(eps), grad_x.data_ptr(),
                                                               grad_params.data_ptr(), int(variant), _stream(x.device))
    _lib.check(rc, "racf_adaptive_mixing_backward")
    return grad_x, grad_params


def refine_bbox(proposal, delta, time_diff, num_ray):
    """(query rays [B,Q,code], reg branch output [B,Q,code], time_diff [B,T] or None) -> (bbox_pred, theta_d2xy(bbox_pred)):
    refine_bbox, the velocity scaling and the polar -> cartesian output transform of a decoder iteration
    (models/racformer_transformer.py:255-279) in one launch."""
    _check(proposal, delta)
    if proposal.shape != delta.shape or proposal.dim() != 3:
        raise RuntimeError("refine_bbox: proposal and delta must both be [B, Q, code]")
    B, Q, code = proposal.shape
    T = 1
    if time_diff is not None:
        _check(proposal, time_diff)
        if time_diff.dim() != 2 or time_diff.shape[0] != B:
            raise RuntimeError("refine_bbox: time_diff must be [B, T]")
        T = time_diff.shape[1]
    pred, pred_xy = torch.empty_like(proposal), torch.empty_like(proposal)
    with torch.cuda.device(proposal.device):
        rc = _lib.load().racf_refine_bbox_forward(proposal.data_ptr(), delta.data_ptr(),
                                                  time_diff.data_ptr() if time_diff is not None else None, B, Q, T, code,
                                                  float(num_ray), pred.data_ptr(), pred_xy.data_ptr(), _stream(proposal.device))
    _lib.check(rc, "racf_refine_bbox_forward")
    return pred, pred_xy


def refine_bbox_backward(proposal, delta, time_diff, num_ray, grad_pred_xy, need_proposal_grad=False):
    """Backward of refine_bbox w.r.t. its second output: (grad_delta, grad_proposal or None)."""
    _check(proposal, delta, grad_pred_xy)
    B, Q, code = proposal.shape
    T = 1 if time_diff is None else time_diff.shape[1]
    g_delta = torch.empty_like(delta)
    g_prop = torch.empty_like(proposal) if need_proposal_grad else None
    with torch.cuda.device(proposal.device):
        rc = _lib.load().racf_refine_bbox_backward(
            proposal.data_ptr(), delta.data_ptr(), time_diff.data_ptr() if time_diff is not None else None,
            grad_pred_xy.data_ptr(), B, Q, T, code, float(num_ray), g_delta.data_ptr(),
            g_prop.data_ptr() if g_prop is not None else None, _stream(proposal.device))
    _lib.check(rc, "racf_refine_bbox_backward")
    return g_delta, g_prop


def sasa_attention(qkv, tau, query_ray, pc_range, num_heads):
    """Attention core of ScaleAdaptiveSelfAttention (models/racformer_transformer.py:283-336) in one launch:
    qkv [B,Q,3E] (in_proj output), tau [B,Q,H] (gen_tau output), query_ray [B,Q,code] -> [B,Q,E] (heads merged, before
    out_proj). The [B,H,Q,Q] distance mask is never materialised."""
    _check(qkv, tau, query_ray)
    B, Q, E3 = qkv.shape
    E = E3 // 3
    if E3 != 3 * E or E % num_heads != 0 or tau.shape != (B, Q, num_heads) or query_ray.shape[:2] != (B, Q):
        raise RuntimeError("sasa_attention: inconsistent input shapes")
    out = torch.empty((B, Q, E), dtype=torch.float32, device=qkv.device)
    with torch.cuda.device(qkv.device):
        rc = _lib.load().racf_sasa_attention_forward(qkv.data_ptr(), tau.data_ptr(), query_ray.data_ptr(), _pc(pc_range),
                                                     B, Q, num_heads, E // num_heads, query_ray.shape[2], out.data_ptr(),
                                                     _stream(qkv.device))
    _lib.check(rc, "racf_sasa_attention_forward")
    return out
