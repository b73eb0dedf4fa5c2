"""Operator API of multi-scale deformable attention -- host-side mirror of the reference's
models/multi_scale_deformable_attn_function.py.

The reference binds mmcv-full 1.6.0's `_ext` (`ext_loader.load_ext('_ext', ['ms_deform_attn_backward',
'ms_deform_attn_forward'])`, :10-12). Here `ext_module` is an object with the same two callables (same
positional order, `im2col_step` keyword) backed by libracformer_ops.so, and the two autograd Functions keep
the reference signatures (:90-163) so models/bev_self_attention.py:199-201 runs unchanged.
No PyTorch/CPU fallback exists in this module.
"""
import ctypes

import torch
from torch.autograd.function import Function, once_differentiable

from . import _lib

_lib.load()


def _require(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    # mmcv: AT_ASSERTM(...is_contiguous()) / is_cuda() for each tensor, then the im2col_step divisibility assert
    _require(value.is_contiguous(), "value tensor has to be contiguous")
    _require(spatial_shapes.is_contiguous(), "spatial_shapes tensor has to be contiguous")
    _require(level_start_index.is_contiguous(), "level_start_index tensor has to be contiguous")
    _require(sampling_loc.is_contiguous(), "sampling_loc tensor has to be contiguous")
    _require(attn_weight.is_contiguous(), "attn_weight tensor has to be contiguous")
    _require(value.is_cuda, "value must be a CUDA tensor")
    _require(spatial_shapes.is_cuda, "spatial_shapes must be a CUDA tensor")
    _require(level_start_index.is_cuda, "level_start_index must be a CUDA tensor")
    _require(sampling_loc.is_cuda, "sampling_loc must be a CUDA tensor")
    _require(attn_weight.is_cuda, "attn_weight must be a CUDA tensor")
    for t in (value, sampling_loc, attn_weight):
        _require(t.dtype == torch.float32, f"expected scalar type Float but found {t.dtype}")
    _require(spatial_shapes.dtype == torch.int64 and level_start_index.dtype == torch.int64,
             "spatial_shapes and level_start_index must be int64 tensors")
    _require(value.dim() == 4, "value must be [bs, num_keys, num_heads, head_dim]")
    B, S, M, D = value.shape
    _require(spatial_shapes.dim() == 2 and spatial_shapes.shape[1] == 2, "spatial_shapes must be [num_levels, 2]")
    L = spatial_shapes.shape[0]
    _require(level_start_index.numel() == L, "level_start_index must be [num_levels]")
    _require(sampling_loc.dim() == 6 and sampling_loc.shape[0] == B and sampling_loc.shape[2] == M
             and sampling_loc.shape[3] == L and sampling_loc.shape[5] == 2,
             "sampling_loc must be [bs, num_queries, num_heads, num_levels, num_points, 2]")
    Q, P = sampling_loc.shape[1], sampling_loc.shape[4]
    _require(tuple(attn_weight.shape) == (B, Q, M, L, P),
             "attn_weight must be [bs, num_queries, num_heads, num_levels, num_points]")
    step = min(B, int(im2col_step))
    _require(step > 0 and B % step == 0, f"batch({B}) must divide im2col_step({step})")
    _check_level_metadata(spatial_shapes, level_start_index, S)
    return B, S, M, D, L, Q, P


_VALIDATED_METADATA = {}


def _check_level_metadata(spatial_shapes, level_start_index, num_keys):
    """Every level must lie inside the value map: level_start_index[l] + H_l * W_l <= num_keys. The kernels compute pixel
    offsets from this device-side metadata, so inconsistent values would read (forward) or `red.add` (backward) out of
    bounds. The check costs one device -> host copy, so it runs once per metadata tensor pair (keyed on pointer and
    version counter; the decoder's shapes are cached constants) and is skipped while a CUDA graph is being captured."""
    key = (spatial_shapes.data_ptr(), spatial_shapes._version, level_start_index.data_ptr(), level_start_index._version,
           int(num_keys), str(spatial_shapes.device))
    if key in _VALIDATED_METADATA or torch.cuda.is_current_stream_capturing():
        return
    hw = spatial_shapes.detach().cpu().tolist()
    lsi = level_start_index.detach().cpu().tolist()
    for (h, w), start in zip(hw, lsi):
        _require(h > 0 and w > 0 and start >= 0 and start + h * w <= num_keys,
                 f"level metadata outside the value map: start {start} + {h} x {w} > num_keys {num_keys}")
    if len(_VALIDATED_METADATA) > 256:
        _VALIDATED_METADATA.clear()
    _VALIDATED_METADATA[key] = True


class _ExtModule:
    """Stand-in for mmcv's `_ext` on this path: ms_deform_attn_forward / ms_deform_attn_backward."""

    @staticmethod
    def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step=64):
        B, S, M, D, L, Q, P = _check(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)
        out = torch.empty((B, Q, M * D), dtype=torch.float32, device=value.device)
        if out.numel() == 0:
            return out
        with torch.cuda.device(value.device):
            rc = _lib.load().racf_msda_forward(
                value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
                attn_weight.data_ptr(), B, S, M, D, L, Q, P, int(im2col_step), out.data_ptr(), _stream(value.device))
        _lib.check(rc, "racf_msda_forward")
        return out

    @staticmethod
    def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                                grad_value, grad_sampling_loc, grad_attn_weight, im2col_step=64):
        B, S, M, D, L, Q, P = _check(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step)
        _require(grad_output.is_contiguous(), "grad_output tensor has to be contiguous")
        _require(grad_output.is_cuda, "grad_output must be a CUDA tensor")
        _require(grad_output.dtype == torch.float32 and grad_output.numel() == B * Q * M * D,
                 "grad_output must be a float tensor [bs, num_queries, embed_dims]")
        for g, ref, name in ((grad_value, value, "grad_value"), (grad_sampling_loc, sampling_loc, "grad_sampling_loc"),
                             (grad_attn_weight, attn_weight, "grad_attn_weight")):
            _require(g.is_cuda and g.is_contiguous() and g.dtype == torch.float32 and g.shape == ref.shape,
                     f"{name} must be a contiguous CUDA float tensor shaped like its input")
        if grad_output.numel() == 0:
            return
        with torch.cuda.device(value.device):
            rc = _lib.load().racf_msda_backward(
                value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
                attn_weight.data_ptr(), grad_output.data_ptr(), B, S, M, D, L, Q, P, int(im2col_step),
                grad_value.data_ptr(), grad_sampling_loc.data_ptr(), grad_attn_weight.data_ptr(),
                _stream(value.device))
        _lib.check(rc, "racf_msda_backward")


ext_module = _ExtModule()


def msda_forward_pair(value_a, loc_a, attn_a, value_b, loc_b, attn_b, spatial_shapes, level_start_index, im2col_step=64):
    """Two forward problems of identical geometry -- the radar and the LSS BEV branch of a decoder iteration
    (models/racformer_transformer.py:229-236) -- in one launch; inference only (no autograd). Returns (out_a, out_b),
    bit-identical to two ext_module.ms_deform_attn_forward calls."""
    B, S, M, D, L, Q, P = _check(value_a, spatial_shapes, level_start_index, loc_a, attn_a, im2col_step)
    _require(_check(value_b, spatial_shapes, level_start_index, loc_b, attn_b, im2col_step) == (B, S, M, D, L, Q, P)
             and value_b.device == value_a.device, "msda_forward_pair: the two problems must have the same shapes and device")
    out_a = torch.empty((B, Q, M * D), dtype=torch.float32, device=value_a.device)
    out_b = torch.empty_like(out_a)
    if out_a.numel() == 0:
        return out_a, out_b
    with torch.cuda.device(value_a.device):
        rc = _lib.load().racf_msda_forward_pair(
            value_a.data_ptr(), loc_a.data_ptr(), attn_a.data_ptr(), out_a.data_ptr(),
            value_b.data_ptr(), loc_b.data_ptr(), attn_b.data_ptr(), out_b.data_ptr(),
            spatial_shapes.data_ptr(), level_start_index.data_ptr(), B, S, M, D, L, Q, P, int(im2col_step),
            _stream(value_a.device))
    _lib.check(rc, "racf_msda_forward_pair")
    return out_a, out_b


def msda_tap_masks(spatial_shapes, sampling_loc):
    """Debug entry: uint8 [bs, Q, M, L, P]; bit0 = tap in range, bit1..4 = corners read."""
    _require(sampling_loc.is_cuda and sampling_loc.is_contiguous() and sampling_loc.dtype == torch.float32,
             "sampling_loc must be a contiguous CUDA float tensor")
    _require(sampling_loc.dim() == 6 and sampling_loc.shape[5] == 2, "sampling_loc must be [bs, Q, M, L, P, 2]")
    B, Q, M, L, P, _ = sampling_loc.shape
    _require(spatial_shapes.is_cuda and spatial_shapes.device == sampling_loc.device and spatial_shapes.dtype == torch.int64
             and spatial_shapes.is_contiguous() and tuple(spatial_shapes.shape) == (L, 2),
             "spatial_shapes must be a contiguous int64 [num_levels, 2] tensor on sampling_loc's device")
    mask = torch.empty((B, Q, M, L, P), dtype=torch.uint8, device=sampling_loc.device)
    with torch.cuda.device(sampling_loc.device):
        rc = _lib.load().racf_msda_tap_masks(spatial_shapes.data_ptr(), sampling_loc.data_ptr(), B, M, L, Q, P,
                                             mask.data_ptr(), _stream(sampling_loc.device))
    _lib.check(rc, "racf_msda_tap_masks")
    return mask


class MultiScaleDeformableAttnFunction_fp32(Function):
    """multi_scale_deformable_attn_function.py:90-163."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        ctx.im2col_step = im2col_step
        output = ext_module.ms_deform_attn_forward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
            im2col_step=ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights = \
            ctx.saved_tensors
        grad_value = torch.zeros_like(value)
        grad_sampling_loc = torch.empty_like(sampling_locations)   # fully overwritten by the kernel
        grad_attn_weight = torch.empty_like(attention_weights)     # fully overwritten by the kernel
        ext_module.ms_deform_attn_backward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
            grad_output.contiguous(), grad_value, grad_sampling_loc, grad_attn_weight, im2col_step=ctx.im2col_step)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None


class MultiScaleDeformableAttnFunction_fp16(Function):
    """multi_scale_deformable_attn_function.py:15-87. The reference casts to fp16 and runs mmcv's half kernel;
    RaCFormer never selects it (models/bev_self_attention.py:194-198 picks _fp32 on both branches). Here the
    arithmetic stays fp32 (the only precision the sm_100a kernels implement) and only the I/O dtype is fp16."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        ctx.im2col_step = im2col_step
        v, loc, aw = value.float().contiguous(), sampling_locations.float().contiguous(), \
            attention_weights.float().contiguous()
        output = ext_module.ms_deform_attn_forward(v, value_spatial_shapes, value_level_start_index, loc, aw,
                                                   im2col_step=im2col_step)
        ctx.save_for_backward(v, value_spatial_shapes, value_level_start_index, loc, aw)
        ctx.io_dtypes = (value.dtype, sampling_locations.dtype, attention_weights.dtype)
        return output.half()

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        v, shapes, lsi, loc, aw = ctx.saved_tensors
        grad_value = torch.zeros_like(v)
        grad_loc = torch.empty_like(loc)
        grad_aw = torch.empty_like(aw)
        ext_module.ms_deform_attn_backward(v, shapes, lsi, loc, aw, grad_output.float().contiguous(), grad_value,
                                           grad_loc, grad_aw, im2col_step=ctx.im2col_step)
        dv, dl, da = ctx.io_dtypes
        return grad_value.to(dv), None, None, grad_loc.to(dl), grad_aw.to(da), None
