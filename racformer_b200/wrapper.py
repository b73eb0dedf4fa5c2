"""Operator API of the MSMV sampling op -- host-side mirror of the reference's models/csrc/wrapper.py.

Same names, argument order and return structure as the reference (wrapper.py:78-156), so that
`from .csrc.wrapper import msmv_sampling, msmv_sampling_v2, MSMV_CUDA` in the unchanged call sites
(models/sparsebev_sampling.py:5, models/racformer_transformer.py:13) can be pointed here.

Differences, all deliberate:
  * the work is done by libracformer_ops.so (hand-written sm_100a kernels behind a C ABI) instead of the
    torch extension `_msmv_sampling_cuda`; the six extension callables are kept as functions of this module;
  * there is NO PyTorch/CPU fallback: `MSMV_CUDA` is always True and a missing library is an ImportError;
  * any number of levels 1..8 runs on the GPU (the reference only has 2/4/5 and falls back otherwise);
  * kernels run on the current PyTorch stream of the tensors' device (the reference uses the legacy stream);
  * shape/dtype mistakes raise RuntimeError instead of reading out of bounds.
"""
import ctypes

import torch

from . import _lib

_lib.load()       # fail loudly at import time if the CUDA library has not been built
MSMV_CUDA = True  # reference: wrapper.py:4-12; selects the channel-last layout at racformer_transformer.py:117-119

MAX_POINT = 128   # msmv_sampling_forward.cu:21, checked at msmv_sampling.cpp:159


def _require(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _check_inputs(feats, loc, w, grad_output=None):
    """The reference's AT_ASSERTMs (msmv_sampling.cpp:140-159, 299-321) plus the shape checks it lacks."""
    _require(len(feats) >= 1, "at least one feature level is required")
    for f in feats:
        _require(f.is_contiguous(), "value tensor has to be contiguous")
    _require(loc.is_contiguous(), "sampling_loc tensor has to be contiguous")
    _require(w.is_contiguous(), "attn_weight tensor has to be contiguous")
    for f in feats:
        _require(f.is_cuda, "value must be a CUDA tensor")
    _require(loc.is_cuda, "sampling_loc must be a CUDA tensor")
    _require(w.is_cuda, "attn_weight must be a CUDA tensor")
    if grad_output is not None:
        _require(grad_output.is_contiguous(), "grad_output tensor has to be contiguous")
        _require(grad_output.is_cuda, "grad_output must be a CUDA tensor")
    for t in list(feats) + [loc, w] + ([grad_output] if grad_output is not None else []):
        _require(t.dtype == torch.float32, f"expected scalar type Float but found {t.dtype}")
        _require(t.device == feats[0].device, "all tensors must be on the same device")
    f0 = feats[0]
    _require(f0.dim() == 5, "value must be [B, N, H, W, C]")
    B, N, _, _, C = f0.shape
    for f in feats:
        _require(f.dim() == 5 and f.shape[0] == B and f.shape[1] == N and f.shape[4] == C,
                 "all feature levels must share B, N and C")
    _require(loc.dim() == 4 and loc.shape[0] == B and loc.shape[3] == 3, "sampling_loc must be [B, Q, P, 3]")
    Q, P = loc.shape[1], loc.shape[2]
    _require(P <= MAX_POINT, "num_point exceed limits")
    _require(tuple(w.shape) == (B, Q, P, len(feats)), "attn_weight must be [B, Q, P, num_levels]")
    if grad_output is not None:
        _require(tuple(grad_output.shape) == (B, Q, C, P), "grad_output must be [B, Q, C, P]")
    return B, N, C, Q, P


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _hw_array(feats):
    flat = []
    for f in feats:
        flat += [f.shape[2], f.shape[3]]
    return (ctypes.c_int * len(flat))(*flat)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def msmv_forward(feats, sampling_loc, attn_weight):
    """feats: list of [B,N,H_l,W_l,C]; sampling_loc [B,Q,P,3]; attn_weight [B,Q,P,L] -> [B,Q,C,P]."""
    feats = list(feats)
    B, N, C, Q, P = _check_inputs(feats, sampling_loc, attn_weight)
    out = torch.empty((B, Q, C, P), dtype=torch.float32, device=feats[0].device)  # kernel writes every element
    if out.numel() == 0:
        return out
    with torch.cuda.device(feats[0].device):
        rc = _lib.load().racf_msmv_forward(
            _ptr_array(feats), _hw_array(feats), len(feats), sampling_loc.data_ptr(), attn_weight.data_ptr(),
            B, C, N, Q, P, out.data_ptr(), _stream(feats[0].device))
    _lib.check(rc, "racf_msmv_forward")
    return out


def msmv_forward_grouped(feats, sampling_loc, attn_weight, num_frames, num_groups):
    """Forward with sampling_4d's un-packing fused in: -> [B, Q, G, T*P, C] (B' = B*T*G). Inference helper, no autograd.
    Falls back to msmv_forward + permute when the fused variant does not exist for the shapes."""
    feats = list(feats)
    Bp, N, C, Q, P = _check_inputs(feats, sampling_loc, attn_weight)
    T, G = int(num_frames), int(num_groups)
    _require(Bp % (T * G) == 0, "batch must be B * num_frames * num_groups")
    B = Bp // (T * G)
    out = torch.empty((B, Q, G, T * P, C), dtype=torch.float32, device=feats[0].device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(feats[0].device):
        rc = _lib.load().racf_msmv_forward_grouped(
            _ptr_array(feats), _hw_array(feats), len(feats), sampling_loc.data_ptr(), attn_weight.data_ptr(),
            Bp, C, N, Q, P, T, G, out.data_ptr(), _stream(feats[0].device))
    if rc == -6:   # RACF_ERR_UNSUPPORTED: generic shapes
        plain = msmv_forward(feats, sampling_loc, attn_weight)
        return plain.reshape(B, T, G, Q, C, P).permute(0, 3, 2, 1, 5, 4).flatten(3, 4).contiguous()
    _lib.check(rc, "racf_msmv_forward_grouped")
    return out


def msmv_backward(grad_output, feats, sampling_loc, attn_weight):
    """-> [grad_feat_0, ..., grad_feat_{L-1}, grad_sampling_loc, grad_attn_weight] (msmv_sampling.cpp:356-358)."""
    feats = list(feats)
    B, N, C, Q, P = _check_inputs(feats, sampling_loc, attn_weight, grad_output)
    grad_feats = [torch.empty_like(f) for f in feats]            # zero-filled by the library on our stream
    grad_loc = torch.empty_like(sampling_loc)                    # fully written by the kernel (z column = 0)
    grad_w = torch.empty_like(attn_weight)
    if grad_output.numel() == 0:
        return [g.zero_() for g in grad_feats] + [grad_loc, grad_w]
    with torch.cuda.device(feats[0].device):
        rc = _lib.load().racf_msmv_backward(
            grad_output.data_ptr(), _ptr_array(feats), _hw_array(feats), len(feats), sampling_loc.data_ptr(),
            attn_weight.data_ptr(), B, C, N, Q, P, _ptr_array(grad_feats), grad_loc.data_ptr(), grad_w.data_ptr(),
            1, _stream(feats[0].device))
    _lib.check(rc, "racf_msmv_backward")
    return grad_feats + [grad_loc, grad_w]


def msmv_tap_masks(feat_hw, sampling_loc, num_views):
    """Debug entry: (view_index int32 [B,Q,P], tap_mask uint8 [B,Q,P,L]); bit0 = tap in range, bit1..4 = corners read."""
    _require(sampling_loc.is_cuda and sampling_loc.is_contiguous() and sampling_loc.dtype == torch.float32,
             "sampling_loc must be a contiguous CUDA float tensor")
    B, Q, P, _ = sampling_loc.shape
    L = len(feat_hw)
    view = torch.empty((B, Q, P), dtype=torch.int32, device=sampling_loc.device)
    mask = torch.empty((B, Q, P, L), dtype=torch.uint8, device=sampling_loc.device)
    flat = [int(v) for hw in feat_hw for v in hw]
    with torch.cuda.device(sampling_loc.device):
        rc = _lib.load().racf_msmv_tap_masks((ctypes.c_int * len(flat))(*flat), L, sampling_loc.data_ptr(), B,
                                             int(num_views), Q, P, view.data_ptr(), mask.data_ptr(),
                                             _stream(sampling_loc.device))
    _lib.check(rc, "racf_msmv_tap_masks")
    return view, mask


# --- the six callables of the reference extension module (msmv_sampling.cpp:498-507) -----------------------
def _ms_deform_attn_cuda_c45_forward(feat_c4, feat_c5, sampling_loc, attn_weight):
    return msmv_forward([feat_c4, feat_c5], sampling_loc, attn_weight)


def _ms_deform_attn_cuda_c45_backward(grad_output, feat_c4, feat_c5, sampling_loc, attn_weight):
    return msmv_backward(grad_output, [feat_c4, feat_c5], sampling_loc, attn_weight)


def _ms_deform_attn_cuda_c2345_forward(feat_c2, feat_c3, feat_c4, feat_c5, sampling_loc, attn_weight):
    return msmv_forward([feat_c2, feat_c3, feat_c4, feat_c5], sampling_loc, attn_weight)


def _ms_deform_attn_cuda_c2345_backward(grad_output, feat_c2, feat_c3, feat_c4, feat_c5, sampling_loc, attn_weight):
    return msmv_backward(grad_output, [feat_c2, feat_c3, feat_c4, feat_c5], sampling_loc, attn_weight)


def _ms_deform_attn_cuda_c23456_forward(feat_c2, feat_c3, feat_c4, feat_c5, feat_c6, sampling_loc, attn_weight):
    return msmv_forward([feat_c2, feat_c3, feat_c4, feat_c5, feat_c6], sampling_loc, attn_weight)


def _ms_deform_attn_cuda_c23456_backward(grad_output, feat_c2, feat_c3, feat_c4, feat_c5, feat_c6, sampling_loc,
                                         attn_weight):
    return msmv_backward(grad_output, [feat_c2, feat_c3, feat_c4, feat_c5, feat_c6], sampling_loc, attn_weight)


# --- autograd Functions (wrapper.py:78-142). Inputs are kept with save_for_backward only, so the reference's
# non-re-entrant activation checkpoint (models/checkpoint.py:381-443) can drop and recompute them. -----------
class MSMVSampling(torch.autograd.Function):
    """Any number of levels: apply(*feats, sampling_locations, scale_weights)."""

    @staticmethod
    def forward(ctx, *args):
        feats, sampling_locations, scale_weights = args[:-2], args[-2], args[-1]
        ctx.save_for_backward(*feats, sampling_locations, scale_weights)
        return msmv_forward(feats, sampling_locations, scale_weights)

    @staticmethod
    def backward(ctx, grad_output):
        saved = ctx.saved_tensors
        feats, sampling_locations, scale_weights = saved[:-2], saved[-2], saved[-1]
        return tuple(msmv_backward(grad_output.contiguous(), feats, sampling_locations, scale_weights))


class MSMVSamplingC45(MSMVSampling):
    @staticmethod
    def forward(ctx, feat_c4, feat_c5, sampling_locations, scale_weights):
        return MSMVSampling.forward(ctx, feat_c4, feat_c5, sampling_locations, scale_weights)


class MSMVSamplingC2345(MSMVSampling):
    @staticmethod
    def forward(ctx, feat_c2, feat_c3, feat_c4, feat_c5, sampling_locations, scale_weights):
        return MSMVSampling.forward(ctx, feat_c2, feat_c3, feat_c4, feat_c5, sampling_locations, scale_weights)


class MSMVSamplingC23456(MSMVSampling):
    @staticmethod
    def forward(ctx, feat_c2, feat_c3, feat_c4, feat_c5, feat_c6, sampling_locations, scale_weights):
        return MSMVSampling.forward(ctx, feat_c2, feat_c3, feat_c4, feat_c5, feat_c6, sampling_locations,
                                    scale_weights)


def msmv_sampling(mlvl_feats, sampling_locations, scale_weights):
    """wrapper.py:145-153. mlvl_feats: list of [B',N,H,W,C] (channel-last); -> [B',Q,C,P]."""
    n = len(mlvl_feats)
    if n == 2:
        return MSMVSamplingC45.apply(*mlvl_feats, sampling_locations, scale_weights)
    if n == 4:
        return MSMVSamplingC2345.apply(*mlvl_feats, sampling_locations, scale_weights)
    if n == 5:
        return MSMVSamplingC23456.apply(*mlvl_feats, sampling_locations, scale_weights)
    return MSMVSampling.apply(*mlvl_feats, sampling_locations, scale_weights)


def msmv_sampling_v2(mlvl_feats, sampling_locations, scale_weights):
    """wrapper.py:41-76,155-156: sample only the level with the largest scale weight, unweighted.

    Selecting one level is the weighted sum with a one-hot weight vector, so this runs on the same kernel.
    The feature tensors use this module's channel-last layout. No gradient reaches scale_weights (argmax).
    """
    idx = torch.argmax(scale_weights, dim=-1, keepdim=True)
    one_hot = torch.zeros_like(scale_weights).scatter_(-1, idx, 1.0)
    return msmv_sampling(mlvl_feats, sampling_locations, one_hot)
