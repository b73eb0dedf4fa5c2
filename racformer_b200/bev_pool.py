"""BEVPoolv2 operator API -- host-side mirror of the reference's models/csrc/bev_pool_v2/bev_pool.py.

Same names and argument order: `bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape,
interval_starts, interval_lengths)`, `QuickCumsumCuda`, and the two extension callables `bev_pool_v2_forward` /
`bev_pool_v2_backward` (src/bev_pool.cpp:30-111) backed by libracformer_ops.so. No CPU fallback.
"""
import ctypes

import torch

from . import _lib

_lib.load()


def _require(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check(floats, ints):
    dev = floats[0].device
    for t in floats:
        _require(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.device == dev,
                 "bev_pool_v2 needs contiguous float32 CUDA tensors")
    for t in ints:
        _require(t.is_cuda and t.dtype == torch.int32 and t.is_contiguous() and t.device == dev,
                 "bev_pool_v2 needs contiguous int32 CUDA index tensors")


def bev_pool_v2_forward(depth, feat, out, ranks_depth, ranks_feat, ranks_bev, interval_lengths, interval_starts):
    """Argument order of the reference binding (bev_pool.cpp:30-39): note lengths before starts. Fills `out` in place."""
    _check([depth, feat, out], [ranks_depth, ranks_feat, ranks_bev, interval_lengths, interval_starts])
    _require(feat.dim() == 5 and out.shape[-1] == feat.shape[-1], "feat must be [b,n,h,w,c], out [b,z,y,x,c]")
    with torch.cuda.device(depth.device):
        rc = _lib.load().racf_bev_pool_v2_forward(
            depth.data_ptr(), feat.data_ptr(), ranks_depth.data_ptr(), ranks_feat.data_ptr(), ranks_bev.data_ptr(),
            interval_starts.data_ptr(), interval_lengths.data_ptr(), interval_lengths.numel(), feat.shape[-1],
            out.data_ptr(), _stream(depth.device))
    _lib.check(rc, "racf_bev_pool_v2_forward")


def bev_pool_v2_backward(out_grad, depth_grad, feat_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev,
                         interval_lengths, interval_starts):
    """bev_pool.cpp:76-104. Intervals are runs of equal ranks_feat. Fills depth_grad / feat_grad in place."""
    _check([out_grad, depth_grad, feat_grad, depth, feat],
           [ranks_depth, ranks_feat, ranks_bev, interval_lengths, interval_starts])
    with torch.cuda.device(depth.device):
        rc = _lib.load().racf_bev_pool_v2_backward(
            out_grad.data_ptr(), depth.data_ptr(), feat.data_ptr(), ranks_depth.data_ptr(), ranks_feat.data_ptr(),
            ranks_bev.data_ptr(), interval_starts.data_ptr(), interval_lengths.data_ptr(), interval_lengths.numel(),
            out_grad.shape[-1], depth_grad.data_ptr(), feat_grad.data_ptr(), _stream(depth.device))
    _lib.check(rc, "racf_bev_pool_v2_backward")


def backward_intervals(ranks_depth, ranks_feat, ranks_bev, stable=False):
    """bev_pool.py:50-60: re-sort the points by ranks_feat and cut them into runs of equal ranks_feat.
    `stable=True` fixes the order inside a run (the reference's argsort is not stable, so its fp32 sums are only
    reproducible up to reordering)."""
    order = torch.argsort(ranks_feat.long(), stable=True) if stable else ranks_feat.argsort()
    ranks_feat, ranks_depth, ranks_bev = ranks_feat[order], ranks_depth[order], ranks_bev[order]
    kept = torch.ones(ranks_bev.shape[0], device=ranks_bev.device, dtype=torch.bool)
    kept[1:] = ranks_feat[1:] != ranks_feat[:-1]
    starts = torch.where(kept)[0].int()
    lengths = torch.zeros_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = ranks_bev.shape[0] - starts[-1]
    return (ranks_depth.contiguous(), ranks_feat.contiguous(), ranks_bev.contiguous(), starts.contiguous(),
            lengths.contiguous())


class QuickCumsumCuda(torch.autograd.Function):
    """bev_pool.py:11-92."""

    @staticmethod
    def forward(ctx, depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts, interval_lengths):
        ranks_bev = ranks_bev.int()
        depth = depth.contiguous().float()
        feat = feat.contiguous().float()
        ranks_depth = ranks_depth.contiguous().int()
        ranks_feat = ranks_feat.contiguous().int()
        interval_lengths = interval_lengths.contiguous().int()
        interval_starts = interval_starts.contiguous().int()
        out = feat.new_zeros(bev_feat_shape)
        bev_pool_v2_forward(depth, feat, out, ranks_depth, ranks_feat, ranks_bev, interval_lengths, interval_starts)
        ctx.save_for_backward(ranks_bev, depth, feat, ranks_feat, ranks_depth)
        return out

    @staticmethod
    def backward(ctx, out_grad):
        ranks_bev, depth, feat, ranks_feat, ranks_depth = ctx.saved_tensors
        ranks_depth, ranks_feat, ranks_bev, starts, lengths = backward_intervals(ranks_depth, ranks_feat, ranks_bev)
        depth_grad = depth.new_zeros(depth.shape)
        feat_grad = feat.new_zeros(feat.shape)
        bev_pool_v2_backward(out_grad.contiguous(), depth_grad, feat_grad, depth, feat, ranks_depth, ranks_feat,
                             ranks_bev, lengths, starts)
        return depth_grad, feat_grad, None, None, None, None, None, None


def bev_pool_v2(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts, interval_lengths):
    """bev_pool.py:95-101: -> [B, C, Z, Y, X]."""
    x = QuickCumsumCuda.apply(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts,
                              interval_lengths)
    return x.permute(0, 4, 1, 2, 3).contiguous()


def forward_intervals(ranks_bev):
    """Runs of equal ranks_bev of points already sorted by it (view_transformer_racformer.py voxel_pooling_prepare_v2)."""
    kept = torch.ones(ranks_bev.shape[0], device=ranks_bev.device, dtype=torch.bool)
    kept[1:] = ranks_bev[1:] != ranks_bev[:-1]
    starts = torch.where(kept)[0].int()
    lengths = torch.zeros_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = ranks_bev.shape[0] - starts[-1]
    return starts, lengths
