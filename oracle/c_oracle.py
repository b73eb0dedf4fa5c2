"""ctypes front-end of oracle/libracf_oracle.so (TEST INFRASTRUCTURE; see racf_oracle.c for what it restates).

All functions take / return CPU torch tensors (fp32, made contiguous here) so tests can feed the same seeded
inputs to the oracle and to the CUDA path.
"""
import ctypes

import torch

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(_build.build())
        for name in ("racf_oracle_msmv_forward", "racf_oracle_msmv_backward", "racf_oracle_msda_forward",
                     "racf_oracle_msda_backward", "racf_oracle_num_threads"):
            getattr(lib, name).restype = ctypes.c_int
        _lib = lib
    return _lib


def num_threads():
    return _load().racf_oracle_num_threads()


def _f32(t):
    return t.detach().to("cpu", torch.float32).contiguous()


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _ptrs(ts):
    return (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def _hw(feats):
    flat = [int(v) for f in feats for v in (f.shape[2], f.shape[3])]
    return (ctypes.c_int * len(flat))(*flat)


def msmv_forward(feats, loc, w, with_masks=False):
    """feats: list of [B,N,H,W,C] (channel-last). -> out [B,Q,C,P] (and view int32 [B,Q,P], mask uint8 [B,Q,P,L])."""
    feats = [_f32(f) for f in feats]
    loc, w = _f32(loc), _f32(w)
    B, N, _, _, C = feats[0].shape
    _, Q, P, _ = loc.shape
    L = len(feats)
    out = torch.empty(B, Q, C, P)
    view = torch.empty(B, Q, P, dtype=torch.int32)
    mask = torch.empty(B, Q, P, L, dtype=torch.uint8)
    rc = _load().racf_oracle_msmv_forward(_ptrs(feats), _hw(feats), L, _p(loc), _p(w), B, C, N, Q, P, _p(out),
                                          _p(view), _p(mask))
    assert rc == 0, rc
    return (out, view, mask) if with_masks else out


def msmv_backward(grad_out, feats, loc, w):
    """-> (list of grad_feats, grad_loc [B,Q,P,3], grad_w [B,Q,P,L]); fp64 accumulation."""
    feats = [_f32(f) for f in feats]
    loc, w, grad_out = _f32(loc), _f32(w), _f32(grad_out)
    B, N, _, _, C = feats[0].shape
    _, Q, P, _ = loc.shape
    L = len(feats)
    gfeats = [torch.empty_like(f) for f in feats]
    gloc = torch.empty_like(loc)
    gw = torch.empty_like(w)
    rc = _load().racf_oracle_msmv_backward(_p(grad_out), _ptrs(feats), _hw(feats), L, _p(loc), _p(w), B, C, N, Q, P,
                                           _ptrs(gfeats), _p(gloc), _p(gw))
    assert rc == 0, rc
    return gfeats, gloc, gw


def msda_forward(value, spatial_shapes, level_start_index, loc, aw, with_masks=False):
    """value [B,S,M,D]; loc [B,Q,M,L,P,2]; aw [B,Q,M,L,P] -> out [B,Q,M*D] (and mask uint8 [B,Q,M,L,P])."""
    value, loc, aw = _f32(value), _f32(loc), _f32(aw)
    shapes = spatial_shapes.detach().to("cpu", torch.int64).contiguous()
    lsi = level_start_index.detach().to("cpu", torch.int64).contiguous()
    B, S, M, D = value.shape
    _, Q, _, L, P, _ = loc.shape
    out = torch.empty(B, Q, M * D)
    mask = torch.empty(B, Q, M, L, P, dtype=torch.uint8)
    rc = _load().racf_oracle_msda_forward(_p(value), _p(shapes), _p(lsi), _p(loc), _p(aw), B, S, M, D, L, Q, P,
                                          _p(out), _p(mask))
    assert rc == 0, rc
    return (out, mask) if with_masks else out


def msda_backward(value, spatial_shapes, level_start_index, loc, aw, grad_out):
    """-> (grad_value, grad_loc, grad_aw); fp64 accumulation."""
    value, loc, aw, grad_out = _f32(value), _f32(loc), _f32(aw), _f32(grad_out)
    shapes = spatial_shapes.detach().to("cpu", torch.int64).contiguous()
    lsi = level_start_index.detach().to("cpu", torch.int64).contiguous()
    B, S, M, D = value.shape
    _, Q, _, L, P, _ = loc.shape
    gv, gl, ga = torch.empty_like(value), torch.empty_like(loc), torch.empty_like(aw)
    rc = _load().racf_oracle_msda_backward(_p(value), _p(shapes), _p(lsi), _p(loc), _p(aw), _p(grad_out), B, S, M,
                                           D, L, Q, P, _p(gv), _p(gl), _p(ga))
    assert rc == 0, rc
    return gv, gl, ga


def _i32(t):
    return t.detach().to("cpu", torch.int32).contiguous()


def bev_pool_v2_forward(depth, feat, ranks_depth, ranks_feat, ranks_bev, bev_feat_shape, interval_starts,
                        interval_lengths):
    """-> out [B,Z,Y,X,C] (zero where no point falls); reference order of summation."""
    depth, feat = _f32(depth), _f32(feat)
    rd, rf, rb, st, ln = (_i32(t) for t in (ranks_depth, ranks_feat, ranks_bev, interval_starts, interval_lengths))
    out = torch.zeros(tuple(bev_feat_shape))
    lib = _load()
    lib.racf_oracle_bev_pool_v2_forward.restype = ctypes.c_int
    rc = lib.racf_oracle_bev_pool_v2_forward(_p(depth), _p(feat), _p(rd), _p(rf), _p(rb), _p(st), _p(ln), st.numel(),
                                             feat.shape[-1], _p(out))
    assert rc == 0
    return out


def bev_pool_v2_backward(out_grad, depth, feat, ranks_depth, ranks_feat, ranks_bev, interval_starts_bp,
                         interval_lengths_bp):
    """Intervals must be runs of equal ranks_feat (see racformer_b200.bev_pool.backward_intervals)."""
    out_grad, depth, feat = _f32(out_grad), _f32(depth), _f32(feat)
    rd, rf, rb, st, ln = (_i32(t) for t in (ranks_depth, ranks_feat, ranks_bev, interval_starts_bp, interval_lengths_bp))
    dg, fg = torch.zeros_like(depth), torch.zeros_like(feat)
    lib = _load()
    lib.racf_oracle_bev_pool_v2_backward.restype = ctypes.c_int
    rc = lib.racf_oracle_bev_pool_v2_backward(_p(out_grad), _p(depth), _p(feat), _p(rd), _p(rf), _p(rb), _p(st), _p(ln),
                                              st.numel(), feat.shape[-1], _p(dg), _p(fg))
    assert rc == 0
    return dg, fg
