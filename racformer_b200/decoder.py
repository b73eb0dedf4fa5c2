"""Host-side RaCFormer query decoder around the B200 sampling ops (the callers either side of the hot path).

A from-scratch restatement of the reference decoder's call sites so that configs 2-4 can be measured on a box that
has no copy of the reference and no mmcv/mmdet:

  RaCFormerTransformer / ...Decoder / ...DecoderLayer   models/racformer_transformer.py:17-279
  ScaleAdaptiveSelfAttention                            models/racformer_transformer.py:282-335
  RaCFormerSampling (+ sampling_4d, make_sample_points) models/racformer_transformer.py:338-427, models/sparsebev_sampling.py:8-134
  BEVSampling (+ BEVSelfAttention)                      models/racformer_transformer.py:429-546, models/bev_self_attention.py:22-225
  AdaptiveMixing                                        models/racformer_transformer.py:549-616
  RadarBEVTemporalEncoder / ConvGRU / ConvGRUCell       models/racformer_transformer.py:618-720
  mmcv MultiheadAttention / FFN / LearnedPositionalEncoding (mmcv-full 1.6.0, third-party): plain-torch stand-ins

Module and parameter names mirror the reference, so `load_state_dict` accepts the reference's
`pts_bbox_head.transformer.*` weights unchanged. The reference's packing quirks are preserved on purpose
(SURVEY.md 7.3-6): sampling locations are packed B*T*G but scale weights B*G*T; MSDA locations/weights are
queue-major T*B while value/output are batch-major B*T; valid_mask is not applied; unseen points go to view 0.

The two hot ops come from a `SamplingOps` object. The default binds the sm_100a kernels (no CPU fallback); tests
inject CPU callables to check this file's host logic against the reference on a machine without a GPU.

`hoist_invariants=True` (default) computes what does not depend on the queries -- the radar temporal encoder, the
BEV positional encoding add and `value_proj` -- once per forward instead of once per decoder iteration. The layer's
parameters are shared by all six iterations (racformer_transformer.py:84-89), so results are identical; set it to
False to reproduce the reference's schedule.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint as _torch_checkpoint
from .caches import cache_epoch as _cache_epoch


# ---------------------------------------------------------------------------------------------- op binding
class SamplingOps:
    """msmv(feats_channel_last, loc, w) -> [B',Q,C,P];  msda(value, shapes, lsi, loc, aw, im2col_step) -> [B,Q,M*D]."""

    def __init__(self, msmv=None, msda=None, msmv_grouped=None):
        self.msda_pair = None      # inference: both BEV branches of an iteration in one launch (racf_msda_forward_pair)
        if msmv is None or msda is None:
            from . import wrapper
            from .multi_scale_deformable_attn_function import MultiScaleDeformableAttnFunction_fp32, msda_forward_pair
            if msmv is None and msmv_grouped is None:
                msmv_grouped = wrapper.msmv_forward_grouped     # inference-only variant with the un-packing fused in
            msmv = msmv or wrapper.msmv_sampling
            if msda is None:
                msda, self.msda_pair = MultiScaleDeformableAttnFunction_fp32.apply, msda_forward_pair
        self.msmv, self.msda, self.msmv_grouped = msmv, msda, msmv_grouped


_CONST_CACHE = {}


def _const_long(values, device):
    """Small int64 device constants (MSDA spatial_shapes / level_start_index), created once per device so that the
    forward issues no host-to-device copy (the reference rebuilds them every call, bev_self_attention.py:189-190) and
    can be captured in a CUDA graph."""
    key = (str(device), values)
    t = _CONST_CACHE.get(key)
    if t is None:
        t = torch.tensor(values, dtype=torch.long, device=device)
        _CONST_CACHE[key] = t
    return t


# ---------------------------------------------------------------------------------------------- coordinate helpers
MAP_SIZE, RAY_R = 102.4, 65.0   # models/bbox/utils.py:82,93


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def decode_bbox(bboxes, pc_range=None):
    """models/bbox/utils.py:66-80: normalised (x,y,z, log w,l,h, sin, cos[, vx, vy]) -> metric box."""
    xyz = bboxes[..., 0:3].clone()
    wlh = bboxes[..., 3:6].exp()
    rot = torch.atan2(bboxes[..., 6:7], bboxes[..., 7:8])
    if pc_range is not None:
        xyz[..., 0] = xyz[..., 0] * (pc_range[3] - pc_range[0]) + pc_range[0]
        xyz[..., 1] = xyz[..., 1] * (pc_range[4] - pc_range[1]) + pc_range[1]
        xyz[..., 2] = xyz[..., 2] * (pc_range[5] - pc_range[2]) + pc_range[2]
    parts = [xyz, wlh, rot]
    if bboxes.shape[-1] > 8:
        parts.append(bboxes[..., 8:10].clone())
    return torch.cat(parts, dim=-1)


def theta_d2xy_coods(td):
    """Polar (theta in turns, d in units of RAY_R) -> normalised cartesian, clamped to [0,1] (bbox/utils.py:82-90)."""
    centre = MAP_SIZE / 2
    ang = td[..., 0:1] * (2 * torch.pi)
    rad = td[..., 1:2] * RAY_R
    xy = torch.cat([(centre + rad * torch.cos(ang)) / MAP_SIZE, (centre + rad * torch.sin(ang)) / MAP_SIZE], dim=-1)
    return torch.cat([torch.clamp(xy, min=0, max=1), td[..., 2:]], dim=-1)


def xy2theta_d_coods(xy):
    """Normalised cartesian -> polar (bbox/utils.py:93-106, norm=True branch)."""
    centre = MAP_SIZE / 2
    dx = xy[..., 0:1] * MAP_SIZE - centre
    dy = xy[..., 1:2] * MAP_SIZE - centre
    dist = torch.sqrt(dx ** 2 + dy ** 2) / RAY_R
    theta = ((torch.atan2(dy, dx) + 2 * torch.pi) % (2 * torch.pi)) / (2 * torch.pi)
    return torch.cat([theta, dist, xy[..., 2:]], dim=-1)


def rotate_about_z(points, angles):
    """models/utils.py:48-82 (VERSION v1.0.0 convention): row-vector points [...,P,3] times R^T, as a bmm."""
    lead = angles.shape[:-1]
    n = points.shape[-2]
    ang = angles[..., 0].reshape(-1)
    s, c = torch.sin(ang), torch.cos(ang)
    one, zero = torch.ones_like(c), torch.zeros_like(c)
    rot_t = torch.stack([c, s, zero, -s, c, zero, zero, zero, one]).transpose(0, 1).reshape(-1, 3, 3)
    return torch.bmm(points.reshape(-1, n, 3), rot_t).reshape(*lead, n, 3)


def make_sample_points(query_bbox, offset, pc_range):
    """models/sparsebev_sampling.py:8-25: box-relative offsets -> lidar-frame points [B,Q,P,3]."""
    box = decode_bbox(query_bbox, pc_range)
    xyz, wlh, ang = box[..., 0:3], box[..., 3:6], box[..., 6:7]
    delta = rotate_about_z(wlh[:, :, None, :] * offset[..., 0:3], ang)
    return xyz[:, :, None, :] + delta


# ---------------------------------------------------------------------------------------------- MSMV call site
def sampling_4d(ops, sample_points, mlvl_feats, scale_weights, lidar2img, image_h, image_w, eps=1e-5, shared_grads=None):
    """models/sparsebev_sampling.py:28-134 (aggregate=True).

    sample_points [B,Q,T,G,P,3]; mlvl_feats[l] [B*T*G,N,H,W,C]; scale_weights [B,Q,G,T,P,L]; lidar2img [B,T*N,4,4].
    Returns [B,Q,G,T*P,C]. Every point is projected into all N views of its frame; the first view that sees it
    (else view 0) is kept and encoded as z = view/(N-1).
    """
    B, Q, T, G, P, _ = sample_points.shape
    N = lidar2img.shape[1] // T
    pts = sample_points.reshape(B, Q, T, G * P, 3)
    pts_h = torch.cat([pts, torch.ones_like(pts[..., :1])], dim=-1)                  # [B,Q,T,GP,4]
    # cam[b,t,n,q,p,:] = lidar2img[b,t,n] @ pts_h[b,q,t,p]: one [Q*GP,4]x[4,4] GEMM per (b,t,n); the reference
    # materialises the broadcast [B,T,N,Q,GP,4,4] operand instead (sparsebev_sampling.py:51-63)
    mats = lidar2img.reshape(B, T, N, 4, 4)
    rows = pts_h.permute(0, 2, 1, 3, 4).reshape(B, T, 1, Q * G * P, 4)
    cam = torch.matmul(rows, mats.transpose(-1, -2)).reshape(B, T, N, Q, G * P, 4)

    homo = cam[..., 2:3]
    uv = cam[..., 0:2] / torch.maximum(homo, torch.zeros_like(homo) + eps)
    uv = torch.cat([uv[..., 0:1] / image_w, uv[..., 1:2] / image_h], dim=-1)
    valid = ((homo > eps) & (uv[..., 1:2] > 0.0) & (uv[..., 1:2] < 1.0) & (uv[..., 0:1] > 0.0)
             & (uv[..., 0:1] < 1.0)).squeeze(-1).float()                             # [B,T,N,Q,GP]
    valid = valid.permute(0, 1, 3, 4, 2)                                             # [B,T,Q,GP,N]
    uv = uv.permute(0, 1, 3, 4, 2, 5)                                                # [B,T,Q,GP,N,2]
    i_view = torch.argmax(valid, dim=-1, keepdim=True)                               # first seeing view, else 0
    uv = torch.gather(uv, 4, i_view[..., None].expand(-1, -1, -1, -1, 1, 2)).squeeze(4)   # [B,T,Q,GP,2]
    loc = torch.cat([uv, i_view.float() / (N - 1)], dim=-1)                          # [B,T,Q,GP,3]
    loc = loc.reshape(B, T, Q, G, P, 3).permute(0, 1, 3, 2, 4, 5).reshape(B * T * G, Q, P, 3)

    w = scale_weights.reshape(B, Q, G, T, P, -1).permute(0, 2, 3, 1, 4, 5).reshape(B * G * T, Q, P, -1)  # quirk (i)

    return _msmv_packed(ops, mlvl_feats, loc.contiguous(), w.contiguous(), B, T, G, shared_grads)


def _msmv_packed(ops, mlvl_feats, loc, w, B, T, G, shared_grads=None):
    """msmv_sampling on the packed tensors (loc [B*T*G,Q,P,3], w [B*G*T,Q,P,L]) + sampling_4d's un-packing
    (sparsebev_sampling.py:122-134) -> [B,Q,G,T*P,C]."""
    Q, P = loc.shape[1], loc.shape[2]
    if shared_grads is not None and getattr(ops, "msmv_grouped", None) is not None:
        from . import training   # training on CUDA: grouped output + grouped grad_out, feature gradients accumulated in place
        if training.msmv_grouped_supported(mlvl_feats, loc, w):
            return training.MSMVGrouped.apply(shared_grads.get(mlvl_feats), T, G, loc, w, *mlvl_feats)
    out = ops.msmv(mlvl_feats, loc, w)                                               # [B*T*G,Q,C,P]
    C = out.shape[2]
    out = out.reshape(B, T, G, Q, C, P).permute(0, 3, 2, 1, 5, 4)                    # [B,Q,G,T,P,C]
    return out.flatten(3, 4)


# ---------------------------------------------------------------------------------------------- mmcv stand-ins
class MultiheadAttention(nn.Module):
    """mmcv.cnn.bricks.transformer.MultiheadAttention(embed_dims, num_heads, attn_drop, batch_first=True)."""

    def __init__(self, embed_dims, num_heads, attn_drop=0.0, proj_drop=0.0, batch_first=True):
        super().__init__()
        self.batch_first = batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, query, attn_mask=None):
        x = query.transpose(0, 1) if self.batch_first else query
        out = self.attn(query=x, key=x, value=x, attn_mask=attn_mask, need_weights=False)[0]
        if self.batch_first:
            out = out.transpose(0, 1)
        return query + self.proj_drop(out)


class FFN(nn.Module):
    """mmcv FFN(embed_dims, feedforward_channels, ffn_drop): Linear-ReLU-Drop-Linear-Drop + identity."""

    def __init__(self, embed_dims, feedforward_channels, ffn_drop=0.0):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(ffn_drop)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop))

    def forward(self, x):
        return x + self.layers(x)


class LearnedPositionalEncoding(nn.Module):
    """mmcv LearnedPositionalEncoding(num_feats, row_num_embed, col_num_embed) -> [bs, 2*num_feats, h, w]."""

    def __init__(self, num_feats, row_num_embed, col_num_embed):
        super().__init__()
        self.row_embed = nn.Embedding(row_num_embed, num_feats)
        self.col_embed = nn.Embedding(col_num_embed, num_feats)
        nn.init.uniform_(self.row_embed.weight)
        nn.init.uniform_(self.col_embed.weight)

    def forward(self, bs, h, w, device):
        x_embed = self.col_embed(torch.arange(w, device=device))
        y_embed = self.row_embed(torch.arange(h, device=device))
        pos = torch.cat([x_embed.unsqueeze(0).expand(h, -1, -1), y_embed.unsqueeze(1).expand(-1, w, -1)], dim=-1)
        return pos.permute(2, 0, 1).unsqueeze(0).expand(bs, -1, -1, -1)


def _xavier_uniform(linear):
    nn.init.xavier_uniform_(linear.weight)
    if linear.bias is not None:
        nn.init.constant_(linear.bias, 0.0)


def _maybe_checkpoint(module, fn, *args):
    """The reference wraps these blocks in a non-re-entrant activation checkpoint while training."""
    if module.training and module.activation_checkpoint and any(torch.is_tensor(a) and a.requires_grad for a in args):
        return _torch_checkpoint(fn, *args, use_reentrant=False)
    return fn(*args)


# ---------------------------------------------------------------------------------------------- modules
class ScaleAdaptiveSelfAttention(nn.Module):
    def __init__(self, embed_dims=256, num_heads=8, dropout=0.1, pc_range=()):
        super().__init__()
        self.pc_range = list(pc_range)
        self.activation_checkpoint = True
        self.attention = MultiheadAttention(embed_dims, num_heads, dropout, batch_first=True)
        self.gen_tau = nn.Linear(embed_dims, num_heads)

    @torch.no_grad()
    def init_weights(self):
        nn.init.zeros_(self.gen_tau.weight)
        nn.init.uniform_(self.gen_tau.bias, 0.0, 2.0)

    @torch.no_grad()
    def calc_bbox_dists(self, bboxes):
        centres = decode_bbox(bboxes, self.pc_range)[..., :2]
        return -torch.norm(centres[:, :, None, :] - centres[:, None, :, :], dim=-1)     # [B,Q,Q]

    def inner_forward(self, query_bbox, query_feat, pre_attn_mask):
        dist = self.calc_bbox_dists(theta_d2xy_coods(query_bbox))
        tau = self.gen_tau(query_feat).permute(0, 2, 1)                                   # [B,8,Q]
        attn_mask = dist[:, None, :, :] * tau[..., None]                                  # [B,8,Q,Q]
        if pre_attn_mask is not None:   # query denoising: blocked pairs
            attn_mask = attn_mask.masked_fill(pre_attn_mask[None, None], float("-inf"))
        return self.attention(query_feat, attn_mask=attn_mask.flatten(0, 1))

    def forward(self, query_bbox, query_feat, pre_attn_mask=None):
        return _maybe_checkpoint(self, lambda qb, qf: self.inner_forward(qb, qf, pre_attn_mask), query_bbox, query_feat)

    fused_training_core = True

    def attention_core(self, query_bbox, qkv, tau, pre_attn_mask, num_heads, dropout_p):
        """softmax(q k^T / sqrt(d) - tau_i |c_i - c_j| [blocked pairs: -inf]) v for projected qkv [B,Q,3E], tau [B,Q,H] ->
        [B,Q,E]: the attention between in_proj and out_proj of inner_forward (racformer_transformer.py:296-336), with
        autograd."""
        B, Q, E3 = qkv.shape
        E = E3 // 3
        if self.fused_training_core and qkv.is_cuda and E == num_heads * 32 and query_bbox.shape[-1] >= 2:
            from . import training   # csrc/sasa_train.cu: no [B,H,Q,Q] tensor in forward or backward
            if training.SasaAttention.supported(qkv, tau, num_heads):
                self._drop_counter = getattr(self, "_drop_counter", 0) + 1
                seed = (torch.initial_seed() * 40503 + self._drop_counter * 2654435761) & 0xffffffff
                return training.SasaAttention.apply(qkv, tau, query_bbox, pre_attn_mask, self.pc_range, num_heads, dropout_p, seed)
        dist = self.calc_bbox_dists(theta_d2xy_coods(query_bbox))                         # [B,Q,Q], no grad
        mask = dist[:, None, :, :] * tau.permute(0, 2, 1)[..., None]                      # [B,H,Q,Q]
        if pre_attn_mask is not None:
            mask = mask.masked_fill(pre_attn_mask[None, None], float("-inf"))
        q, k, v = (t.reshape(B, Q, num_heads, E // num_heads).transpose(1, 2) for t in qkv.split(E, dim=-1))
        out = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=dropout_p)
        return out.transpose(1, 2).reshape(B, Q, E)


def _use_fused_points(module, *tensors):
    """The fused CUDA point kernels are forward-only: use them when autograd is off and everything lives on a GPU."""
    return (getattr(module, "fused_points", True) and not torch.is_grad_enabled()
            and all(t.is_cuda and t.dtype == torch.float32 for t in tensors))


def _use_fused_points_train(module, *tensors):
    """Training on CUDA: the same point kernels with a fused backward (racformer_b200/training.py, csrc/points_train.cu)."""
    return (getattr(module, "fused_points", True) and getattr(module, "fused_points_train", True) and torch.is_grad_enabled()
            and all(t.is_cuda and t.dtype == torch.float32 for t in tensors))


def _depth_base(d_region, depth_num, device):
    key = ("linspace", str(device), float(d_region), int(depth_num))
    t = _CONST_CACHE.get(key)
    if t is None:
        t = torch.linspace(-d_region, d_region, depth_num, device=device)
        _CONST_CACHE[key] = t
    return t


def _polar_depth_offsets(module, query_feat, d_region, ray_logits=None):
    """linspace(-d, d, D) + learned jitter, shared by both samplers (racformer_transformer.py:395-396, 513-514).
    ray_logits: ray_points_offset(query_feat) when the caller has already computed it."""
    D = module.depth_num
    base = torch.linspace(-d_region, d_region, D, device=query_feat.device, dtype=query_feat.dtype).view(1, 1, D)
    ray = module.ray_points_offset(query_feat) if ray_logits is None else ray_logits
    return base + (ray.sigmoid() * 2 - 1) * d_region / D / 2


class RaCFormerSampling(nn.Module):
    """Image branch: builds [B,Q,T,G,P*D,3] lidar-frame points and per-level weights, then MSMV sampling."""

    def __init__(self, embed_dims=256, num_frames=4, num_groups=4, num_points=8, num_levels=4, depth_num=15, pc_range=()):
        super().__init__()
        self.num_frames, self.num_points, self.num_groups = num_frames, num_points, num_groups
        self.num_levels, self.depth_num, self.pc_range = num_levels, depth_num, list(pc_range)
        self.activation_checkpoint = True
        self.ray_points_offset = nn.Linear(embed_dims, depth_num)
        self.sampling_offset = nn.Linear(embed_dims, depth_num * num_groups * num_points * 3)
        self.scale_weights = nn.Linear(embed_dims, num_groups * num_frames * depth_num * num_points * num_levels)

    @torch.no_grad()
    def init_weights(self):
        nn.init.zeros_(self.sampling_offset.weight)
        nn.init.uniform_(self.sampling_offset.bias, -0.5, 0.5)

    def inner_forward(self, ops, query_ray, query_feat, mlvl_feats, meta, d_region, heads=None):
        B, Q, _ = query_ray.shape
        T, G, Pn, D, pr = self.num_frames, self.num_groups, self.num_points, self.depth_num, self.pc_range
        if _use_fused_points(self, query_ray, query_feat, meta["lidar2img"]):
            from . import points   # one kernel instead of the ~150 PyTorch launches below
            off, ray, sw = heads if heads is not None else (self.sampling_offset(query_feat),
                                                            self.ray_points_offset(query_feat),
                                                            self.scale_weights(query_feat))
            loc, w = points.msmv_points(
                query_ray.contiguous(), off, ray, sw, meta["time_diff"], meta["lidar2img"],
                _depth_base(d_region, D, query_feat.device), pr, d_region, meta["image_w"], meta["image_h"],
                T, G, Pn, D, self.num_levels)
            if getattr(ops, "msmv_grouped", None) is not None:
                return ops.msmv_grouped(mlvl_feats, loc, w, T, G)                      # [B,Q,G,T*P,C] directly
            out = ops.msmv(mlvl_feats, loc, w)                                         # [B*T*G,Q,C,P]
            C = out.shape[2]
            return out.reshape(B, T, G, Q, C, Pn * D).permute(0, 3, 2, 1, 5, 4).flatten(3, 4)
        if _use_fused_points_train(self, query_ray, query_feat, meta["lidar2img"], meta["time_diff"]):
            from . import training   # one forward + one backward launch instead of ~400 autograd nodes
            off, ray, sw = heads if heads is not None else (self.sampling_offset(query_feat),
                                                            self.ray_points_offset(query_feat),
                                                            self.scale_weights(query_feat))
            loc, w = training.MSMVPoints.apply(
                query_ray, off, ray, sw, meta["time_diff"], meta["lidar2img"], _depth_base(d_region, D, query_feat.device),
                (tuple(pr), d_region, meta["image_w"], meta["image_h"], T, G, Pn, D, self.num_levels))
            return _msmv_packed(ops, mlvl_feats, loc, w, B, T, G, meta.get("shared_grads"))
        query_bbox = theta_d2xy_coods(query_ray)
        off, ray, sw = heads if heads is not None else (self.sampling_offset(query_feat), None, None)
        offset = off.view(B, Q, G * Pn * D, 3)
        pts = make_sample_points(query_bbox, offset, pr).reshape(B, Q, 1, G, Pn * D, 3).expand(B, Q, T, G, Pn * D, 3)
        # ego-motion-free warp by the query velocity
        shift = (query_ray[..., 8:].detach()[:, :, None, :] * meta["time_diff"][:, None, :, None])[:, :, :, None, None, :]
        x = (pts[..., 0:1] - shift[..., 0:1] - pr[0]) / (pr[3] - pr[0])
        y = (pts[..., 1:2] - shift[..., 1:2] - pr[1]) / (pr[4] - pr[1])
        polar = xy2theta_d_coods(torch.cat([x, y, pts[..., 2:3]], dim=-1)).reshape(B, Q, T, G, Pn, D, 3)
        depth = _polar_depth_offsets(self, query_feat, d_region, ray).view(B, Q, 1, 1, 1, D, 1)
        polar = torch.cat([polar[..., 0:1], polar[..., 1:2] + depth, polar[..., 2:]], dim=-1)
        cart = theta_d2xy_coods(polar.reshape(B, Q, T, G, Pn * D, 3))
        pts = torch.cat([cart[..., 0:1] * (pr[3] - pr[0]) + pr[0], cart[..., 1:2] * (pr[4] - pr[1]) + pr[1],
                         cart[..., 2:]], dim=-1)
        w = (self.scale_weights(query_feat) if sw is None else sw).view(B, Q, G, T, D * Pn, self.num_levels)
        w = torch.softmax(w, dim=-1)
        return sampling_4d(ops, pts, mlvl_feats, w, meta["lidar2img"], meta["image_h"], meta["image_w"],
                           shared_grads=meta.get("shared_grads"))

    def forward(self, ops, query_ray, query_feat, mlvl_feats, meta, d_region=0.1, heads=None):
        fn = lambda qr, qf, *feats: self.inner_forward(ops, qr, qf, list(feats), meta, d_region, heads)
        return _maybe_checkpoint(self, fn, query_ray, query_feat, *mlvl_feats)


class BEVSelfAttention(nn.Module):
    """models/bev_self_attention.py:22-225: value_proj -> MSDA over the T BEV maps -> softmax queue fusion -> output_proj."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, num_bev_queue=2, im2col_step=64,
                 dropout=0.1, queue_weight=False):
        super().__init__()
        self.embed_dims, self.num_heads, self.num_levels, self.num_points = embed_dims, num_heads, num_levels, num_points
        self.num_bev_queue, self.im2col_step, self.queue_weight = num_bev_queue, im2col_step, queue_weight
        self.dropout = nn.Dropout(dropout)
        if queue_weight:
            self.bev_queue_weight = nn.Linear(embed_dims, num_bev_queue)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        _xavier_uniform(self.value_proj)
        _xavier_uniform(self.output_proj)
        if self.queue_weight:
            _xavier_uniform(self.bev_queue_weight)

    tensor_core_value_proj = True     # inference on CUDA: csrc/linear.cu (see AdaptiveMixing.gemm_precision)

    def project_value(self, bev, pos=None, channel_bias=None, addend=None):
        """([B,T,C,H,W] + pos [C,H,W]) -> value [B*T, H*W, heads, C/heads] (bev_self_attention.py:162-174).
        Query-independent. Inference on CUDA: the add, the permute + copy and the operand split are one kernel and the
        projection (with its bias) runs on the tcgen05 Linear kernel; otherwise the PyTorch ops of the reference."""
        B, T, C = bev.shape[:3]
        if (self.tensor_core_value_proj and not torch.is_grad_enabled() and bev.is_cuda and bev.dtype == torch.float32
                and C % 8 == 0 and (pos is None or pos.numel() == bev[0, 0].numel())):   # C % 8: 16-byte chunks of the split
            from . import linear
            if getattr(self, "_split_value_proj", None) is None:
                self._split_value_proj = linear.SplitLinear(self.value_proj, max_order=linear.SIX_TERMS)
            pixels = bev.permute(0, 1, 3, 4, 2)
            if addend is not None:          # prepared by the caller: [S, C] for channel-last maps, [C, S] otherwise
                if pixels.is_contiguous():
                    x3 = linear.split_tiled(pixels.reshape(-1, C), addend)
                else:
                    x3 = linear.split_bf16x3_chw_to_hwc(bev.reshape(B * T, C, -1).contiguous(), addend, tiled=True)
            elif pixels.is_contiguous():      # channel-last maps (the radar temporal encoder's output): already [B*T*S, C]
                # [C, S] -> [S, C]: a view of LearnedPositionalEncoding's own [H, W, C] layout (no copy); a deferred
                # per-channel bias (the temporal encoder's last convolution) joins the same row-periodic addend
                addend = None if pos is None else pos.reshape(C, -1).t()
                if channel_bias is not None:
                    addend = (channel_bias.expand(pixels[0, 0].numel() // C, C) if addend is None else addend + channel_bias)
                x3 = linear.split_tiled(pixels.reshape(-1, C), None if addend is None else addend.contiguous())
            else:
                cpos = None if pos is None else pos.reshape(C, -1)
                if channel_bias is not None:
                    cpos = channel_bias[:, None].expand(C, bev[0, 0, 0].numel()) if cpos is None else cpos + channel_bias[:, None]
                x3 = linear.split_bf16x3_chw_to_hwc(bev.reshape(B * T, C, -1).contiguous(),
                                                    None if cpos is None else cpos.contiguous(), tiled=True)
            v = self._split_value_proj(x3=x3)
            return v.reshape(B * T, x3.rows // (B * T), self.num_heads, -1)
        if channel_bias is not None:
            bev = bev + channel_bias.view(1, 1, C, 1, 1)
        if pos is not None:
            bev = bev + pos.view(1, 1, C, *bev.shape[3:])
        pixels = bev.reshape(B * T, C, -1).permute(0, 2, 1)
        if (self.tensor_core_value_proj and torch.is_grad_enabled() and bev.is_cuda and bev.dtype == torch.float32
                and C % 8 == 0 and pixels.shape[1] % 8 == 0):
            from . import linear   # training on CUDA: forward, input and weight gradient on the tcgen05 Linear kernel
            if getattr(self, "_train_value_proj", None) is None:
                self._train_value_proj = linear.TrainableSplitLinear(self.value_proj, max_order=linear.SIX_TERMS)
            v = self._train_value_proj(pixels)
        else:
            v = self.value_proj(pixels)
        return v.reshape(B * T, v.shape[1], self.num_heads, -1)

    def forward(self, ops, query, value, sampling_locations, attention_weights, spatial_shapes, raw=False, shared_grads=None,
                queue_logits=None):
        B, Q, C = query.shape
        T, M, L, P = self.num_bev_queue, self.num_heads, self.num_levels, self.num_points
        loc = sampling_locations.view(B, Q, M, T, L, P, 2).permute(3, 0, 1, 2, 4, 5, 6).reshape(B * T, Q, M, L, P, 2)
        aw = attention_weights.view(B, Q, M, T, L, P).permute(3, 0, 1, 2, 4, 5).reshape(B * T, Q, M, L, P)   # quirk (ii)
        return self.attend(ops, query, value, loc, aw, spatial_shapes, raw=raw, shared_grads=shared_grads,
                           queue_logits=queue_logits)

    def attend(self, ops, query, value, loc, aw, spatial_shapes, queue_logits=None, raw=False, shared_grads=None):
        """loc [T*B,Q,M,L,P,2] / aw [T*B,Q,M,L,P] already in the queue-major packing; queue_logits: bev_queue_weight(query)
        when the caller has already computed it (the decoder layer's stacked head launch). raw: return the MSDA output
        [B*T,Q,C] and the queue logits (None: plain mean) -- the caller's row program (csrc/rowops.cu) does the queue
        fusion, output_proj and residual."""
        B, Q, C = query.shape
        T = self.num_bev_queue
        shapes = _const_long((tuple(int(v) for v in spatial_shapes),), value.device)
        lsi = _const_long((0,), value.device)
        if (shared_grads is not None and value.is_cuda and value.requires_grad and value.dtype == torch.float32
                and value.is_contiguous()):
            from . import training   # training on CUDA: the value gradient of all iterations accumulates in one buffer
            out = training.MSDAShared.apply(shared_grads.get([value]), value, shapes, lsi, loc.contiguous(),
                                            aw.contiguous(), self.im2col_step)
        else:
            out = ops.msda(value, shapes, lsi, loc.contiguous(), aw.contiguous(), self.im2col_step)   # [B*T,Q,C]
        if raw:
            if not self.queue_weight:
                return out, None
            return out, (queue_logits if queue_logits is not None else self.bev_queue_weight(query))
        out = out.permute(1, 2, 0).reshape(Q, C, B, T)
        if self.queue_weight:
            logits = queue_logits if queue_logits is not None else self.bev_queue_weight(query)
            qw = torch.softmax(logits.permute(1, 0, 2).reshape(Q, 1, B, T), dim=-1)
            out = torch.sum(out * qw, dim=-1)
        else:
            out = torch.sum(out, dim=-1) / T
        out = self.output_proj(out.permute(2, 0, 1))
        return self.dropout(out) + query


class ConvGRUCell(nn.Module):
    def __init__(self, input_channels, hidden_channels, kernel_size):
        super().__init__()
        self.hidden_channels = hidden_channels
        self.gates_conv = nn.Conv2d(input_channels + hidden_channels, 3 * hidden_channels, kernel_size,
                                    padding=kernel_size // 2)
        self.matching_layer = nn.Conv2d(hidden_channels, input_channels, 1)

    fused_gates = True      # inference on CUDA, channel-last tensors: the nine elementwise ops below as one launch

    def forward(self, x, h_prev):
        gates = self.gates_conv(torch.cat([x, self.matching_layer(h_prev)], dim=1))
        if self.fused_gates and not torch.is_grad_enabled() and gates.is_cuda and gates.dim() == 4:
            from . import points
            h = points.convgru_gates(gates, h_prev)
            if h is not None:
                return h
        z, r, cand = torch.split(gates, self.hidden_channels, dim=1)
        z, r = torch.sigmoid(z), torch.sigmoid(r)
        cand = torch.tanh(cand + r * h_prev)
        return (1 - z) * h_prev + z * cand


class ConvGRU(nn.Module):
    """racformer_transformer.py:665-694: recurrence over the first min(T,4) frames, zeros afterwards; frames >= 2
    run without gradient."""

    def __init__(self, input_channels, hidden_channels, kernel_size):
        super().__init__()
        self.convGRUCell = ConvGRUCell(input_channels, hidden_channels, kernel_size)
        self.hidden_channels = hidden_channels

    def forward(self, x):
        B, T, _, H, W = x.shape
        # channel-last frames (the inference path of RadarBEVTemporalEncoder) keep the whole recurrence channel-last: no cuDNN
        # NCHW <-> NHWC conversions around the two convolutions of every step
        cl = x.is_cuda and x[:, 0].dim() == 4 and x[:, 0].is_contiguous(memory_format=torch.channels_last)
        h = torch.zeros(B, self.hidden_channels, H, W, device=x.device, dtype=x.dtype)
        if cl:
            h = h.contiguous(memory_format=torch.channels_last)
        zeros = h.clone()
        steps = 4 if T > 4 else T
        out = []
        for t in range(T):
            if t >= steps:
                out.append(zeros)
                continue
            if t > 1:
                with torch.no_grad():
                    h = self.convGRUCell(x[:, t], h)
            else:
                h = self.convGRUCell(x[:, t], h)
            out.append(h)
        return torch.stack(out, dim=1)


class RadarBEVTemporalEncoder(nn.Module):
    def __init__(self, embed_dims=256, hidden_dims=64, num_frames=8, kernel_size=3, downsample_ratio=2):
        super().__init__()
        self.hidden_dims, self.downsample_ratio = hidden_dims, downsample_ratio
        self.activation_checkpoint = True
        self.convGRU = ConvGRU(hidden_dims, hidden_dims, kernel_size)
        self.temporal_fusion = nn.Conv2d(embed_dims + hidden_dims, embed_dims, kernel_size, padding=kernel_size // 2)
        self.downsample = nn.Conv2d(embed_dims, hidden_dims, kernel_size=3, stride=downsample_ratio, padding=1)
        self.upsample = nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True),
                                      nn.Conv2d(hidden_dims, hidden_dims, kernel_size=3, padding=1))

    channels_last = True    # inference on CUDA: run the convolutions on channel-last tensors (no cuDNN layout round trips)

    def can_defer_bias(self, bev):
        """The channel-last inference path can hand temporal_fusion's bias to the consumer instead of adding it itself."""
        return self.channels_last and not torch.is_grad_enabled() and bev.is_cuda and bev.dtype == torch.float32

    def inner_forward(self, bev, defer_bias=False):
        """defer_bias (only with can_defer_bias): returns (output WITHOUT temporal_fusion's bias, that bias) -- cuDNN adds a
        convolution's bias in a separate pass over the 134 MB output; the caller (BEVSampling.prepare_value) adds it where it
        adds the positional encoding, inside value_proj's operand split."""
        B, T, C, H, W = bev.shape
        r = self.downsample_ratio
        if self.channels_last and not torch.is_grad_enabled() and bev.is_cuda:
            # Same operators on NHWC memory: cuDNN's fp32 / TF32 kernels are NHWC kernels, so NCHW tensors are converted
            # in and out around every convolution (0.27 ms per forward at f8). The result stays channel-last -- it is
            # exactly the [B*T*H*W, C] pixel matrix value_proj consumes (BEVSelfAttention.project_value).
            if bev.dtype == torch.float32 and bev.is_contiguous():
                from . import points   # one tiled transpose writes x and the first C channels of the concatenation
                x, both = points.to_channels_last(bev.flatten(0, 1), self.hidden_dims)
            else:
                x = bev.flatten(0, 1).contiguous(memory_format=torch.channels_last)
                both = torch.empty((B * T, C + self.hidden_dims, H, W), dtype=bev.dtype, device=bev.device,
                                   memory_format=torch.channels_last)
                both[:, :C].copy_(x)
            down = self.downsample(x).unflatten(0, (B, T))
            state = self.convGRU(down).flatten(0, 1).contiguous(memory_format=torch.channels_last)
            up = None
            if bev.dtype == torch.float32:
                from . import points   # 2x bilinear upsampling of the channel-last state in one launch (ATen's NHWC kernel: 70 us)
                up = points.upsample2x_bilinear(state)
            if up is not None:      # bias add and the copy into the concatenation buffer as one pass
                uc = self.upsample[1]
                torch.add(F.conv2d(up, uc.weight, None, uc.stride, uc.padding), uc.bias.view(1, -1, 1, 1), out=both[:, C:])
            else:
                both[:, C:].copy_(self.upsample(state))
            if defer_bias:
                tf = self.temporal_fusion
                return F.conv2d(both, tf.weight, None, tf.stride, tf.padding).unflatten(0, (B, T)), tf.bias
            return self.temporal_fusion(both).unflatten(0, (B, T))
        down = self.downsample(bev.flatten(0, 1)).reshape(B, T, self.hidden_dims, H // r, W // r)
        hid = self.upsample(self.convGRU(down).flatten(0, 1)).reshape(B, T, self.hidden_dims, H, W)
        return self.temporal_fusion(torch.cat([bev, hid], dim=2).flatten(0, 1)).reshape(B, T, C, H, W)

    def forward(self, bev):
        return _maybe_checkpoint(self, self.inner_forward, bev)


class BEVSampling(nn.Module):
    """BEV branch (radar or LSS): polar sampling points in the BEV plane + MSDA over the T-frame BEV queue."""

    def __init__(self, embed_dims=256, num_frames=4, num_points=8, num_heads=4, num_levels=4, pc_range=(),
                 spatial_shapes=(128, 128), depth_num=30, temp_radar=False):
        super().__init__()
        self.num_frames, self.num_points, self.num_heads, self.num_levels = num_frames, num_points, num_heads, num_levels
        self.embed_dims, self.pc_range, self.depth_num, self.temp_radar = embed_dims, list(pc_range), depth_num, temp_radar
        self.activation_checkpoint = True
        self.ray_points_offset = nn.Linear(embed_dims, depth_num)
        self.sampling_offset = nn.Linear(embed_dims, depth_num * num_heads * num_points * 2)
        self.scale_weights = nn.Linear(embed_dims, num_heads * num_levels * depth_num * num_points)
        # the reference hard-codes num_feats=128 (embed_dims 256); embed_dims // 2 is the same there and keeps other widths valid
        self.positional_encoding = LearnedPositionalEncoding(embed_dims // 2, row_num_embed=spatial_shapes[1],
                                                             col_num_embed=spatial_shapes[0])
        self.attention = BEVSelfAttention(embed_dims, num_heads=4, num_levels=1, num_points=num_points * depth_num,
                                          num_bev_queue=num_frames, queue_weight=True)
        if temp_radar:
            self.temporal_encoder = RadarBEVTemporalEncoder(embed_dims, 64, num_frames)

    @torch.no_grad()
    def init_weights(self):
        nn.init.zeros_(self.sampling_offset.weight)
        nn.init.uniform_(self.sampling_offset.bias, -0.5, 0.5)
        self.attention.init_weights()

    def prepare_value(self, bev_feats):
        """Everything that does not depend on the queries: temporal encoder (radar), + positional encoding, value_proj."""
        bias = None
        if self.temp_radar:
            enc = self.temporal_encoder
            if bev_feats.shape[0] == 1 and enc.can_defer_bias(bev_feats):
                bev_feats, bias = enc.inner_forward(bev_feats, defer_bias=True)
            else:
                bev_feats = enc(bev_feats)
        B, T, C, H, W = bev_feats.shape
        if (B == 1 and not torch.is_grad_enabled() and bev_feats.is_cuda and bev_feats.dtype == torch.float32
                and self.attention.tensor_core_value_proj and C % 8 == 0):
            # the addend of value_proj's operand split (positional encoding [+ deferred bias]) depends on weights only: built once
            # per weight version in the layout the split reads ([S, C] rows for channel-last maps, [C, S] otherwise)
            rows = bev_feats.permute(0, 1, 3, 4, 2).is_contiguous()
            pe = self.positional_encoding
            key = (pe.row_embed.weight.data_ptr(), pe.row_embed.weight._version, pe.col_embed.weight.data_ptr(),
                   pe.col_embed.weight._version, None if bias is None else (bias.data_ptr(), bias._version), H, W, rows,
                   str(bev_feats.device), _cache_epoch())
            hit = getattr(self, "_addend", None)
            if hit is None or hit[0] != key:
                pos = pe(1, H, W, bev_feats.device).to(bev_feats.dtype).reshape(C, H * W)
                if bias is not None:
                    pos = pos + bias.detach()[:, None]
                hit = self._addend = (key, (pos.t() if rows else pos).contiguous())
            return self.attention.project_value(bev_feats, addend=hit[1]), (H, W)
        pos = self.positional_encoding(B, H, W, bev_feats.device).to(bev_feats.dtype)
        if B == 1:
            return self.attention.project_value(bev_feats, pos.reshape(C, H, W), channel_bias=bias), (H, W)
        return self.attention.project_value(bev_feats + pos.view(B, 1, C, H, W)), (H, W)

    def fused_point_tensors(self, query_ray, query_feat, meta, d_region, heads=None):
        """Sampling locations, attention weights (queue-major packing) and queue logits from the fused point kernel."""
        from . import points
        T, M, Pn, D, pr = self.num_frames, self.num_heads, self.num_points, self.depth_num, self.pc_range
        off, ray, sw, qw = heads if heads is not None else (self.sampling_offset(query_feat),
                                                            self.ray_points_offset(query_feat),
                                                            self.scale_weights(query_feat), None)
        loc, aw = points.bev_points(query_ray.contiguous(), off, ray, sw, meta["time_diff"],
                                    _depth_base(d_region, D, query_feat.device), pr, d_region, T, M, Pn, D)
        return loc, aw, qw

    def sample(self, ops, query_ray, query_feat, value, hw, meta, d_region, heads=None, raw=False):
        B, Q, _ = query_ray.shape
        T, M, Pn, D, pr = self.num_frames, self.num_heads, self.num_points, self.depth_num, self.pc_range
        if self.num_levels == 1 and _use_fused_points(self, query_ray, query_feat, value):
            loc, aw, qw = self.fused_point_tensors(query_ray, query_feat, meta, d_region, heads)
            return self.attention.attend(ops, query_feat, value, loc, aw, hw, queue_logits=qw, raw=raw,
                                         shared_grads=meta.get("shared_grads"))
        if self.num_levels == 1 and _use_fused_points_train(self, query_ray, query_feat, value, meta["time_diff"]):
            from . import training
            off, ray, sw, qw = heads if heads is not None else (self.sampling_offset(query_feat),
                                                                self.ray_points_offset(query_feat),
                                                                self.scale_weights(query_feat), None)
            loc, aw = training.BEVPoints.apply(query_ray, off, ray, sw, meta["time_diff"],
                                               _depth_base(d_region, D, query_feat.device), (tuple(pr), d_region, T, M, Pn, D))
            return self.attention.attend(ops, query_feat, value, loc, aw, hw, queue_logits=qw, raw=raw,
                                         shared_grads=meta.get("shared_grads"))
        query_bbox = theta_d2xy_coods(query_ray)
        off, ray, sw, qw = heads if heads is not None else (self.sampling_offset(query_feat), None, None, None)
        offset = off.view(B, Q, M * Pn * D, 2)
        offset = torch.cat([offset, torch.zeros_like(offset[..., 0:1])], dim=-1)
        pts = make_sample_points(query_bbox, offset, pr).reshape(B, Q, 1, M, Pn * D, 3).expand(B, Q, T, M, Pn * D, 3)
        shift = (query_ray[..., 8:].detach()[:, :, None, :] * meta["time_diff"][:, None, :, None])[:, :, :, None, None, :]
        x = (pts[..., 0:1] - shift[..., 0:1] - pr[0]) / (pr[3] - pr[0])
        y = (pts[..., 1:2] - shift[..., 1:2] - pr[1]) / (pr[4] - pr[1])
        polar = xy2theta_d_coods(torch.cat([x, y], dim=-1)).reshape(B, Q, T, M, Pn, D, 2)
        depth = _polar_depth_offsets(self, query_feat, d_region, ray).view(B, Q, 1, 1, 1, D, 1)
        polar = torch.cat([polar[..., 0:1], polar[..., 1:2] + depth], dim=-1).reshape(B, Q, T, M, Pn * D, 2)
        loc = theta_d2xy_coods(polar).permute(0, 1, 3, 2, 4, 5).contiguous()               # [B,Q,M,T,P,2]
        w = (self.scale_weights(query_feat) if sw is None else sw).view(B, Q, M, 1, self.num_levels, D * Pn)
        w = torch.softmax(w, dim=-1).expand(B, Q, M, T, self.num_levels, D * Pn).contiguous()
        return self.attention(ops, query_feat, value, loc, w, hw, raw=raw, shared_grads=meta.get("shared_grads"),
                              queue_logits=qw)

    def forward(self, ops, query_ray, query_feat, bev_feats, meta, d_region=0.1, prepared=None, heads=None, raw=False):
        def fn(qr, qf, bev):
            value, hw = prepared if prepared is not None else self.prepare_value(bev)
            return self.sample(ops, qr, qf, value, hw, meta, d_region, heads, raw=raw)
        return _maybe_checkpoint(self, fn, query_ray, query_feat, bev_feats)


def _tf32_round(x):
    """Round-to-nearest-even to TF32's 10 explicit mantissa bits, kept in fp32 storage."""
    i = x.view(torch.int32)
    return ((i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF).view(torch.float32)


def _tf32_split(x):
    """x ~= hi + lo with hi, lo exactly representable in TF32 (|x - hi - lo| <= 2^-22 |x|)."""
    hi = _tf32_round(x)
    return hi, _tf32_round(x - hi)


class _SplitTF32Linear:
    """OPT-IN (inference): y = x @ W^T + b on TF32 tensor cores at near-fp32 accuracy.

    Operand splitting: x = xh + xl, W = Wh + Wl, y ~= xh Wh^T + xh Wl^T + xl Wh^T, evaluated as ONE TF32 GEMM over the
    K-concatenated operands [xh, xh, xl] x [Wh, Wl, Wh]; every product of two TF32 numbers is exact in fp32. Long
    reductions are cut into `chunk`-sized pieces that are summed in fp32 outside the tensor cores, because their
    internal accumulation truncates (K = 32768 in one go gave 3e-4 error, profiles/r01_tf32x3_experiment.json).
    Measured on B200: 2.6e-6 mean relative error vs an fp64 product (fp32 SGEMM: 2.5e-7; plain TF32: 2.9e-4) at
    4.5x the SGEMM speed. Not the default: the reference computes these layers with fp32 SGEMM.
    """

    def __init__(self, linear, chunk=512):
        self.linear, self.chunk, self._key, self._w3 = linear, chunk, None, None

    def _weights(self):
        lin = self.linear
        key = (lin.weight.data_ptr(), lin.weight._version, lin.bias.data_ptr(), lin.bias._version, _cache_epoch())
        if self._key != key:
            w = torch.cat([lin.weight.detach(), lin.bias.detach()[:, None]], dim=1)          # bias as one more K column
            k = w.shape[1]
            nchunk = (k + self.chunk - 1) // self.chunk
            kpad = ((k + nchunk * 4 - 1) // (nchunk * 4)) * nchunk * 4
            w = F.pad(w, (0, kpad - k))
            wh, wl = _tf32_split(w.reshape(w.shape[0], nchunk, kpad // nchunk))
            self._w3 = torch.cat([wh, wl, wh], dim=2).permute(1, 2, 0).contiguous()          # [chunks, 3*kc, out]
            self._key, self._kpad, self._nchunk = key, kpad, nchunk
        return self._w3

    def __call__(self, x):
        w3 = self._weights()
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        x2 = F.pad(torch.cat([x2, torch.ones_like(x2[:, :1])], dim=1), (0, self._kpad - x2.shape[1] - 1))
        xh, xl = _tf32_split(x2.reshape(x2.shape[0], self._nchunk, -1))
        x3 = torch.cat([xh, xh, xl], dim=2).permute(1, 0, 2)                                 # [chunks, rows, 3*kc]
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            y = torch.bmm(x3, w3)                                                            # [chunks, rows, out]
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old
        y = y.sum(0) if self._nchunk > 1 else y[0]
        return y.reshape(*lead, -1)


class AdaptiveMixing(nn.Module):
    """racformer_transformer.py:549-616 (AdaMixer): query-generated channel and point mixing of the sampled features."""

    def __init__(self, in_dim, in_points, n_groups=1, out_points=None):
        super().__init__()
        self.in_dim, self.in_points, self.n_groups = in_dim, in_points, n_groups
        self.out_points = out_points if out_points is not None else in_points
        self.eff_in_dim = self.eff_out_dim = in_dim // n_groups
        self.m_parameters = self.eff_in_dim * self.eff_out_dim
        self.s_parameters = self.in_points * self.out_points
        self.activation_checkpoint = True
        self.fold_bias = True
        self.fused_core = True              # inference: csrc/mixing.cu instead of 2 bmm + 2 layer_norm + 2 relu
        self._folded = None
        # the two large Linear layers in inference on CUDA: "bf16x6" / "bf16x9" = csrc/linear.cu (tcgen05; every fp32
        # operand split exactly into three bf16 pieces, the six largest / all nine piece products accumulated in fp32 --
        # measured closer to an fp64 product than cuBLAS SGEMM, profiles/r01_linear_bf16x3_check.json); "fp32" = cuBLAS
        # SGEMM; "tf32x3" = the earlier cuBLAS TF32 operand-split experiment (less accurate, opt-in only)
        self.gemm_precision = "bf16x6"
        self._split_gen = self._split_out = None
        self._split = {}
        self.parameter_generator = nn.Linear(in_dim, n_groups * (self.m_parameters + self.s_parameters))
        self.out_proj = nn.Linear(self.eff_out_dim * self.out_points * n_groups, in_dim)

    @torch.no_grad()
    def init_weights(self):
        nn.init.zeros_(self.parameter_generator.weight)

    def _tensor_core_linear(self, query):
        """True when the two large Linear layers run on csrc/linear.cu (tcgen05, exact bf16 operand splitting)."""
        return (self.gemm_precision in ("bf16x9", "bf16x6") and not torch.is_grad_enabled() and query.is_cuda
                and query.dtype == torch.float32)

    train_tensor_cores = True      # training on CUDA: the two large Linear layers with autograd on csrc/linear.cu

    def _train_linear(self, name, x):
        """The tensor-core Linear with autograd (linear.TrainableSplitLinear) for `name` when it applies to x, else None."""
        if not (self.train_tensor_cores and self.gemm_precision in ("bf16x9", "bf16x6") and torch.is_grad_enabled()
                and x.is_cuda and x.dtype == torch.float32):
            return None
        from . import linear
        order = linear.ALL_TERMS if self.gemm_precision == "bf16x9" else linear.SIX_TERMS
        cur = self._split.get(("train", name))
        if cur is None or cur.max_order != order:
            cur = self._split[("train", name)] = linear.TrainableSplitLinear(getattr(self, name), max_order=order)
        return cur if cur.supports(x) else None

    def _split_linear(self, name):
        from . import linear
        order = linear.ALL_TERMS if self.gemm_precision == "bf16x9" else linear.SIX_TERMS
        cur = self._split.get(name)
        if cur is None or cur.max_order != order:
            cur = self._split[name] = linear.SplitLinear(getattr(self, name), max_order=order)
        return cur

    def _generate(self, query):
        """parameter_generator(query). Inference on CUDA: the tcgen05 kernel (bias added in its epilogue). With
        gemm_precision == "fp32" the bias is folded into the cuBLAS GEMM as an extra K column ([q, 1] @ [W, b]^T): cuBLAS
        otherwise adds it in a separate pass over the 236 MB output (0.16 ms per layer on B200,
        profiles/r01_decoder_forward_kernel_breakdown_fused.json). Same sum, rounded inside the accumulator."""
        lin = self.parameter_generator
        if self._tensor_core_linear(query):
            x3 = getattr(self, "_query_x3", None)      # the caller has already split these query features (stacked heads)
            if x3 is not None and x3.rows == query.numel() // query.shape[-1]:
                return self._split_linear("parameter_generator")(x3=x3, lead=tuple(query.shape[:-1]))
            return self._split_linear("parameter_generator")(query)
        trainable = self._train_linear("parameter_generator", query)
        if trainable is not None:
            return trainable(query)
        if torch.is_grad_enabled() or not query.is_cuda or lin.bias is None or not self.fold_bias:
            return lin(query)
        if self.gemm_precision == "tf32x3":
            if self._split_gen is None:
                self._split_gen = _SplitTF32Linear(lin)
            return self._split_gen(query)
        key = (lin.weight.data_ptr(), lin.weight._version, lin.bias.data_ptr(), lin.bias._version, _cache_epoch())
        if self._folded is None or self._folded[0] != key:
            self._folded = (key, torch.cat([lin.weight.detach(), lin.bias.detach()[:, None]], dim=1).contiguous())
        ones = torch.ones_like(query[..., :1])
        return F.linear(torch.cat([query, ones], dim=-1), self._folded[1])

    def inner_forward(self, x, query):
        B, Q, G, P, C = x.shape
        params = self._generate(query).reshape(B * Q, G, -1)
        if self.fused_core and not torch.is_grad_enabled() and x.is_cuda and x.dtype == torch.float32:
            from . import points   # one kernel for matmul-LN-ReLU-matmul-LN-ReLU
            split = self._tensor_core_linear(query)      # emit the bf16 pieces out_proj consumes, no fp32 round trip
            core = points.adaptive_mixing_core(x.reshape(B * Q * G, P, C).contiguous(),
                                               params.reshape(B * Q * G, -1).contiguous(), self.out_points, split=split,
                                               tiled_groups=G if split else 0, tensor_cores=split)
            if core is not None:
                if split:
                    return query + self._split_linear("out_proj")(x3=core, lead=(B, Q))
                return query + self._project(core.reshape(B, Q, -1))
        if self.fused_core and torch.is_grad_enabled() and x.is_cuda and x.dtype == torch.float32:
            from . import training   # training on CUDA: one forward and one (recomputing) backward kernel for the core
            xs, ps = x.reshape(B * Q * G, P, C), params.reshape(B * Q * G, -1)
            if training.AdaptiveMixingCore.supported(xs, ps, self.out_points):
                core = training.AdaptiveMixingCore.apply(xs, ps, self.out_points)
                return query + self._project(core.reshape(B, Q, -1))
        m, s = params.split([self.m_parameters, self.s_parameters], 2)
        m = m.reshape(B * Q, G, self.eff_in_dim, self.eff_out_dim)
        s = s.reshape(B * Q, G, self.out_points, self.in_points)
        out = torch.matmul(x.reshape(B * Q, G, P, C), m)
        out = F.relu(F.layer_norm(out, [out.size(-2), out.size(-1)]))
        out = torch.matmul(s, out)
        out = F.relu(F.layer_norm(out, [out.size(-2), out.size(-1)]))
        return query + self._project(out.reshape(B, Q, -1))

    def _project(self, out):
        if self._tensor_core_linear(out):
            return self._split_linear("out_proj")(out)
        trainable = self._train_linear("out_proj", out)
        if trainable is not None:
            return trainable(out)
        if self.gemm_precision == "tf32x3" and not torch.is_grad_enabled() and out.is_cuda:
            if self._split_out is None:
                self._split_out = _SplitTF32Linear(self.out_proj)
            return self._split_out(out)
        return self.out_proj(out)

    def forward(self, x, query, query_x3=None):
        """query_x3: linear.TiledOperand of `query` when the caller has one (inference; the split is then not repeated)."""
        self._query_x3 = query_x3
        try:
            return _maybe_checkpoint(self, self.inner_forward, x, query)
        finally:
            self._query_x3 = None


class RaCFormerTransformerDecoderLayer(nn.Module):
    def __init__(self, embed_dims, num_frames=8, num_points=4, num_points_bev=4, num_levels=4, num_classes=10,
                 code_size=10, num_cls_fcs=2, num_reg_fcs=2, img_depth_num=3, bev_depth_num=5, num_ray=150,
                 pc_range=(), d_region_list=(0.15, 0.1, 0.1, 0.08, 0.08, 0.05), spatial_shapes=(128, 128)):
        super().__init__()
        self.embed_dims, self.num_classes, self.code_size = embed_dims, num_classes, code_size
        self.pc_range, self.d_region_list, self.num_ray = list(pc_range), list(d_region_list), num_ray
        self.position_encoder = nn.Sequential(
            nn.Linear(3, embed_dims), nn.LayerNorm(embed_dims), nn.ReLU(inplace=True),
            nn.Linear(embed_dims, embed_dims), nn.LayerNorm(embed_dims), nn.ReLU(inplace=True))
        self.self_attn = ScaleAdaptiveSelfAttention(embed_dims, num_heads=8, dropout=0.1, pc_range=pc_range)
        self.sampling = RaCFormerSampling(embed_dims, num_frames=num_frames, num_groups=4, num_points=num_points,
                                          num_levels=num_levels, depth_num=img_depth_num, pc_range=pc_range)
        self.sampling_radar_bev = BEVSampling(embed_dims, num_frames=num_frames, num_heads=4, num_points=num_points_bev,
                                              num_levels=1, pc_range=pc_range, depth_num=bev_depth_num,
                                              spatial_shapes=spatial_shapes, temp_radar=True)
        self.sampling_lss_bev = BEVSampling(embed_dims, num_frames=num_frames, num_heads=4, num_points=num_points_bev,
                                            num_levels=1, pc_range=pc_range, depth_num=bev_depth_num,
                                            spatial_shapes=spatial_shapes)
        self.mixing = AdaptiveMixing(in_dim=embed_dims, in_points=num_points * num_frames * img_depth_num, n_groups=4,
                                     out_points=128)
        self.ffn = FFN(embed_dims, feedforward_channels=512, ffn_drop=0.1)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(embed_dims), nn.LayerNorm(embed_dims), nn.LayerNorm(embed_dims)
        self.fusion = nn.Linear(embed_dims * 3, embed_dims)
        self.norm_radar_bev, self.norm_lss_bev = nn.LayerNorm(embed_dims), nn.LayerNorm(embed_dims)
        self.norm_fusion = nn.LayerNorm(embed_dims)
        cls_branch = []
        for _ in range(num_cls_fcs):
            cls_branch += [nn.Linear(embed_dims, embed_dims), nn.LayerNorm(embed_dims), nn.ReLU(inplace=True)]
        cls_branch.append(nn.Linear(embed_dims, num_classes))
        self.cls_branch = nn.Sequential(*cls_branch)
        reg_branch = []
        for _ in range(num_reg_fcs):
            reg_branch += [nn.Linear(embed_dims, embed_dims), nn.ReLU(inplace=True)]
        reg_branch.append(nn.Linear(embed_dims, code_size))
        self.reg_branch = nn.Sequential(*reg_branch)

    @torch.no_grad()
    def init_weights(self):
        self.self_attn.init_weights()
        self.sampling.init_weights()
        self.mixing.init_weights()
        self.sampling_radar_bev.init_weights()
        self.sampling_lss_bev.init_weights()
        nn.init.constant_(self.cls_branch[-1].bias, float(-math.log((1 - 0.01) / 0.01)))
        _xavier_uniform(self.fusion)

    stacked_heads = True    # inference on CUDA: the 11 sampling heads that read query_feat run as one tcgen05 launch

    def _sampling_heads(self, query_feat):
        """sampling_offset / ray_points_offset / scale_weights (/ bev_queue_weight) of the radar-BEV, LSS-BEV and image
        samplers all read the same query features: one stacked Linear launch (csrc/linear.cu) instead of 11 small SGEMMs
        with separate bias kernels. Returns (radar 4, lss 4, image 3) or None when the fused point kernels are not in use."""
        radar, lss, img = self.sampling_radar_bev, self.sampling_lss_bev, self.sampling
        if (not self.stacked_heads or torch.is_grad_enabled() or not query_feat.is_cuda or query_feat.dtype != torch.float32
                or self.embed_dims > 512 or radar.num_levels != 1 or lss.num_levels != 1
                or not all(getattr(m, "fused_points", True) for m in (radar, lss, img))
                or not (radar.attention.queue_weight and lss.attention.queue_weight)):
            return None
        if getattr(self, "_heads", None) is None:
            from . import linear
            self._heads = linear.MultiSplitLinear(
                [radar.sampling_offset, radar.ray_points_offset, radar.scale_weights, radar.attention.bev_queue_weight,
                 lss.sampling_offset, lss.ray_points_offset, lss.scale_weights, lss.attention.bev_queue_weight,
                 img.sampling_offset, img.ray_points_offset, img.scale_weights])
        B, Q, _ = query_feat.shape
        from . import linear
        x3 = linear.split_tiled(query_feat.reshape(-1, query_feat.shape[-1]).contiguous())
        self._heads_x3 = x3         # parameter_generator of the mixing block reads the same rows: _forward hands the split on
        outs = [o.view(B, Q, -1) for o in self._heads(x3=x3)]
        return outs[0:4], outs[4:8], outs[8:11]

    def _sampling_heads_train(self, query_feat):
        """Training counterpart of _sampling_heads: the eleven sampling heads as ONE row chain with autograd (one forward
        launch, one backward launch that also sums their input gradients, eleven weight-gradient GEMMs)."""
        from . import rowtrain
        radar, lss, img = self.sampling_radar_bev, self.sampling_lss_bev, self.sampling
        if not (radar.attention.queue_weight and lss.attention.queue_weight):
            return None
        lins = [radar.sampling_offset, radar.ray_points_offset, radar.scale_weights, radar.attention.bev_queue_weight,
                lss.sampling_offset, lss.ray_points_offset, lss.scale_weights, lss.attention.bev_queue_weight,
                img.sampling_offset, img.ray_points_offset, img.scale_weights]
        B, Q, E = query_feat.shape
        width = max([E] + [-(-lin.out_features // 4) * 4 for lin in lins])
        p = rowtrain.RowChain(B * Q, width=width, num_bufs=2)
        p.load(0, query_feat.contiguous())
        handles = []
        for lin in lins:
            p.linear(1, 0, lin)
            handles.append(p.store(1, lin.out_features))
        res = p.run()
        outs = [res[h].view(B, Q, -1) for h in handles]
        return outs[0:4], outs[4:8], outs[8:11]

    def refine_bbox(self, proposal, delta):
        dz = torch.sigmoid(delta[..., 1:3] + inverse_sigmoid(proposal[..., 1:3]))
        theta = proposal[..., 0:1] + (torch.sigmoid(delta[..., 0:1]) * 2 - 1) / self.num_ray
        return torch.cat([theta, dz, delta[..., 3:]], dim=-1)

    fused_refine_train = True   # training on CUDA: refine_bbox + output transform as one forward and one backward launch
    row_programs = True     # inference on CUDA: the row-wise operator chains run as row programs (csrc/rowops.cu)

    def _rows_ok(self, query_feat):
        """The row programs are forward-only fp32 CUDA kernels (dropout must be the identity)."""
        return (self.row_programs and not torch.is_grad_enabled() and not self.training and query_feat.is_cuda
                and query_feat.dtype == torch.float32 and self.embed_dims % 4 == 0)

    train_row_programs = True     # training on CUDA: the same chains with autograd (racformer_b200/rowtrain.py)

    def _train_rows_ok(self, query_feat):
        return (self.train_row_programs and torch.is_grad_enabled() and query_feat.is_cuda and query_feat.dtype == torch.float32
                and self.embed_dims % 4 == 0 and self.embed_dims <= 384
                and not (self.training and self.self_attn.activation_checkpoint))

    def _seed(self):
        """Seeds of the dropout masks of the row chains: a per-layer counter mixed with torch's seed (reproducible runs)."""
        self._drop_counter = getattr(self, "_drop_counter", 0) + 1
        return (torch.initial_seed() * 2654435761 + self._drop_counter * 40503) & 0x7fffffff

    def _self_attn_rows_train(self, query_bbox, query_feat, attn_mask):
        """Training counterpart of _self_attn_rows: two row chains with autograd around the attention core (the
        scale-adaptive mask + attention of ScaleAdaptiveSelfAttention, racformer_transformer.py:296-336)."""
        from . import rowtrain
        mha = self.self_attn.attention.attn
        E, H = self.embed_dims, mha.num_heads
        if not mha._qkv_same_embed_dim or mha.in_proj_bias is None or mha.bias_k is not None or mha.add_zero_attn:
            return None
        B, Q, _ = query_feat.shape
        pe = self.position_encoder
        query_bbox = query_bbox.contiguous()
        p = rowtrain.RowChain(B * Q, width=3 * E, num_bufs=3)
        p.load(0, query_bbox, n=3)
        p.linear(1, 0, pe[0])
        p.layernorm(1, pe[1], relu=True)
        p.linear(2, 1, pe[3])
        p.layernorm(2, pe[4], relu=True)
        p.load(0, query_feat.contiguous())
        p.add(0, 2, E)
        h_qpos = p.store(0, E)
        p.linear(1, 0, mha, weight=mha.in_proj_weight, bias=mha.in_proj_bias)
        h_qkv = p.store(1, 3 * E)
        p.linear(2, 0, self.self_attn.gen_tau)
        h_tau = p.store(2, H)
        outs = p.run()
        qpos, qkv, tau = outs[h_qpos], outs[h_qkv], outs[h_tau]
        att = self.self_attn.attention_core(query_bbox, qkv.view(B, Q, 3 * E), tau.view(B, Q, H), attn_mask, H,
                                            mha.dropout if self.training else 0.0)
        p = rowtrain.RowChain(B * Q, width=E, num_bufs=2)
        p.load(0, att.reshape(B * Q, E))
        p.linear(1, 0, mha.out_proj)
        p.dropout(1, E, self.self_attn.attention.proj_drop.p if self.training else 0.0, self._seed())
        p.load(0, qpos)
        p.add(1, 0, E)
        p.layernorm(1, self.norm1)
        h = p.store(1, E)
        return p.run()[h].view(B, Q, E)

    def _pos_encode_rows(self, query_bbox, query_feat):
        """query_feat + position_encoder(query_bbox[..., :3]) as one launch (Linear-LN-ReLU-Linear-LN-ReLU + add)."""
        from . import rowops
        B, Q, E = query_feat.shape
        pe = self.position_encoder
        p = rowops.RowProgram(B * Q, width=E, num_bufs=3)
        p.load(0, query_bbox.contiguous(), n=3)
        p.linear(1, 0, pe[0])
        p.layernorm(1, pe[1], relu=True)
        p.linear(2, 1, pe[3])
        p.layernorm(2, pe[4], relu=True)
        p.load(0, query_feat.contiguous())
        p.add(0, 2, E)
        out = p.store(0, E)
        p.run()
        return out.view(B, Q, E)

    fused_self_attention = True     # with row_programs: csrc/sasa.cu instead of the [B,8,Q,Q] mask + generic attention

    def _self_attn_rows(self, query_bbox, query_feat):
        """norm1(self_attn(query_bbox, query_feat + position_encoder(query_bbox[..., :3]))) in three launches: a row
        program (position encoder, add, in_proj, gen_tau), the scale-adaptive attention core, a row program (out_proj,
        residual, norm1). None when the attention module is not the plain 32-wide-head configuration."""
        mha = self.self_attn.attention.attn
        E, H = self.embed_dims, mha.num_heads
        if (not self.fused_self_attention or not mha._qkv_same_embed_dim or mha.in_proj_bias is None or E != H * 32
                or mha.bias_k is not None or mha.add_zero_attn or query_bbox.shape[-1] < 2):
            return None
        from . import points, rowops
        B, Q, _ = query_feat.shape
        pe = self.position_encoder
        query_bbox = query_bbox.contiguous()
        p = rowops.RowProgram(B * Q, width=3 * E, num_bufs=3)
        p.load(0, query_bbox, n=3)
        p.linear(1, 0, pe[0])
        p.layernorm(1, pe[1], relu=True)
        p.linear(2, 1, pe[3])
        p.layernorm(2, pe[4], relu=True)
        p.load(0, query_feat.contiguous())
        p.add(0, 2, E)
        qpos = p.store(0, E)
        p.linear(1, 0, mha, weight=mha.in_proj_weight, bias=mha.in_proj_bias)
        qkv = p.store(1, 3 * E)
        p.linear(2, 0, self.self_attn.gen_tau)
        tau = p.store(2, H)
        p.run()
        att = points.sasa_attention(qkv.view(B, Q, 3 * E), tau.view(B, Q, H), query_bbox, self.pc_range, H)
        p = rowops.RowProgram(B * Q, width=E, num_bufs=2)
        p.load(0, att)
        p.linear(1, 0, mha.out_proj)
        p.load(0, qpos)
        p.add(1, 0, E)
        p.layernorm(1, self.norm1)
        out = p.store(1, E)
        p.run()
        return out.view(B, Q, E)

    def _tail_rows(self, mixed, query_feat, radar, lss, train=False):
        """Everything after the sampling ops of an iteration as ONE launch: norm2(mixed); per BEV branch the softmax queue
        fusion, output_proj, residual and norm; cat + fusion + norm_fusion; FFN + norm3; cls and reg branches.
        radar / lss = (MSDA output [B*T,Q,C], queue logits [B,Q,T]). -> (query_feat, cls_score, reg delta).
        train: the same chain with autograd (forward + generated backward row program) and the reference's dropouts."""
        from . import rowops
        B, Q, E = query_feat.shape
        if train:
            from . import rowtrain
            p = rowtrain.RowChain(B * Q, width=3 * E, num_bufs=3)
            drop = (lambda buf, n, prob, col=0: p.dropout(buf, n, prob if self.training else 0.0, self._seed(), col))
        else:
            p = rowops.RowProgram(B * Q, width=3 * E, num_bufs=3)
            drop = (lambda *a, **k: None)
        p.load(0, mixed.contiguous())
        p.layernorm(0, self.norm2)
        p.load(1, query_feat.contiguous())
        for i, (branch, norm, (values, logits)) in enumerate(((self.sampling_radar_bev, self.norm_radar_bev, radar),
                                                               (self.sampling_lss_bev, self.norm_lss_bev, lss))):
            col = (i + 1) * E
            p.load_queue(2, values, logits, Q, branch.attention.num_bev_queue)
            p.linear(0, 2, branch.attention.output_proj, dst_col=col)
            drop(0, E, branch.attention.dropout.p, col)
            p.add(0, 1, E, dst_col=col)
            p.layernorm(0, norm, col=col)
        p.linear(1, 0, self.fusion)
        p.layernorm(1, self.norm_fusion)
        p.linear(2, 1, self.ffn.layers[0][0], relu=True)
        drop(2, self.ffn.layers[0][0].out_features, self.ffn.layers[0][2].p)
        p.linear(0, 2, self.ffn.layers[1])
        drop(0, E, self.ffn.layers[2].p)
        p.add(0, 1, E)
        p.layernorm(0, self.norm3)
        out_feat = p.store(0, E)
        outs = []
        for branch in (self.cls_branch, self.reg_branch):
            mods = list(branch)
            src, i = 0, 0
            free = [1, 2]
            while i < len(mods):
                lin = mods[i]
                i += 1
                dst = free[0] if src != free[0] else free[1]
                relu_now = i < len(mods) and isinstance(mods[i], nn.ReLU)
                p.linear(dst, src, lin, relu=relu_now)
                if relu_now:
                    i += 1
                elif i < len(mods) and isinstance(mods[i], nn.LayerNorm):
                    relu_after = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                    p.layernorm(dst, mods[i], relu=relu_after)
                    i += 2 if relu_after else 1
                src = dst
            outs.append(p.store(src, mods[-1].out_features))
        res = p.run()
        if train:
            out_feat, outs = res[out_feat], [res[h] for h in outs]
        return out_feat.view(B, Q, E), outs[0].view(B, Q, -1), outs[1].view(B, Q, -1)

    def forward(self, ops, query_bbox, query_feat, mlvl_feats, lss_bev_feats, radar_bev_feats, attn_mask, meta, layer=0,
                prepared=None):
        rows = self._rows_ok(query_feat)
        if rows:
            from ._lib import Unsupported
            try:
                return self._forward(ops, query_bbox, query_feat, mlvl_feats, lss_bev_feats, radar_bev_feats, attn_mask,
                                     meta, layer, prepared, True)
            except Unsupported:
                # a fused call-site kernel does not exist for this configuration (embed_dims too wide for the shared-memory
                # row buffers, parameters that are not 16-byte aligned such as views into a flat buffer, ...): the kernels
                # are pure functions of their inputs, so the iteration is simply redone with the PyTorch operator chain
                self.row_programs = False
        return self._forward(ops, query_bbox, query_feat, mlvl_feats, lss_bev_feats, radar_bev_feats, attn_mask, meta,
                             layer, prepared, False)

    bev_pair_launch = True     # inference: the radar and the LSS MSDA of an iteration as one launch

    def _bev_pair(self, ops, query_bbox, query_feat, prep_radar, prep_lss, meta, d_region, heads):
        """(radar, lss) = ((MSDA output [B*T,Q,C], queue logits), ...) as BEVSampling.forward(raw=True) returns them, with the
        two MSDA forwards -- same geometry, different value maps / locations / weights -- in ONE launch
        (racf_msda_forward_pair). None when the configuration does not allow it (the caller runs the branches one by one)."""
        r, l = self.sampling_radar_bev, self.sampling_lss_bev
        if (not self.bev_pair_launch or getattr(ops, "msda_pair", None) is None or prep_radar is None or prep_lss is None
                or torch.is_grad_enabled() or r.num_levels != 1 or l.num_levels != 1):
            return None
        (value_r, hw_r), (value_l, hw_l) = prep_radar, prep_lss
        if (tuple(hw_r) != tuple(hw_l) or value_r.shape != value_l.shape
                or (r.num_frames, r.num_heads, r.num_points, r.depth_num) != (l.num_frames, l.num_heads, l.num_points, l.depth_num)
                or not (_use_fused_points(r, query_bbox, query_feat, value_r) and _use_fused_points(l, query_bbox, query_feat, value_l))
                or r.attention.im2col_step != l.attention.im2col_step):
            return None
        loc_r, aw_r, qw_r = r.fused_point_tensors(query_bbox, query_feat, meta, d_region, heads[0])
        loc_l, aw_l, qw_l = l.fused_point_tensors(query_bbox, query_feat, meta, d_region, heads[1])
        shapes = _const_long((tuple(int(v) for v in hw_r),), value_r.device)
        lsi = _const_long((0,), value_r.device)
        out_r, out_l = ops.msda_pair(value_r, loc_r.contiguous(), aw_r.contiguous(), value_l, loc_l.contiguous(), aw_l.contiguous(),
                                     shapes, lsi, r.attention.im2col_step)

        def logits(branch, qw):
            if not branch.attention.queue_weight:
                return None
            return qw if qw is not None else branch.attention.bev_queue_weight(query_feat)
        return (out_r, logits(r, qw_r)), (out_l, logits(l, qw_l))

    def _forward(self, ops, query_bbox, query_feat, mlvl_feats, lss_bev_feats, radar_bev_feats, attn_mask, meta, layer,
                 prepared, rows):
        d_region = self.d_region_list[layer]
        train_rows = not rows and self._train_rows_ok(query_feat)
        if train_rows:
            attended = self._self_attn_rows_train(query_bbox, query_feat, attn_mask)
        else:
            attended = self._self_attn_rows(query_bbox, query_feat) if rows and attn_mask is None else None
        if attended is not None:
            query_feat = attended
        else:
            if rows:
                query_feat = self._pos_encode_rows(query_bbox, query_feat)
            else:
                query_feat = query_feat + self.position_encoder(query_bbox[..., :3])
            query_feat = self.norm1(self.self_attn(query_bbox, query_feat, attn_mask))
        prep_radar, prep_lss = prepared if prepared is not None else (None, None)
        self._heads_x3 = None
        heads = (self._sampling_heads_train(query_feat) if train_rows else self._sampling_heads(query_feat)) or (None, None, None)
        pair = self._bev_pair(ops, query_bbox, query_feat, prep_radar, prep_lss, meta, d_region, heads) if rows else None
        if pair is not None:
            radar, lss = pair
        else:
            radar = self.sampling_radar_bev(ops, query_bbox, query_feat, radar_bev_feats, meta, d_region=d_region,
                                            prepared=prep_radar, heads=heads[0], raw=rows or train_rows)
            lss = self.sampling_lss_bev(ops, query_bbox, query_feat, lss_bev_feats, meta, d_region=d_region,
                                        prepared=prep_lss, heads=heads[1], raw=rows or train_rows)
        sampled = self.sampling(ops, query_bbox, query_feat, mlvl_feats, meta, d_region=d_region, heads=heads[2])
        mixed = self.mixing(sampled, query_feat, query_x3=self._heads_x3)
        self._heads_x3 = None
        if rows or train_rows:
            query_feat, cls_score, delta = self._tail_rows(mixed, query_feat, radar, lss, train=train_rows)
        else:
            radar, lss = self.norm_radar_bev(radar), self.norm_lss_bev(lss)
            query_feat = self.norm2(mixed)
            query_feat = self.norm_fusion(self.fusion(torch.cat((query_feat, radar, lss), dim=-1)))
            query_feat = self.norm3(self.ffn(query_feat))
            cls_score = self.cls_branch(query_feat)
            delta = self.reg_branch(query_feat)
        time_diff = meta["time_diff"]
        if rows and self.code_size >= 8:
            from . import points   # refine + velocity scaling + output transform: one launch
            bbox_pred, bbox_xy = points.refine_bbox(query_bbox.contiguous(), delta.contiguous(), time_diff.contiguous(),
                                                    self.num_ray)
            return query_feat, cls_score, bbox_pred, bbox_xy
        if (self.fused_refine_train and torch.is_grad_enabled() and self.code_size >= 8 and delta.is_cuda
                and delta.dtype == torch.float32 and query_bbox.dtype == torch.float32):
            from . import training   # the same launch with a backward kernel (csrc/points_train.cu)
            bbox_pred, bbox_xy = training.RefineBBox.apply(query_bbox, delta, time_diff, self.num_ray)
            return query_feat, cls_score, bbox_pred, bbox_xy
        bbox_pred = self.refine_bbox(query_bbox, delta)
        if time_diff.shape[1] > 1:   # relative -> absolute velocity
            td = torch.where(time_diff < 1e-5, torch.ones_like(time_diff), time_diff)
            bbox_pred = torch.cat([bbox_pred[..., :8], bbox_pred[..., 8:] / td[:, 1:2, None]], dim=-1)
        return query_feat, cls_score, bbox_pred, None


def to_sampling_layout(feat, num_cams, num_groups=4):
    """[B, T*N, G*C, H, W] FPN level -> channel-last sampling layout [B*T*G, N, H, W, C] (racformer_transformer.py:112-124)."""
    B, TN, GC, H, W = feat.shape
    N, T, G, C = num_cams, TN // num_cams, num_groups, GC // num_groups
    if C == 64 and feat.is_cuda and feat.dtype == torch.float16 and not (torch.is_grad_enabled() and feat.requires_grad):
        # fp16 FPN outputs (the reference's image branch is fp16 with fp32 outputs, models/racformer.py:106): the cast up
        # to fp32 is fused into the re-layout kernel; same values as `.float()` followed by the fp32 path
        from . import points
        return points.to_sampling_layout_f16(feat, num_cams, num_groups)
    if C == 64 and feat.is_cuda and feat.dtype == torch.float32 and feat.is_contiguous():
        if torch.is_grad_enabled() and feat.requires_grad:
            from . import training   # the same tiled-transpose kernels in both directions, as an autograd Function
            return training.SamplingLayout.apply(feat, num_cams, num_groups)
        from . import points   # tiled-transpose kernel; the PyTorch permute below is the generic path
        return points.to_sampling_layout(feat, num_cams, num_groups)
    return feat.reshape(B, T, N, G, C, H, W).permute(0, 1, 3, 2, 5, 6, 4).reshape(B * T * G, N, H, W, C).contiguous()


class RaCFormerTransformerDecoder(nn.Module):
    def __init__(self, embed_dims, num_frames=8, num_points=4, num_points_bev=4, num_layers=6, num_levels=4,
                 num_classes=10, code_size=10, img_depth_num=3, bev_depth_num=5, pc_range=(), num_ray=150,
                 d_region_list=(0.15, 0.1, 0.1, 0.08, 0.08, 0.05), spatial_shapes=(128, 128), num_cams=6):
        super().__init__()
        self.num_layers, self.pc_range, self.num_cams = num_layers, list(pc_range), num_cams
        self.decoder_layer = RaCFormerTransformerDecoderLayer(   # ONE layer, shared by all iterations
            embed_dims, num_frames, num_points, num_points_bev, num_levels, num_classes, code_size,
            img_depth_num=img_depth_num, bev_depth_num=bev_depth_num, num_ray=num_ray, pc_range=pc_range,
            d_region_list=d_region_list, spatial_shapes=spatial_shapes)

    shared_gradients = True     # training on CUDA: see forward()

    @torch.no_grad()
    def init_weights(self):
        self.decoder_layer.init_weights()

    def build_meta(self, img_metas, batch, device):
        """Per-forward host metadata -> device tensors (racformer_transformer.py:98-110)."""
        stamps = np.array([m["img_timestamp"] for m in img_metas], dtype=np.float64).reshape(batch, -1, self.num_cams)
        time_diff = np.mean(stamps[:, :1, :] - stamps, axis=-1).astype(np.float32)
        lidar2img = np.asarray([m["lidar2img"] for m in img_metas]).astype(np.float32)
        image_h, image_w, _ = img_metas[0]["img_shape"][0]
        return {"time_diff": torch.from_numpy(time_diff).to(device), "lidar2img": torch.from_numpy(lidar2img).to(device),
                "image_h": image_h, "image_w": image_w}

    def forward(self, ops, query_bbox, query_feat, mlvl_feats, lss_bev_feats, radar_bev_feats, attn_mask, img_metas,
                hoist_invariants=True, feats_in_sampling_layout=False):
        meta = img_metas if isinstance(img_metas, dict) else self.build_meta(img_metas, query_bbox.shape[0], query_bbox.device)
        feats = list(mlvl_feats) if feats_in_sampling_layout else [to_sampling_layout(f, self.num_cams) for f in mlvl_feats]
        layer = self.decoder_layer
        if (self.shared_gradients and torch.is_grad_enabled() and query_bbox.is_cuda and getattr(ops, "msmv_grouped", None) is not None
                and not (self.training and layer.sampling.activation_checkpoint)):
            # training on CUDA: the six iterations scatter their feature / value gradients into shared buffers
            # (racformer_b200/training.py); not with activation checkpointing, whose recompute re-runs the forwards
            from . import training
            meta = dict(meta, shared_grads=training.SharedGrads())
        prepared = None
        if hoist_invariants:
            prepared = (layer.sampling_radar_bev.prepare_value(radar_bev_feats),
                        layer.sampling_lss_bev.prepare_value(lss_bev_feats))
        cls_scores, bbox_preds = [], []
        for i in range(self.num_layers):
            query_feat, cls_score, bbox_pred, bbox_xy = layer(ops, query_bbox, query_feat, feats, lss_bev_feats,
                                                              radar_bev_feats, attn_mask, meta, layer=i, prepared=prepared)
            query_bbox = bbox_pred.clone().detach() if bbox_xy is None else bbox_pred   # fused: a fresh no-grad tensor
            cls_scores.append(cls_score)
            bbox_preds.append(bbox_xy if bbox_xy is not None else theta_d2xy_coods(bbox_pred))
        return torch.stack(cls_scores), torch.stack(bbox_preds)


class RaCFormerTransformer(nn.Module):
    """Same constructor and forward signature as the reference's RaCFormerTransformer (racformer_transformer.py:17-58)."""

    def __init__(self, embed_dims, num_frames=8, num_points=4, num_points_bev=4, num_layers=6, num_levels=4,
                 num_classes=10, code_size=10, img_depth_num=3, bev_depth_num=5, pc_range=(), num_ray=150,
                 d_region_list=(0.15, 0.1, 0.1, 0.08, 0.08, 0.05), spatial_shapes=(128, 128), init_cfg=None, num_cams=6,
                 ops=None, hoist_invariants=True):
        super().__init__()
        assert init_cfg is None
        self.embed_dims, self.pc_range, self.num_cams = embed_dims, list(pc_range), num_cams
        self.hoist_invariants = hoist_invariants
        self._ops = ops
        self.decoder = RaCFormerTransformerDecoder(
            embed_dims, num_frames, num_points, num_points_bev, num_layers, num_levels, num_classes, code_size,
            img_depth_num=img_depth_num, bev_depth_num=bev_depth_num, pc_range=pc_range, num_ray=num_ray,
            d_region_list=d_region_list, spatial_shapes=spatial_shapes, num_cams=num_cams)

    @property
    def ops(self):
        if self._ops is None:
            self._ops = SamplingOps()   # binds libracformer_ops.so; raises if it is missing
        return self._ops

    @torch.no_grad()
    def init_weights(self):
        self.decoder.init_weights()

    def load_state_dict(self, *args, **kwargs):
        from .caches import invalidate_weight_caches   # in-place parameter copies under no_grad: drop the operand splits
        invalidate_weight_caches()
        return super().load_state_dict(*args, **kwargs)

    def train(self, mode=True):
        from .caches import invalidate_weight_caches   # EMA / weight-averaging utilities write through .data between modes
        if mode != self.training:
            invalidate_weight_caches()
        return super().train(mode)

    def set_fused_points(self, enabled):
        """Toggle the fused CUDA point-generation / box-refinement kernels, inference and training (eager PyTorch chain
        when False)."""
        for m in self.modules():
            if isinstance(m, (RaCFormerSampling, BEVSampling)):
                m.fused_points = enabled
            elif isinstance(m, RaCFormerTransformerDecoderLayer):
                m.fused_refine_train = enabled

    def set_mixing_precision(self, precision):
        """"bf16x6" (default) / "bf16x9": tcgen05 kernel with exact bf16 operand splitting (fp32-grade, csrc/linear.cu);
        "fp32": cuBLAS SGEMM like the reference; "tf32x3": opt-in cuBLAS TF32 operand split (see _SplitTF32Linear)."""
        assert precision in ("fp32", "tf32x3", "bf16x6", "bf16x9")
        tensor_core = precision.startswith("bf16")
        for m in self.modules():
            if isinstance(m, AdaptiveMixing):
                m.gemm_precision = precision
            elif isinstance(m, BEVSelfAttention):            # value_proj and the stacked sampling heads follow
                m.tensor_core_value_proj = tensor_core
            elif isinstance(m, RaCFormerTransformerDecoderLayer):
                m.stacked_heads = tensor_core

    def set_activation_checkpoint(self, enabled):
        for m in self.modules():
            if hasattr(m, "activation_checkpoint"):
                m.activation_checkpoint = enabled

    def forward(self, query_bbox, query_feat, mlvl_feats, lss_bev_feats, radar_bev_feats, attn_mask, img_metas,
                feats_in_sampling_layout=False):
        cls_scores, bbox_preds = self.decoder(self.ops, query_bbox, query_feat, mlvl_feats, lss_bev_feats,
                                              radar_bev_feats, attn_mask, img_metas,
                                              hoist_invariants=self.hoist_invariants,
                                              feats_in_sampling_layout=feats_in_sampling_layout)
        return torch.nan_to_num(cls_scores), torch.nan_to_num(bbox_preds)


def generate_query_points(num_query=900, num_clusters=6):
    """models/racformer_head.py:68-79: the polar (angle, distance) grid the queries are initialised on."""
    num_angles = num_query // num_clusters
    angles = torch.linspace(0, 1, num_angles + 1)[:-1].view(num_angles, 1).expand(num_angles, num_clusters)
    dists = torch.linspace(0, 1, num_clusters + 2, dtype=torch.float)[1:-1].view(1, num_clusters).expand(num_angles, num_clusters)
    return torch.stack([angles, dists], dim=-1).flatten(0, 1)


def initial_query_bbox(num_query=900, num_clusters=6):
    """models/racformer_head.py:51-66: init_query_bbox weights before training (z=0.5, h=0.2, zero velocity)."""
    q = torch.zeros(num_query, 10)
    q[:, :2] = generate_query_points(num_query, num_clusters)
    q[:, 2] = 0.5
    q[:, 5] = 0.2
    return q
