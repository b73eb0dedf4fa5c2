"""Restatement of the reference's own PyTorch (grid_sample) implementations of the two ops.
TEST INFRASTRUCTURE and the CPU baseline of bench.py; never imported by racformer_b200/.

  msmv_sampling_torch   follows models/csrc/wrapper.py:15-39 (`msmv_sampling_pytorch`): the dispatcher's fallback
                        (wrapper.py:152-153). Features are CHANNEL-FIRST [B', C, N, H, W] there.
  msda_torch            follows mmcv-full 1.6.0 `multi_scale_deformable_attn_pytorch`
                        (mmcv/ops/multi_scale_deform_attn.py), BEVSelfAttention's own CPU fallback
                        (models/bev_self_attention.py:202-204). mmcv is third-party and absent from
                        /root/reference; the algorithm restated is the published Deformable-DETR one and is
                        cross-checked in tests/golden/make_golden.py against the arithmetic-identical copy in
                        transformers.models.mask2former.

Validated against the real reference by tests/golden/make_golden.py (which imports /root/reference) and, on
every run, against the committed fixtures by tests/test_oracle.py.
"""
import torch
import torch.nn.functional as F


def msmv_sampling_torch(mlvl_feats, sampling_locations, scale_weights):
    """mlvl_feats[l]: [B, C, N, H_l, W_l]; sampling_locations [B, Q, P, 3] in [0,1]; scale_weights [B, Q, P, L].

    Each level is sampled with a 3-D (view, y, x) trilinear grid_sample, zero padding, align_corners=True,
    scaled by its weight and accumulated in level order. Returns [B, Q, C, P].
    """
    num_levels = len(mlvl_feats)
    if scale_weights.shape[-1] != num_levels:
        raise AssertionError("one scale weight per level is required")
    B, C = mlvl_feats[0].shape[:2]
    Q, P = sampling_locations.shape[1:3]
    grid = (2.0 * sampling_locations - 1.0).unsqueeze(3)            # [B, Q, P, 1, 3] in [-1, 1]
    total = torch.zeros(B, C, Q, P, dtype=mlvl_feats[0].dtype, device=mlvl_feats[0].device)
    for level in range(num_levels):
        sampled = F.grid_sample(mlvl_feats[level], grid, mode="bilinear", padding_mode="zeros",
                                align_corners=True).squeeze(-1)     # [B, C, Q, P]
        total = total + sampled * scale_weights[:, :, :, level].unsqueeze(1)
    return total.transpose(1, 2)                                    # [B, Q, C, P]


def msmv_sampling_torch_channel_last(feats_cl, sampling_locations, scale_weights):
    """Same op for the CUDA layout [B', N, H, W, C] (what racformer_transformer.py:117-119 produces)."""
    return msmv_sampling_torch([f.permute(0, 4, 1, 2, 3) for f in feats_cl], sampling_locations, scale_weights)


def msmv_sampling_v2_torch(mlvl_feats, sampling_locations, scale_weights):
    """wrapper.py:41-76: keep only the level with the largest scale weight (unweighted). Channel-first feats."""
    grid = (2.0 * sampling_locations - 1.0).unsqueeze(3)
    per_level = [F.grid_sample(f, grid, mode="bilinear", padding_mode="zeros", align_corners=True).squeeze(-1)
                 for f in mlvl_feats]
    stacked = torch.stack(per_level, dim=-1)                         # [B, C, Q, P, L]
    best = scale_weights.argmax(dim=-1)                              # [B, Q, P]
    C = stacked.shape[1]
    idx = best[:, None, :, :, None].expand(-1, C, -1, -1, 1)
    return stacked.gather(-1, idx).squeeze(-1).transpose(1, 2)       # [B, Q, C, P]


def msda_torch(value, spatial_shapes, sampling_locations, attention_weights):
    """value [B, S, M, D]; spatial_shapes: iterable of (H, W); sampling_locations [B, Q, M, L, P, 2] in [0,1];
    attention_weights [B, Q, M, L, P]. Bilinear, zero padding, align_corners=False. Returns [B, Q, M*D]."""
    B, _, M, D = value.shape
    _, Q, _, L, P, _ = sampling_locations.shape
    shapes = [(int(h), int(w)) for h, w in (spatial_shapes.tolist() if torch.is_tensor(spatial_shapes)
                                            else spatial_shapes)]
    per_level_value = value.split([h * w for h, w in shapes], dim=1)
    grids = 2.0 * sampling_locations - 1.0
    sampled = []
    for level, (h, w) in enumerate(shapes):
        # [B, H*W, M, D] -> [B*M, D, H, W]
        v = per_level_value[level].permute(0, 2, 3, 1).reshape(B * M, D, h, w)
        # [B, Q, M, P, 2] -> [B*M, Q, P, 2]
        g = grids[:, :, :, level].permute(0, 2, 1, 3, 4).reshape(B * M, Q, P, 2)
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    stacked = torch.stack(sampled, dim=-2).reshape(B * M, D, Q, L * P)         # [B*M, D, Q, L*P]
    weights = attention_weights.permute(0, 2, 1, 3, 4).reshape(B * M, 1, Q, L * P)
    out = (stacked * weights).sum(-1)                                           # [B*M, D, Q]
    return out.reshape(B, M * D, Q).transpose(1, 2).contiguous()
