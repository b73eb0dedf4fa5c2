"""Seeded synthetic inputs for the decoder harness: a 6-camera ring, timestamps,
FPN / BEV features and the polar query grid. There is no dataset or checkpoint in this environment."""
import math
import zlib

import numpy as np
import torch

from .decoder import initial_query_bbox

PC_RANGE = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]        # configs/racformer_r50_nuimg_704x256_f8.py:23
D_REGION_LIST = [0.08, 0.07, 0.06, 0.05, 0.04, 0.03]    # :36
CAMERA_YAWS_DEG = (55.0, 0.0, -55.0, 110.0, 180.0, -110.0)
F8_LEVEL_SHAPES = [(64, 176), (32, 88), (16, 44), (8, 22)]


def lidar2img_matrix(yaw_rad, focal=557.0, cx=352.0, cy=128.0, cam_height=1.5):
    """Pinhole camera looking along `yaw` in the lidar frame (x fwd, y left, z up) -> 4x4 lidar-to-pixel matrix."""
    c, s = math.cos(yaw_rad), math.sin(yaw_rad)
    rot = np.array([[s, -c, 0.0], [0.0, 0.0, -1.0], [c, s, 0.0]], dtype=np.float64)   # rows: camera x, y, z axes
    ext = np.eye(4)
    ext[:3, :3] = rot
    ext[:3, 3] = -rot @ np.array([0.0, 0.0, cam_height])
    intr = np.eye(4)
    intr[0, 0] = intr[1, 1] = focal
    intr[0, 2], intr[1, 2] = cx, cy
    return intr @ ext


def make_img_metas(batch=1, num_frames=8, num_cams=6, image_hw=(256, 704), frame_dt=0.5, focal=557.0):
    """Frame-major, cameras inner (loaders/pipelines/loading.py:662-678); same rig for every frame."""
    h, w = image_hw
    yaws = CAMERA_YAWS_DEG[:num_cams] if num_cams <= len(CAMERA_YAWS_DEG) else \
        tuple(360.0 * i / num_cams for i in range(num_cams))
    mats = [lidar2img_matrix(math.radians(a), focal=focal * w / 704.0, cx=w / 2.0, cy=h / 2.0) for a in yaws] * num_frames
    meta = dict(lidar2img=mats, img_timestamp=[-frame_dt * (i // num_cams) for i in range(num_frames * num_cams)],
                img_shape=[(h, w, 3)] * (num_frames * num_cams))
    return [dict(meta) for _ in range(batch)]


def make_decoder_inputs(seed=0, batch=1, num_frames=8, num_cams=6, embed_dims=256, num_query=900, num_clusters=6,
                        level_shapes=None, bev_hw=(128, 128), image_hw=(256, 704), device="cpu"):
    """Returns dict(query_bbox, query_feat, mlvl_feats, lss_bev, radar_bev, img_metas)."""
    level_shapes = level_shapes or F8_LEVEL_SHAPES
    g = torch.Generator().manual_seed(seed)
    feats = [torch.randn(batch, num_frames * num_cams, embed_dims, h, w, generator=g) for h, w in level_shapes]
    lss = torch.randn(batch, num_frames, embed_dims, *bev_hw, generator=g)
    radar = torch.randn(batch, num_frames, embed_dims, *bev_hw, generator=g)
    qb = initial_query_bbox(num_query, num_clusters)[None].repeat(batch, 1, 1)
    qf = torch.randn(batch, num_query, embed_dims, generator=g) * 0.1
    mv = lambda t: t.to(device)
    return dict(query_bbox=mv(qb), query_feat=mv(qf), mlvl_feats=[mv(f) for f in feats], lss_bev=mv(lss),
                radar_bev=mv(radar), img_metas=make_img_metas(batch, num_frames, num_cams, image_hw))


@torch.no_grad()
def fill_parameters_by_name(module, seed=0):
    """Deterministic weights that depend only on (seed, parameter name, shape): two module trees with the same
    parameter names -- this repo's decoder and the reference's -- get bit-identical weights without a checkpoint."""
    for name, p in sorted(module.state_dict().items()):
        if not p.is_floating_point():
            continue
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))
        if p.dim() >= 2:
            fan_in = p[0].numel()
            val = torch.randn(p.shape, generator=g) / math.sqrt(fan_in)
            if name.endswith("embed.weight"):
                val = torch.rand(p.shape, generator=g)
        elif name.endswith("weight"):           # LayerNorm scale
            val = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
        else:                                   # biases
            val = 0.1 * torch.randn(p.shape, generator=g)
        if "sampling_offset.bias" in name:      # spread the sample points like the reference's init does
            val = torch.rand(p.shape, generator=g) - 0.5
        p.copy_(val.to(p.dtype))


def make_lss_pool_case(seed, B, N, D, H, W, C, bev=(128, 128), device="cpu"):
    """Random frustum-to-BEV assignment shaped like voxel_pooling_prepare_v2 (models/necks/view_transformer_racformer.py:
    202-260) for BEVPoolv2: every (b,n,d,h,w) point gets a random BEV cell or falls outside the grid (~23 %); kept points
    are sorted by cell and cut into intervals."""
    g = torch.Generator().manual_seed(seed)
    n_pts = B * N * D * H * W
    depth = torch.rand(B, N, D, H, W, generator=g)
    feat = torch.randn(B, N, H, W, C, generator=g)
    ranks_depth = torch.arange(n_pts, dtype=torch.int32)
    ranks_feat = torch.arange(n_pts // D, dtype=torch.int32).reshape(B, N, 1, H, W).expand(B, N, D, H, W).flatten()
    cell = torch.randint(0, int(bev[0] * bev[1] * 1.3), (n_pts,), generator=g)
    batch = torch.arange(B).reshape(B, 1).expand(B, n_pts // B).flatten()
    kept = cell < bev[0] * bev[1]
    ranks_bev = (batch * bev[0] * bev[1] + cell)[kept].int()
    ranks_depth, ranks_feat = ranks_depth[kept], ranks_feat[kept].contiguous()
    order = torch.argsort(ranks_bev.long(), stable=True)
    ranks_bev, ranks_depth, ranks_feat = ranks_bev[order], ranks_depth[order], ranks_feat[order]
    keep = torch.ones(ranks_bev.shape[0], dtype=torch.bool)
    keep[1:] = ranks_bev[1:] != ranks_bev[:-1]
    starts = torch.where(keep)[0].int()
    lengths = torch.zeros_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = ranks_bev.shape[0] - starts[-1]
    mv = lambda t: t.to(device)
    return dict(depth=mv(depth), feat=mv(feat), ranks_depth=mv(ranks_depth), ranks_feat=mv(ranks_feat),
                ranks_bev=mv(ranks_bev), starts=mv(starts), lengths=mv(lengths), shape=(B, 1, bev[1], bev[0], C))


def make_op_inputs(case="allvalid", device="cuda", num_views=6, batch=1, num_query=900):
    """Config 1: the op-level inputs of both sampling ops at the f8 shapes, seeded on the device so that
    bench.py, tools/op_timing.py and the ncu captures under profiles/ all see the same tensors.
    case "allvalid": MSMV xy ~ U(0,1), MSDA xy ~ U(0,1); "mixed": U(-0.1,1.1) (~75 % of corners valid) / U(-0.05,1.05)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(0)
    Bp, N, C, Q, P = 32 * batch, num_views, 64, num_query, 12
    lo, hi = (0.0, 1.0) if case == "allvalid" else (-0.1, 1.1)
    feats = [torch.randn(Bp, N, h, w, C, device=dev, generator=g) for h, w in F8_LEVEL_SHAPES]
    xy = torch.rand(Bp, Q, P, 2, device=dev, generator=g) * (hi - lo) + lo
    view = torch.randint(0, N, (Bp, Q, P, 1), device=dev, generator=g).float() / (N - 1)
    loc = torch.cat([xy, view], -1).contiguous()
    w = torch.softmax(torch.randn(Bp, Q, P, 4, device=dev, generator=g), -1).contiguous()
    grad_out = torch.randn(Bp, Q, C, P, device=dev, generator=g)
    B, S, M, D, MP = 8 * batch, 128 * 128, 4, 64, 20
    mlo, mhi = (0.0, 1.0) if case == "allvalid" else (-0.05, 1.05)
    value = torch.randn(B, S, M, D, device=dev, generator=g)
    mloc = (torch.rand(B, Q, M, 1, MP, 2, device=dev, generator=g) * (mhi - mlo) + mlo).contiguous()
    aw = torch.softmax(torch.randn(B, Q, M, 1, MP, device=dev, generator=g), -1).contiguous()
    mgrad = torch.randn(B, Q, M * D, device=dev, generator=g)
    sp = torch.tensor([[128, 128]], dtype=torch.long, device=dev)
    lsi = torch.tensor([0], dtype=torch.long, device=dev)
    return dict(feats=feats, loc=loc, w=w, g=grad_out, value=value, sp=sp, lsi=lsi, mloc=mloc, aw=aw, mg=mgrad,
                level_shapes=list(F8_LEVEL_SHAPES), num_views=N)
