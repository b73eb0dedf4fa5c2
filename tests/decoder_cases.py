"""Reduced decoder configurations shared by the fixture generator and the tests (small enough for CPU + a fixture)."""
import torch

from racformer_b200.synthetic import D_REGION_LIST, PC_RANGE, make_decoder_inputs

SMALL = dict(embed_dims=256, num_frames=2, num_points=4, num_points_bev=4, num_layers=2, num_levels=4, num_classes=10,
             code_size=10, img_depth_num=3, bev_depth_num=5, pc_range=PC_RANGE, num_ray=4,
             d_region_list=D_REGION_LIST, spatial_shapes=(16, 16), num_cams=3)
SMALL_INPUTS = dict(batch=1, num_frames=2, num_cams=3, num_query=24, num_clusters=6,
                    level_shapes=[(8, 22), (4, 11), (2, 6), (1, 3)], bev_hw=(16, 16), image_hw=(32, 88))


def small_inputs(seed=0, device="cpu", batch=1):
    kw = dict(SMALL_INPUTS)
    kw["batch"] = batch
    d = make_decoder_inputs(seed=seed, device=device, **kw)
    # give the queries some velocity and spread so the temporal warp and all cameras are exercised
    g = torch.Generator().manual_seed(seed + 99)
    d["query_bbox"][..., 8:10] = (torch.randn(d["query_bbox"][..., 8:10].shape, generator=g) * 0.5).to(device)
    d["query_bbox"][..., 3:6] = (torch.randn(d["query_bbox"][..., 3:6].shape, generator=g) * 0.3 + 0.5).to(device)
    return d


def cpu_oracle_ops():
    """SamplingOps backed by the oracle's torch ports (CPU): lets the decoder's HOST logic be compared with the reference."""
    from oracle import reference_port
    from racformer_b200.decoder import SamplingOps

    def msda(value, shapes, lsi, loc, aw, im2col_step):
        return reference_port.msda_torch(value, shapes, loc, aw)

    return SamplingOps(msmv=reference_port.msmv_sampling_torch_channel_last, msda=msda)
