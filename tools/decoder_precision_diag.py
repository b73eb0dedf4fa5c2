"""Diagnostic: full-size (f8, 2 iterations) decoder on the CUDA path with the Linear layers as cuBLAS SGEMM / bf16x6 /
bf16x9, each against the CPU reference path and against each other: fraction of outputs outside rtol = atol = 1e-3 and
the number of queries they belong to (border-flip outliers, see tests/test_decoder.py)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from racformer_b200.decoder import RaCFormerTransformer  # noqa: E402
from racformer_b200.synthetic import D_REGION_LIST, PC_RANGE, fill_parameters_by_name, make_decoder_inputs  # noqa: E402
from tests.decoder_cases import cpu_oracle_ops  # noqa: E402

cfg = dict(embed_dims=256, num_frames=8, num_points=4, num_points_bev=4, num_layers=2, num_levels=4, num_classes=10,
           code_size=10, img_depth_num=3, bev_depth_num=5, pc_range=PC_RANGE, num_ray=150,
           d_region_list=D_REGION_LIST, spatial_shapes=(128, 128), num_cams=6)
torch.manual_seed(0)
cpu_model = RaCFormerTransformer(**cfg, ops=cpu_oracle_ops(), hoist_invariants=False)
cpu_model.init_weights()
fill_parameters_by_name(cpu_model, seed=1)
cpu_model.eval()
gpu_model = RaCFormerTransformer(**cfg)
gpu_model.load_state_dict(cpu_model.state_dict())
gpu_model.eval().cuda()
torch.backends.cudnn.allow_tf32 = False
res = {}
for seed in (12, 13):
    d = make_decoder_inputs(seed=seed)
    g = torch.Generator().manual_seed(5)
    d["query_bbox"][..., 8:10] = torch.randn(1, 900, 2, generator=g) * 0.5
    with torch.no_grad():
        ref = cpu_model(d["query_bbox"], d["query_feat"], d["mlvl_feats"], d["lss_bev"], d["radar_bev"], None, d["img_metas"])
    dg = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items() if k != "mlvl_feats"}
    dg["mlvl_feats"] = [f.cuda() for f in d["mlvl_feats"]]
    outs = {}
    for prec in ("fp32", "bf16x6", "bf16x9"):
        gpu_model.set_mixing_precision(prec)
        with torch.no_grad():
            outs[prec] = [o.cpu().double() for o in gpu_model(dg["query_bbox"], dg["query_feat"], dg["mlvl_feats"], dg["lss_bev"],
                                                               dg["radar_bev"], None, dg["img_metas"])]

    def cmp(a, b):
        bad = (a - b).abs() > 1e-3 + 1e-3 * b.abs()
        return {"frac": float(bad.float().mean()), "queries": int(bad.any(-1).sum()), "worst": float((a - b).abs().max()),
                "median": float((a - b).abs().median())}
    r = {}
    for prec in outs:
        r[prec + "_vs_cpu"] = {"cls": cmp(outs[prec][0], ref[0].double()), "box": cmp(outs[prec][1], ref[1].double())}
    r["bf16x6_vs_fp32_gpu"] = {"cls": cmp(outs["bf16x6"][0], outs["fp32"][0]), "box": cmp(outs["bf16x6"][1], outs["fp32"][1])}
    r["bf16x9_vs_bf16x6_gpu"] = {"cls": cmp(outs["bf16x9"][0], outs["bf16x6"][0]), "box": cmp(outs["bf16x9"][1], outs["bf16x6"][1])}
    res[f"seed{seed}"] = r
print(json.dumps(res, indent=1))
