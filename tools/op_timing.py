"""Time the four hot-path kernels at the racformer_r50_nuimg_704x256_f8 shapes (CUDA events, L2 flushed between
iterations) and print one JSON object. Also the short command profiled with ncu (see profiles/).

    python tools/op_timing.py [--iters 20] [--case allvalid|mixed] [--ops msmv_fwd,msmv_bwd,msda_fwd,msda_bwd]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from racformer_b200 import wrapper  # noqa: E402
from racformer_b200.multi_scale_deformable_attn_function import ext_module  # noqa: E402
from racformer_b200.roofline import msda_bytes, msmv_bytes  # noqa: E402

F8_SHAPES = [(64, 176), (32, 88), (16, 44), (8, 22)]


def make_inputs(case):
    from racformer_b200.synthetic import make_op_inputs
    return make_op_inputs(case, "cuda")


def time_op(fn, iters, warm, flush):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return {"best_ms": ts[0], "median_ms": ts[len(ts) // 2], "mean_ms": sum(ts) / len(ts)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--case", default="allvalid")
    ap.add_argument("--ops", default="msmv_fwd,msmv_bwd,msda_fwd,msda_bwd,bev_pool_fwd,bev_pool_bwd")
    ap.add_argument("--no-flush", action="store_true")
    args = ap.parse_args()
    d = make_inputs(args.case)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    peaks = {"hbm_gbs": 6541.8}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    gv = torch.empty_like(d["value"])
    gl, ga = torch.empty_like(d["mloc"]), torch.empty_like(d["aw"])

    def msda_bwd():
        gv.zero_()
        ext_module.ms_deform_attn_backward(d["value"], d["sp"], d["lsi"], d["mloc"], d["aw"], d["mg"], gv, gl, ga,
                                           im2col_step=64)

    ops = {
        "msmv_fwd": lambda: wrapper.msmv_forward(d["feats"], d["loc"], d["w"]),
        "msmv_bwd": lambda: wrapper.msmv_backward(d["g"], d["feats"], d["loc"], d["w"]),
        "msda_fwd": lambda: ext_module.ms_deform_attn_forward(d["value"], d["sp"], d["lsi"], d["mloc"], d["aw"], im2col_step=64),
        "msda_bwd": msda_bwd,
    }
    # BEVPoolv2 at the LSS shapes of the f8 config: 6 cams x 96 depth bins x 16x44, C = 256, 128x128 BEV
    from racformer_b200 import bev_pool
    from racformer_b200.synthetic import make_lss_pool_case
    pc = make_lss_pool_case(0, 1, 6, 96, 16, 44, 256, (128, 128), device="cuda")
    p_out = torch.zeros(pc["shape"], device="cuda")
    p_og = torch.randn(pc["shape"], device="cuda")
    prd, prf, prb, pst, pln = bev_pool.backward_intervals(pc["ranks_depth"], pc["ranks_feat"], pc["ranks_bev"], stable=True)
    p_dg, p_fg = torch.zeros_like(pc["depth"]), torch.zeros_like(pc["feat"])
    ops["bev_pool_fwd"] = lambda: bev_pool.bev_pool_v2_forward(pc["depth"], pc["feat"], p_out, pc["ranks_depth"], pc["ranks_feat"],
                                                               pc["ranks_bev"], pc["lengths"], pc["starts"])
    ops["bev_pool_bwd"] = lambda: bev_pool.bev_pool_v2_backward(p_og, p_dg, p_fg, pc["depth"], pc["feat"], prd, prf, prb, pln, pst)
    npts = pc["ranks_bev"].numel()
    pool_bytes = npts * (256 * 4 + 4 + 12) + pc["starts"].numel() * (256 * 4 + 8)
    _, mask = wrapper.msmv_tap_masks(F8_SHAPES, d["loc"], 6)
    from racformer_b200.multi_scale_deformable_attn_function import msda_tap_masks
    mmask = msda_tap_masks(d["sp"], d["mloc"])
    algo = {}
    algo["msmv_fwd"], algo["msmv_bwd"] = msmv_bytes(mask, C=64, L=4, feat_bytes=sum(f.numel() * 4 for f in d["feats"]))
    algo["msda_fwd"], algo["msda_bwd"] = msda_bytes(mmask, D=64, value_bytes=d["value"].numel() * 4)
    # per point: feature row (C*4) + depth + 3 ranks; per interval: output row + start/length. Backward reads out_grad rows.
    algo["bev_pool_fwd"] = algo["bev_pool_bwd"] = pool_bytes
    res = {"case": args.case, "gpu": torch.cuda.get_device_name(0), "l2_flush": not args.no_flush, "peak_hbm_gbs": peaks["hbm_gbs"]}
    for name in args.ops.split(","):
        t = time_op(ops[name], args.iters, args.warmup, flush)
        gbs = algo[name] / (t["median_ms"] * 1e-3) / 1e9
        res[name] = dict(t, algo_bytes=algo[name], gbs=gbs, frac_of_measured_peak=gbs / peaks["hbm_gbs"])
    print(json.dumps(res))


if __name__ == "__main__":
    main()
