"""TEST/BENCH INFRASTRUCTURE (may use oracle/): time the REFERENCE implementations on the GPU box beside ours.

    python tests/bench_reference_on_gpu.py > gpurun_out/reference_on_b200.json

  * the reference's unmodified CUDA extensions from oracle/_ref (built by `python -m oracle.build_ref`), when present;
  * the reference's PyTorch grid_sample paths (oracle.reference_port) on the GPU and on the host cores.
No reference source is read here; /root/reference is not needed at run time.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import build_ref, reference_port  # noqa: E402
from racformer_b200 import bev_pool, wrapper  # noqa: E402
from racformer_b200.multi_scale_deformable_attn_function import ext_module  # noqa: E402
from racformer_b200.synthetic import make_lss_pool_case  # noqa: E402
from tests.helpers import F8_SHAPES, make_msda_inputs, make_msmv_inputs  # noqa: E402


def gpu_ms(fn, iters=10, warm=3, flush=None):
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
    return sorted(ts)[len(ts) // 2]


def cpu_ms(fn, iters=3):
    fn()
    ts = []
    for _ in range(iters):
        t = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t) * 1e3)
    return min(ts)


def main():
    res = {"gpu": torch.cuda.get_device_name(0), "host_cores": os.cpu_count()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    feats, loc, w, g = make_msmv_inputs(0, Bp=32, N=6, C=64, Q=900, P=12, shapes=F8_SHAPES, lo=-0.1, hi=1.1, device="cuda")
    res["msmv_fwd_ours_us"] = 1e3 * gpu_ms(lambda: wrapper.msmv_forward(feats, loc, w), flush=flush)
    res["msmv_bwd_ours_us"] = 1e3 * gpu_ms(lambda: wrapper.msmv_backward(g, feats, loc, w), flush=flush)
    ext = build_ref.load_prebuilt()
    if ext is not None:
        res["msmv_fwd_reference_ext_us"] = 1e3 * gpu_ms(lambda: ext._ms_deform_attn_cuda_c2345_forward(*feats, loc, w), flush=flush)
        res["msmv_bwd_reference_ext_us"] = 1e3 * gpu_ms(lambda: ext._ms_deform_attn_cuda_c2345_backward(g, *feats, loc, w), flush=flush)
    feats_cf = [f.permute(0, 4, 1, 2, 3).contiguous() for f in feats]
    res["msmv_fwd_torch_gridsample_gpu_us"] = 1e3 * gpu_ms(lambda: reference_port.msmv_sampling_torch(feats_cf, loc, w), iters=5, flush=flush)
    torch.set_num_threads(os.cpu_count())
    fc, lc, wc = [f.cpu() for f in feats_cf], loc.cpu(), w.cpu()
    res["msmv_fwd_torch_gridsample_cpu_ms"] = cpu_ms(lambda: reference_port.msmv_sampling_torch(fc, lc, wc))
    del feats, feats_cf, fc
    torch.cuda.empty_cache()

    value, sp, lsi, mloc, aw, mg = make_msda_inputs(0, B=8, M=4, D=64, Q=900, P=20, shapes=[(128, 128)], device="cuda")
    res["msda_fwd_ours_us"] = 1e3 * gpu_ms(lambda: ext_module.ms_deform_attn_forward(value, sp, lsi, mloc, aw, im2col_step=64), flush=flush)
    res["msda_fwd_torch_gridsample_gpu_us"] = 1e3 * gpu_ms(lambda: reference_port.msda_torch(value, [(128, 128)], mloc, aw), iters=5, flush=flush)
    vc, lc2, ac = value.cpu(), mloc.cpu(), aw.cpu()
    res["msda_fwd_torch_gridsample_cpu_ms"] = cpu_ms(lambda: reference_port.msda_torch(vc, [(128, 128)], lc2, ac))

    pc = make_lss_pool_case(0, 1, 6, 96, 16, 44, 256, (128, 128), device="cuda")
    out = torch.zeros(pc["shape"], device="cuda")
    og = torch.randn(pc["shape"], device="cuda")
    rd, rf, rb, st, ln = bev_pool.backward_intervals(pc["ranks_depth"], pc["ranks_feat"], pc["ranks_bev"], stable=True)
    dg, fg = torch.zeros_like(pc["depth"]), torch.zeros_like(pc["feat"])
    fwd_args = (pc["depth"], pc["feat"], out, pc["ranks_depth"], pc["ranks_feat"], pc["ranks_bev"], pc["lengths"], pc["starts"])
    bwd_args = (og, dg, fg, pc["depth"], pc["feat"], rd, rf, rb, ln, st)
    res["bev_pool_fwd_ours_us"] = 1e3 * gpu_ms(lambda: bev_pool.bev_pool_v2_forward(*fwd_args), flush=flush)
    res["bev_pool_bwd_ours_us"] = 1e3 * gpu_ms(lambda: bev_pool.bev_pool_v2_backward(*bwd_args), flush=flush)
    pext = build_ref.load_prebuilt_bev_pool()
    if pext is not None:
        res["bev_pool_fwd_reference_ext_us"] = 1e3 * gpu_ms(lambda: pext.bev_pool_v2_forward(*fwd_args), flush=flush)
        res["bev_pool_bwd_reference_ext_us"] = 1e3 * gpu_ms(lambda: pext.bev_pool_v2_backward(*bwd_args), iters=3, warm=1, flush=flush)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
