"""Does a row program's time depend on how many CTAs stream the weights at once? Tail program of one decoder iteration
(csrc/rowops.cu) under a CUDA graph at 8 rows per CTA, for 1 .. 148 CTAs (Q = 8 * CTAs rows). Constant time = every CTA is
bound by its own instruction stream / its SM's ingest rate; growing time = the shared L2 -> SM path (then a cluster with
TMA multicast would pay).    python tools/rowops_cta_scaling.py > gpurun_out/rowops_cta_scaling.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from racformer_b200 import rowops  # noqa: E402
from racformer_b200.decoder import RaCFormerTransformerDecoderLayer  # noqa: E402
from racformer_b200.synthetic import PC_RANGE  # noqa: E402


def timed(fn, iters=50, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


dev = torch.device("cuda", 0)
torch.manual_seed(0)
E, T, B = 256, 8, 1
layer = RaCFormerTransformerDecoderLayer(E, num_frames=T, pc_range=PC_RANGE).to(dev).eval()
layer.init_weights()
res = {}
auto = rowops.choose_rows_per_cta
with torch.no_grad():
    for rpc in (8, 4):
        rowops.choose_rows_per_cta = lambda *a, _r=rpc, **k: _r
        for ctas in (1, 8, 32, 64, 113, 148):
            Q = rpc * ctas
            mixed, qf = torch.randn(B, Q, E, device=dev), torch.randn(B, Q, E, device=dev)
            radar = (torch.randn(B * T, Q, E, device=dev), torch.randn(B, Q, T, device=dev))
            lss = (torch.randn(B * T, Q, E, device=dev), torch.randn(B, Q, T, device=dev))
            g, s = torch.cuda.CUDAGraph(), torch.cuda.Stream()
            with torch.cuda.stream(s):
                layer._tail_rows(mixed, qf, radar, lss)
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=s):
                    keep = layer._tail_rows(mixed, qf, radar, lss)
            res[f"tail_rows{rpc}_{ctas}_ctas_us"] = timed(g.replay)
rowops.choose_rows_per_cta = auto
print(json.dumps(res, indent=1))
