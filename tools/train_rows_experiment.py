"""Training step time with the row programs' automatic rows-per-CTA choice vs the fixed 8 rows of round 1.

    python tools/train_rows_experiment.py [workload] > gpurun_out/train_rows.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_workloads  # noqa: E402
from racformer_b200 import rowops  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "decoder_train_f8"
wl = bench_workloads.build(name, torch.device("cuda", 0))


def timed(steps=5, warmup=2):
    for _ in range(warmup):
        wl.step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        wl.step()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


auto = rowops.choose_rows_per_cta
res = {"workload": name, "auto_choice_2440_rows": {f"{nb}x{w}": auto(2440, nb, w, "cuda:0") for nb, w in ((3, 768), (2, 256), (2, 1024))}}
res["auto_ms"] = timed()
for r in (8, 7, 6, 5):
    rowops.choose_rows_per_cta = lambda rows, nb, w, dev=None, _r=r: _r if _r * (nb * w * 4 + 3072) <= 98 * 1024 else 4
    res[f"rows{r}_ms"] = timed()
rowops.choose_rows_per_cta = auto
res["auto_again_ms"] = timed()
print(json.dumps(res, indent=1))
