"""What launches the PyTorch (aten) kernels of a workload step: torch.profiler with shapes, aten ops grouped by (op, input
shapes) -- the shapes identify the call site -- with the GPU time of the kernels they launch.

    python tools/profile_callsites.py decoder_train_f8 > gpurun_out/callsites.json"""
import collections
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "decoder_train_f8"
wl = bench_workloads.build(name, torch.device("cuda", 0))
wl.use_graph = False
for _ in range(3):
    wl.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    wl.step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_type.name != "CPU" or not e.name.startswith("aten::"):
        continue
    # only leaf-ish ops that own kernels directly
    t = sum(k.duration for k in e.kernels) if getattr(e, "kernels", None) else 0.0
    if t <= 0:
        continue
    frame = str(getattr(e, "input_shapes", "?"))[:110]
    d = agg[(e.name, frame)]
    d[0] += t
    d[1] += len(e.kernels)
rows = sorted(agg.items(), key=lambda kv: -kv[1][0])
print(json.dumps({"workload": name, "total_aten_gpu_ms": sum(v[0] for v in agg.values()) / 1e3,
                  "callsites": [{"op": k[0], "shapes": k[1], "gpu_ms": v[0] / 1e3, "kernels": v[1]} for k, v in rows[:70]]}, indent=1))
