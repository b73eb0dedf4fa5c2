"""torch.profiler breakdown of one eager decoder forward by ATen op and input shape (which PyTorch ops the GPU time
outside this library's kernels belongs to). Usage: python tools/profile_decoder_ops.py [workload] > out.json"""
import json
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "decoder_forward_f8"
wl = bench_workloads.build(name, torch.device("cuda", 0))
wl.use_graph = False
for _ in range(3):
    wl.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    wl.step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    t = e.self_device_time_total
    if t > 0:
        rows.append({"op": e.key, "shapes": str(e.input_shapes)[:160], "self_device_us": t, "calls": e.count})
rows.sort(key=lambda r: -r["self_device_us"])
total = sum(r["self_device_us"] for r in rows)
print(json.dumps({"workload": name, "total_device_ms": total / 1e3, "ops": rows[:70]}, indent=1))
