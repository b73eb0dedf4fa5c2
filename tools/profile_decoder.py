"""torch.profiler breakdown of one decoder forward (GPU kernel time by kernel, and total GPU-busy vs step time)."""
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
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    wl.step()
b.record()
torch.cuda.synchronize()
step_ms = a.elapsed_time(b) / 5
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    wl.step()
    torch.cuda.synchronize()
rows = []
total = 0.0
for e in prof.key_averages():
    t = getattr(e, "device_time_total", 0) or getattr(e, "cuda_time_total", 0)
    if e.device_type.name == "CUDA" or (t and e.key.startswith(("void", "sm", "cutlass", "ampere", "cudnn", "racf", "nvjet", "Memcpy", "Memset", "msda", "msmv"))):
        pass
evs = [e for e in prof.events() if e.device_type.name == "CUDA"]
agg = {}
for e in evs:
    k = e.name[:160]
    d = agg.setdefault(k, [0.0, 0])
    d[0] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    d[1] += 1
total = sum(v[0] for v in agg.values())
top = sorted(agg.items(), key=lambda kv: -kv[1][0])[:90]
single = sorted(evs, key=lambda e: -(e.device_time if hasattr(e, "device_time") else e.cuda_time))[:30]
hist = {"<3us": 0, "3-10us": 0, "10-50us": 0, "50-200us": 0, ">200us": 0}
hist_ms = dict.fromkeys(hist, 0.0)
for e in evs:
    t = e.device_time if hasattr(e, "device_time") else e.cuda_time
    k = "<3us" if t < 3 else "3-10us" if t < 10 else "10-50us" if t < 50 else "50-200us" if t < 200 else ">200us"
    hist[k] += 1
    hist_ms[k] += t / 1e3
def family(n):
    if "racf::" in n:
        return "ours (libracformer_ops.so)"
    if n.startswith(("Memcpy", "Memset")):
        return "memcpy / memset"
    if "at::native" in n or "at_cuda_detail" in n or "Optimizer" in n:
        return "pytorch (aten elementwise / reduce / copy kernels)"
    return "library (cuDNN / cuBLAS / CUTLASS / NCCL)"


shares = {}
for k, v in agg.items():
    f = shares.setdefault(family(k), {"ms": 0.0, "launches": 0})
    f["ms"] += v[0] / 1e3
    f["launches"] += v[1]
for f in shares.values():
    f["share_of_gpu_time"] = f["ms"] / (total / 1e3)
ours = {}
for k, v in agg.items():
    if "racf::" in k:
        short = k.split("racf::")[1].split("(")[0].split("<")[0]
        o = ours.setdefault(short, {"ms": 0.0, "launches": 0})
        o["ms"] += v[0] / 1e3
        o["launches"] += v[1]
print(json.dumps({"workload": name, "launch_shares": shares,
                  "own_kernels": dict(sorted(ours.items(), key=lambda kv: -kv[1]["ms"])), "step_ms": step_ms, "gpu_busy_ms": total / 1e3, "num_gpu_kernels": sum(v[1] for v in agg.values()),
                  "duration_histogram_calls": hist, "duration_histogram_ms": hist_ms,
                  "largest_single_kernels": [{"name": e.name[:70], "us": (e.device_time if hasattr(e, "device_time") else e.cuda_time)} for e in single],
                  "top": [{"name": k, "ms": v[0] / 1e3, "calls": v[1]} for k, v in top]}, indent=1))
