"""ncu target: ONE eager decoder forward of bench.py's default workload (decoder_forward_f8, same seeded inputs and
weights), bracketed by cudaProfilerStart/Stop after a warm-up forward, so that `ncu --profile-from-start off -k regex:...`
captures the step's kernels on the decoder's own sampling locations.

    python tools/decoder_ncu_target.py [workload]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "decoder_forward_f8"
wl = bench_workloads.build(name, torch.device("cuda", 0), seed=0)
wl.use_graph = False
wl.step()
wl.step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
wl.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
