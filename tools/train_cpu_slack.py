"""How much of a training step is the host's enqueue time? After a device synchronisation, the wall time until `step()` returns
(everything enqueued, nothing waited for except where the launch queue fills) vs the device time of the step.

    python tools/train_cpu_slack.py [workload] > gpurun_out/train_cpu_slack.json"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_workloads  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "decoder_train_f8"
wl = bench_workloads.build(name, torch.device("cuda", 0))
for _ in range(4):
    wl.step()
torch.cuda.synchronize()
enq, tot = [], []
for _ in range(8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    wl.step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    enq.append((t1 - t0) * 1e3)
    tot.append((t2 - t0) * 1e3)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(8):
    wl.step()
b.record()
torch.cuda.synchronize()
print(json.dumps({"workload": name, "host_enqueue_ms": enq, "single_step_total_ms": tot, "back_to_back_ms_per_step": a.elapsed_time(b) / 8,
                  "cpu_count": os.cpu_count(), "loadavg": os.getloadavg()}, indent=1))
