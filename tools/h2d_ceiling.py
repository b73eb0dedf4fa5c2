"""Concurrent pinned host->device copy ceiling of this box: plain cudaMemcpyAsync (torch `copy_(non_blocking=True)` from
pinned memory = one cudaMemcpyAsync per buffer) of `--mb` MB per GPU on 1, 2, 4, ... GPUs at once, one stream per GPU.

The end-to-end number of bench.py moves ~1 GB of fp32 features per sample over PCIe; this is its denominator.
Also reports where the pinned pages live (/proc/self/numa_maps), the NUMA topology, and -- with --numa-first-touch -- the
same copy from pages that were first touched by a thread pinned to each NUMA node in turn and then cudaHostRegister-ed.

    python tools/h2d_ceiling.py [--mb 1024] [--reps 5] [--numa-first-touch] > profiles/r02_h2d_ceiling.json
"""
import argparse
import ctypes
import glob
import json
import os
import re
import subprocess
import time

import torch


def numa_nodes():
    nodes = {}
    for path in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        n = int(re.search(r"node(\d+)$", path).group(1))
        try:
            nodes[n] = open(os.path.join(path, "cpulist")).read().strip()
        except OSError:
            pass
    return nodes


def parse_cpulist(text):
    cpus = []
    for part in text.split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus += list(range(int(a), int(b or a) + 1))
    return cpus


def pages_by_node(tensor):
    """NUMA placement of a host tensor's pages from /proc/self/numa_maps (N<node>=<pages> of the mapping holding it)."""
    addr = tensor.data_ptr()
    try:
        best = None
        for line in open("/proc/self/numa_maps"):
            f = line.split()
            start = int(f[0], 16)
            if start <= addr and (best is None or start > best[0]):
                best = (start, {m.group(1): int(m.group(2)) for m in (re.match(r"N(\d+)=(\d+)", x) for x in f) if m})
        return best[1] if best else None
    except OSError:
        return None


def copy_rate(host, dev, streams, k, reps):
    for d in range(k):
        with torch.cuda.device(d), torch.cuda.stream(streams[d]):
            dev[d].copy_(host[d], non_blocking=True)
    for d in range(k):
        torch.cuda.synchronize(d)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
    t0 = time.perf_counter()
    for d in range(k):
        with torch.cuda.device(d), torch.cuda.stream(streams[d]):
            ev[d][0].record()
            for _ in range(reps):
                dev[d].copy_(host[d], non_blocking=True)
            ev[d][1].record()
    for d in range(k):
        torch.cuda.synchronize(d)
    wall = time.perf_counter() - t0
    nbytes = host[0].numel() * host[0].element_size()
    per_gpu = [nbytes * reps / (ev[d][0].elapsed_time(ev[d][1]) * 1e-3) / 1e9 for d in range(k)]
    return {"gpus": k, "aggregate_gbs": k * nbytes * reps / wall / 1e9, "per_gpu_gbs_min": min(per_gpu),
            "per_gpu_gbs_max": max(per_gpu), "wall_ms": 1e3 * wall}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--numa-first-touch", action="store_true")
    args = ap.parse_args()
    n = torch.cuda.device_count()
    nbytes = args.mb << 20
    out = {"what": "concurrent pinned H2D, one cudaMemcpyAsync per buffer per repetition, one stream per GPU",
           "gpus_visible": n, "mb_per_gpu": args.mb, "reps": args.reps, "numa_nodes": numa_nodes(), "host_cpus": os.cpu_count()}
    try:
        out["nvidia_smi_topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except Exception as exc:  # noqa: BLE001
        out["nvidia_smi_topo"] = str(exc)
    streams = []
    for d in range(n):
        with torch.cuda.device(d):
            streams.append(torch.cuda.Stream(device=d))
    dev = [torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n)]
    host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(n)]
    for h in host:
        h.fill_(1)
    out["cudaHostAlloc_pages_by_node"] = pages_by_node(host[0])
    ks = [k for k in (1, 2, 4, 8) if k <= n]
    out["cudaHostAlloc"] = [copy_rate(host, dev, streams, k, args.reps) for k in ks]
    if args.numa_first_touch and len(out["numa_nodes"]) > 1:
        cudart = ctypes.CDLL("libcudart.so")
        saved = os.sched_getaffinity(0)
        del host
        res = {}
        for node, cpulist in out["numa_nodes"].items():
            try:
                os.sched_setaffinity(0, parse_cpulist(cpulist))
            except OSError as exc:
                res[str(node)] = {"error": str(exc)}
                continue
            bufs = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(n)]
            for b in bufs:
                b.fill_(1)                                         # first touch on this node
                rc = cudart.cudaHostRegister(ctypes.c_void_p(b.data_ptr()), ctypes.c_size_t(nbytes), 0)
                assert rc == 0, rc
            res[str(node)] = {"pages_by_node": pages_by_node(bufs[0]),
                              "rates": [copy_rate(bufs, dev, streams, k, args.reps) for k in ks]}
            for b in bufs:
                cudart.cudaHostUnregister(ctypes.c_void_p(b.data_ptr()))
            del bufs
        os.sched_setaffinity(0, saved)
        out["cudaHostRegister_first_touch_on_node"] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
