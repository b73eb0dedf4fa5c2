"""Measure the gather speed of light on this GPU: random coalesced 512-byte row reads (see tools/csrc/ceiling.cu)."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import tools_lib  # noqa: E402

lib = tools_lib.load()
dev = torch.device("cuda", 0)
sink = torch.zeros(4, device=dev)
res = []
for mb in (32, 64, 134, 735, 2048):
    n_rows = mb * (1 << 20) // 512
    buf = torch.randn(n_rows * 128, device=dev)
    total_rows = int(1.5e9 // 512)     # ~1.5 GB of requests, like one MSMV forward at the f8 shapes
    for ilp in (2, 4, 8, 16):
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            assert lib.racf_bench_gather_ceiling(buf.data_ptr(), n_rows, total_rows, ilp, sink.data_ptr(), st) == 0
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            lib.racf_bench_gather_ceiling(buf.data_ptr(), n_rows, total_rows, ilp, sink.data_ptr(), st)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        res.append({"footprint_mb": mb, "ilp": ilp, "ms": ms, "gbs": total_rows * 512 / ms / 1e6})
    # scatter: vector reductions on random rows; (a) warm = buffer resident / dirty in L2 where it fits,
    # (b) after a memset of the buffer (what the backward pass sees: zero-fill, then read-modify-write)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    total_rows = int(1.4e9 // 512)
    for mode in ("warm", "after_memset"):
        for _ in range(2):
            lib.racf_bench_scatter_ceiling(buf.data_ptr(), n_rows, total_rows, st)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            if mode == "after_memset":
                buf.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            lib.racf_bench_scatter_ceiling(buf.data_ptr(), n_rows, total_rows, st)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
        res.append({"footprint_mb": mb, "scatter": mode, "ms": ms, "gbs": total_rows * 512 / ms / 1e6})
    del buf
print(json.dumps(res, indent=1))
