"""AdaptiveMixing core backward: tcgen05 kernel (csrc/mixing_bwd_tc.cu, variant 2) vs the CUDA-core kernel (csrc/mixing_bwd.cu,
variant 1) vs fp64 autograd of the PyTorch chain -- error of g_x / g_M / g_S relative to each gradient's largest magnitude,
and time per launch at the training shapes.    python tools/mixing_bwd_check.py > gpurun_out/mixing_bwd_check.json"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from racformer_b200 import points  # noqa: E402

C, P_out = 64, 128


def chain(xx, pp, p_in):
    QG = xx.shape[0]
    m, s = pp.split([C * C, P_out * p_in], 1)
    t = F.relu(F.layer_norm(torch.matmul(xx, m.reshape(QG, C, C)), [p_in, C]))
    return F.relu(F.layer_norm(torch.matmul(s.reshape(QG, P_out, p_in), t), [P_out, C]))


def timed(fn, iters=10, warmup=2):
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


res = {}
g = torch.Generator(device="cuda").manual_seed(11)
for QG, p_in in ((3, 96), (40, 96), (449, 96), (40, 32), (40, 48), (40, 16), (40, 80), (40, 64)):
    x = torch.randn(QG, p_in, C, device="cuda", generator=g)
    params = torch.randn(QG, C * C + P_out * p_in, device="cuda", generator=g) * 0.2
    gy = torch.randn(QG, P_out, C, device="cuda", generator=g)
    x64, p64 = x.double().requires_grad_(), params.double().requires_grad_()
    chain(x64, p64, p_in).backward(gy.double())
    entry = {}
    for name, variant in (("cuda_cores", 1), ("tcgen05", 2)):
        gx, gp = points.adaptive_mixing_core_backward(x, params, gy, P_out, variant=variant)
        torch.cuda.synchronize()
        for what, got, want in (("g_x", gx, x64.grad), ("g_M", gp[:, :C * C], p64.grad[:, :C * C]), ("g_S", gp[:, C * C:], p64.grad[:, C * C:])):
            entry[f"{name}_{what}_rel_err"] = float((got.double() - want).abs().max() / want.abs().max())
    res[f"QG{QG}_P{p_in}"] = entry
    print(f"QG{QG}_P{p_in}", entry, file=sys.stderr)
QG, p_in = 9760, 96
x = torch.randn(QG, p_in, C, device="cuda", generator=g)
params = torch.randn(QG, C * C + P_out * p_in, device="cuda", generator=g) * 0.2
gy = torch.randn(QG, P_out, C, device="cuda", generator=g)
for name, variant in (("cuda_cores", 1), ("tcgen05", 2)):
    res[f"{name}_us_QG{QG}_P{p_in}"] = timed(lambda: points.adaptive_mixing_core_backward(x, params, gy, P_out, variant=variant))
print(json.dumps(res, indent=1))
