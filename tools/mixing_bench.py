"""AdaptiveMixing core at f8 shapes (900 queries x 4 groups, P_in = 96): CUDA-core kernel vs tcgen05 kernel vs the
PyTorch op chain; time (CUDA events, 20 iterations) and max error against an fp64 evaluation."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from racformer_b200 import points  # noqa: E402

QG, P_in, C, P_out = 3600, 96, 64, 128
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(QG, P_in, C, device="cuda", generator=g)
params = torch.randn(QG, C * C + P_out * P_in, device="cuda", generator=g) * 0.2


def chain(dt):
    m, s = params.to(dt).split([C * C, P_out * P_in], 1)
    t = F.relu(F.layer_norm(torch.matmul(x.to(dt), m.reshape(QG, C, C)), [P_in, C]))
    return F.relu(F.layer_norm(torch.matmul(s.reshape(QG, P_out, P_in), t), [P_out, C]))


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


ref64 = chain(torch.float64)
res = {"shape": {"QG": QG, "P_in": P_in, "C": C, "P_out": P_out}}
for name, fn in (("cuda_cores", lambda: points.adaptive_mixing_core(x, params, P_out, tensor_cores=False)),
                 ("tcgen05", lambda: points.adaptive_mixing_core(x, params, P_out)),
                 ("tcgen05_tiled_out", lambda: points.adaptive_mixing_core(x, params, P_out, split=True, tiled_groups=4)),
                 ("pytorch_chain", lambda: chain(torch.float32))):
    out = fn()
    res[name] = {"us": timed(fn)}
    if torch.is_tensor(out):
        res[name]["max_err_vs_fp64"] = float((out.double() - ref64).abs().max())
        res[name]["mean_err_vs_fp64"] = float((out.double() - ref64).abs().mean())
print(json.dumps(res, indent=1))
