"""ncu target / timing of the AdaptiveMixing-core backward (csrc/mixing_bwd.cu) at the training shapes
(B = 2, 1220 queries, 4 groups -> 9760 items, P_in = 96)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from racformer_b200 import points  # noqa: E402

QG, P_in, C, P_out = 9760, 96, 64, 128
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(QG, P_in, C, device="cuda", generator=g)
params = torch.randn(QG, C * C + P_out * P_in, device="cuda", generator=g) * 0.2
gy = torch.randn(QG, P_out, C, device="cuda", generator=g)
for _ in range(2):
    points.adaptive_mixing_core_backward(x, params, gy, P_out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    points.adaptive_mixing_core_backward(x, params, gy, P_out)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
flops = QG * 2.0 * (2 * (P_in * C * C + P_out * P_in * C) + P_in * C * C + P_out * P_in * C)
print(json.dumps({"items": QG, "ms": ms, "fp32_tflops": flops / ms / 1e9}))
