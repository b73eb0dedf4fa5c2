"""ncu target: one row program that is only a chain of 17 256x256 Linear operators (900 rows), to profile the Linear operator
of csrc/rowops.cu in isolation."""
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from racformer_b200 import rowops  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
lin = nn.Linear(256, 256).to(dev)
x = torch.randn(900, 256, device=dev)
with torch.no_grad():
    for _ in range(3):
        p = rowops.RowProgram(900, width=256, num_bufs=2)
        p.load(0, x)
        for i in range(17):
            p.linear((i + 1) & 1, i & 1, lin)
        p.store(1, 256)
        p.run()
torch.cuda.synchronize()
print("ok")
