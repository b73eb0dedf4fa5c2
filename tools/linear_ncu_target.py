"""ncu target: the tcgen05 Linear at the f8 call-site shapes (tiled operands), one kernel variant.
   python tools/linear_ncu_target.py [wide|narrow]   (wide = csrc/linear_wide.cu, narrow = csrc/linear.cu 128 x 128 tiles)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from racformer_b200 import linear  # noqa: E402

linear.WIDE_TILES = (sys.argv[1] if len(sys.argv) > 1 else "wide") == "wide"
torch.manual_seed(0)
for (M, N, K) in ((900, 65536, 256), (900, 256, 32768), (131072, 256, 256)):
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    a3, w3 = linear.split_tiled(a), linear.split_tiled(w)
    for _ in range(2):
        linear.linear_bf16x3(a3, w3, b)
torch.cuda.synchronize()
