import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from racformer_b200 import linear
M, N, K = 900, 65536, 256
a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / 16; b = torch.randn(N, device="cuda")
a3, w3 = linear.split_tiled(a), linear.split_tiled(w)
for _ in range(3): y = linear.linear_bf16x3(a3, w3, b)
torch.cuda.synchronize()
