"""ncu target: a few eager launches of the decoder's row programs and self-attention core at the f8 shapes (see
tools/rowops_bench.py)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from racformer_b200.decoder import RaCFormerTransformerDecoderLayer  # noqa: E402
from racformer_b200.synthetic import PC_RANGE  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, Q, E, T = 1, 900, 256, 8
layer = RaCFormerTransformerDecoderLayer(E, num_frames=T, pc_range=PC_RANGE).to(dev).eval()
layer.init_weights()
mixed, qf = torch.randn(B, Q, E, device=dev), torch.randn(B, Q, E, device=dev)
radar = (torch.randn(B * T, Q, E, device=dev), torch.randn(B, Q, T, device=dev))
lss = (torch.randn(B * T, Q, E, device=dev), torch.randn(B, Q, T, device=dev))
qb = torch.rand(B, Q, 10, device=dev)
with torch.no_grad():
    for _ in range(3):
        layer._tail_rows(mixed, qf, radar, lss)
        layer._self_attn_rows(qb, qf)
torch.cuda.synchronize()
print("ok")
