import torch, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from racformer_b200 import points, linear
g = torch.Generator(device="cuda").manual_seed(1)
QG, P_in, C, P_out = 48, 96, 64, 128
x = torch.randn(QG, P_in, C, device="cuda", generator=g)
params = torch.randn(QG, C * C + P_out * P_in, device="cuda", generator=g) * 0.2
full = points.adaptive_mixing_core(x, params, P_out)
full2 = points.adaptive_mixing_core(x, params, P_out)
t = points.adaptive_mixing_core(x, params, P_out, split=True, tiled_groups=4)
a = linear.untile(t).double().sum(0).reshape(QG, P_out, C)
b = full.double()
bad = (a != b)
print("equal runs", torch.equal(full, full2), "mismatches", int(bad.sum()), "of", bad.numel(), "max diff", float((a-b).abs().max()))
idx = bad.nonzero()
print(idx[:20].tolist())
print("items", sorted(set(idx[:,0].tolist()))[:50])
print("rows", sorted(set(idx[:,1].tolist()))[:130])
print("cols", sorted(set(idx[:,2].tolist()))[:70])
