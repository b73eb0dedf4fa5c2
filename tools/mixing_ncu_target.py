"""ncu target: the AdaptiveMixing core at the f8 shapes (3600 items, P_in = 96), tiled output, one kernel variant.
   python tools/mixing_ncu_target.py [variant]   (1 = phase-serial csrc/mixing_tc.cu, 2 = warp-specialised csrc/mixing_ws.cu)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from racformer_b200 import points  # noqa: E402
from tools.mixing_bench import make, P_out  # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
x, params = make(3600, 96)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
for _ in range(reps):
    points.adaptive_mixing_core(x, params, P_out, split=True, tiled_groups=4, variant=variant)
torch.cuda.synchronize()
