import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.mixing_bench import make, chain, P_out
from racformer_b200 import points
for QG in (449, 3600):
    x, params = make(QG, 96, seed=QG)
    ref = chain(x, params, torch.float64)
    for rep in range(2):
        o = points.adaptive_mixing_core(x, params, P_out, variant=2)
        err = (o.double() - ref).abs().amax(dim=(1, 2))
        bad = (err > 1e-4).nonzero().flatten().tolist()
        print(QG, "rep", rep, "bad items:", len(bad), [(b % 148, b // 148) for b in bad[:40]])
        if bad:
            b = bad[0]
            e = (o[b].double() - ref[b]).abs()
            print(" first bad item", b, "rows with err", (e.amax(1) > 1e-4).nonzero().flatten().tolist()[:40], "cols", (e.amax(0) > 1e-4).nonzero().flatten().tolist()[:70])
            print(" nan?", bool(torch.isnan(o[b]).any()), "max", float(e.max()))
