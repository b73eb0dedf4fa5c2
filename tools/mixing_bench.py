"""AdaptiveMixing core at f8 shapes (900 queries x 4 groups, P_in = 96): CUDA-core kernel vs the two tcgen05 kernels
(phase-serial csrc/mixing_tc.cu, warp-specialised csrc/mixing_ws.cu) vs the PyTorch op chain; time (CUDA events, L2 flushed
by the 490 MB working set itself) and max error against an fp64 evaluation. `--check` also runs small / ragged shapes."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from racformer_b200 import points  # noqa: E402

C, P_out = 64, 128


def make(QG, P_in, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(QG, P_in, C, device="cuda", generator=g)
    params = torch.randn(QG, C * C + P_out * P_in, device="cuda", generator=g) * 0.2
    return x, params


def chain(x, params, dt):
    QG, P_in, _ = x.shape
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


def tiled_to_dense(op, rows, K):
    """TiledOperand buffer -> fp32 [rows, K] (sum of the three pieces), undoing the 64-byte swizzle."""
    buf = op.buf.view(torch.bfloat16).reshape(-1, K // 32, 3, 128, 4, 8).float().sum(2)       # [rt, kb, r, chunk, 8]
    r = torch.arange(128, device=buf.device)
    c = torch.arange(4, device=buf.device)
    phys = c[None, :] ^ ((r[:, None] >> 1) & 3)                                                  # logical chunk c of row r sits at phys
    dense = torch.gather(buf, 3, phys[None, None, :, :, None].expand(buf.shape[0], buf.shape[1], 128, 4, 8))
    return dense.permute(0, 2, 1, 3, 4).reshape(-1, K)[:rows]


def main():
    res = {}
    if "--check" in sys.argv:
        for QG, P_in in ((4, 96), (148 * 3 + 5, 96), (300, 64), (40, 48), (40, 80), (40, 32), (40, 16), (8, 128)):
            x, params = make(QG, P_in, seed=QG + P_in)
            ref = chain(x, params, torch.float64)
            row = {}
            for v in (1, 2):
                if v == 2 and P_in > 96:
                    continue
                o = points.adaptive_mixing_core(x, params, P_out, variant=v)
                row[f"v{v}_max_err"] = float((o.double() - ref).abs().max())
                if QG % 4 == 0:
                    ot = points.adaptive_mixing_core(x, params, P_out, split=True, tiled_groups=4, variant=v)
                    d = tiled_to_dense(ot, QG // 4, 4 * P_out * C).reshape(QG, P_out, C)
                    row[f"v{v}_tiled_max_err"] = float((d.double() - ref).abs().max())
            torch.cuda.synchronize()
            res[f"check_QG{QG}_P{P_in}"] = row
    QG, P_in = 3600, 96
    x, params = make(QG, P_in)
    ref64 = chain(x, params, torch.float64)
    res["shape"] = {"QG": QG, "P_in": P_in, "C": C, "P_out": P_out}
    for name, fn in (("cuda_cores", lambda: points.adaptive_mixing_core(x, params, P_out, tensor_cores=False)),
                     ("tcgen05_phase_serial", lambda: points.adaptive_mixing_core(x, params, P_out, variant=1)),
                     ("tcgen05_phase_serial_tiled_out", lambda: points.adaptive_mixing_core(x, params, P_out, split=True, tiled_groups=4, variant=1)),
                     ("tcgen05_warp_specialised", lambda: points.adaptive_mixing_core(x, params, P_out, variant=2)),
                     ("tcgen05_warp_specialised_tiled_out", lambda: points.adaptive_mixing_core(x, params, P_out, split=True, tiled_groups=4, variant=2)),
                     ("pytorch_chain", lambda: chain(x, params, torch.float32))):
        out = fn()
        res[name] = {"us": timed(fn)}
        if torch.is_tensor(out):
            res[name]["max_err_vs_fp64"] = float((out.double() - ref64).abs().max())
            res[name]["mean_err_vs_fp64"] = float((out.double() - ref64).abs().mean())
        else:
            d = tiled_to_dense(out, QG // 4, 4 * P_out * C).reshape(QG, P_out, C)
            res[name]["max_err_vs_fp64"] = float((d.double() - ref64).abs().max())
    hbm_bytes = QG * (P_in * C * 4 + (C * C + P_out * P_in) * 4 + 3 * P_out * C * 2)
    res["algorithmic_hbm_bytes_tiled_out"] = hbm_bytes
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
