import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from racformer_b200 import linear
def timed(fn, iters=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3
res = {}
for (M, N, K, bias) in ((128, 256, 32, False), (77, 200, 40, True), (300, 384, 256, True), (900, 640, 96, True), (900, 65536, 256, True), (900, 256, 32768, True), (131072, 256, 256, True)):
    torch.manual_seed(0)
    a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda") if bias else None
    a3, w3 = linear.split_tiled(a), linear.split_tiled(w)
    linear.WIDE_TILES = False
    y2 = linear.linear_bf16x3(a3, w3, b)
    linear.WIDE_TILES = True
    y3 = linear.linear_bf16x3(a3, w3, b)
    torch.cuda.synchronize()
    rows = slice(0, min(M, 256))
    ref = a[rows].double() @ w.double().t() + (b.double() if b is not None else 0)
    e = {"equal": bool(torch.equal(y2, y3)), "max_diff": float((y2 - y3).abs().max()), "err3_vs_fp64": float((y3[rows].double() - ref).abs().max())}
    if M * N >= 900 * 256:
        linear.WIDE_TILES = False
        e["us_v2"] = timed(lambda: linear.linear_bf16x3(a3, w3, b))
        linear.WIDE_TILES = True
        e["us_v3"] = timed(lambda: linear.linear_bf16x3(a3, w3, b))
    res[f"{M}x{N}x{K}"] = e
    print(f"{M}x{N}x{K}", e, flush=True)
json.dump(res, open("gpurun_out/t32_lin.json", "w"), indent=1)
