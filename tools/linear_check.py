"""Accuracy and speed of the tcgen05 bf16x3 Linear (csrc/linear.cu) against an fp64 product and cuBLAS SGEMM, on a B200.

    python tools/linear_check.py [--out gpurun_out/linear_check.json]

Every case runs in its own process under a timeout, so a faulting variant cannot take the others down.
"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = [   # (M, N, K, bias)
    (128, 128, 32, False), (128, 128, 256, True), (77, 200, 40, True), (900, 256, 256, True),
    (900, 65536, 256, True),     # AdaptiveMixing.parameter_generator at f8
    (900, 256, 32768, True),     # AdaptiveMixing.out_proj at f8
]


def run_case(M, N, K, bias, variant, max_order, iters):
    import torch
    from racformer_b200 import linear
    torch.manual_seed(0)
    dev = "cuda"
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev) if bias else None
    a3, w3 = (linear.split_tiled(a), linear.split_tiled(w)) if variant == 2 else (linear.split_bf16x3(a), linear.split_bf16x3(w))
    pa, pw = (linear.untile(a3), linear.untile(w3)) if variant == 2 else (a3, w3)
    exact_split = bool((pa.double().sum(0) == a.double()).all()) and bool((pw.double().sum(0) == w.double()).all())
    del pa, pw
    y = linear.linear_bf16x3(a3, w3, b, max_order=max_order, variant=variant)
    torch.cuda.synchronize()
    rows = slice(0, min(M, 256))
    ref = a[rows].double() @ w.double().t()
    scale = a[rows].double().abs() @ w.double().abs().t()     # sum_k |a||w|: the natural error scale of a dot product
    if b is not None:
        ref += b.double()
        scale += b.double().abs()
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    y32 = torch.nn.functional.linear(a, w, b)
    torch.backends.cuda.matmul.allow_tf32 = old
    err = ((y[rows].double() - ref).abs() / scale)
    err32 = ((y32[rows].double() - ref).abs() / scale)
    rel = (y[rows].double() - ref).abs().max().item() / ref.abs().max().item()
    res = {"M": M, "N": N, "K": K, "bias": bias, "variant": variant, "max_order": max_order, "exact_split": exact_split,
           "split_k": linear.plan(M, N, K)[0],
           "err_over_sum_abs_max": err.max().item(), "err_over_sum_abs_mean": err.mean().item(),
           "sgemm_err_over_sum_abs_max": err32.max().item(), "sgemm_err_over_sum_abs_mean": err32.mean().item(),
           "max_abs_err_over_max_abs_ref": rel,
           "signed_mean_err_over_scale": ((y[rows].double() - ref) / scale).mean().item()}
    if iters > 0:
        def timed(fn):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters * 1e3
        res["us"] = timed(lambda: linear.linear_bf16x3(a3, w3, b, max_order=max_order, variant=variant))
        res["us_split_a"] = timed(lambda: linear.split_tiled(a) if variant == 2 else linear.split_bf16x3(a))
        torch.backends.cuda.matmul.allow_tf32 = False
        res["us_sgemm"] = timed(lambda: torch.nn.functional.linear(a, w, b))
        terms = {4: 9, 3: 8, 2: 6, 1: 3, 0: 1}[max_order]
        res["bf16_tflops"] = 2.0 * M * N * K * terms / res["us"] * 1e-6
    return res


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--case":
        M, N, K, bias, variant, max_order, iters = [int(v) for v in sys.argv[2:9]]
        print("RESULT " + json.dumps(run_case(M, N, K, bool(bias), variant, max_order, iters)))
        sys.exit(0)
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    results = []
    variants = [int(v) for v in sys.argv[sys.argv.index("--variants") + 1].split(",")] if "--variants" in sys.argv else (0, 1, 2)
    for variant in variants:
        for (M, N, K, bias) in CASES:
            for max_order in ((4, 2, 0) if N * K >= 1 << 22 or (M, N, K) == (128, 128, 256) else (4,)):
                iters = 20 if N * K >= 1 << 22 else 0
                cmd = [sys.executable, os.path.abspath(__file__), "--case"] + [str(int(v)) for v in
                                                                              (M, N, K, bias, variant, max_order, iters)]
                try:
                    p = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
                    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                    r = json.loads(line[0][7:]) if line else {"M": M, "N": N, "K": K, "variant": variant,
                                                              "max_order": max_order, "failed": p.returncode,
                                                              "stderr": p.stderr[-600:]}
                except subprocess.TimeoutExpired:
                    r = {"M": M, "N": N, "K": K, "variant": variant, "max_order": max_order, "failed": "timeout"}
                results.append(r)
                print(json.dumps(r), flush=True)
    if out:
        with open(out, "w") as fh:
            json.dump(results, fh, indent=1)
