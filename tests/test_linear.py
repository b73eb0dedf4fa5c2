"""The tcgen05 Linear layers (csrc/linear.cu) behind racformer_b200.linear.

Reference computation: F.linear in fp32 (AdaptiveMixing.parameter_generator / out_proj, models/racformer_transformer.py:560-566,
called at :592 and :606). A floating-point kernel, so the checker is a torch reference: the fp64 product is the truth, and
cuBLAS SGEMM (what the reference runs on a GPU) sets the bar -- the tensor-core path must be at least as close to fp64.

Error measure: |y - y64| / (sum_k |a_k||w_k| + |b|), the natural scale of a dot product's rounding error.
Stated tolerance: max 6e-7, mean 4e-8 (fp32 unit roundoff is 6e-8; SGEMM measures 3.6e-7 / 1.9e-8 on these shapes).
"""
import pytest
import torch

from racformer_b200 import _lib

MAX_TOL, MEAN_TOL = 6e-7, 4e-8


def _normalised_error(y, a, w, b):
    ref = a.double() @ w.double().t()
    scale = a.double().abs() @ w.double().abs().t()
    if b is not None:
        ref += b.double()
        scale += b.double().abs()
    return (y.double() - ref).abs() / scale


def test_plan_bounds_k_per_accumulator():
    """Host logic (no GPU): the K split keeps every TMEM accumulator at <= 512 of K and sizes the workspace."""
    import ctypes
    lib = _lib.load()
    for (M, N, K) in [(900, 65536, 256), (900, 256, 32768), (5, 7, 8), (128, 128, 513)]:
        s, ws = ctypes.c_int(0), ctypes.c_longlong(0)
        assert lib.racf_linear_bf16x3_plan(M, N, K, ctypes.byref(s), ctypes.byref(ws)) == 0
        assert s.value == (K + 511) // 512
        assert ws.value == (s.value * M * N * 4 if s.value > 1 else 0)
    s, ws = ctypes.c_int(0), ctypes.c_longlong(0)
    assert lib.racf_linear_bf16x3_plan(0, 4, 8, ctypes.byref(s), ctypes.byref(ws)) == -3
    assert lib.racf_linear_bf16x3_plan(4, 4, 8, None, None) == -1


def test_forward_rejects_bad_arguments_without_touching_the_gpu():
    """Argument errors are reported before any CUDA call (so this runs on a CPU-only box)."""
    lib = _lib.load()
    p = 0x1000   # never dereferenced
    assert lib.racf_linear_bf16x3_forward(None, p, None, 4, 4, 8, 4, 1, 0, None, p, None) == -1
    assert lib.racf_linear_bf16x3_forward(p, p, None, 4, 4, 12, 4, 1, 0, None, p, None) == -6      # K % 8
    assert lib.racf_linear_bf16x3_forward(p, p, None, 4, 4, 8, 5, 1, 0, None, p, None) == -6       # max_order
    assert lib.racf_linear_bf16x3_forward(p, p, None, 4, 0, 8, 4, 1, 0, None, p, None) == -3
    assert lib.racf_linear_bf16x3_forward(p, p, None, 4, 4, 1024, 4, 2, 0, None, p, None) == -1    # split without workspace
    assert lib.racf_split_bf16x3(None, 4, p, None) == -1
    assert lib.racf_split_bf16x3(p, 6, p, None) == -6


@pytest.mark.gpu
def test_split_is_exact_and_ordered():
    from racformer_b200 import linear
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(1000, 64, device="cuda", generator=g) * torch.logspace(-20, 20, 64, device="cuda")
    x[0, :8] = torch.tensor([0.0, -0.0, 1.0, -1.0, 3.0e38, 1.0e-30, 1.0 + 2.0 ** -23, 65504.0], device="cuda")
    p = linear.split_bf16x3(x)
    assert p.shape == (3, 1000, 64) and p.dtype == torch.bfloat16
    assert torch.equal(p.double().sum(0), x.double())                       # exact: 8 + 8 + 8 significand bits
    assert bool((p[1].float().abs() <= p[0].float().abs() * 2.0 ** -8 + 1e-45).all())   # each piece 2^-8 of the previous
    assert bool((p[2].float().abs() <= p[1].float().abs() * 2.0 ** -8 + 1e-45).all())


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(128, 128, 32), (77, 200, 40), (1, 8, 8), (300, 132, 264), (900, 256, 256), (257, 1030, 520)])
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_linear_matches_fp64_like_sgemm(shape, variant):
    """Ragged M / N / K (TMA zero-fills the tails; the tiled format pads them), bias on, all nine terms and the six-term
    default. Variant 2 = the pre-tiled operand format (bulk copies), 0 / 1 = plain pieces through tensor maps."""
    from racformer_b200 import linear
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    a = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    a3, w3 = (linear.split_tiled(a), linear.split_tiled(w)) if variant == 2 else (linear.split_bf16x3(a), linear.split_bf16x3(w))
    for order in (linear.ALL_TERMS, linear.SIX_TERMS):
        y = linear.linear_bf16x3(a3, w3, b, max_order=order, variant=variant)
        err = _normalised_error(y, a, w, b)
        assert err.max().item() < MAX_TOL and err.mean().item() < MEAN_TOL, (shape, order, err.max().item(), err.mean().item())
        assert torch.equal(y, linear.linear_bf16x3(a3, w3, b, max_order=order, variant=variant)), "deterministic"
    y0 = linear.linear_bf16x3(a3, w3, None, variant=variant)
    assert _normalised_error(y0, a, w, None).max().item() < MAX_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(5, 12), (128, 32), (300, 100), (900, 256)])
def test_tiled_format_holds_the_same_pieces(shape):
    """split_tiled writes exactly the pieces of split_bf16x3, at the offsets csrc/linear_tiled.cuh defines (any K: the
    tail of the last 32-wide K block is zero), and the tiled Linear works for a K that is not a multiple of 8."""
    from racformer_b200 import linear
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(*shape, device="cuda", generator=g)
    t = linear.split_tiled(x)
    assert t.buf.numel() * 2 == linear.tiled_bytes(*shape) == -(-shape[0] // 128) * -(-shape[1] // 32) * 24576
    assert torch.equal(linear.untile(t).double().sum(0), x.double())
    if shape[1] % 4 == 0:
        assert torch.equal(linear.untile(t), linear.split_bf16x3(x))
    if shape[1] % 32:
        padded = linear.untile(linear.TiledOperand(t.buf, shape[0], -(-shape[1] // 32) * 32))
        assert bool((padded[:, :, shape[1]:] == 0).all())
    w = torch.randn(40, shape[1], device="cuda", generator=g)
    y = linear.linear_bf16x3(t, linear.split_tiled(w))
    assert _normalised_error(y, x, w, None).max().item() < MAX_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(900, 65536, 256), (900, 256, 32768)])
def test_linear_f8_shapes_no_worse_than_sgemm(shape):
    """AdaptiveMixing's two layers at racformer_r50_nuimg_704x256_f8 sizes: closer to fp64 than (or as close as) cuBLAS
    SGEMM, and inside the stated tolerance. out_proj exercises the K split (64 partial tiles + fp32 reduction)."""
    from racformer_b200 import linear
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    lin = torch.nn.Linear(K, N, device="cuda")
    with torch.no_grad():
        lin.weight.copy_(w)
        lin.bias.copy_(b)
        fast = linear.SplitLinear(lin, max_order=linear.SIX_TERMS)
        y = fast(a)
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        y32 = lin(a)
        torch.backends.cuda.matmul.allow_tf32 = old
    rows = slice(0, 200)
    err = _normalised_error(y[rows], a[rows], w, b)
    err32 = _normalised_error(y32[rows], a[rows], w, b)
    assert err.max().item() < MAX_TOL and err.mean().item() < MEAN_TOL
    assert err.mean().item() <= 1.25 * err32.mean().item(), (err.mean().item(), err32.mean().item())
    torch.testing.assert_close(y, y32, rtol=1e-5, atol=2e-5)
    # weight cache: re-split when the parameter changes in place
    with torch.no_grad():
        lin.weight.mul_(2.0)
        torch.testing.assert_close(fast(a) - b, 2.0 * (y - b), rtol=1e-5, atol=4e-5)


@pytest.mark.gpu
def test_mixing_core_split_output_is_the_fp32_output():
    """racf_adaptive_mixing_forward_split: the three bf16 pieces sum to the fp32 kernel's result bit for bit."""
    from racformer_b200 import points
    g = torch.Generator(device="cuda").manual_seed(1)
    QG, P_in, C, P_out = 48, 96, 64, 128
    x = torch.randn(QG, P_in, C, device="cuda", generator=g)
    params = torch.randn(QG, C * C + P_out * P_in, device="cuda", generator=g) * 0.2
    full = points.adaptive_mixing_core(x, params, P_out, tensor_cores=False)
    pieces = points.adaptive_mixing_core(x, params, P_out, split=True)
    assert pieces.shape == (3, QG, P_out, C) and pieces.dtype == torch.bfloat16
    assert torch.equal(pieces.double().sum(0), full.double())
    # ... and as the pre-tiled A operand of out_proj: [QG / 4 queries, 4 groups * 128 points * 64 channels]
    from racformer_b200 import linear
    tiled = points.adaptive_mixing_core(x, params, P_out, split=True, tiled_groups=4, tensor_cores=False)
    assert (tiled.rows, tiled.K) == (QG // 4, 4 * P_out * C)
    assert torch.equal(linear.untile(tiled), pieces.reshape(3, QG // 4, -1))
    # the tensor-core kernel (csrc/mixing_tc.cu): its tiled pieces are the split of its own fp32 result
    full_tc = points.adaptive_mixing_core(x, params, P_out)
    tiled_tc = points.adaptive_mixing_core(x, params, P_out, split=True, tiled_groups=4)
    assert torch.equal(linear.untile(tiled_tc).double().sum(0), full_tc.reshape(QG // 4, -1).double())
    assert float((full_tc - full).abs().max()) <= 2e-5


@pytest.mark.gpu
def test_linear_python_layer_rejects_bad_inputs():
    from racformer_b200 import linear
    a3 = torch.zeros(3, 8, 16, dtype=torch.bfloat16, device="cuda")
    w3 = torch.zeros(3, 8, 24, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError):
        linear.linear_bf16x3(a3, w3)                                   # K mismatch
    with pytest.raises(RuntimeError):
        linear.linear_bf16x3(a3.float(), a3.float())                   # not bf16 pieces
    with pytest.raises(RuntimeError):
        linear.split_bf16x3(torch.zeros(8, 16))                        # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        linear.linear_bf16x3(a3, a3, bias=torch.zeros(3, device="cuda"))


@pytest.mark.gpu
def test_chw_to_hwc_split_and_tensor_core_value_proj():
    """racf_split_bf16x3_chw_to_hwc == split(permute(x + pos)) exactly, and BEVSelfAttention.project_value on the
    tensor-core path equals the reference's add / permute / F.linear chain (models/bev_self_attention.py:162-174)."""
    from racformer_b200 import linear
    from racformer_b200.decoder import BEVSelfAttention
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(3, 40, 50, device="cuda", generator=g)
    pos = torch.rand(40, 50, device="cuda", generator=g)
    p = linear.split_bf16x3_chw_to_hwc(x, pos)
    assert p.shape == (3, 150, 40)
    assert torch.equal(p.double().sum(0), (x + pos).permute(0, 2, 1).reshape(150, 40).double())
    assert torch.equal(linear.split_bf16x3_chw_to_hwc(x).double().sum(0), x.permute(0, 2, 1).reshape(150, 40).double())
    tiled = linear.split_bf16x3_chw_to_hwc(x, pos, tiled=True)
    assert (tiled.rows, tiled.K) == (150, 40) and torch.equal(linear.untile(tiled), p)
    wide = linear.untile(linear.TiledOperand(tiled.buf, 150, 64))
    assert bool((wide[:, :, 40:] == 0).all())                      # K tail of the tiled format is zero-filled

    attn = BEVSelfAttention(embed_dims=64, num_heads=4, num_levels=1, num_points=4, num_bev_queue=2).cuda().eval()
    bev = torch.randn(1, 2, 64, 9, 13, device="cuda", generator=g)
    pos = torch.rand(64, 9, 13, device="cuda", generator=g)
    with torch.no_grad():
        fast = attn.project_value(bev, pos)
        attn.tensor_core_value_proj = False
        slow = attn.project_value(bev, pos)
    assert fast.shape == slow.shape == (2, 117, 4, 16)
    torch.testing.assert_close(fast, slow, rtol=1e-5, atol=1e-5)
    assert attn._split_value_proj is not None


@pytest.mark.gpu
def test_multi_linear_equals_the_separate_layers():
    """racf_linear_bf16x3_multi_forward: the sampling heads that share the query features, one launch, each output
    dense in its own tensor -- equal to the per-layer tensor-core results bit for bit, and to F.linear to fp32 noise."""
    from racformer_b200 import linear
    torch.manual_seed(4)
    sizes = [160, 5, 80, 8, 144, 3, 1536, 256]
    lins = [torch.nn.Linear(256, n, device="cuda") for n in sizes]
    lins[3].bias = None
    x = torch.randn(2, 450, 256, device="cuda")
    with torch.no_grad():
        multi = linear.MultiSplitLinear(lins)
        outs = multi(x)
        assert [tuple(o.shape) for o in outs] == [(900, n) for n in sizes]
        for lin, o in zip(lins, outs):
            single = linear.SplitLinear(lin, max_order=linear.SIX_TERMS)(x).reshape(900, -1)
            assert torch.equal(o, single)
            torch.testing.assert_close(o, lin(x).reshape(900, -1), rtol=1e-5, atol=1e-5)
        lins[1].weight.add_(1.0)              # the stacked weight is rebuilt when any layer changes
        torch.testing.assert_close(multi(x)[1], lins[1](x).reshape(900, -1), rtol=1e-5, atol=2e-5)
    with pytest.raises(RuntimeError):
        linear.MultiSplitLinear([torch.nn.Linear(256, 8), torch.nn.Linear(128, 8)])


@pytest.mark.gpu
def test_split_tiled_with_row_periodic_addend_is_the_split_of_the_sum():
    """racf_split_bf16x3_tiled_add == racf_split_bf16x3_tiled of (x + addend broadcast over row blocks), bit for bit."""
    from racformer_b200 import linear
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(3 * 70, 40, device=dev, generator=g)
    add = torch.randn(70, 40, device=dev, generator=g)
    got = linear.split_tiled(x, add)
    ref = linear.split_tiled((x.view(3, 70, 40) + add).reshape(210, 40).contiguous())
    assert torch.equal(linear.untile(got), linear.untile(ref))


@pytest.mark.gpu
@pytest.mark.parametrize("rows,K,N", [(304, 256, 1024), (64, 512, 256), (2440, 256, 384)])
def test_trainable_split_linear_gradients_are_fp32_grade(rows, K, N):
    """TrainableSplitLinear (forward, grad_x, grad_W, grad_b as bf16x3 tcgen05 GEMMs) vs an fp64 evaluation, next to the
    errors of nn.Linear's cuBLAS SGEMM path on the same inputs."""
    import torch.nn as nn
    from racformer_b200 import linear
    dev = torch.device("cuda", 0)
    torch.manual_seed(rows + K)
    lin = nn.Linear(K, N).to(dev)
    x = torch.randn(2, rows // 2, K, device=dev)
    g = torch.randn(2, rows // 2, N, device=dev)

    def run(fn, dtype):
        xx = x.detach().to(dtype).clone().requires_grad_()
        w = lin.weight.detach().to(dtype).requires_grad_()
        b = lin.bias.detach().to(dtype).requires_grad_()
        y = fn(xx, w, b)
        y.backward(g.to(dtype))
        return y.detach(), xx.grad, w.grad, b.grad

    ref = run(lambda xx, w, b: torch.nn.functional.linear(xx, w, b), torch.float64)
    sgemm = run(lambda xx, w, b: torch.nn.functional.linear(xx, w, b), torch.float32)
    tsl = linear.TrainableSplitLinear(lin, max_order=linear.SIX_TERMS)
    assert tsl.supports(x)
    xx = x.detach().clone().requires_grad_()
    lin.zero_grad()
    y = tsl(xx)
    y.backward(g)
    got = (y.detach(), xx.grad, lin.weight.grad, lin.bias.grad)
    for name, a, s32, r in zip(("y", "grad_x", "grad_w", "grad_b"), got, sgemm, ref):
        scale = float(r.abs().max())
        err, err_sgemm = float((a.double() - r).abs().max()) / scale, float((s32.double() - r).abs().max()) / scale
        assert err <= max(2 * err_sgemm, 2e-6), (name, err, err_sgemm)
    # the cached W / W^T pieces follow the parameter
    with torch.no_grad():
        lin.weight.mul_(2.0)
    y2 = tsl(x)
    assert torch.allclose(y2, torch.nn.functional.linear(x, lin.weight, lin.bias), rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_decoder_training_gradients_tensor_core_linears_match_sgemm():
    """One training-mode decoder iteration on the GPU: AdaptiveMixing's Linear layers on the tcgen05 autograd path vs
    cuBLAS SGEMM (everything else identical) -- outputs and all parameter / input gradients."""
    from racformer_b200.decoder import RaCFormerTransformer
    from racformer_b200.synthetic import fill_parameters_by_name
    from tests.decoder_cases import SMALL, small_inputs
    dev = torch.device("cuda", 0)
    cfg = dict(SMALL, num_layers=1)
    model = RaCFormerTransformer(**cfg).to(dev)
    fill_parameters_by_name(model, seed=3)
    model.train()
    for m in model.modules():                      # deterministic: no dropout
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    inp = small_inputs(batch=2)
    results = []
    for precision in ("bf16x6", "fp32"):
        model.set_mixing_precision(precision)
        model.zero_grad()
        qf = inp["query_feat"].to(dev).requires_grad_()
        feats = [f.to(dev).requires_grad_() for f in inp["mlvl_feats"]]
        cls, box = model(inp["query_bbox"].to(dev), qf, feats, inp["lss_bev"].to(dev), inp["radar_bev"].to(dev), None,
                         inp["img_metas"])
        (cls.square().sum() + box.square().sum()).backward()
        mix = model.decoder.decoder_layer.mixing
        results.append((cls.detach(), box.detach(), qf.grad, feats[0].grad, mix.parameter_generator.weight.grad.clone(),
                        mix.out_proj.weight.grad.clone()))
    for name, a, b in zip(("cls", "box", "grad_query", "grad_feat0", "grad_param_gen_w", "grad_out_proj_w"), *results):
        tol = 1e-4 * float(b.abs().max()) + 1e-6
        assert float((a - b).abs().max()) <= tol, (name, float((a - b).abs().max()), tol)


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K,bias", [(128, 256, 32, False), (77, 200, 40, True), (300, 384, 256, True), (900, 640, 96, True),
                                        (130, 1152, 512, True), (257, 129, 64, False)])
def test_wide_tile_linear_is_bit_identical_to_the_128x128_kernel(M, N, K, bias):
    """csrc/linear_wide.cu (128 x 256 tiles on persistent CTAs, N = 256 MMAs, variant 3 -- the default for tiled operands with
    N > 128) issues the same piece products in the same order per accumulator as csrc/linear.cu: identical bits while a
    work item is one K pass (K <= 512), including odd numbers of 128-column blocks, ragged M / N / K and the bias."""
    from racformer_b200 import linear
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda") if bias else None
    a3, w3 = linear.split_tiled(a), linear.split_tiled(w)
    try:
        linear.WIDE_TILES = False
        narrow = linear.linear_bf16x3(a3, w3, b)
        linear.WIDE_TILES = True
        wide = linear.linear_bf16x3(a3, w3, b)
    finally:
        linear.WIDE_TILES = True
    assert torch.equal(narrow, wide)
    ref = a.double() @ w.double().t() + (b.double() if bias else 0)
    scale = a.double().abs() @ w.double().abs().t() + (b.double().abs() if bias else 0)
    assert float(((wide.double() - ref).abs() / scale).max()) <= 6e-7


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K", [(900, 256, 32768), (200, 384, 2048), (64, 256, 1536)])
def test_wide_tile_linear_multi_pass_split_k_is_fp32_grade_and_deterministic(M, N, K):
    """K > 512: a work item of the wide kernel adds several 512-wide passes in registers (fewer K splits than the 128 x 128
    kernel, other grouping of the same fp32 additions): same error bar against an fp64 product, run-to-run identical."""
    from racformer_b200 import linear
    torch.manual_seed(K)
    a = torch.randn(M, K, device="cuda")
    w = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    a3, w3 = linear.split_tiled(a), linear.split_tiled(w)
    y = linear.linear_bf16x3(a3, w3, b)
    assert torch.equal(y, linear.linear_bf16x3(a3, w3, b))
    try:
        linear.WIDE_TILES = False
        narrow = linear.linear_bf16x3(a3, w3, b)
    finally:
        linear.WIDE_TILES = True
    ref = a.double() @ w.double().t() + b.double()
    scale = a.double().abs() @ w.double().abs().t() + b.double().abs()
    err, err_n = float(((y.double() - ref).abs() / scale).max()), float(((narrow.double() - ref).abs() / scale).max())
    assert err <= 6e-7 and err <= 2.0 * err_n + 1e-8, (err, err_n)
