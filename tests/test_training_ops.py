"""Training-mode bindings (racformer_b200/training.py) against the plain autograd path of the same ops (GPU)."""
import pytest
import torch

from tests.helpers import assert_close, make_msda_inputs, make_msmv_inputs

pytestmark = pytest.mark.gpu


def _dev(ts):
    return [t.cuda() for t in ts]


def test_sampling_layout_function_matches_permute_autograd():
    from racformer_b200 import training
    g = torch.Generator(device="cuda").manual_seed(0)
    B, T, N, G, C, H, W = 2, 2, 3, 4, 64, 6, 13
    x = torch.randn(B, T * N, G * C, H, W, device="cuda", generator=g, requires_grad=True)
    y = training.SamplingLayout.apply(x, N, G)
    want = x.reshape(B, T, N, G, C, H, W).permute(0, 1, 3, 2, 5, 6, 4).reshape(B * T * G, N, H, W, C)
    assert torch.equal(y, want)
    gy = torch.randn(y.shape, device="cuda", generator=g)
    (gx,) = torch.autograd.grad(y, x, gy)
    (gx_want,) = torch.autograd.grad(want, x, gy)
    assert torch.equal(gx, gx_want)


@pytest.mark.parametrize("levels", [2, 4, 5])
def test_msmv_grouped_function_matches_plain_function_and_accumulates(levels):
    """MSMVGrouped (grouped output, grouped grad_out, shared feature-gradient buffer) == MSMVSampling + permute, and two
    consumers of the same pyramid leave the SUM of their feature gradients in the one buffer autograd receives."""
    from racformer_b200 import training, wrapper
    B, T, G, N, Q, P = 2, 2, 4, 3, 37, 12
    shapes = [(16, 44), (8, 22), (4, 11), (2, 6), (1, 3)][:levels]
    feats, loc, w, _ = make_msmv_inputs(7, Bp=B * T * G, N=N, C=64, Q=Q, P=P, shapes=shapes, lo=-0.1, hi=1.1)
    loc2 = loc.flip(1).contiguous()
    gen = torch.Generator().manual_seed(1)
    gos = [torch.randn(B, Q, G, T * P, 64, generator=gen).cuda() for _ in range(2)]

    def run(fused):
        f = [t.cuda().requires_grad_() for t in feats]
        ls = [loc.cuda().requires_grad_(), loc2.cuda().requires_grad_()]
        ws = [w.cuda().requires_grad_(), w.cuda().requires_grad_()]
        shared = training.SharedGrads() if fused else None
        outs = []
        for l, ww in zip(ls, ws):
            if fused:
                outs.append(training.MSMVGrouped.apply(shared.get(f), T, G, l, ww, *f))
            else:
                o = wrapper.msmv_sampling(f, l, ww)                                   # [B*T*G,Q,C,P]
                outs.append(o.reshape(B, T, G, Q, 64, P).permute(0, 3, 2, 1, 5, 4).flatten(3, 4))
        loss = sum((o * g).sum() for o, g in zip(outs, gos))
        loss.backward()
        return outs, [t.grad for t in f], [t.grad for t in ls], [t.grad for t in ws]

    o1, gf1, gl1, gw1 = run(True)
    o0, gf0, gl0, gw0 = run(False)
    for a, b in zip(o1, o0):
        assert torch.equal(a, b)
    for a, b in zip(gl1 + gw1, gl0 + gw0):
        assert_close(a, b, 1e-5, 1e-6 * float(b.abs().max()), "grouped grad_loc / grad_w")
    for a, b in zip(gf1, gf0):
        assert_close(a, b, 1e-4, 1e-5 * float(b.abs().max()), "shared feature gradient")


def test_msda_shared_function_matches_plain_function_and_accumulates():
    from racformer_b200 import training
    from racformer_b200.multi_scale_deformable_attn_function import MultiScaleDeformableAttnFunction_fp32
    value, sp, lsi, loc, aw, g = make_msda_inputs(3, B=2, M=4, D=64, Q=50, P=20, shapes=[(16, 16)], lo=-0.05, hi=1.05)
    sp, lsi, g = sp.cuda(), lsi.cuda(), g.cuda()

    def run(fused):
        v = value.cuda().requires_grad_()
        ls = [loc.cuda().requires_grad_(), loc.flip(1).contiguous().cuda().requires_grad_()]
        aws = [aw.cuda().requires_grad_(), aw.cuda().requires_grad_()]
        shared = training.SharedGrads()
        outs = []
        for l, a in zip(ls, aws):
            if fused:
                outs.append(training.MSDAShared.apply(shared.get([v]), v, sp, lsi, l, a, 64))
            else:
                outs.append(MultiScaleDeformableAttnFunction_fp32.apply(v, sp, lsi, l, a, 64))
        (outs[0] * g).sum().add((outs[1] * g * 0.5).sum()).backward()
        return outs, v.grad, [t.grad for t in ls + aws]

    o1, gv1, r1 = run(True)
    o0, gv0, r0 = run(False)
    for a, b in zip(o1, o0):
        assert torch.equal(a, b)
    for a, b in zip(r1, r0):
        assert_close(a, b, 1e-5, 1e-6 * float(b.abs().max()), "msda shared grad_loc / grad_attn")
    assert_close(gv1, gv0, 1e-4, 1e-5 * float(gv0.abs().max()), "msda shared grad_value")


def test_decoder_training_fused_bindings_match_plain_autograd():
    """One training iteration pair of the small decoder: shared-gradient / grouped bindings on vs off."""
    from racformer_b200.decoder import RaCFormerTransformer
    from racformer_b200.synthetic import fill_parameters_by_name
    from tests.decoder_cases import SMALL, small_inputs
    res = []
    for fused in (True, False):
        torch.manual_seed(0)
        model = RaCFormerTransformer(**SMALL)
        model.init_weights()
        fill_parameters_by_name(model, seed=3)
        model = model.cuda().train()
        model.set_activation_checkpoint(False)
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
            if isinstance(m, torch.nn.MultiheadAttention):
                m.dropout = 0.0
        model.decoder.shared_gradients = fused
        inp = small_inputs(seed=5, device="cuda", batch=2)
        leaves = [inp["lss_bev"].requires_grad_(), inp["radar_bev"].requires_grad_()] + [f.requires_grad_() for f in inp["mlvl_feats"]]
        qf = inp["query_feat"].requires_grad_()
        cls, box = model(inp["query_bbox"], qf, inp["mlvl_feats"], inp["lss_bev"], inp["radar_bev"], None, inp["img_metas"])
        gen = torch.Generator().manual_seed(9)
        loss = (cls * torch.randn(cls.shape, generator=gen).cuda()).sum() + (box * torch.randn(box.shape, generator=gen).cuda()).sum()
        loss.backward()
        res.append((cls.detach(), box.detach(), [t.grad for t in leaves + [qf]], {n: p.grad for n, p in model.named_parameters()}))
    (c1, b1, g1, p1), (c0, b0, g0, p0) = res
    assert_close(c1, c0, 1e-5, 1e-5, "cls")
    assert_close(b1, b0, 1e-5, 1e-5, "box")
    for a, b in zip(g1, g0):
        assert_close(a, b, 1e-3, 1e-4 * float(b.abs().max()), "input gradient")
    for n in p0:
        if p0[n] is None:
            assert p1[n] is None or float(p1[n].abs().max()) == 0.0, n
            continue
        assert_close(p1[n], p0[n], 1e-3, 1e-4 * float(p0[n].abs().max()) + 1e-7, f"grad of {n}")


@pytest.mark.parametrize("p_in", [96, 32, 128, 48])
def test_adaptive_mixing_core_function_gradients_vs_fp64_autograd(p_in):
    """AdaptiveMixingCore (tcgen05 forward, recomputing CUDA-core backward) against the PyTorch chain of
    models/racformer_transformer.py:592-604 evaluated in fp64 with autograd."""
    import torch.nn.functional as F
    from racformer_b200 import training
    g = torch.Generator(device="cuda").manual_seed(11)
    QG, C, P_out = 40, 64, 128
    x = torch.randn(QG, p_in, C, device="cuda", generator=g)
    params = torch.randn(QG, C * C + P_out * p_in, device="cuda", generator=g) * 0.2
    gy = torch.randn(QG, P_out, C, device="cuda", generator=g)

    def chain(xx, pp):
        m, s = pp.split([C * C, P_out * p_in], 1)
        t = F.relu(F.layer_norm(torch.matmul(xx, m.reshape(QG, C, C)), [p_in, C]))
        return F.relu(F.layer_norm(torch.matmul(s.reshape(QG, P_out, p_in), t), [P_out, C]))

    x1, p1 = x.clone().requires_grad_(), params.clone().requires_grad_()
    y1 = training.AdaptiveMixingCore.apply(x1, p1, P_out)
    y1.backward(gy)
    x64, p64 = x.double().requires_grad_(), params.double().requires_grad_()
    y64 = chain(x64, p64)
    y64.backward(gy.double())
    x32, p32 = x.clone().requires_grad_(), params.clone().requires_grad_()
    chain(x32, p32).backward(gy)
    assert float((y1.double() - y64).abs().max()) < 2e-5
    for got, want, eager, name in ((x1.grad, x64.grad, x32.grad, "grad_x"), (p1.grad, p64.grad, p32.grad, "grad_params")):
        scale = float(want.abs().max())
        err = float((got.double() - want).abs().max()) / scale
        err_eager = float((eager.double() - want).abs().max()) / scale
        # elements whose pre-ReLU value is within rounding of 0 may take the other branch: same bar as PyTorch's own fp32 chain
        assert err <= max(3 * err_eager, 2e-5), (name, err, err_eager)


@pytest.mark.parametrize("QG,p_in", [(3, 96), (449, 96), (40, 16), (40, 80)])
def test_adaptive_mixing_backward_tensor_core_and_cuda_core_kernels_agree_with_fp64(QG, p_in):
    """Both backward kernels (csrc/mixing_bwd_tc.cu: all six products as tcgen05 MMAs on exact bf16x3 pieces; csrc/mixing_bwd.cu:
    fp32 FMA) against fp64 autograd of the chain, per gradient: 2e-6 of the gradient's largest magnitude."""
    import torch.nn.functional as F
    from racformer_b200 import points
    g = torch.Generator(device="cuda").manual_seed(5)
    C, P_out = 64, 128
    x = torch.randn(QG, p_in, C, device="cuda", generator=g)
    params = torch.randn(QG, C * C + P_out * p_in, device="cuda", generator=g) * 0.2
    gy = torch.randn(QG, P_out, C, device="cuda", generator=g)
    x64, p64 = x.double().requires_grad_(), params.double().requires_grad_()
    m, s = p64.split([C * C, P_out * p_in], 1)
    t = F.relu(F.layer_norm(torch.matmul(x64, m.reshape(QG, C, C)), [p_in, C]))
    F.relu(F.layer_norm(torch.matmul(s.reshape(QG, P_out, p_in), t), [P_out, C])).backward(gy.double())
    for variant in (1, 2):
        gx, gp = points.adaptive_mixing_core_backward(x, params, gy, P_out, variant=variant)
        for name, got, want in (("g_x", gx, x64.grad), ("g_M", gp[:, :C * C], p64.grad[:, :C * C]),
                                ("g_S", gp[:, C * C:], p64.grad[:, C * C:])):
            err = float((got.double() - want).abs().max() / want.abs().max())
            assert err < 2e-6, (variant, name, err)
    with pytest.raises(RuntimeError):       # the tensor-core kernel holds P_in <= 96 rows per tile
        points.adaptive_mixing_core_backward(torch.randn(2, 128, C, device="cuda"), torch.randn(2, C * C + P_out * 128, device="cuda"),
                                             torch.randn(2, P_out, C, device="cuda"), P_out, variant=2)


def _row_chain_case(dev, seed=0, rows=200, E=256, T=4, B=2):
    import torch.nn as nn
    torch.manual_seed(seed)
    mods = dict(l1=nn.Linear(3, E), n1=nn.LayerNorm(E), l2=nn.Linear(E, 2 * E), l3=nn.Linear(2 * E, E), n2=nn.LayerNorm(E),
                wide=nn.Linear(3 * E, E), narrow=nn.Linear(E, 10), proj=nn.Linear(E, E))
    for m in mods.values():
        m.to(dev)
        for p in m.parameters():
            p.data.normal_(0, 0.2) if p.dim() > 1 else p.data.normal_(0.5 if isinstance(m, nn.LayerNorm) and p is m.weight else 0, 0.3)
    g = torch.Generator(device=dev).manual_seed(seed + 1)
    Q = rows // B
    inp = dict(box=torch.rand(rows, 10, device=dev, generator=g), feat=torch.randn(rows, E, device=dev, generator=g),
               values=torch.randn(B * T, Q, E, device=dev, generator=g), logits=torch.randn(rows, T, device=dev, generator=g))
    return mods, inp, Q, T


def _row_chain_reference(mods, inp, Q, T, dt):
    import torch.nn.functional as F
    c = lambda t: t.to(dt)
    lin = lambda m, x: F.linear(x, c(m.weight), c(m.bias))
    ln = lambda m, x: F.layer_norm(x, m.normalized_shape, c(m.weight), c(m.bias), m.eps)
    box, feat, values, logits = (inp[k] for k in ("box", "feat", "values", "logits"))
    rows, E = feat.shape
    B = rows // Q
    a = F.relu(ln(mods["n1"], lin(mods["l1"], box[:, :3])))                    # partial-column load, K = 3 Linear, LN + ReLU
    hid = F.relu(lin(mods["l2"], a))                                           # Linear with fused ReLU (wide N)
    ffn = lin(mods["l3"], hid) + feat                                          # residual add of a loaded tensor
    out0 = ln(mods["n2"], ffn)
    w = torch.softmax(logits.view(B, Q, T), -1)                                # queue fusion
    fused = (values.view(B, T, Q, E) * w.permute(0, 2, 1)[..., None]).sum(1).reshape(rows, E)
    branch = lin(mods["proj"], fused)
    cat = torch.cat([out0, branch, feat], -1)                                  # concatenation by column offsets
    out1 = lin(mods["wide"], cat)
    out2 = lin(mods["narrow"], out1)                                           # narrow output
    return out0, out1, out2


def _row_chain_fused(mods, inp, Q, T):
    from racformer_b200 import rowtrain
    box, feat, values, logits = (inp[k] for k in ("box", "feat", "values", "logits"))
    rows, E = feat.shape
    p = rowtrain.RowChain(rows, width=3 * E, num_bufs=3)
    p.load(0, box, n=3)
    p.linear(1, 0, mods["l1"])
    p.layernorm(1, mods["n1"], relu=True)
    p.linear(2, 1, mods["l2"], relu=True)
    p.linear(0, 2, mods["l3"])
    p.load(1, feat)
    p.add(0, 1, E)
    p.layernorm(0, mods["n2"])
    h0 = p.store(0, E)
    p.load_queue(2, values, logits, Q, T)
    p.linear(0, 2, mods["proj"], dst_col=E)
    p.load(0, feat, dst_col=2 * E)
    p.linear(1, 0, mods["wide"])
    h1 = p.store(1, E)
    p.linear(2, 1, mods["narrow"])
    h2 = p.store(2, 10)
    outs = p.run()
    return outs[h0], outs[h1], outs[h2]


def test_row_chain_forward_and_generated_backward_match_fp64_autograd():
    """rowtrain.RowChain: every operator kind of the decoder's chains in one synthetic program; outputs, input gradients
    and parameter gradients against an fp64 PyTorch evaluation, with PyTorch's own fp32 chain as the yardstick."""
    dev = "cuda"
    mods, inp, Q, T = _row_chain_case(dev)
    gen = torch.Generator(device=dev).manual_seed(5)

    def run(kind):
        for m in mods.values():
            m.zero_grad(set_to_none=True)
        leaves = {k: inp[k].clone().requires_grad_(k != "box") for k in inp}
        if kind == "fused":
            outs = _row_chain_fused(mods, leaves, Q, T)
        else:
            dt = torch.float64 if kind == "fp64" else torch.float32
            outs = _row_chain_reference(mods, {k: v.to(dt) for k, v in leaves.items()}, Q, T, dt)
        gs = [torch.randn(o.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(9 + i)) for i, o in enumerate(outs)]
        sum((o * g.to(o.dtype)).sum() for o, g in zip(outs, gs)).backward()
        grads = {k: v.grad for k, v in leaves.items() if v.grad is not None}
        grads.update({f"{n}.{pn}": p.grad.clone() for n, m in mods.items() for pn, p in m.named_parameters()})
        return [o.detach() for o in outs], grads

    o_f, g_f = run("fused")
    o_64, g_64 = run("fp64")
    o_32, g_32 = run("fp32")
    for a, b, c in zip(o_f, o_64, o_32):
        err, ref = float((a.double() - b).abs().max()), float((c.double() - b).abs().max())
        assert err <= max(3 * ref, 2e-5), ("output", err, ref)
    assert set(g_f) == set(g_64)
    for k in g_64:
        scale = float(g_64[k].abs().max()) + 1e-12
        err = float((g_f[k].double() - g_64[k].double()).abs().max()) / scale
        ref = float((g_32[k].double() - g_64[k].double()).abs().max()) / scale
        assert err <= max(3 * ref, 2e-5), (k, err, ref)


def test_row_chain_dropout_mask_is_replayed_in_backward():
    """DROPOUT: keeps ~ (1 - p) of the elements scaled by 1 / (1 - p); the backward program applies the identical mask."""
    from racformer_b200 import rowtrain
    rows, E, p_drop = 128, 256, 0.25
    x = torch.ones(rows, E, device="cuda", requires_grad=True)
    chain = rowtrain.RowChain(rows, width=E, num_bufs=2)
    chain.load(0, x)
    chain.dropout(0, E, p_drop, seed=1234)
    h = chain.store(0, E)
    y = chain.run()[h]
    keep = y > 0
    assert abs(float(keep.float().mean()) - (1 - p_drop)) < 0.02
    assert torch.allclose(y[keep], torch.full_like(y[keep], 1 / (1 - p_drop)))
    y.backward(torch.full_like(y, 2.0))
    assert torch.equal(x.grad > 0, keep) and torch.allclose(x.grad[keep], torch.full_like(x.grad[keep], 2 / (1 - p_drop)))
    chain2 = rowtrain.RowChain(rows, width=E, num_bufs=2)
    chain2.load(0, x.detach())
    chain2.dropout(0, E, p_drop, seed=99)
    h2 = chain2.store(0, E)
    assert not torch.equal(chain2.run()[h2] > 0, keep)                            # another seed, another mask


def test_decoder_training_row_chains_match_pytorch_modules():
    """Small decoder, one training step with dropout off: row chains with autograd (train_row_programs) vs the PyTorch modules."""
    from racformer_b200.decoder import RaCFormerTransformer
    from racformer_b200.synthetic import fill_parameters_by_name
    from tests.decoder_cases import SMALL, small_inputs
    res = []
    for fused in (True, False):
        torch.manual_seed(0)
        model = RaCFormerTransformer(**SMALL)
        model.init_weights()
        fill_parameters_by_name(model, seed=3)
        model = model.cuda().train()
        model.set_activation_checkpoint(False)
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
            if isinstance(m, torch.nn.MultiheadAttention):
                m.dropout = 0.0
        model.decoder.decoder_layer.train_row_programs = fused
        inp = small_inputs(seed=5, device="cuda", batch=2)
        leaves = [inp["lss_bev"].requires_grad_(), inp["radar_bev"].requires_grad_()] + [f.requires_grad_() for f in inp["mlvl_feats"]]
        qf = inp["query_feat"].requires_grad_()
        q = qf.shape[1]
        mask = torch.zeros(q, q, dtype=torch.bool, device="cuda")
        mask[q // 2:, :q // 2] = True                                            # a denoising-style blocked region
        cls, box = model(inp["query_bbox"], qf, inp["mlvl_feats"], inp["lss_bev"], inp["radar_bev"], mask, inp["img_metas"])
        gen = torch.Generator().manual_seed(9)
        loss = (cls * torch.randn(cls.shape, generator=gen).cuda()).sum() + (box * torch.randn(box.shape, generator=gen).cuda()).sum()
        loss.backward()
        res.append((cls.detach(), box.detach(), [t.grad for t in leaves + [qf]], {n: p.grad for n, p in model.named_parameters()}))
    (c1, b1, g1, p1), (c0, b0, g0, p0) = res
    assert_close(c1, c0, 1e-4, 2e-4, "cls")
    assert_close(b1, b0, 1e-4, 2e-4, "box")
    # two decoder iterations: a last-ulp difference in the first iteration's output moves second-iteration sample points
    # across bilinear cell borders, which moves gradient mass between neighbouring cells / parameters -- the same bar as the
    # comparison with the unchanged reference in tests/test_decoder.py (3e-2 * max|g| for two iterations)
    for a, b in zip(g1, g0):
        assert_close(a, b, 2e-3, 3e-2 * float(b.abs().max()), "input gradient")
    for n in p0:
        if p0[n] is None:
            assert p1[n] is None or float(p1[n].abs().max()) == 0.0, n
            continue
        assert p1[n] is not None, n
        assert_close(p1[n], p0[n], 2e-3, 3e-2 * float(p0[n].abs().max()) + 1e-7, f"grad of {n}")


def _sasa_case(B=2, Q=150, H=8, dn=40, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    E = H * 32
    qkv = torch.randn(B, Q, 3 * E, device="cuda", generator=g) * 0.7
    tau = torch.rand(B, Q, H, device="cuda", generator=g) * 2
    ray = torch.rand(B, Q, 10, device="cuda", generator=g)
    blocked = torch.zeros(Q, Q, dtype=torch.bool, device="cuda")
    blocked[dn:, :dn] = True                       # matching queries do not see the denoising queries (a whole key slice)
    blocked[:dn // 2, dn // 2:dn] = True           # denoising groups do not see each other
    blocked[dn // 2:dn, :dn // 2] = True
    return qkv, tau, ray, blocked


@pytest.mark.parametrize("masked", [False, True])
def test_sasa_training_core_matches_pytorch_attention_with_autograd(masked):
    """csrc/sasa_train.cu (forward + two backward kernels) vs ScaleAdaptiveSelfAttention's PyTorch path (materialised
    [B,H,Q,Q] mask + scaled_dot_product_attention) in fp64 with autograd; PyTorch's own fp32 path is the yardstick."""
    from racformer_b200.decoder import ScaleAdaptiveSelfAttention
    from racformer_b200.synthetic import PC_RANGE
    qkv, tau, ray, blocked = _sasa_case()
    mask = blocked if masked else None
    sa = ScaleAdaptiveSelfAttention(256, 8, 0.1, PC_RANGE).cuda()
    gout = torch.randn(qkv.shape[0], qkv.shape[1], 256, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))

    def run(fused, dt):
        sa.fused_training_core = fused
        a, t = qkv.detach().clone().to(dt).requires_grad_(), tau.detach().clone().to(dt).requires_grad_()
        out = sa.attention_core(ray.to(dt), a, t, mask, 8, 0.0)
        out.backward(gout.to(dt))
        return out.detach(), a.grad, t.grad

    got, ref64, ref32 = run(True, torch.float32), run(False, torch.float64), run(False, torch.float32)
    for name, a, b, c in zip(("out", "grad_qkv", "grad_tau"), got, ref64, ref32):
        scale = float(b.abs().max())
        err, yard = float((a.double() - b).abs().max()) / scale, float((c.double() - b).abs().max()) / scale
        assert err <= max(4 * yard, 2e-5), (name, err, yard)


def test_sasa_training_core_dropout_is_consistent_between_forward_and_backward():
    """With attention dropout the op is a deterministic function of (inputs, seed): its analytic gradient must agree with a
    central finite difference of the kernel itself, the mean output must stay near the no-dropout output, and another seed
    must give another mask."""
    from racformer_b200 import training
    from racformer_b200.synthetic import PC_RANGE
    qkv, tau, ray, blocked = _sasa_case(B=1, Q=96, dn=24, seed=4)
    gen = torch.Generator(device="cuda").manual_seed(8)
    gout = torch.randn(1, 96, 256, device="cuda", generator=gen)
    f = lambda a, t, seed=77, p=0.3: training.SasaAttention.apply(a, t, ray, blocked, PC_RANGE, 8, p, seed)
    a, t = qkv.clone().requires_grad_(), tau.clone().requires_grad_()
    out = f(a, t)
    (out * gout).sum().backward()
    da, dt_ = torch.randn(qkv.shape, device="cuda", generator=gen), torch.randn(tau.shape, device="cuda", generator=gen)
    eps = 1e-2
    with torch.no_grad():
        fd = ((f(qkv + eps * da, tau + eps * dt_).double() - f(qkv - eps * da, tau - eps * dt_).double()) * gout.double()).sum() / (2 * eps)
    analytic = (a.grad.double() * da.double()).sum() + (t.grad.double() * dt_.double()).sum()
    assert abs(float(fd - analytic)) <= 2e-2 * abs(float(analytic)) + 1e-3, (float(fd), float(analytic))
    with torch.no_grad():
        plain = f(qkv, tau, p=0.0)
        mean = torch.stack([f(qkv, tau, seed=s) for s in range(40)]).mean(0)
        assert float((mean - plain).abs().mean()) < 0.15 * float(plain.abs().mean())
        assert not torch.equal(f(qkv, tau, seed=1), f(qkv, tau, seed=2))
