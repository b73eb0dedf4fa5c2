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


@pytest.mark.parametrize("p_in", [96, 32, 128])
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
