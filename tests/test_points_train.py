"""GPU: the training-mode point-generation / box-refinement Functions (racformer_b200/training.py MSMVPoints, BEVPoints,
RefineBBox; kernels csrc/points.cu + csrc/points_train.cu) against fp64 autograd of the harness's PyTorch chain, which
tests/test_decoder.py pins to the unchanged reference (models/racformer_transformer.py:361-408, 493-529, 255-279).

Tolerance: gradients within 3e-4 of the reference gradient tensor's largest magnitude (fp32 kernels vs fp64 autograd).
Points whose camera view differs between the fp32 kernel and the fp64 chain (projections within an ulp of an image
border) or that no camera sees get no incoming gradient, as msmv_sampling's backward gives them none."""
import pytest
import torch

from racformer_b200.decoder import BEVSampling, RaCFormerSampling, RaCFormerTransformer
from racformer_b200.synthetic import PC_RANGE, fill_parameters_by_name
from tests.test_points_bwd_host import _CaptureOps, _lidar2img, _rays

pytestmark = pytest.mark.gpu


def _close(got, want, name, tol=3e-4):
    scale = float(want.abs().max())
    assert scale > 0, f"{name}: reference gradient is identically zero"
    err = float((got.double() - want).abs().max())
    assert err <= tol * scale, f"{name}: max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("shape", [(2, 37, 3, 2, 2, 3, 3, 4), (1, 50, 8, 4, 4, 3, 6, 4), (2, 5, 2, 4, 8, 15, 2, 2)])
def test_msmv_points_function_gradients_vs_fp64_chain(shape):
    from racformer_b200 import training
    B, Q, T, G, Pn, D, N, L = shape
    P = Pn * D
    g = torch.Generator().manual_seed(11)
    mod = RaCFormerSampling(embed_dims=16, num_frames=T, num_groups=G, num_points=Pn, num_levels=L, depth_num=D, pc_range=PC_RANGE)
    mod.fused_points = False
    ray64 = _rays(B, Q, g).float().double()
    off64 = (torch.rand(B, Q, G * P * 3, generator=g) - 0.5).double()
    logit64 = torch.randn(B, Q, D, generator=g).double()
    sw64 = torch.randn(B, Q, G * T * P * L, generator=g).double()
    time_diff = (torch.rand(B, T, generator=g) * 0.5)
    time_diff[:, 0] = 0
    l2i = _lidar2img(B, T, N, g).float()
    d_region = 0.08
    depth_base = torch.linspace(-d_region, d_region, D, device="cuda")
    geom = (tuple(PC_RANGE), d_region, 704.0, 256.0, T, G, Pn, D, L)

    leaves32 = [t.float().cuda().requires_grad_() for t in (ray64, off64, logit64, sw64)]
    loc, w = training.MSMVPoints.apply(*leaves32, time_diff.cuda(), l2i.cuda(), depth_base, geom)

    leaves64 = [t.cuda().requires_grad_() for t in (ray64, off64, logit64, sw64)]
    meta = {"time_diff": time_diff.double().cuda(), "lidar2img": l2i.double().cuda(), "image_h": 256.0, "image_w": 704.0}
    ops = _CaptureOps()
    mod.inner_forward(ops, leaves64[0], torch.zeros(B, Q, 16, dtype=torch.float64, device="cuda"),
                      [torch.zeros(B * T * G, N, 2, 2, 4, dtype=torch.float64, device="cuda")], meta, d_region,
                      heads=tuple(leaves64[1:]))
    loc64, w64 = ops.loc, ops.w
    assert loc.shape == loc64.shape and w.shape == w64.shape
    same_view = (loc[..., 2].double() - loc64[..., 2]).abs() < 1e-6
    seen = (loc64[..., 0] > 0) & (loc64[..., 0] < 1) & (loc64[..., 1] > 0) & (loc64[..., 1] < 1)
    live = same_view & seen
    assert float(same_view.double().mean()) > 0.999 and float(live.double().mean()) > 0.05
    assert float((loc[..., :2].double() - loc64[..., :2])[live].abs().max()) < 2e-5
    assert float((w.double() - w64).abs().max()) < 2e-6

    g_loc = torch.randn(loc.shape, generator=g).cuda() * live.unsqueeze(-1)
    g_loc[..., 2] = 0
    g_w = torch.randn(w.shape, generator=g).cuda()
    torch.autograd.backward([loc, w], [g_loc, g_w])
    torch.autograd.backward([loc64, w64], [g_loc.double(), g_w.double()])
    for a, b, name in zip(leaves32, leaves64, ("grad_ray", "grad_offset", "grad_ray_logit", "grad_scale_raw")):
        _close(a.grad, b.grad, name)


@pytest.mark.parametrize("shape", [(2, 41, 3, 4, 2, 3), (1, 30, 8, 4, 4, 5), (2, 6, 2, 4, 8, 30)])
def test_bev_points_function_gradients_vs_fp64_chain(shape):
    from racformer_b200 import training
    B, Q, T, M, Pn, D = shape
    P = Pn * D
    g = torch.Generator().manual_seed(12)
    mod = BEVSampling(embed_dims=16, num_frames=T, num_points=Pn, num_heads=M, num_levels=1, pc_range=PC_RANGE, depth_num=D)
    mod.fused_points = False
    ray64 = _rays(B, Q, g).float().double()
    off64 = (torch.rand(B, Q, M * P * 2, generator=g) - 0.5).double()
    logit64 = torch.randn(B, Q, D, generator=g).double()
    sw64 = torch.randn(B, Q, M * P, generator=g).double()
    time_diff = torch.rand(B, T, generator=g) * 0.5
    time_diff[:, 0] = 0
    d_region = 0.08
    depth_base = torch.linspace(-d_region, d_region, D, device="cuda")

    leaves32 = [t.float().cuda().requires_grad_() for t in (ray64, off64, logit64, sw64)]
    loc, aw = training.BEVPoints.apply(*leaves32, time_diff.cuda(), depth_base, (tuple(PC_RANGE), d_region, T, M, Pn, D))

    leaves64 = [t.cuda().requires_grad_() for t in (ray64, off64, logit64, sw64)]
    ops = _CaptureOps()
    mod.sample(ops, leaves64[0], torch.zeros(B, Q, M * 8, dtype=torch.float64, device="cuda"),
               torch.zeros(B * T, 16, M, 8, dtype=torch.float64, device="cuda"), (4, 4),
               {"time_diff": time_diff.double().cuda()}, d_region,
               heads=(leaves64[1], leaves64[2], leaves64[3], torch.zeros(B, Q, T, dtype=torch.float64, device="cuda")), raw=True)
    loc64, aw64 = ops.loc, ops.aw
    assert loc.shape == loc64.shape and aw.shape == aw64.shape
    assert float((loc.double() - loc64).abs().max()) < 2e-6 and float((aw.double() - aw64).abs().max()) < 2e-6
    g_loc = torch.randn(loc.shape, generator=g).cuda()
    g_aw = torch.randn(aw.shape, generator=g).cuda()
    torch.autograd.backward([loc, aw], [g_loc, g_aw])
    torch.autograd.backward([loc64, aw64], [g_loc.double(), g_aw.double()])
    for a, b, name in zip(leaves32, leaves64, ("grad_ray", "grad_offset", "grad_ray_logit", "grad_attn_raw")):
        _close(a.grad, b.grad, name)


@pytest.mark.parametrize("frames", [1, 4])
def test_refine_bbox_function_vs_fp64_chain(frames):
    from racformer_b200 import training
    from racformer_b200.decoder import RaCFormerTransformerDecoderLayer, theta_d2xy_coods
    B, Q = 2, 77
    g = torch.Generator().manual_seed(13)
    prop64 = _rays(B, Q, g).float().double()
    prop64[..., 2] = torch.rand(B, Q, generator=g).double()
    delta64 = torch.randn(B, Q, 10, generator=g).double()
    time_diff = torch.rand(B, frames, generator=g) * 0.5 + 0.05
    time_diff[:, 0] = 0
    p32, d32 = prop64.float().cuda().requires_grad_(), delta64.float().cuda().requires_grad_()
    pred, xy = training.RefineBBox.apply(p32, d32, time_diff.cuda(), 150)
    assert not pred.requires_grad

    p64, d64 = prop64.cuda().requires_grad_(), delta64.cuda().requires_grad_()
    import types
    pred64 = RaCFormerTransformerDecoderLayer.refine_bbox(types.SimpleNamespace(num_ray=150), p64, d64)
    if frames > 1:
        td = time_diff.double().cuda()
        td = torch.where(td < 1e-5, torch.ones_like(td), td)
        pred64 = torch.cat([pred64[..., :8], pred64[..., 8:] / td[:, 1:2, None]], dim=-1)
    xy64 = theta_d2xy_coods(pred64)
    assert float((pred.double() - pred64).abs().max()) < 1e-5 and float((xy.double() - xy64).abs().max()) < 1e-5
    gxy = torch.randn(xy.shape, generator=g).cuda()
    xy.backward(gxy)
    xy64.backward(gxy.double())
    _close(d32.grad, d64.grad, "grad_delta")
    _close(p32.grad, p64.grad, "grad_proposal")


def test_decoder_training_fused_points_match_pytorch_chain():
    """Small decoder, one training step with dropout off: fused point / refinement Functions on vs the PyTorch op chain."""
    from tests.decoder_cases import SMALL, small_inputs
    from tests.helpers import assert_close
    res = []
    for fused in (True, False):
        torch.manual_seed(0)
        model = RaCFormerTransformer(**SMALL)
        model.init_weights()
        fill_parameters_by_name(model, seed=3)
        model = model.cuda().train()
        model.set_activation_checkpoint(False)
        model.set_fused_points(fused)
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
            if isinstance(m, torch.nn.MultiheadAttention):
                m.dropout = 0.0
        inp = small_inputs(seed=5, device="cuda", batch=2)
        leaves = [inp["lss_bev"].requires_grad_(), inp["radar_bev"].requires_grad_()] + [f.requires_grad_() for f in inp["mlvl_feats"]]
        qf = inp["query_feat"].requires_grad_()
        cls, box = model(inp["query_bbox"], qf, inp["mlvl_feats"], inp["lss_bev"], inp["radar_bev"], None, inp["img_metas"])
        gen = torch.Generator().manual_seed(9)
        loss = (cls * torch.randn(cls.shape, generator=gen).cuda()).sum() + (box * torch.randn(box.shape, generator=gen).cuda()).sum()
        loss.backward()
        res.append((cls.detach(), box.detach(), [t.grad for t in leaves + [qf]], {n: p.grad for n, p in model.named_parameters()}))
    (c1, b1, g1, p1), (c0, b0, g0, p0) = res
    assert_close(c1, c0, 1e-4, 2e-4, "cls")
    assert_close(b1, b0, 1e-4, 2e-4, "box")
    # same bar as the row-chain comparison (tests/test_training_ops.py): last-ulp differences in iteration 1 move
    # iteration-2 sample points across bilinear cell borders
    for a, b in zip(g1, g0):
        assert_close(a, b, 2e-3, 3e-2 * float(b.abs().max()), "input gradient")
    for n in p0:
        if p0[n] is None:
            assert p1[n] is None or float(p1[n].abs().max()) == 0.0, n
            continue
        assert p1[n] is not None, n
        assert_close(p1[n], p0[n], 2e-3, 3e-2 * float(p0[n].abs().max()) + 1e-7, f"grad of {n}")
