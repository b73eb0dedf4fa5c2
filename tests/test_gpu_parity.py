"""GPU parity tests: the sm_100a kernels, called through the C ABI via the reference-shaped Python API, against
(i) the committed fixtures generated from the reference, (ii) the exact-arithmetic C oracle on seeded inputs,
(iii) size-independent properties at the full racformer_r50_nuimg_704x256_f8 shapes, and (iv) the reference's own
CUDA extension where oracle/_ref was built.

Tolerances (also in DESIGN.md):
  forward  vs exact-arithmetic oracle : rtol 1e-5, atol 1e-6 * max|feat|
  forward  vs reference grid_sample   : rtol 1e-5, atol 5e-5 * max|feat|   (reference-vs-reference gap)
  backward vs fp64-accumulating oracle: rtol 1e-4, atol 1e-5 * max|expected| (fp32 atomics / shuffle trees reorder sums)
  view indices and validity masks     : bit-exact
"""
import pytest
import torch

from tests.helpers import (BWD_ATOL, BWD_RTOL, F8_SHAPES, FWD_ATOL_EXACT, FWD_ATOL_GRIDSAMPLE, FWD_RTOL, MSDA_CASES,
                           MSMV_CASES, assert_close, load_golden, make_msda_inputs, make_msmv_inputs, msmv_feats,
                           near_integer_pixel_msda, near_integer_pixel_msmv)

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _scale(t):
    return max(float(t.abs().max()), 1e-30)


def _ops():
    from racformer_b200 import wrapper
    from racformer_b200 import multi_scale_deformable_attn_function as msda
    return wrapper, msda


def _oracle():
    from oracle import c_oracle
    return c_oracle


# ------------------------------------------------------------------------------------------------ MSMV
def _check_msmv_against_oracle(feats, loc, w, grad_out, what):
    wrapper, _ = _ops()
    co = _oracle()
    fmax = max(_scale(f) for f in feats)
    f_d = [f.to(DEV) for f in feats]
    out = wrapper.msmv_forward(f_d, loc.to(DEV), w.to(DEV))
    ref_out, ref_view, ref_mask = co.msmv_forward(feats, loc, w, with_masks=True)
    assert_close(out, ref_out, FWD_RTOL, FWD_ATOL_EXACT * fmax, f"{what} forward vs oracle")
    view, mask = wrapper.msmv_tap_masks([f.shape[2:4] for f in feats], loc.to(DEV), feats[0].shape[1])
    assert torch.equal(view.cpu(), ref_view), f"{what}: view indices differ"
    assert torch.equal(mask.cpu(), ref_mask), f"{what}: validity masks differ"
    grads = wrapper.msmv_backward(grad_out.to(DEV), f_d, loc.to(DEV), w.to(DEV))
    rgf, rgl, rgw = co.msmv_backward(grad_out, feats, loc, w)
    for i, (g, r) in enumerate(zip(grads[:len(feats)], rgf)):
        assert_close(g, r, BWD_RTOL, BWD_ATOL * _scale(r), f"{what} grad_feat{i} vs oracle")
    assert_close(grads[-2], rgl, BWD_RTOL, BWD_ATOL * _scale(rgl), f"{what} grad_loc vs oracle")
    assert_close(grads[-1], rgw, BWD_RTOL, BWD_ATOL * _scale(rgw), f"{what} grad_w vs oracle")
    assert bool((grads[-2][..., 2] == 0).all()), "grad of the view coordinate must be exactly 0"
    return out, grads


@pytest.mark.parametrize("case", MSMV_CASES)
def test_msmv_golden_fixture(case):
    d = load_golden(case)
    feats = msmv_feats(d)
    out, grads = _check_msmv_against_oracle(feats, d["loc"], d["w"], d["grad_out"], case)
    fmax = max(_scale(f) for f in feats)
    assert_close(out, d["out"], FWD_RTOL, FWD_ATOL_GRIDSAMPLE * fmax, f"{case} forward vs reference grid_sample")
    for i in range(len(feats)):
        assert_close(grads[i], d[f"grad_feat{i}"], BWD_RTOL, 1e-4 * _scale(d[f"grad_feat{i}"]), f"{case} grad_feat{i} vs reference")
    assert_close(grads[-1], d["grad_w"], BWD_RTOL, 1e-4 * _scale(d["grad_w"]), f"{case} grad_w vs reference")
    ok = ~near_integer_pixel_msmv(d["loc"], [f.shape[2:4] for f in feats])
    assert_close(grads[-2].cpu()[..., :2][ok], d["grad_loc"][..., :2][ok], BWD_RTOL,
                 1e-4 * _scale(d["grad_loc"][..., :2][ok]), f"{case} grad_loc vs reference")
    wrapper, _ = _ops()
    v2 = wrapper.msmv_sampling_v2([f.to(DEV) for f in feats], d["loc"].to(DEV), d["w"].to(DEV))
    assert_close(v2, d["out_v2"], FWD_RTOL, FWD_ATOL_GRIDSAMPLE * fmax, f"{case} v2 vs reference")


@pytest.mark.parametrize("levels,C,P,lo,hi", [
    (4, 64, 12, 0.0, 1.0),      # fast path, RaCFormer point count, everything valid
    (4, 64, 12, -0.1, 1.1),     # mixed validity
    (4, 64, 7, -0.3, 1.3),      # ragged P (scalar store path)
    (2, 64, 12, -0.1, 1.1),     # c45
    (5, 64, 4, -0.1, 1.1),      # c23456
    (3, 64, 5, -0.1, 1.1),      # level count without a fast kernel -> generic path
    (4, 32, 6, -0.1, 1.1),      # generic path (C != 64)
    (4, 64, 12, -5.0, -2.0),    # nothing valid at all: zeros everywhere
    (4, 64, 128, -0.1, 1.1),    # maximum point count (msmv_sampling.cpp:159), several staged batches per query
    (4, 64, 16, -0.1, 1.1),     # exactly one staged batch incl. the run-ahead pad point
    (4, 64, 20, -0.1, 1.1),     # more than one batch -> non-persistent kernel
    (5, 64, 12, -0.1, 1.1),     # c23456 with the RaCFormer point count (12-point batches)
    (1, 64, 3, -0.1, 1.1),      # single level (generic path)
    (8, 4, 2, -0.1, 1.1),       # maximum level count, tiny channel count
    (4, 1, 1, 0.0, 1.0),        # degenerate C = 1, P = 1
])
def test_msmv_seeded_vs_oracle(levels, C, P, lo, hi):
    shapes = [(16, 44), (8, 22), (4, 11), (2, 6), (1, 3), (3, 2), (2, 2), (1, 1)][:levels]
    feats, loc, w, g = make_msmv_inputs(11 + levels + P, Bp=3, N=6, C=C, Q=37, P=P, shapes=shapes, lo=lo, hi=hi)
    out, grads = _check_msmv_against_oracle(feats, loc, w, g, f"L{levels} C{C} P{P} [{lo},{hi}]")
    if hi < 0:
        assert float(out.abs().max()) == 0.0
        assert all(float(t.abs().max()) == 0.0 for t in grads)


@pytest.mark.parametrize("Bp,N,Q", [(1, 2, 1), (1, 2, 9), (7, 2, 1), (2, 6, 257)])
def test_msmv_small_and_ragged_batches(Bp, N, Q):
    """Single query / single batch element / fewer queries than a CTA holds / a query count that is not a multiple of it."""
    feats, loc, w, g = make_msmv_inputs(Bp * 31 + Q, Bp=Bp, N=N, C=64, Q=Q, P=12, shapes=[(8, 22), (4, 11), (2, 6), (1, 3)],
                                        lo=-0.1, hi=1.1)
    _check_msmv_against_oracle(feats, loc, w, g, f"Bp{Bp} N{N} Q{Q}")


def test_msmv_special_coordinates_and_views():
    """NaN / Inf / out-of-range view coordinates contribute exactly zero and never fault."""
    feats, loc, w, g = make_msmv_inputs(5, Bp=2, N=3, C=64, Q=8, P=4, shapes=[(8, 12), (4, 6), (2, 3), (1, 2)])
    loc[0, 0, 0, 0] = float("nan")
    loc[0, 0, 1, 1] = float("inf")
    loc[0, 0, 2, 0] = -float("inf")
    loc[0, 1, 0, 2] = 1.5      # view index 3 >= N: reference would read out of bounds; we define it as "no contribution"
    loc[0, 1, 1, 2] = -0.5     # view index -1
    out, grads = _check_msmv_against_oracle(feats, loc, w, g, "special coordinates")
    assert torch.isfinite(out).all()
    assert float(out[0, 0, :, :3].abs().max()) == 0.0
    assert float(out[0, 1, :, :2].abs().max()) == 0.0


def test_msmv_f8_full_size_forward_vs_oracle_and_properties():
    """racformer_r50_nuimg_704x256_f8, batch 1: B'=32, N=6, C=64, Q=900, P=12 (mixed case)."""
    wrapper, _ = _ops()
    co = _oracle()
    feats, loc, w, g = make_msmv_inputs(0, Bp=32, N=6, C=64, Q=900, P=12, shapes=F8_SHAPES, lo=-0.1, hi=1.1)
    f_d, loc_d, w_d, g_d = [f.to(DEV) for f in feats], loc.to(DEV), w.to(DEV), g.to(DEV)
    out = wrapper.msmv_forward(f_d, loc_d, w_d)
    assert torch.equal(out, wrapper.msmv_forward(f_d, loc_d, w_d)), "forward must be deterministic"
    ref_out, ref_view, ref_mask = co.msmv_forward(feats, loc, w, with_masks=True)
    assert_close(out, ref_out, FWD_RTOL, FWD_ATOL_EXACT * max(_scale(f) for f in feats), "f8 forward vs oracle")
    view, mask = wrapper.msmv_tap_masks(F8_SHAPES, loc_d, 6)
    assert torch.equal(view.cpu(), ref_view) and torch.equal(mask.cpu(), ref_mask)
    # adjoint identities: the op is linear in the features and in the weights
    grads = wrapper.msmv_backward(g_d, f_d, loc_d, w_d)
    lhs = float((out.double() * g_d.double()).sum())
    rhs_feat = sum(float((gf.double() * f.double()).sum()) for gf, f in zip(grads[:4], f_d))
    rhs_w = float((grads[-1].double() * w_d.double()).sum())
    assert rhs_feat == pytest.approx(lhs, rel=1e-5, abs=1e-2)
    assert rhs_w == pytest.approx(lhs, rel=1e-5, abs=1e-2)
    # grad_loc / grad_w against the oracle on the first two batch elements (full rows, cheap)
    sub = slice(0, 2)
    rgf, rgl, rgw = co.msmv_backward(g[sub], [f[sub] for f in feats], loc[sub], w[sub])
    assert_close(grads[-2][sub], rgl, BWD_RTOL, BWD_ATOL * _scale(rgl), "f8 grad_loc vs oracle")
    assert_close(grads[-1][sub], rgw, BWD_RTOL, BWD_ATOL * _scale(rgw), "f8 grad_w vs oracle")
    for i in range(4):
        assert_close(grads[i][sub], rgf[i], BWD_RTOL, BWD_ATOL * _scale(rgf[i]), f"f8 grad_feat{i} vs oracle")
    # constant features + weights summing to one -> constant output wherever all four corners exist on all levels
    ones = [torch.ones_like(f) for f in f_d]
    o1 = wrapper.msmv_forward(ones, loc_d, w_d)
    all_valid = (mask == 31).all(-1)[:, :, None, :].expand_as(o1)
    assert float((o1[all_valid] - 1).abs().max()) < 1e-5


def test_msmv_training_and_3cam_shapes():
    """B=2 training shape (B'=64, Q=1220) and the 3cam variant (N=3): adjoint identity + determinism."""
    wrapper, _ = _ops()
    for Bp, N, Q in ((64, 6, 1220), (32, 3, 900)):
        feats, loc, w, g = make_msmv_inputs(3, Bp=Bp, N=N, C=64, Q=Q, P=12, shapes=F8_SHAPES, device=DEV)
        out = wrapper.msmv_forward(feats, loc, w)
        grads = wrapper.msmv_backward(g, feats, loc, w)
        lhs = float((out.double() * g.double()).sum())
        rhs = sum(float((gf.double() * f.double()).sum()) for gf, f in zip(grads[:4], feats))
        assert rhs == pytest.approx(lhs, rel=1e-5, abs=1e-2)
        del feats, grads, out
        torch.cuda.empty_cache()


def test_msmv_autograd_function_and_streams():
    wrapper, _ = _ops()
    co = _oracle()
    feats, loc, w, g = make_msmv_inputs(7, Bp=2, N=6, C=64, Q=20, P=12, shapes=[(16, 44), (8, 22), (4, 11), (2, 6)],
                                        lo=-0.1, hi=1.1)
    f_d = [f.to(DEV).requires_grad_() for f in feats]
    loc_d, w_d = loc.to(DEV).requires_grad_(), w.to(DEV).requires_grad_()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        out = wrapper.msmv_sampling(f_d, loc_d, w_d)
        # a non-contiguous upstream gradient exercises the .contiguous() in backward (wrapper.py:93)
        (out.permute(0, 1, 3, 2) * g.to(DEV).permute(0, 1, 3, 2)).sum().backward()
    torch.cuda.current_stream().wait_stream(side)
    rgf, rgl, rgw = co.msmv_backward(g, feats, loc, w)
    for f, r in zip(f_d, rgf):
        assert_close(f.grad, r, BWD_RTOL, BWD_ATOL * _scale(r), "autograd grad_feat")
    assert_close(loc_d.grad, rgl, BWD_RTOL, BWD_ATOL * _scale(rgl), "autograd grad_loc")
    assert_close(w_d.grad, rgw, BWD_RTOL, BWD_ATOL * _scale(rgw), "autograd grad_w")


def test_msmv_empty_and_errors():
    wrapper, _ = _ops()
    feats = [torch.zeros(2, 6, 4, 4, 64, device=DEV), torch.zeros(2, 6, 2, 2, 64, device=DEV)]
    out = wrapper.msmv_sampling(feats, torch.zeros(2, 0, 12, 3, device=DEV), torch.zeros(2, 0, 12, 2, device=DEV))
    assert out.shape == (2, 0, 64, 12)
    with pytest.raises(RuntimeError, match="num_point exceed limits"):
        wrapper.msmv_sampling(feats, torch.zeros(2, 1, 129, 3, device=DEV), torch.zeros(2, 1, 129, 2, device=DEV))
    with pytest.raises(RuntimeError, match="attn_weight must be"):
        wrapper.msmv_sampling(feats, torch.zeros(2, 1, 4, 3, device=DEV), torch.zeros(2, 1, 4, 4, device=DEV))
    with pytest.raises(RuntimeError, match="Float"):
        wrapper.msmv_sampling([f.half() for f in feats], torch.zeros(2, 1, 4, 3, device=DEV),
                              torch.zeros(2, 1, 4, 2, device=DEV))


def test_msmv_vs_reference_cuda_extension():
    """The reference's own kernel (oracle/_ref, built by oracle/build_ref.py from /root/reference) on the same inputs."""
    from oracle import build_ref
    ext = build_ref.load_prebuilt()
    if ext is None:
        pytest.skip("oracle/_ref/_msmv_sampling_cuda.so not built")
    wrapper, _ = _ops()
    feats, loc, w, g = make_msmv_inputs(0, Bp=8, N=6, C=64, Q=300, P=12, shapes=F8_SHAPES, lo=-0.1, hi=1.1, device=DEV)
    out = wrapper.msmv_forward(feats, loc, w)
    ref = ext._ms_deform_attn_cuda_c2345_forward(*feats, loc, w)
    torch.cuda.synchronize()
    assert_close(out, ref, FWD_RTOL, FWD_ATOL_EXACT * max(_scale(f) for f in feats), "forward vs reference CUDA ext")
    grads = wrapper.msmv_backward(g, feats, loc, w)
    rgrads = ext._ms_deform_attn_cuda_c2345_backward(g, *feats, loc, w)
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(grads, rgrads)):
        assert_close(a, b, BWD_RTOL, 2e-5 * _scale(b), f"backward[{i}] vs reference CUDA ext")


# ------------------------------------------------------------------------------------------------ MSDA
def _check_msda_against_oracle(value, shapes, lsi, loc, aw, grad_out, what, im2col_step=64):
    _, msda = _ops()
    co = _oracle()
    vmax = _scale(value)
    args_d = [t.to(DEV) for t in (value, shapes, lsi, loc, aw)]
    out = msda.ext_module.ms_deform_attn_forward(*args_d, im2col_step=im2col_step)
    ref_out, ref_mask = co.msda_forward(value, shapes, lsi, loc, aw, with_masks=True)
    assert_close(out, ref_out, FWD_RTOL, FWD_ATOL_EXACT * vmax, f"{what} forward vs oracle")
    mask = msda.msda_tap_masks(args_d[1], args_d[3])
    assert torch.equal(mask.cpu(), ref_mask), f"{what}: validity masks differ"
    gv = torch.zeros_like(args_d[0])
    gl = torch.full_like(args_d[3], 7.0)   # must be fully overwritten
    ga = torch.full_like(args_d[4], 7.0)
    msda.ext_module.ms_deform_attn_backward(*args_d, grad_out.to(DEV), gv, gl, ga, im2col_step=im2col_step)
    rgv, rgl, rga = co.msda_backward(value, shapes, lsi, loc, aw, grad_out)
    assert_close(gv, rgv, BWD_RTOL, BWD_ATOL * _scale(rgv), f"{what} grad_value vs oracle")
    assert_close(gl, rgl, BWD_RTOL, BWD_ATOL * _scale(rgl), f"{what} grad_loc vs oracle")
    assert_close(ga, rga, BWD_RTOL, BWD_ATOL * _scale(rga), f"{what} grad_attn vs oracle")
    return out, (gv, gl, ga)


@pytest.mark.parametrize("case", MSDA_CASES)
def test_msda_golden_fixture(case):
    d = load_golden(case)
    out, (gv, gl, ga) = _check_msda_against_oracle(d["value"], d["spatial_shapes"], d["level_start_index"], d["loc"],
                                                   d["aw"], d["grad_out"], case)
    assert_close(out, d["out"], FWD_RTOL, FWD_ATOL_GRIDSAMPLE * _scale(d["value"]), f"{case} forward vs grid_sample")
    assert_close(gv, d["grad_value"], BWD_RTOL, 1e-4 * _scale(d["grad_value"]), f"{case} grad_value vs grid_sample")
    assert_close(ga, d["grad_aw"], BWD_RTOL, 1e-4 * _scale(d["grad_aw"]), f"{case} grad_attn vs grid_sample")
    ok = ~near_integer_pixel_msda(d["loc"], d["spatial_shapes"].tolist())
    assert_close(gl.cpu()[ok], d["grad_loc"][ok], BWD_RTOL, 1e-4 * _scale(d["grad_loc"][ok]), f"{case} grad_loc vs grid_sample")


@pytest.mark.parametrize("shapes,M,D,P,lo,hi", [
    ([(32, 32)], 4, 64, 20, 0.0, 1.0),                 # RaCFormer configuration, scaled-down map
    ([(32, 32)], 4, 64, 20, -0.05, 1.05),
    ([(16, 20), (8, 10), (4, 5), (2, 3)], 8, 64, 4, -0.1, 1.1),   # Deformable-DETR style, 4 levels
    ([(12, 7)], 2, 64, 40, -0.1, 1.1),                 # more than 32 taps per (b,q,m): two staging rounds
    ([(16, 20), (8, 10)], 8, 32, 4, -0.1, 1.1),        # generic path
    ([(9, 9)], 3, 16, 5, -2.0, -1.0),                  # nothing valid
    ([(5, 3)], 1, 64, 1, -0.1, 1.1),                   # one head, one point
    ([(7, 9), (4, 4)], 3, 64, 16, -0.1, 1.1),          # exactly 32 taps per row
    ([(7, 9), (4, 4)], 3, 7, 3, -0.1, 1.1),            # odd head_dim (generic path)
])
def test_msda_seeded_vs_oracle(shapes, M, D, P, lo, hi):
    value, sp, lsi, loc, aw, g = make_msda_inputs(23 + P, B=4, M=M, D=D, Q=33, P=P, shapes=shapes, lo=lo, hi=hi)
    _check_msda_against_oracle(value, sp, lsi, loc, aw, g, f"{shapes} M{M} D{D} P{P}", im2col_step=2)


def test_msda_f8_full_size():
    """BEV cross-attention at the f8 shapes: value [8,16384,4,64], Q=900, 20 points, 128x128 map."""
    _, msda = _ops()
    co = _oracle()
    value, sp, lsi, loc, aw, g = make_msda_inputs(0, B=8, M=4, D=64, Q=900, P=20, shapes=[(128, 128)], lo=-0.05, hi=1.05)
    args_d = [t.to(DEV) for t in (value, sp, lsi, loc, aw)]
    out = msda.ext_module.ms_deform_attn_forward(*args_d, im2col_step=64)
    assert torch.equal(out, msda.ext_module.ms_deform_attn_forward(*args_d, im2col_step=64))
    ref_out, ref_mask = co.msda_forward(value, sp, lsi, loc, aw, with_masks=True)
    assert_close(out, ref_out, FWD_RTOL, FWD_ATOL_EXACT * _scale(value), "f8 msda forward vs oracle")
    assert torch.equal(msda.msda_tap_masks(args_d[1], args_d[3]).cpu(), ref_mask)
    gv, gl, ga = torch.zeros_like(args_d[0]), torch.empty_like(args_d[3]), torch.empty_like(args_d[4])
    msda.ext_module.ms_deform_attn_backward(*args_d, g.to(DEV), gv, gl, ga, im2col_step=64)
    rgv, rgl, rga = co.msda_backward(value, sp, lsi, loc, aw, g)
    assert_close(gv, rgv, BWD_RTOL, BWD_ATOL * _scale(rgv), "f8 msda grad_value vs oracle")
    assert_close(gl, rgl, BWD_RTOL, BWD_ATOL * _scale(rgl), "f8 msda grad_loc vs oracle")
    assert_close(ga, rga, BWD_RTOL, BWD_ATOL * _scale(rga), "f8 msda grad_attn vs oracle")


def test_msda_autograd_function_matches_torch_port():
    """MultiScaleDeformableAttnFunction_fp32.apply end to end vs the grid_sample port (autograd of both)."""
    _, msda = _ops()
    from oracle import reference_port
    value, sp, lsi, loc, aw, g = make_msda_inputs(9, B=2, M=4, D=64, Q=50, P=20, shapes=[(32, 32)], lo=0.02, hi=0.98)
    leaf = [t.to(DEV).requires_grad_() for t in (value, loc, aw)]
    out = msda.MultiScaleDeformableAttnFunction_fp32.apply(leaf[0], sp.to(DEV), lsi.to(DEV), leaf[1], leaf[2], 64)
    out.backward(g.to(DEV))
    cpu_leaf = [t.clone().requires_grad_() for t in (value, loc, aw)]
    ref = reference_port.msda_torch(cpu_leaf[0], sp, cpu_leaf[1], cpu_leaf[2])
    ref.backward(g)
    assert_close(out, ref, FWD_RTOL, FWD_ATOL_GRIDSAMPLE * _scale(value), "Function forward vs port")
    assert_close(leaf[0].grad, cpu_leaf[0].grad, BWD_RTOL, 1e-4 * _scale(cpu_leaf[0].grad), "Function grad_value vs port")
    assert_close(leaf[2].grad, cpu_leaf[2].grad, BWD_RTOL, 1e-4 * _scale(cpu_leaf[2].grad), "Function grad_attn vs port")
    ok = ~near_integer_pixel_msda(loc, sp.tolist())
    assert_close(leaf[1].grad.cpu()[ok], cpu_leaf[1].grad[ok], BWD_RTOL, 1e-4 * _scale(cpu_leaf[1].grad), "Function grad_loc vs port")


def test_msda_errors():
    _, msda = _ops()
    value = torch.zeros(6, 16, 1, 64, device=DEV)
    sp, lsi = torch.tensor([[4, 4]], device=DEV), torch.tensor([0], device=DEV)
    loc, aw = torch.zeros(6, 1, 1, 1, 1, 2, device=DEV), torch.zeros(6, 1, 1, 1, 1, device=DEV)
    with pytest.raises(RuntimeError, match="im2col_step"):
        msda.ext_module.ms_deform_attn_forward(value, sp, lsi, loc, aw, im2col_step=4)
    with pytest.raises(RuntimeError, match="contiguous"):
        strided = torch.zeros(6, 16, 1, 128, device=DEV)[..., ::2]
        msda.ext_module.ms_deform_attn_forward(strided, sp, lsi, loc, aw, im2col_step=64)
    # level metadata that does not fit the value map (would read / red.add out of bounds) is refused
    with pytest.raises(RuntimeError, match="outside the value map"):
        msda.ext_module.ms_deform_attn_forward(value, torch.tensor([[5, 4]], device=DEV), lsi, loc, aw, im2col_step=64)
    with pytest.raises(RuntimeError, match="outside the value map"):
        msda.ext_module.ms_deform_attn_forward(value, sp, torch.tensor([1], device=DEV), loc, aw, im2col_step=64)
    with pytest.raises(RuntimeError, match="spatial_shapes"):
        msda.msda_tap_masks(sp.int(), loc)


# ------------------------------------------------------------------------------------------------ guard bands
def _guarded(shape, pad=1024, fill=float("nan")):
    """A tensor living inside a larger sentinel-filled allocation (compute-sanitizer is closed on this pool)."""
    n = 1
    for s in shape:
        n *= s
    big = torch.full((n + 2 * pad,), fill, device=DEV)
    return big, big[pad:pad + n].view(shape)


def test_no_out_of_bounds_access_guard_bands():
    """Boundary-heavy coordinates (pixel -1, -0.5, 0, W-1, W-0.5, W, corners of the first/last view and batch element):
    forward must never read the NaN guard bands around the feature maps (outputs stay finite and equal to the
    unguarded run) and backward must never write outside the gradient buffers (sentinels intact)."""
    import ctypes
    from racformer_b200 import _lib, wrapper
    shapes = [(8, 12), (4, 6), (2, 3), (1, 2)]
    Bp, N, C, Q, P = 2, 3, 64, 16, 12
    feats, loc, w, g = make_msmv_inputs(17, Bp=Bp, N=N, C=C, Q=Q, P=P, shapes=shapes, lo=-0.2, hi=1.2)
    h0, w0 = shapes[0]
    edge_x = torch.tensor([-1.0, -0.5, 0.0, w0 - 1.0, w0 - 0.5, float(w0), -1e-7, w0 - 1 + 1e-7]) / (w0 - 1)
    edge_y = torch.tensor([-1.0, -0.5, 0.0, h0 - 1.0, h0 - 0.5, float(h0), -1e-7, h0 - 1 + 1e-7]) / (h0 - 1)
    for b in (0, Bp - 1):
        for k in range(8):
            loc[b, k, :, 0] = edge_x[k]
            loc[b, k, :, 1] = edge_y[(k + torch.arange(P)) % 8]
            loc[b, k, :, 2] = (0.0 if b == 0 else 1.0)          # first view of the first element / last of the last
    guards, gfeats = zip(*[_guarded(f.shape) for f in feats])
    for gf, f in zip(gfeats, feats):
        gf.copy_(f)
    loc_d, w_d, g_d = loc.to(DEV), w.to(DEV), g.to(DEV)
    out_guarded = wrapper.msmv_forward(list(gfeats), loc_d, w_d)
    out_plain = wrapper.msmv_forward([f.to(DEV) for f in feats], loc_d, w_d)
    assert torch.isfinite(out_guarded).all(), "forward read a guard band"
    assert torch.equal(out_guarded, out_plain)
    # backward through the C ABI with caller-owned, guarded gradient buffers
    gg_guards, gg = zip(*[_guarded(f.shape, fill=12345.0) for f in feats])
    for t in gg:
        t.zero_()
    gl_guard, gl = _guarded(loc.shape, fill=12345.0)
    gw_guard, gw = _guarded(w.shape, fill=12345.0)
    lib = _lib.load()
    ptrs = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in gfeats])
    gptrs = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in gg])
    hw = (ctypes.c_int * 8)(*[d for s in shapes for d in s])
    rc = lib.racf_msmv_backward(g_d.data_ptr(), ptrs, hw, 4, loc_d.data_ptr(), w_d.data_ptr(), Bp, C, N, Q, P, gptrs,
                                gl.data_ptr(), gw.data_ptr(), 0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    pad = 1024
    for big in list(gg_guards) + [gl_guard, gw_guard]:
        assert bool((big[:pad] == 12345.0).all()) and bool((big[-pad:] == 12345.0).all()), "backward wrote out of bounds"
    ref = wrapper.msmv_backward(g_d, [f.to(DEV) for f in feats], loc_d, w_d)
    for a, b in zip(list(gg) + [gl, gw], ref):
        assert_close(a, b, BWD_RTOL, BWD_ATOL * _scale(b), "guarded backward vs plain backward")
    # MSDA: same idea on the value map
    value, sp, lsi, mloc, aw, mg = make_msda_inputs(3, B=2, M=4, D=64, Q=16, P=20, shapes=[(6, 5)], lo=-0.2, hi=1.2)
    mloc[0, 0, 0, 0, :6, 0] = torch.tensor([-0.5, 0.0, 0.5, 4.5, 5.0, 5.5]) / 5
    mloc[0, 0, 0, 0, :6, 1] = torch.tensor([-0.5, 0.0, 0.5, 5.5, 6.0, 6.5]) / 6
    mloc[-1, -1, -1, 0, :6] = mloc[0, 0, 0, 0, :6]
    _, msda = _ops()
    vguard, vg = _guarded(value.shape)
    vg.copy_(value)
    args = [sp.to(DEV), lsi.to(DEV), mloc.to(DEV), aw.to(DEV)]
    mo = msda.ext_module.ms_deform_attn_forward(vg, *args, im2col_step=64)
    assert torch.isfinite(mo).all() and torch.equal(mo, msda.ext_module.ms_deform_attn_forward(value.to(DEV), *args, im2col_step=64))
    gvguard, gv = _guarded(value.shape, fill=12345.0)
    gv.zero_()
    gl2, ga2 = torch.empty_like(args[2]), torch.empty_like(args[3])
    msda.ext_module.ms_deform_attn_backward(vg, *args, mg.to(DEV), gv, gl2, ga2, im2col_step=64)
    torch.cuda.synchronize()
    assert bool((gvguard[:pad] == 12345.0).all()) and bool((gvguard[-pad:] == 12345.0).all())


def test_ops_on_a_non_current_device():
    """Tensors on cuda:1 while cuda:0 is current: the wrappers must launch under a device guard on that device's stream."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    wrapper, msda = _ops()
    co = _oracle()
    assert torch.cuda.current_device() == 0
    feats, loc, w, g = make_msmv_inputs(2, Bp=2, N=3, C=64, Q=9, P=12, shapes=[(8, 22), (4, 11), (2, 6), (1, 3)], lo=-0.1, hi=1.1)
    d1 = "cuda:1"
    f1 = [f.to(d1).requires_grad_() for f in feats]
    loc1, w1 = loc.to(d1).requires_grad_(), w.to(d1).requires_grad_()
    out = wrapper.msmv_sampling(f1, loc1, w1)
    out.backward(g.to(d1))
    assert out.device == torch.device(d1) and torch.cuda.current_device() == 0
    assert_close(out, co.msmv_forward(feats, loc, w), FWD_RTOL, FWD_ATOL_EXACT * max(_scale(f) for f in feats), "cuda:1 forward")
    rgf, rgl, rgw = co.msmv_backward(g, feats, loc, w)
    assert_close(loc1.grad, rgl, BWD_RTOL, BWD_ATOL * _scale(rgl), "cuda:1 grad_loc")
    assert_close(f1[0].grad, rgf[0], BWD_RTOL, BWD_ATOL * _scale(rgf[0]), "cuda:1 grad_feat0")
    value, sp, lsi, mloc, aw, mg = make_msda_inputs(4, B=2, M=4, D=64, Q=9, P=20, shapes=[(16, 16)])
    mo = msda.MultiScaleDeformableAttnFunction_fp32.apply(value.to(d1), sp.to(d1), lsi.to(d1), mloc.to(d1), aw.to(d1), 64)
    assert_close(mo, co.msda_forward(value, sp, lsi, mloc, aw), FWD_RTOL, FWD_ATOL_EXACT * _scale(value), "cuda:1 msda forward")


def test_ops_are_cuda_graph_capture_safe_and_repeatable():
    """Forward + backward of both ops captured in one CUDA graph (incl. the library's own zero-fill memsets), replayed
    several times: every replay must reproduce the eager forward bit for bit and the eager backward within the atomic
    reordering tolerance; the library allocates nothing and never synchronises, so capture must succeed as is."""
    wrapper, msda = _ops()
    feats, loc, w, g = make_msmv_inputs(8, Bp=4, N=6, C=64, Q=64, P=12, shapes=[(16, 44), (8, 22), (4, 11), (2, 6)],
                                        lo=-0.1, hi=1.1, device=DEV)
    value, sp, lsi, mloc, aw, mg = make_msda_inputs(8, B=2, M=4, D=64, Q=64, P=20, shapes=[(32, 32)], lo=-0.05, hi=1.05,
                                                    device=DEV)
    eager_out = wrapper.msmv_forward(feats, loc, w)
    eager_grads = wrapper.msmv_backward(g, feats, loc, w)
    eager_mout = msda.ext_module.ms_deform_attn_forward(value, sp, lsi, mloc, aw, im2col_step=64)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            wrapper.msmv_backward(g, feats, loc, w)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    gv = torch.zeros_like(value)
    gl, ga = torch.empty_like(mloc), torch.empty_like(aw)
    with torch.cuda.graph(graph):
        out = wrapper.msmv_forward(feats, loc, w)
        grads = wrapper.msmv_backward(g, feats, loc, w)
        mout = msda.ext_module.ms_deform_attn_forward(value, sp, lsi, mloc, aw, im2col_step=64)
        gv.zero_()
        msda.ext_module.ms_deform_attn_backward(value, sp, lsi, mloc, aw, mg, gv, gl, ga, im2col_step=64)
    first = None
    for _ in range(5):
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, eager_out) and torch.equal(mout, eager_mout)
        for a, b in zip(grads, eager_grads):
            assert_close(a, b, BWD_RTOL, BWD_ATOL * _scale(b), "captured backward vs eager backward")
        if first is None:
            first = [t.clone() for t in grads] + [gv.clone()]
        else:   # run-to-run spread of the atomics stays inside the same tolerance
            for a, b in zip(list(grads) + [gv], first):
                assert_close(a, b, BWD_RTOL, BWD_ATOL * _scale(b), "replay vs first replay")
