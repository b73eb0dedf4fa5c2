"""CPU: pin the oracle (C exact-arithmetic restatement + torch port) to the fixtures generated from the reference."""
import pytest
import torch

from oracle import c_oracle, reference_port
from tests.helpers import (BWD_ATOL, BWD_RTOL, FWD_ATOL_GRIDSAMPLE, FWD_RTOL, MSDA_CASES, MSMV_CASES, assert_close,
                           load_golden, msmv_feats, near_integer_pixel_msda, near_integer_pixel_msmv)


def _scale(t):
    return float(t.abs().max())


@pytest.mark.parametrize("case", MSMV_CASES)
def test_msmv_c_oracle_matches_reference_fixture(case):
    d = load_golden(case)
    feats = msmv_feats(d)
    fmax = max(_scale(f) for f in feats)
    out, view, mask = c_oracle.msmv_forward(feats, d["loc"], d["w"], with_masks=True)
    assert_close(out, d["out"], FWD_RTOL, FWD_ATOL_GRIDSAMPLE * fmax, f"{case} forward")
    n_views = int(d["num_views"])
    assert torch.equal(view.long(), (d["loc"][..., 2] * (n_views - 1)).round().long())
    gfeats, gloc, gw = c_oracle.msmv_backward(d["grad_out"], feats, d["loc"], d["w"])
    for i, g in enumerate(gfeats):
        assert_close(g, d[f"grad_feat{i}"], BWD_RTOL, BWD_ATOL * max(_scale(d[f"grad_feat{i}"]), 1.0) * 10,
                     f"{case} grad_feat{i}")
    assert_close(gw, d["grad_w"], BWD_RTOL, 1e-4 * _scale(d["grad_w"]), f"{case} grad_w")
    ok = ~near_integer_pixel_msmv(d["loc"], [f.shape[2:4] for f in feats])
    assert ok.float().mean() > 0.5
    assert_close(gloc[..., :2][ok], d["grad_loc"][..., :2][ok], BWD_RTOL, 1e-4 * _scale(d["grad_loc"][..., :2][ok]),
                 f"{case} grad_loc")
    assert (gloc[..., 2] == 0).all()   # reference quirk (v): the view coordinate never receives a gradient


@pytest.mark.parametrize("case", MSMV_CASES)
def test_msmv_torch_port_matches_reference_fixture(case):
    d = load_golden(case)
    feats_cf = [f.permute(0, 4, 1, 2, 3).contiguous().requires_grad_() for f in msmv_feats(d)]
    loc, w = d["loc"].clone().requires_grad_(), d["w"].clone().requires_grad_()
    out = reference_port.msmv_sampling_torch(feats_cf, loc, w)
    assert_close(out, d["out"], 1e-6, 1e-6, f"{case} forward (port)")
    out.backward(d["grad_out"])
    assert_close(loc.grad, d["grad_loc"], 1e-5, 1e-5 * _scale(d["grad_loc"]), f"{case} grad_loc (port)")
    assert_close(w.grad, d["grad_w"], 1e-5, 1e-5 * _scale(d["grad_w"]), f"{case} grad_w (port)")
    for i, f in enumerate(feats_cf):
        assert_close(f.grad.permute(0, 2, 3, 4, 1), d[f"grad_feat{i}"], 1e-5, 1e-6, f"{case} grad_feat{i} (port)")
    v2 = reference_port.msmv_sampling_v2_torch([f.detach() for f in feats_cf], d["loc"], d["w"])
    assert_close(v2, d["out_v2"], 1e-6, 1e-6, f"{case} v2 (port)")


@pytest.mark.parametrize("case", MSDA_CASES)
def test_msda_c_oracle_matches_fixture(case):
    d = load_golden(case)
    vmax = _scale(d["value"])
    out, mask = c_oracle.msda_forward(d["value"], d["spatial_shapes"], d["level_start_index"], d["loc"], d["aw"],
                                      with_masks=True)
    assert_close(out, d["out"], FWD_RTOL, FWD_ATOL_GRIDSAMPLE * vmax, f"{case} forward")
    gv, gl, ga = c_oracle.msda_backward(d["value"], d["spatial_shapes"], d["level_start_index"], d["loc"], d["aw"],
                                        d["grad_out"])
    assert_close(gv, d["grad_value"], BWD_RTOL, BWD_ATOL * max(_scale(d["grad_value"]), 1.0) * 10, f"{case} grad_value")
    assert_close(ga, d["grad_aw"], BWD_RTOL, 1e-4 * _scale(d["grad_aw"]), f"{case} grad_aw")
    ok = ~near_integer_pixel_msda(d["loc"], d["spatial_shapes"].tolist())
    assert_close(gl[ok], d["grad_loc"][ok], BWD_RTOL, 1e-4 * _scale(d["grad_loc"][ok]), f"{case} grad_loc")


@pytest.mark.parametrize("case", MSDA_CASES)
def test_msda_torch_port_matches_fixture(case):
    d = load_golden(case)
    value, loc, aw = d["value"].clone().requires_grad_(), d["loc"].clone().requires_grad_(), d["aw"].clone().requires_grad_()
    out = reference_port.msda_torch(value, d["spatial_shapes"], loc, aw)
    assert_close(out, d["out"], 1e-6, 1e-6, f"{case} forward (port)")
    out.backward(d["grad_out"])
    assert_close(value.grad, d["grad_value"], 1e-5, 1e-6, f"{case} grad_value (port)")
    assert_close(loc.grad, d["grad_loc"], 1e-5, 1e-5 * _scale(d["grad_loc"]), f"{case} grad_loc (port)")
    assert_close(aw.grad, d["grad_aw"], 1e-5, 1e-5 * _scale(d["grad_aw"]), f"{case} grad_aw (port)")


def test_msmv_masks_known_answers():
    """Hand-checked corner cases of the validity logic (msmv_sampling_forward.cu:48-67,126)."""
    H, W = 4, 6
    feats = [torch.ones(1, 2, H, W, 4)]
    pix = [(-1.0, 0.0), (-0.5, 0.0), (0.0, 0.0), (W - 1.0, H - 1.0), (W - 0.5, H - 0.5), (float(W), 0.0),
           (2.5, 1.5), (float("nan"), 0.0)]
    loc = torch.tensor([[x / (W - 1), y / (H - 1), 1.0] for x, y in pix]).view(1, 1, len(pix), 3)
    w = torch.ones(1, 1, len(pix), 1)
    out, view, mask = c_oracle.msmv_forward(feats, loc, w, with_masks=True)
    assert view.flatten().tolist() == [1] * len(pix)
    # bit0 in-range, bit1 TL, bit2 TR, bit3 BL, bit4 BR
    assert mask.flatten().tolist() == [0, 1 | 4 | 16, 1 | 2 | 4 | 8 | 16, 1 | 2, 1 | 2, 0, 31, 0]
    got = out[0, 0, 0].tolist()
    want = [0.0, 0.5, 1.0, 1.0, 0.25, 0.0, 1.0, 0.0]
    assert got == pytest.approx(want, abs=1e-6)
