"""CPU: randomized cross-checks inside the oracle package -- the exact-arithmetic C restatement of the CUDA kernels
against the restatement of the reference's grid_sample paths, on random shapes (hypothesis), plus algebraic properties
the operators must satisfy (linearity / adjointness, constant-feature invariance, mask semantics)."""
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import c_oracle, reference_port
from tests.helpers import make_msda_inputs, make_msmv_inputs

SETTINGS = dict(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow])


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), Bp=st.integers(1, 3), N=st.integers(2, 6), C=st.sampled_from([1, 3, 8, 64]),
       Q=st.integers(1, 6), P=st.integers(1, 9), L=st.integers(1, 5), lo=st.sampled_from([0.0, -0.15]),
       hi=st.sampled_from([1.0, 1.15]))
def test_msmv_c_oracle_equals_gridsample_port(seed, Bp, N, C, Q, P, L, lo, hi):
    shapes = [(7, 9), (4, 5), (3, 3), (2, 2), (1, 2)][:L]
    feats, loc, w, g = make_msmv_inputs(seed, Bp=Bp, N=N, C=C, Q=Q, P=P, shapes=shapes, lo=lo, hi=hi)
    out, view, mask = c_oracle.msmv_forward(feats, loc, w, with_masks=True)
    port = reference_port.msmv_sampling_torch_channel_last(feats, loc, w)
    assert torch.allclose(out, port, rtol=1e-5, atol=5e-5 * max(float(f.abs().max()) for f in feats))
    # a tap that is out of range on a level contributes nothing: zeroing that level's weight must not change anything
    w2 = w.clone()
    w2[(mask & 1) == 0] = 0.0
    assert torch.equal(c_oracle.msmv_forward(feats, loc, w2), out)
    # adjoint identities (the op is linear in the features and in the weights)
    gf, gl, gw = c_oracle.msmv_backward(g, feats, loc, w)
    lhs = float((out.double() * g.double()).sum())
    rhs_f = sum(float((a.double() * b.double()).sum()) for a, b in zip(gf, feats))
    rhs_w = float((gw.double() * w.double()).sum())
    scale = 1e-4 * (1.0 + abs(lhs)) + 1e-5 * float(g.abs().sum())
    assert abs(lhs - rhs_f) <= scale and abs(lhs - rhs_w) <= scale
    assert bool((gl[..., 2] == 0).all())


@settings(**SETTINGS)
@given(seed=st.integers(0, 10_000), B=st.integers(1, 3), M=st.integers(1, 4), D=st.sampled_from([1, 4, 32, 64]),
       Q=st.integers(1, 6), P=st.integers(1, 7), L=st.integers(1, 3))
def test_msda_c_oracle_equals_gridsample_port(seed, B, M, D, Q, P, L):
    shapes = [(6, 5), (3, 4), (2, 2)][:L]
    value, sp, lsi, loc, aw, g = make_msda_inputs(seed, B=B, M=M, D=D, Q=Q, P=P, shapes=shapes, lo=-0.1, hi=1.1)
    out = c_oracle.msda_forward(value, sp, lsi, loc, aw)
    port = reference_port.msda_torch(value, sp, loc, aw)
    assert torch.allclose(out, port, rtol=1e-5, atol=5e-5 * float(value.abs().max()))
    gv, gl, ga = c_oracle.msda_backward(value, sp, lsi, loc, aw, g)
    lhs = float((out.double() * g.double()).sum())
    scale = 1e-4 * (1.0 + abs(lhs)) + 1e-5 * float(g.abs().sum())
    assert abs(lhs - float((gv.double() * value.double()).sum())) <= scale
    assert abs(lhs - float((ga.double() * aw.double()).sum())) <= scale


def test_constant_features_give_constant_output_where_everything_is_valid():
    feats, loc, w, _ = make_msmv_inputs(1, Bp=2, N=3, C=4, Q=6, P=5, shapes=[(6, 8), (3, 4)], lo=0.0, hi=1.0)
    ones = [torch.full_like(f, 2.5) for f in feats]
    out, _, mask = c_oracle.msmv_forward(ones, loc, w, with_masks=True)
    ok = (mask == 31).all(-1)[:, :, None, :].expand_as(out)
    assert ok.any() and torch.allclose(out[ok], torch.full_like(out[ok], 2.5), atol=1e-5)
