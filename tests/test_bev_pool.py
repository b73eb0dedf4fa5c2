"""BEVPoolv2. The reference pins this op with a known-answer test
(models/csrc/bev_pool_v2/bev_pool.py:147-178: loss 4.4, depth/feat gradients); it is restated here for the oracle (CPU)
and for the CUDA path (GPU), plus seeded LSS-shaped cases against the oracle."""
import numpy as np
import pytest
import torch

from oracle import c_oracle


def _kat_inputs(device="cpu"):
    depth = torch.tensor([0.3, 0.4, 0.2, 0.1, 0.7, 0.6, 0.8, 0.9], device=device).view(1, 1, 2, 2, 2)
    feat = torch.ones(1, 1, 2, 2, 2, device=device)
    ranks_depth = torch.tensor([0, 4, 1, 6], dtype=torch.int32, device=device)
    ranks_feat = torch.tensor([0, 0, 1, 2], dtype=torch.int32, device=device)
    ranks_bev = torch.tensor([0, 0, 1, 1], dtype=torch.int32, device=device)
    starts = torch.tensor([0, 2], dtype=torch.int32, device=device)
    lengths = torch.tensor([2, 2], dtype=torch.int32, device=device)
    return depth, feat, ranks_depth, ranks_feat, ranks_bev, starts, lengths


KAT_GRAD_DEPTH = torch.tensor([2., 2., 0., 0., 2., 0., 2., 0.]).view(1, 1, 2, 2, 2)
KAT_GRAD_FEAT = torch.tensor([1.0, 1.0, 0.4, 0.4, 0.8, 0.8, 0., 0.]).view(1, 1, 2, 2, 2)


def _bp_intervals(ranks_depth, ranks_feat, ranks_bev):
    order = torch.argsort(ranks_feat.long(), stable=True)
    rf, rd, rb = ranks_feat[order], ranks_depth[order], ranks_bev[order]
    kept = torch.ones(rb.shape[0], dtype=torch.bool)
    kept[1:] = rf[1:] != rf[:-1]
    starts = torch.where(kept)[0].int()
    lengths = torch.zeros_like(starts)
    lengths[:-1] = starts[1:] - starts[:-1]
    lengths[-1] = rb.shape[0] - starts[-1]
    return rd, rf, rb, starts, lengths


def test_oracle_reproduces_reference_known_answer():
    depth, feat, rd, rf, rb, starts, lengths = _kat_inputs()
    out = c_oracle.bev_pool_v2_forward(depth, feat, rd, rf, rb, (1, 1, 2, 2, 2), starts, lengths)
    assert float(out.sum()) == pytest.approx(4.4, abs=1e-6)
    rd2, rf2, rb2, s2, l2 = _bp_intervals(rd, rf, rb)
    dg, fg = c_oracle.bev_pool_v2_backward(torch.ones_like(out), depth, feat, rd2, rf2, rb2, s2, l2)
    assert torch.allclose(dg, KAT_GRAD_DEPTH) and torch.allclose(fg, KAT_GRAD_FEAT)


from racformer_b200.synthetic import make_lss_pool_case as make_lss_case  # noqa: E402


@pytest.mark.gpu
def test_cuda_reproduces_reference_known_answer():
    from racformer_b200.bev_pool import bev_pool_v2
    depth, feat, rd, rf, rb, starts, lengths = _kat_inputs("cuda")
    depth.requires_grad_()
    feat.requires_grad_()
    bev = bev_pool_v2(depth, feat, rd, rf, rb, (1, 1, 2, 2, 2), starts, lengths)
    assert bev.shape == (1, 2, 1, 2, 2)
    loss = bev.sum()
    loss.backward()
    assert float(loss) == pytest.approx(4.4, abs=1e-6)
    assert depth.grad.cpu().allclose(KAT_GRAD_DEPTH) and feat.grad.cpu().allclose(KAT_GRAD_FEAT)


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,D,H,W,C,bev", [
    (1, 6, 96, 16, 44, 256, (128, 128)),     # LSS view transformer at 704x256, numC_Trans = 256 (configs/...f8.py:55-63)
    (2, 3, 16, 4, 11, 64, (32, 32)),
    (1, 2, 8, 3, 5, 80, (16, 16)),           # C not a multiple of 128
    (1, 2, 8, 3, 5, 6, (8, 8)),              # C % 4 != 0 -> generic kernels
    (1, 1, 40, 2, 2, 1024, (4, 4)),          # long intervals (> 32 points) and the widest vector path
])
def test_cuda_vs_oracle(B, N, D, H, W, C, bev):
    from racformer_b200 import bev_pool
    case = make_lss_case(B + C, B, N, D, H, W, C, bev)
    dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in case.items()}
    out = torch.zeros(case["shape"], device="cuda")
    bev_pool.bev_pool_v2_forward(dev["depth"], dev["feat"], out, dev["ranks_depth"], dev["ranks_feat"], dev["ranks_bev"],
                                 dev["lengths"], dev["starts"])
    ref = c_oracle.bev_pool_v2_forward(case["depth"], case["feat"], case["ranks_depth"], case["ranks_feat"],
                                       case["ranks_bev"], case["shape"], case["starts"], case["lengths"])
    assert torch.equal(out.cpu(), ref), "forward must be bit-identical (same summation order, fused multiply-add)"
    # backward with a fixed (stable) re-sorting so that the oracle sees the same point order
    g = torch.Generator().manual_seed(3)
    og = torch.randn(case["shape"], generator=g)
    rd, rf, rb, st, ln = _bp_intervals(case["ranks_depth"], case["ranks_feat"], case["ranks_bev"])
    dg, fg = torch.zeros_like(dev["depth"]), torch.zeros_like(dev["feat"])
    bev_pool.bev_pool_v2_backward(og.cuda(), dg, fg, dev["depth"], dev["feat"], rd.cuda(), rf.cuda(), rb.cuda(),
                                  ln.cuda(), st.cuda())
    rdg, rfg = c_oracle.bev_pool_v2_backward(og, case["depth"], case["feat"], rd, rf, rb, st, ln)
    assert torch.equal(fg.cpu(), rfg), "feat_grad must be bit-identical (sequential fused sum over the run)"
    scale = float(rdg.abs().max())
    assert float((dg.cpu() - rdg).abs().max()) <= 1e-5 * scale + 1e-6, "depth_grad: shuffle-tree dot vs fp64 dot"
    # autograd wrapper = same numbers (its argsort need not be stable: compare with tolerance)
    d2, f2 = dev["depth"].clone().requires_grad_(), dev["feat"].clone().requires_grad_()
    y = bev_pool.bev_pool_v2(d2, f2, dev["ranks_depth"], dev["ranks_feat"], dev["ranks_bev"], case["shape"],
                             dev["starts"], dev["lengths"])
    assert torch.equal(y, out.permute(0, 4, 1, 2, 3))
    y.backward(og.cuda().permute(0, 4, 1, 2, 3))
    assert torch.allclose(f2.grad.cpu(), rfg, rtol=1e-5, atol=1e-5 * float(rfg.abs().max()))
    assert torch.allclose(d2.grad.cpu(), rdg, rtol=1e-5, atol=1e-5 * scale)


@pytest.mark.gpu
def test_cuda_rejects_bad_inputs():
    from racformer_b200 import bev_pool
    depth, feat, rd, rf, rb, starts, lengths = _kat_inputs("cuda")
    out = torch.zeros(1, 1, 2, 2, 2, device="cuda")
    with pytest.raises(RuntimeError, match="int32"):
        bev_pool.bev_pool_v2_forward(depth, feat, out, rd.long(), rf, rb, lengths, starts)
    with pytest.raises(RuntimeError, match="float32 CUDA"):
        bev_pool.bev_pool_v2_forward(depth.cpu(), feat, out, rd, rf, rb, lengths, starts)


@pytest.mark.gpu
def test_cuda_vs_reference_extension():
    """The reference's own bev_pool_v2_ext (oracle/_ref, built by oracle/build_ref.py) on LSS-shaped inputs."""
    from oracle import build_ref
    ext = build_ref.load_prebuilt_bev_pool()
    if ext is None:
        pytest.skip("oracle/_ref/bev_pool_v2_ext.so not built")
    from racformer_b200 import bev_pool
    case = make_lss_case(11, 1, 6, 96, 16, 44, 256, (128, 128), device="cuda")
    out = torch.zeros(case["shape"], device="cuda")
    ref = torch.zeros(case["shape"], device="cuda")
    args = (case["ranks_depth"], case["ranks_feat"], case["ranks_bev"], case["lengths"], case["starts"])
    bev_pool.bev_pool_v2_forward(case["depth"], case["feat"], out, *args)
    ext.bev_pool_v2_forward(case["depth"], case["feat"], ref, *args)
    torch.cuda.synchronize()
    assert torch.equal(out, ref), "forward must be bit-identical to the reference kernel"
    og = torch.randn(case["shape"], device="cuda")
    rd, rf, rb, st, ln = (t.cuda() for t in _bp_intervals(case["ranks_depth"].cpu(), case["ranks_feat"].cpu(), case["ranks_bev"].cpu()))
    dg, fg = torch.zeros_like(case["depth"]), torch.zeros_like(case["feat"])
    rdg, rfg = torch.zeros_like(case["depth"]), torch.zeros_like(case["feat"])
    bev_pool.bev_pool_v2_backward(og, dg, fg, case["depth"], case["feat"], rd, rf, rb, ln, st)
    ext.bev_pool_v2_backward(og, rdg, rfg, case["depth"], case["feat"], rd, rf, rb, ln, st)
    torch.cuda.synchronize()
    assert torch.equal(fg, rfg)
    assert float((dg - rdg).abs().max()) <= 1e-5 * float(rdg.abs().max())
