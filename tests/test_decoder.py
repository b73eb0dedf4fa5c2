"""Decoder harness (racformer_b200/decoder.py) against the UNCHANGED reference decoder.

CPU: the harness' host logic with the oracle's torch ops injected vs (i) the committed fixture produced by the
reference, (ii) the live reference when /root/reference is present (forward, gradients, denoising mask, B=2 quirks,
state_dict compatibility). GPU: the same fixture through the sm_100a kernels.
"""
import pytest
import torch

from racformer_b200.decoder import RaCFormerTransformer
from racformer_b200.synthetic import fill_parameters_by_name
from tests import reference_shim
from tests.decoder_cases import SMALL, cpu_oracle_ops, small_inputs
from tests.helpers import load_golden

DEC_RTOL, DEC_ATOL = 1e-4, 2e-4   # fp32 through 2 decoder iterations (LayerNorms, softmaxes, atan2/sin/cos)


def _close(a, b, what, rtol=DEC_RTOL, atol=DEC_ATOL, max_outlier_frac=0.0):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bad = (a - b).abs() > atol + rtol * b.abs()
    frac = float(bad.float().mean())
    assert frac <= max_outlier_frac, f"{what}: {frac:.2e} of elements differ, worst {float((a - b).abs().max()):.3g}"


def _fixture_inputs(device="cpu"):
    d = load_golden("decoder_small")
    feats = [d[f"feat{i}"].to(device) for i in range(4)]
    metas = small_inputs(seed=5)["img_metas"]
    return d, feats, metas


def _my_model(ops=None, **kw):
    model = RaCFormerTransformer(**SMALL, ops=ops, **kw)
    model.init_weights()
    fill_parameters_by_name(model, seed=3)
    return model.eval()


@pytest.mark.parametrize("hoist", [True, False])
def test_decoder_matches_reference_fixture_cpu(hoist):
    d, feats, metas = _fixture_inputs()
    model = _my_model(ops=cpu_oracle_ops(), hoist_invariants=hoist)
    with torch.no_grad():
        cls, box = model(d["query_bbox"], d["query_feat"], feats, d["lss_bev"], d["radar_bev"], None, metas)
    _close(cls, d["cls_scores"], "cls_scores vs reference fixture")
    _close(box, d["bbox_preds"], "bbox_preds vs reference fixture")


@pytest.mark.gpu
def test_decoder_matches_reference_fixture_gpu():
    d, feats, metas = _fixture_inputs("cuda")
    model = _my_model().cuda()
    with torch.no_grad():
        cls, box = model(d["query_bbox"].cuda(), d["query_feat"].cuda(), feats, d["lss_bev"].cuda(),
                         d["radar_bev"].cuda(), None, metas)
    # cuDNN convolutions may use TF32 (PyTorch default, as for the reference on GPU): looser than the CPU comparison
    _close(cls, d["cls_scores"], "cls_scores vs reference fixture (GPU)", rtol=2e-3, atol=5e-3, max_outlier_frac=0.01)
    _close(box, d["bbox_preds"], "bbox_preds vs reference fixture (GPU)", rtol=2e-3, atol=5e-3, max_outlier_frac=0.01)


@pytest.mark.gpu
def test_decoder_gpu_strict_fp32_and_hoisting_equivalence():
    """With TF32 off, the CUDA path must meet the CPU tolerance, and hoisting must not change the result."""
    d, feats, metas = _fixture_inputs("cuda")
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        outs = []
        for hoist in (True, False):
            model = _my_model(hoist_invariants=hoist).cuda()
            with torch.no_grad():
                outs.append(model(d["query_bbox"].cuda(), d["query_feat"].cuda(), feats, d["lss_bev"].cuda(),
                                  d["radar_bev"].cuda(), None, metas))
        _close(outs[0][0], d["cls_scores"], "cls (GPU, fp32 conv)", rtol=1e-3, atol=1e-3, max_outlier_frac=0.005)
        _close(outs[0][1], d["bbox_preds"], "box (GPU, fp32 conv)", rtol=1e-3, atol=1e-3, max_outlier_frac=0.005)
        _close(outs[0][0], outs[1][0], "hoisted vs per-layer cls", rtol=1e-5, atol=1e-5)
        _close(outs[0][1], outs[1][1], "hoisted vs per-layer box", rtol=1e-5, atol=1e-5)
    finally:
        torch.backends.cudnn.allow_tf32 = old


needs_ref = pytest.mark.skipif(not reference_shim.available(), reason="/root/reference not present")


def _ref_model():
    rt = reference_shim.load_reference_transformer_module()
    model = rt.RaCFormerTransformer(**SMALL)
    model.init_weights()
    fill_parameters_by_name(model, seed=3)
    return model


@needs_ref
def test_state_dict_is_interchangeable_with_reference():
    ref, mine = _ref_model(), _my_model(ops=cpu_oracle_ops())
    rs, ms = ref.state_dict(), mine.state_dict()
    assert set(rs) == set(ms), (sorted(set(rs) - set(ms)), sorted(set(ms) - set(rs)))
    for k in rs:
        assert rs[k].shape == ms[k].shape, k
    mine.load_state_dict(rs, strict=True)


def _ref_model_layers(num_layers):
    rt = reference_shim.load_reference_transformer_module()
    model = rt.RaCFormerTransformer(**dict(SMALL, num_layers=num_layers))
    model.init_weights()
    fill_parameters_by_name(model, seed=3)
    return model


@needs_ref
@pytest.mark.parametrize("batch,with_dn_mask,num_layers,grad_tol", [(1, False, 1, 1e-4), (2, True, 1, 1e-4),
                                                                    (2, True, 2, 3e-2)])
def test_decoder_forward_and_gradients_match_live_reference(batch, with_dn_mask, num_layers, grad_tol):
    """Autograd through the whole decoder (eval mode so dropout is off): outputs and the gradients w.r.t. query
    features, FPN features, BEV maps and parameters match the unchanged reference. B=2 exercises the reference's
    B*T*G / B*G*T and queue-major packing quirks; the mask exercises query denoising.

    One decoder iteration: gradients agree to 1e-4 of their max. Two iterations: fp32 rounding of the first
    iteration's outputs (1e-5) is amplified by the second iteration's sampling derivatives -- measured against an
    fp64 evaluation of the same graph, the reference itself and this harness are both ~1e-2 off (relative to max|g|),
    so they are compared at 3e-2."""
    ref = _ref_model_layers(num_layers).eval()
    mine = RaCFormerTransformer(**dict(SMALL, num_layers=num_layers), ops=cpu_oracle_ops())
    mine.init_weights()
    fill_parameters_by_name(mine, seed=3)
    mine.eval()
    d = small_inputs(seed=11, batch=batch)
    Q = d["query_bbox"].shape[1]
    mask = None
    if with_dn_mask:
        mask = torch.zeros(Q, Q, dtype=torch.bool)
        mask[: Q // 2, Q // 2:] = True

    def run(model, is_ref):
        qf = d["query_feat"].clone().requires_grad_()
        feats = [f.clone().requires_grad_() for f in d["mlvl_feats"]]
        lss, radar = d["lss_bev"].clone().requires_grad_(), d["radar_bev"].clone().requires_grad_()
        metas = [dict(m) for m in d["img_metas"]]
        feats_in = list(feats)   # the reference mutates the list it is given (quirk vii)
        cls, box = model(d["query_bbox"].clone(), qf, feats_in, lss, radar, mask, metas)
        g = torch.Generator().manual_seed(0)
        loss = (cls * torch.randn(cls.shape, generator=g)).sum() + (box * torch.randn(box.shape, generator=g)).sum()
        model.zero_grad()
        loss.backward()
        grads = {"query_feat": qf.grad, "lss": lss.grad, "radar": radar.grad}
        for i, f in enumerate(feats):
            grads[f"feat{i}"] = f.grad
        for name in ("decoder.decoder_layer.sampling.scale_weights.weight", "decoder.decoder_layer.mixing.out_proj.bias",
                     "decoder.decoder_layer.sampling_radar_bev.attention.value_proj.weight",
                     "decoder.decoder_layer.sampling_lss_bev.sampling_offset.bias",
                     "decoder.decoder_layer.sampling_radar_bev.temporal_encoder.temporal_fusion.weight"):
            grads[name] = dict(model.named_parameters())[name].grad
        return cls, box, grads

    rc, rb, rg = run(ref, True)
    mc, mb, mg = run(mine, False)
    _close(mc, rc, "cls vs live reference")
    _close(mb, rb, "box vs live reference")
    for k in rg:
        assert rg[k] is not None and mg[k] is not None, k
        scale = float(rg[k].abs().max())
        _close(mg[k], rg[k], f"grad {k} vs live reference", rtol=0.0, atol=grad_tol * max(scale, 1e-6))


@pytest.mark.gpu
def test_cuda_graph_replay_equals_eager_and_tracks_new_inputs():
    """The captured decoder forward (racformer_b200/graphs.py) reproduces the eager result bit for bit, also after
    new inputs are loaded into its static buffers."""
    from racformer_b200.graphs import GraphedDecoderForward
    model = _my_model().cuda()
    d = small_inputs(seed=5, device="cuda")
    graphed = GraphedDecoderForward(model, d)
    with torch.no_grad():
        eager = model(d["query_bbox"], d["query_feat"], d["mlvl_feats"], d["lss_bev"], d["radar_bev"], None, d["img_metas"])
    out = graphed()
    assert torch.equal(out[0], eager[0]) and torch.equal(out[1], eager[1])
    d2 = small_inputs(seed=6, device="cuda")
    with torch.no_grad():
        eager2 = model(d2["query_bbox"], d2["query_feat"], d2["mlvl_feats"], d2["lss_bev"], d2["radar_bev"], None, d2["img_metas"])
    out2 = graphed({k: v for k, v in d2.items() if k != "img_metas"})
    assert torch.equal(out2[0], eager2[0]) and torch.equal(out2[1], eager2[1])
    assert not torch.equal(eager[0], eager2[0])


@pytest.mark.gpu
def test_pipelined_serving_loop_returns_each_samples_result():
    """H2D/compute overlap must not mix samples up: results equal the eager forward of the corresponding inputs."""
    from racformer_b200.graphs import PipelinedDecoderForward
    model = _my_model().cuda()
    samples = [small_inputs(seed=s, device="cuda") for s in (5, 6, 7, 8, 9)]
    pipe = PipelinedDecoderForward(model, samples[0], depth=2)
    hosts = [{k: (v.cpu().pin_memory() if torch.is_tensor(v) else v) for k, v in s.items() if k not in ("mlvl_feats", "img_metas")}
             for s in samples]
    for h, s in zip(hosts, samples):
        h["mlvl_feats"] = [f.cpu().pin_memory() for f in s["mlvl_feats"]]
    got, prev = [], None
    for h in hosts:
        t = pipe.submit(h)
        if prev is not None:
            got.append([o.clone() for o in pipe.result(prev)])
        prev = t
    got.append([o.clone() for o in pipe.result(prev)])
    for s, (cls, box) in zip(samples, got):
        with torch.no_grad():
            e_cls, e_box = model(s["query_bbox"], s["query_feat"], s["mlvl_feats"], s["lss_bev"], s["radar_bev"], None, s["img_metas"])
        assert torch.equal(cls, e_cls.cpu()) and torch.equal(box, e_box.cpu())


@pytest.mark.gpu
def test_optin_split_tf32_mixing_stays_at_fp32_noise_level():
    """Opt-in `set_mixing_precision("tf32x3")`: the decoder output moves by no more than the fp32 reordering noise that
    already separates two fp32 implementations (DEC_RTOL/DEC_ATOL of the CPU comparison with the reference)."""
    d, feats, metas = _fixture_inputs("cuda")
    args = (d["query_bbox"].cuda(), d["query_feat"].cuda(), feats, d["lss_bev"].cuda(), d["radar_bev"].cuda(), None, metas)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        model = _my_model().cuda()
        with torch.no_grad():
            model.set_mixing_precision("fp32")
            ref = model(*args)
            model.set_mixing_precision("tf32x3")
            got = model(*args)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert not torch.equal(ref[0], got[0])            # the other code path really ran
    _close(got[0], ref[0], "cls tf32x3 vs fp32", rtol=DEC_RTOL, atol=DEC_ATOL)
    _close(got[1], ref[1], "box tf32x3 vs fp32", rtol=DEC_RTOL, atol=DEC_ATOL)
    assert torch.backends.cuda.matmul.allow_tf32 is False   # the global flag is restored


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["bf16x6", "bf16x9"])
def test_tensor_core_mixing_linears_match_sgemm_decoder(precision):
    """Default inference path: AdaptiveMixing's parameter_generator / out_proj on csrc/linear.cu (tcgen05, exact bf16
    operand splitting) against the same decoder with cuBLAS SGEMM for those layers. Both are fp32-grade evaluations of
    the same sums (tests/test_linear.py bounds each against fp64), so the decoder outputs agree to the reordering noise
    that separates two fp32 implementations."""
    d, feats, metas = _fixture_inputs("cuda")
    args = (d["query_bbox"].cuda(), d["query_feat"].cuda(), feats, d["lss_bev"].cuda(), d["radar_bev"].cuda(), None, metas)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        model = _my_model().cuda()
        with torch.no_grad():
            model.set_mixing_precision("fp32")
            ref = model(*args)
            model.set_mixing_precision(precision)
            got = model(*args)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    mix = model.decoder.decoder_layer.mixing
    assert set(mix._split) == {"parameter_generator", "out_proj"}   # the tensor-core path really ran
    _close(got[0], ref[0], f"cls {precision} vs sgemm", rtol=DEC_RTOL, atol=DEC_ATOL)
    _close(got[1], ref[1], f"box {precision} vs sgemm", rtol=DEC_RTOL, atol=DEC_ATOL)


@pytest.mark.gpu
def test_decoder_full_f8_size_cuda_path_vs_cpu_reference_path():
    """racformer_r50_nuimg_704x256_f8 shapes, batch 1, 2 decoder iterations: the production path (sm_100a sampling
    kernels, fused point kernels, grouped output, hoisting, CUDA graph) against the same decoder on the reference's
    PyTorch CPU path (oracle ports, reference schedule), which tests above pin to the unchanged reference."""
    from racformer_b200.graphs import GraphedDecoderForward
    from racformer_b200.synthetic import D_REGION_LIST, PC_RANGE, make_decoder_inputs
    cfg = dict(embed_dims=256, num_frames=8, num_points=4, num_points_bev=4, num_layers=2, num_levels=4, num_classes=10,
               code_size=10, img_depth_num=3, bev_depth_num=5, pc_range=PC_RANGE, num_ray=150,
               d_region_list=D_REGION_LIST, spatial_shapes=(128, 128), num_cams=6)
    torch.manual_seed(0)
    cpu_model = RaCFormerTransformer(**cfg, ops=cpu_oracle_ops(), hoist_invariants=False)
    cpu_model.init_weights()
    fill_parameters_by_name(cpu_model, seed=1)
    cpu_model.eval()
    gpu_model = RaCFormerTransformer(**cfg)
    gpu_model.load_state_dict(cpu_model.state_dict())
    gpu_model.eval().cuda()
    d = make_decoder_inputs(seed=12)
    g = torch.Generator().manual_seed(5)
    d["query_bbox"][..., 8:10] = torch.randn(1, 900, 2, generator=g) * 0.5
    torch.set_num_threads(max(1, (torch.get_num_threads())))
    with torch.no_grad():
        ref_cls, ref_box = cpu_model(d["query_bbox"], d["query_feat"], d["mlvl_feats"], d["lss_bev"], d["radar_bev"], None,
                                     d["img_metas"])
    dg = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items() if k != "mlvl_feats"}
    dg["mlvl_feats"] = [f.cuda() for f in d["mlvl_feats"]]
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        cls, box = GraphedDecoderForward(gpu_model, dg)()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    # fp32 end to end. The view selection and the in-image test of a sample point are discontinuous: a point within
    # rounding noise of an image border flips, and then ALL outputs of that query move (by up to ~0.06), whichever fp32
    # implementation of the dense layers produced the noise (profiles/r01_decoder_precision_diag.json: 3 such queries
    # with cuBLAS SGEMM, 4 with the tensor-core Linear layers on this seed, 0 / 1 on another). So: at most 6 of the
    # 2 x 900 (iteration, query) rows may contain an element outside rtol = atol = 1e-3, every other row is inside,
    # and the typical difference is at rounding level.
    for got, ref, what in ((cls, ref_cls, "cls"), (box, ref_box, "box")):
        a, b = got.detach().cpu().double(), ref.detach().cpu().double()
        bad_rows = ((a - b).abs() > 1e-3 + 1e-3 * b.abs()).any(-1)
        assert int(bad_rows.sum()) <= 6, f"f8 {what}: {int(bad_rows.sum())} query rows differ, worst {float((a - b).abs().max()):.3g}"
        assert float((a - b).abs().median()) < 5e-5, what


def _forward_both_ways(model, inp):
    """(fused call-site kernels where they exist, pure PyTorch operator chain) for the same model and inputs."""
    outs = []
    for fused in (True, False):
        layer = model.decoder.decoder_layer
        layer.row_programs, layer.stacked_heads = fused, fused
        with torch.no_grad():
            outs.append(model(inp["query_bbox"], inp["query_feat"], inp["mlvl_feats"], inp["lss_bev"], inp["radar_bev"],
                              None, inp["img_metas"]))
    return outs


@pytest.mark.gpu
@pytest.mark.parametrize("embed_dims", [384, 512, 768])
def test_decoder_wide_embed_dims_run_and_match_eager_gpu(embed_dims):
    """embed_dims beyond the row programs' shared-memory budget at 8 rows per CTA: the default eval() forward
    must pick a configuration that exists (4 rows per CTA) or fall back to the PyTorch chain -- never raise."""
    from tests.decoder_cases import SMALL_INPUTS
    from racformer_b200.synthetic import make_decoder_inputs
    cfg = dict(SMALL, embed_dims=embed_dims)
    model = RaCFormerTransformer(**cfg)
    model.init_weights()
    fill_parameters_by_name(model, seed=3)
    model = model.eval().cuda()
    inp = make_decoder_inputs(seed=5, device="cuda", embed_dims=embed_dims, **SMALL_INPUTS)
    fused, eager = _forward_both_ways(model, inp)
    _close(fused[0], eager[0], f"cls, embed_dims={embed_dims}", rtol=1e-3, atol=1e-3, max_outlier_frac=0.005)
    _close(fused[1], eager[1], f"box, embed_dims={embed_dims}", rtol=1e-3, atol=1e-3, max_outlier_frac=0.005)


@pytest.mark.gpu
def test_decoder_parameters_as_views_into_a_flat_buffer_gpu():
    """Parameters that are 4-byte-aligned views into one flat buffer (DDP / flat-parameter optimisers): kernels that need
    16-byte aligned weights must report 'unsupported' and the forward must fall back, with the same result."""
    model = _my_model().cuda()
    ref_inp = small_inputs(seed=5, device="cuda")
    with torch.no_grad():
        want = model(ref_inp["query_bbox"], ref_inp["query_feat"], ref_inp["mlvl_feats"], ref_inp["lss_bev"],
                     ref_inp["radar_bev"], None, ref_inp["img_metas"])
    params = list(model.parameters())
    flat = torch.empty(sum(p.numel() + 1 for p in params) + 1, device="cuda")
    off = 1                                           # every view starts at an odd float offset: 4-byte aligned only
    for p in params:
        view = flat[off:off + p.numel()].view_as(p)
        view.copy_(p.data)
        p.data = view
        off += p.numel() + 1
    with torch.no_grad():
        got = model(ref_inp["query_bbox"], ref_inp["query_feat"], ref_inp["mlvl_feats"], ref_inp["lss_bev"],
                    ref_inp["radar_bev"], None, ref_inp["img_metas"])
    _close(got[0], want[0], "cls, flat-buffer parameter views", rtol=1e-3, atol=1e-3, max_outlier_frac=0.005)
    _close(got[1], want[1], "box, flat-buffer parameter views", rtol=1e-3, atol=1e-3, max_outlier_frac=0.005)


@pytest.mark.gpu
def test_weight_caches_follow_data_updates_after_invalidation_and_graph_refuses_stale_replay():
    """updates through `.data` do not bump autograd's version counter, so the bf16x3 operand splits / transposes
    need the explicit epoch (racformer_b200/caches.py); a graph captured before the invalidation must not replay."""
    from racformer_b200 import caches
    from racformer_b200.graphs import GraphedDecoderForward
    model = _my_model().cuda()
    inp = small_inputs(seed=2, device="cuda")
    args = (inp["query_bbox"], inp["query_feat"], inp["mlvl_feats"], inp["lss_bev"], inp["radar_bev"], None, inp["img_metas"])
    with torch.no_grad():
        before = [t.clone() for t in model(*args)]
        graphed = GraphedDecoderForward(model, inp)
        for p in model.parameters():
            p.data.mul_(1.05)                      # EMA-style update: invisible to the version counter
        caches.invalidate_weight_caches()
        after = model(*args)
        model.set_mixing_precision("fp32")          # cuBLAS SGEMM path: no weight-derived caches in the mixing layers
        model.decoder.decoder_layer.row_programs = False
        model.decoder.decoder_layer.stacked_heads = False
        plain = model(*args)
    assert float((after[0] - before[0]).abs().max()) > 1e-4, "the update must change the output"
    for a, b in zip(after, plain):
        bad = (a - b).abs() > 1e-3 + 1e-3 * b.abs()
        assert float(bad.float().mean()) <= 5e-3
    with pytest.raises(RuntimeError, match="invalidated after capture"):
        graphed()
