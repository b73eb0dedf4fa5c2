"""GPU: the fused sampling-point kernels (SURVEY 8f-2 "next" row) against the eager PyTorch op chain of the harness,
which tests/test_decoder.py pins to the unchanged reference.

Bar: the packed tensors handed to the sampling ops agree -- coordinates and weights to 2e-6 absolute + 4e-6 relative (normalised
units; the kernel mirrors PyTorch's fp32 op order, libm calls are the same CUDA functions), the selected camera view
exactly, except for points whose projection lies within 1e-5 of an image border in some view (there a last-ulp
difference can legitimately flip the visibility test); at most 1e-4 of the points may be in that set.
"""
import pytest
import torch

from racformer_b200.decoder import RaCFormerTransformer, SamplingOps
from racformer_b200.synthetic import D_REGION_LIST, PC_RANGE, fill_parameters_by_name, make_decoder_inputs
from tests.decoder_cases import SMALL, small_inputs

pytestmark = pytest.mark.gpu


class Spy(SamplingOps):
    def __init__(self):
        super().__init__()
        self.msmv_args, self.msda_args = [], []
        real_msmv, real_msda = self.msmv, self.msda

        def msmv(feats, loc, w):
            self.msmv_args.append((loc.clone(), w.clone()))
            return real_msmv(feats, loc, w)

        def msda(value, shapes, lsi, loc, aw, step):
            self.msda_args.append((loc.clone(), aw.clone()))
            return real_msda(value, shapes, lsi, loc, aw, step)

        real_grouped = self.msmv_grouped

        def grouped(feats, loc, w, T, G):
            self.msmv_args.append((loc.clone(), w.clone()))
            return real_grouped(feats, loc, w, T, G)

        self.msmv, self.msda, self.msmv_grouped = msmv, msda, grouped
        self.msda_pair = None      # record every MSDA call: the paired launch (tested on its own below) bypasses self.msda


def _run(cfg, inputs, fused, seed=3):
    spy = Spy()
    model = RaCFormerTransformer(**cfg, ops=spy)
    model.init_weights()
    fill_parameters_by_name(model, seed=seed)
    model.eval().cuda()
    model.set_fused_points(fused)
    with torch.no_grad():
        out = model(inputs["query_bbox"], inputs["query_feat"], inputs["mlvl_feats"], inputs["lss_bev"],
                    inputs["radar_bev"], None, inputs["img_metas"])
    return spy, out


def _compare(cfg, inputs, num_views):
    eager, out_e = _run(cfg, inputs, fused=False)
    fused, out_f = _run(cfg, inputs, fused=True)
    # only the first decoder iteration has bit-identical inputs in both runs; later ones inherit tiny differences
    (loc_e, w_e), (loc_f, w_f) = eager.msmv_args[0], fused.msmv_args[0]
    assert loc_e.shape == loc_f.shape and w_e.shape == w_f.shape
    view_e = (loc_e[..., 2] * (num_views - 1)).round().long()
    view_f = (loc_f[..., 2] * (num_views - 1)).round().long()
    same = view_e == view_f
    assert float((~same).float().mean()) <= 1e-4, "too many view-selection differences"
    # Points no camera sees keep view 0 with ill-conditioned, far out-of-range coordinates (division by eps = 1e-5);
    # they contribute exactly zero as long as every tap stays out of range, so they are compared through the
    # sampling kernel's own validity masks. Points with at least one in-range tap must agree in their coordinates.
    from racformer_b200 import wrapper
    hw = [tuple(f.shape[-2:]) for f in inputs["mlvl_feats"]]
    _, mask_e = wrapper.msmv_tap_masks(hw, loc_e.contiguous(), num_views)
    _, mask_f = wrapper.msmv_tap_masks(hw, loc_f.contiguous(), num_views)
    assert float((mask_e != mask_f).any(-1).float().mean()) <= 1e-4
    live = ((mask_e & 1).bool().any(-1) | (mask_f & 1).bool().any(-1)) & same
    assert float(live.float().mean()) > 0.3
    xy_e, xy_f = loc_e[..., :2][live], loc_f[..., :2][live]
    # projections with a small camera depth are ill-conditioned (rounding of cx, cy is magnified by 1/cz, which also
    # makes the point land far outside the image): all live points must agree to 2e-5 of their larger coordinate,
    # and at least 99 % of them to the tight 2e-6 + 4e-6 relative bound
    diff = (xy_e - xy_f).abs()
    excess = diff - (2e-6 + 2e-5 * xy_e.abs().max(-1, keepdim=True).values)
    worst = int(torch.argmax(excess.max(-1).values))
    assert float(excess.max()) <= 0, (
        f"{int((excess > 0).any(-1).sum())}/{excess.shape[0]} live points differ; worst eager={xy_e[worst].tolist()} "
        f"fused={xy_f[worst].tolist()}")
    tight = (diff <= 2e-6 + 4e-6 * xy_e.abs()).all(-1)
    assert float(tight.float().mean()) >= 0.99
    assert float((w_e - w_f).abs().max()) <= 2e-6
    for (le, ae), (lf, af) in zip(eager.msda_args[:2], fused.msda_args[:2]):   # radar + lss branch, iteration 0
        assert le.shape == lf.shape and ae.shape == af.shape
        assert float((le - lf).abs().max()) <= 2e-6
        assert float((ae - af).abs().max()) <= 2e-6
    # whole decoder output: same tolerance as the decoder-vs-reference comparison on GPU
    for a, b in zip(out_e, out_f):
        bad = (a - b).abs() > 1e-3 + 1e-3 * b.abs()
        assert float(bad.float().mean()) <= 5e-3
    return float((~same).float().mean())


def test_fused_points_small_config_with_velocity_and_batch2():
    d = small_inputs(seed=21, device="cuda", batch=2)     # B=2: the B*T*G / B*G*T and T*B packings differ from B=1
    _compare(SMALL, d, num_views=3)


def test_fused_points_f8_shapes():
    cfg = dict(embed_dims=256, num_frames=8, num_points=4, num_points_bev=4, num_layers=1, num_levels=4, num_classes=10,
               code_size=10, img_depth_num=3, bev_depth_num=5, pc_range=PC_RANGE, num_ray=150,
               d_region_list=D_REGION_LIST, spatial_shapes=(128, 128), num_cams=6)
    d = make_decoder_inputs(seed=4, device="cuda")
    g = torch.Generator().manual_seed(1)
    d["query_bbox"][..., 8:10] = (torch.randn(1, 900, 2, generator=g) * 0.5).cuda()
    _compare(cfg, d, num_views=6)


def test_grouped_output_variant_is_bit_identical_to_permuted_reference_layout():
    """racf_msmv_forward_grouped = racf_msmv_forward + sampling_4d's un-packing (sparsebev_sampling.py:128-131)."""
    from racformer_b200 import wrapper
    from tests.helpers import make_msmv_inputs
    B, T, G, Q, P, C = 2, 3, 4, 50, 12, 64
    feats, loc, w, _ = make_msmv_inputs(3, Bp=B * T * G, N=6, C=C, Q=Q, P=P, shapes=[(16, 44), (8, 22), (4, 11), (2, 6)],
                                        lo=-0.1, hi=1.1, device="cuda")
    plain = wrapper.msmv_forward(feats, loc, w)
    want = plain.reshape(B, T, G, Q, C, P).permute(0, 3, 2, 1, 5, 4).flatten(3, 4)
    got = wrapper.msmv_forward_grouped(feats, loc, w, T, G)
    assert got.shape == (B, Q, G, T * P, C) and torch.equal(got, want)
    # ragged point count and a level count without a fast kernel (falls back to forward + permute)
    feats, loc, w, _ = make_msmv_inputs(4, Bp=T * G, N=3, C=C, Q=7, P=5, shapes=[(8, 22), (4, 11), (2, 6)], device="cuda")
    plain = wrapper.msmv_forward(feats, loc, w)
    want = plain.reshape(1, T, G, 7, C, 5).permute(0, 3, 2, 1, 5, 4).flatten(3, 4)
    assert torch.equal(wrapper.msmv_forward_grouped(feats, loc, w, T, G), want)
    feats, loc, w, _ = make_msmv_inputs(5, Bp=T * G, N=3, C=C, Q=7, P=6, shapes=[(8, 22), (4, 11)], device="cuda")
    plain = wrapper.msmv_forward(feats, loc, w)
    want = plain.reshape(1, T, G, 7, C, 6).permute(0, 3, 2, 1, 5, 4).flatten(3, 4)
    assert torch.equal(wrapper.msmv_forward_grouped(feats, loc, w, T, G), want)


def test_channel_last_relayout_kernel_equals_permute():
    """racf_to_sampling_layout = reshape/permute/contiguous of racformer_transformer.py:112-124, bit for bit."""
    from racformer_b200 import points
    for (B, T, N, G, H, W) in ((1, 8, 6, 4, 8, 22), (2, 2, 3, 4, 5, 7), (1, 1, 6, 4, 64, 176)):
        feat = torch.randn(B, T * N, G * 64, H, W, device="cuda")
        want = feat.reshape(B, T, N, G, 64, H, W).permute(0, 1, 3, 2, 5, 6, 4).reshape(B * T * G, N, H, W, 64).contiguous()
        assert torch.equal(points.to_sampling_layout(feat, N, G), want)


@pytest.mark.parametrize("tensor_cores", [False, True])
@pytest.mark.parametrize("p_in", [96, 24, 128, 4, 48, 16, 112])
def test_fused_adaptive_mixing_core_equals_pytorch_chain(p_in, tensor_cores):
    """csrc/mixing.cu (CUDA cores, fp32 FMA) and csrc/mixing_tc.cu (tcgen05, exact bf16x3 operand split, fp32 accumulation;
    used when P_in % 16 == 0) vs the op chain of AdaptiveMixing.inner_forward (racformer_transformer.py:592-604): all are
    fp32-grade evaluations, so they agree to rounding / summation order (outputs are layer-normalised, O(1)); against an
    fp64 evaluation of the same chain the kernels must be as close as PyTorch's own fp32 chain is."""
    import torch.nn.functional as F
    from racformer_b200 import points
    g = torch.Generator().manual_seed(p_in)
    QG, C, P_out = 300, 64, 128
    x = torch.randn(QG, p_in, C, generator=g).cuda()
    params = (torch.randn(QG, C * C + P_out * p_in, generator=g) * 0.2).cuda()
    m, s = params.split([C * C, P_out * p_in], 1)
    t = torch.matmul(x, m.reshape(QG, C, C))
    t = F.relu(F.layer_norm(t, [p_in, C]))
    ref = torch.matmul(s.reshape(QG, P_out, p_in), t)
    ref = F.relu(F.layer_norm(ref, [P_out, C]))
    got = points.adaptive_mixing_core(x, params, P_out, tensor_cores=tensor_cores)
    assert got is not None and got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 2e-5
    assert torch.equal(got, points.adaptive_mixing_core(x, params, P_out, tensor_cores=tensor_cores)), "deterministic"
    t64 = torch.matmul(x.double(), m.reshape(QG, C, C).double())
    t64 = F.relu(F.layer_norm(t64, [p_in, C]))
    ref64 = F.relu(F.layer_norm(torch.matmul(s.reshape(QG, P_out, p_in).double(), t64), [P_out, C]))
    err_kernel, err_torch = float((got.double() - ref64).abs().max()), float((ref.double() - ref64).abs().max())
    assert err_kernel <= max(2.0 * err_torch, 5e-6), (err_kernel, err_torch)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("qg,p_in", [(4, 96), (148 * 7 + 8, 96), (1000, 64), (40, 48), (40, 80), (52, 16), (300, 32)])
def test_tensor_core_mixing_kernels_agree_and_are_fp32_grade(qg, p_in, variant):
    """Both tcgen05 kernels -- phase-serial csrc/mixing_tc.cu (variant 1) and warp-specialised csrc/mixing_ws.cu (variant 2,
    the default for P_in <= 96) -- on item counts that give a CTA 1..8 items (ring wrap-around, accumulator double buffers,
    every slot kind of ragged P_in), fp32 and tiled output: error against an fp64 evaluation of
    AdaptiveMixing.inner_forward (racformer_transformer.py:592-604) within the same bar as the PyTorch fp32 chain, the
    tiled pieces are the exact split of the kernel's own fp32 result, and the launch is deterministic."""
    import torch.nn.functional as F
    from racformer_b200 import linear, points
    g = torch.Generator(device="cuda").manual_seed(qg * 131 + p_in)
    C, P_out = 64, 128
    x = torch.randn(qg, p_in, C, device="cuda", generator=g)
    params = torch.randn(qg, C * C + P_out * p_in, device="cuda", generator=g) * 0.2
    m, s = params.double().split([C * C, P_out * p_in], 1)
    t64 = F.relu(F.layer_norm(torch.matmul(x.double(), m.reshape(qg, C, C)), [p_in, C]))
    ref64 = F.relu(F.layer_norm(torch.matmul(s.reshape(qg, P_out, p_in), t64), [P_out, C]))
    got = points.adaptive_mixing_core(x, params, P_out, variant=variant)
    assert float((got.double() - ref64).abs().max()) <= 5e-6
    assert torch.equal(got, points.adaptive_mixing_core(x, params, P_out, variant=variant)), "deterministic"
    tiled = points.adaptive_mixing_core(x, params, P_out, split=True, tiled_groups=4, variant=variant)
    assert torch.equal(linear.untile(tiled).double().sum(0), got.reshape(qg // 4, -1).double())
    # the pieces are the round-to-nearest split (largest piece = bf16(x)), as racf_split_bf16x3 produces it
    assert torch.equal(linear.untile(tiled)[0], got.reshape(qg // 4, -1).to(torch.bfloat16))


@pytest.mark.gpu
def test_mixing_variant_selection_errors():
    from racformer_b200 import _lib, points
    x = torch.randn(8, 128, 64, device="cuda")
    params = torch.randn(8, 64 * 64 + 128 * 128, device="cuda")
    assert points.adaptive_mixing_core(x, params, 128, variant=0) is not None        # falls back to the phase-serial kernel
    with pytest.raises(RuntimeError):
        points.adaptive_mixing_core(x, params, 128, variant=2)                       # warp-specialised: P_in <= 96 only
    with pytest.raises(RuntimeError):
        points.adaptive_mixing_core(x[:, :96].contiguous(), params[:, :64 * 64 + 128 * 96].contiguous(), 128, variant=7)


def test_fused_adaptive_mixing_core_declines_unsupported_shapes():
    from racformer_b200 import points
    x = torch.randn(4, 10, 64, device="cuda")          # in_points not a multiple of 4
    assert points.adaptive_mixing_core(x, torch.randn(4, 64 * 64 + 128 * 10, device="cuda"), 128) is None
    x = torch.randn(4, 8, 32, device="cuda")           # C != 64
    assert points.adaptive_mixing_core(x, torch.randn(4, 32 * 32 + 128 * 8, device="cuda"), 128) is None


@pytest.mark.gpu
@pytest.mark.parametrize("shape,extra", [((8, 256, 16, 16), 64), ((2, 37, 5, 7), 3), ((1, 64, 9, 33), 0)])
def test_to_channels_last_is_a_pure_copy(shape, extra):
    """racf_chw_to_hwc: same values as .contiguous(memory_format=channels_last), also into the concatenation buffer."""
    from racformer_b200 import points
    x = torch.randn(*shape, device="cuda")
    dense, both = points.to_channels_last(x, extra)
    assert dense.is_contiguous(memory_format=torch.channels_last) and torch.equal(dense, x)
    if extra:
        assert both.shape == (shape[0], shape[1] + extra, shape[2], shape[3])
        assert both.is_contiguous(memory_format=torch.channels_last) and torch.equal(both[:, :shape[1]], x)
    else:
        assert both is None


@pytest.mark.gpu
@pytest.mark.parametrize("channels_last", [False, True])
def test_to_sampling_layout_f16_is_bit_identical_to_upcast_then_relayout(channels_last):
    """racf_to_sampling_layout_f16: fp16 storage (NCHW or cuDNN's NHWC) -> fp32 sampling layout == .float() + permute."""
    from racformer_b200 import points
    from racformer_b200.decoder import to_sampling_layout
    g = torch.Generator(device="cuda").manual_seed(3)
    B, T, N, G, C, H, W = 2, 3, 3, 4, 64, 5, 11
    x = torch.randn(B, T * N, G * C, H, W, device="cuda", generator=g).half()
    if channels_last:
        x = x.flatten(0, 1).contiguous(memory_format=torch.channels_last).view(B, T * N, G * C, H, W)
        assert not x.is_contiguous()
    got = points.to_sampling_layout_f16(x, N, G)
    want = x.float().reshape(B, T, N, G, C, H, W).permute(0, 1, 3, 2, 5, 6, 4).reshape(B * T * G, N, H, W, C).contiguous()
    assert torch.equal(got, want)
    assert torch.equal(to_sampling_layout(x, N, G), want)


def test_msda_forward_pair_equals_two_forwards_bit_for_bit():
    """racf_msda_forward_pair: the radar and the LSS BEV branch of an iteration in one launch (grid.y = 2)."""
    from racformer_b200.multi_scale_deformable_attn_function import ext_module, msda_forward_pair
    g = torch.Generator(device="cuda").manual_seed(3)
    B, H, W, M, D, Q, P = 8, 32, 48, 4, 64, 77, 20
    shapes = torch.tensor([[H, W]], dtype=torch.long, device="cuda")
    lsi = torch.zeros(1, dtype=torch.long, device="cuda")
    probs = []
    for _ in range(2):
        value = torch.randn(B, H * W, M, D, device="cuda", generator=g)
        loc = torch.rand(B, Q, M, 1, P, 2, device="cuda", generator=g) * 1.2 - 0.1        # some taps outside the map
        aw = torch.softmax(torch.randn(B, Q, M, 1, P, device="cuda", generator=g), -1)
        probs.append((value, loc, aw))
    (va, la, aa), (vb, lb, ab) = probs
    out_a, out_b = msda_forward_pair(va, la, aa, vb, lb, ab, shapes, lsi, 64)
    assert torch.equal(out_a, ext_module.ms_deform_attn_forward(va, shapes, lsi, la, aa, 64))
    assert torch.equal(out_b, ext_module.ms_deform_attn_forward(vb, shapes, lsi, lb, ab, 64))
    with pytest.raises(RuntimeError):
        msda_forward_pair(va, la, aa, vb[:, :100].contiguous(), lb, ab, shapes, lsi, 64)


def test_decoder_layer_paired_bev_launch_equals_separate_launches():
    """DecoderLayer._bev_pair (one MSDA launch for both BEV branches) against the branch-by-branch path: identical outputs."""
    model = RaCFormerTransformer(**SMALL)
    model.init_weights()
    fill_parameters_by_name(model, seed=3)
    model.eval().cuda()
    d = small_inputs(device="cuda")
    args = (d["query_bbox"], d["query_feat"], d["mlvl_feats"], d["lss_bev"], d["radar_bev"], None, d["img_metas"])
    layer = model.decoder.decoder_layer
    assert layer.bev_pair_launch and model.ops.msda_pair is not None
    with torch.no_grad():
        paired = model(*args)
        layer.bev_pair_launch = False
        separate = model(*args)
        layer.bev_pair_launch = True
    for a, b in zip(paired, separate):
        assert torch.equal(a, b)


def test_convgru_gates_kernel_matches_pytorch_ops():
    """racf_convgru_gates_forward (one launch) against ConvGRUCell's nine elementwise ops (racformer_transformer.py:713-720) on
    channel-last tensors: same libm calls, separate multiplies / adds -> agreement to the last bit or two."""
    from racformer_b200 import points
    g = torch.Generator(device="cuda").manual_seed(3)
    N, hc, H, W = 2, 64, 17, 23
    gates = (torch.randn(N, 3 * hc, H, W, device="cuda", generator=g) * 2).contiguous(memory_format=torch.channels_last)
    h_prev = torch.randn(N, hc, H, W, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
    got = points.convgru_gates(gates, h_prev)
    z, r, cand = torch.split(gates, hc, dim=1)
    z, r = torch.sigmoid(z), torch.sigmoid(r)
    want = (1 - z) * h_prev + z * torch.tanh(cand + r * h_prev)
    assert got.shape == want.shape and got.is_contiguous(memory_format=torch.channels_last)
    assert float((got - want).abs().max()) <= 2e-7
    assert points.convgru_gates(gates.contiguous(), h_prev) is None          # NCHW tensors: the caller keeps the PyTorch ops
    # the encoder with and without the fused gates / channel-last recurrence
    from racformer_b200.decoder import RadarBEVTemporalEncoder
    torch.manual_seed(0)
    enc = RadarBEVTemporalEncoder(64, 16, num_frames=4).cuda().eval()
    bev = torch.randn(1, 4, 64, 16, 16, device="cuda", generator=g)
    with torch.no_grad():
        a = enc(bev)
        enc.convGRU.convGRUCell.fused_gates = False
        b = enc(bev)
    assert float((a - b).abs().max()) <= 1e-5


def test_upsample2x_bilinear_kernel_matches_aten():
    from racformer_b200 import points
    g = torch.Generator(device="cuda").manual_seed(4)
    for N, C, H, W in ((2, 64, 13, 9), (8, 64, 64, 64), (1, 8, 1, 5)):
        x = torch.randn(N, C, H, W, device="cuda", generator=g).contiguous(memory_format=torch.channels_last)
        got = points.upsample2x_bilinear(x)
        want = torch.nn.functional.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
        assert got.shape == want.shape and got.is_contiguous(memory_format=torch.channels_last)
        assert float((got - want).abs().max()) <= 1e-6
    assert points.upsample2x_bilinear(torch.randn(2, 8, 4, 4, device="cuda")) is None      # NCHW: PyTorch ops


def test_deferred_temporal_fusion_bias_reaches_value_proj():
    """BEVSampling.prepare_value at batch 1 skips cuDNN's separate bias pass of the temporal encoder's last convolution and adds
    the bias inside value_proj's operand split, next to the positional encoding: same value map as the plain order."""
    from racformer_b200.decoder import BEVSampling
    torch.manual_seed(0)
    m = BEVSampling(embed_dims=64, num_frames=4, num_points=2, num_heads=4, num_levels=1, pc_range=PC_RANGE,
                    spatial_shapes=(16, 16), depth_num=3, temp_radar=True).cuda().eval()
    with torch.no_grad():
        for p in m.parameters():
            p.normal_(0, 0.2)
        bev = torch.randn(1, 4, 64, 16, 16, device="cuda")
        value, hw = m.prepare_value(bev)
        feats = m.temporal_encoder(bev)                                    # bias added by the convolution
        pos = m.positional_encoding(1, 16, 16, bev.device)
        want = m.attention.project_value(feats, pos.reshape(64, 16, 16))
    assert tuple(hw) == (16, 16) and value.shape == want.shape
    assert float((value - want).abs().max()) <= 2e-5 * float(want.abs().max())
    assert float(m.temporal_encoder.temporal_fusion.bias.detach().abs().max()) > 0.05      # the test has power
