"""Row programs (racformer_b200/csrc/rowops.cu) against the PyTorch operator chains they replace
(models/racformer_transformer.py:204-262, models/bev_self_attention.py:206-225) and against fp64."""
import ctypes

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from racformer_b200 import _lib


def test_row_program_argument_errors_are_codes():
    from racformer_b200.rowops import LINEAR, LOAD, RowOp
    lib = _lib.load()
    ops = (RowOp * 1)(RowOp(kind=LOAD, dst=0, n=4, ld=4, p0=16))
    assert lib.racf_row_program_forward(None, 1, 8, 8, 1, 4, None) == -1
    assert lib.racf_row_program_forward(ops, 0, 8, 8, 1, 4, None) == -3
    assert lib.racf_row_program_forward(ops, 129, 8, 8, 1, 4, None) == -3         # more than RACF_ROW_MAX_OPS
    assert lib.racf_row_program_forward(ops, 1, 8, 3, 1, 4, None) == -6          # rows per CTA: 4 .. 8
    assert lib.racf_row_program_forward(ops, 1, 8, 9, 1, 4, None) == -6
    assert lib.racf_row_program_forward(ops, 1, 8, 8, 1, 6, None) == -6          # width % 4
    bad = (RowOp * 1)(RowOp(kind=LOAD, dst=1, n=4, ld=4, p0=16))
    assert lib.racf_row_program_forward(bad, 1, 8, 8, 1, 4, None) == -3          # buffer index out of range
    bad = (RowOp * 1)(RowOp(kind=LOAD, dst=0, n=8, ld=8, p0=16))
    assert lib.racf_row_program_forward(bad, 1, 8, 8, 1, 4, None) == -3          # wider than the buffer
    bad = (RowOp * 1)(RowOp(kind=LINEAR, dst=0, src=0, n=4, k=4, p0=16))
    assert lib.racf_row_program_forward(bad, 1, 8, 8, 1, 4, None) == -6          # in-place Linear
    bad = (RowOp * 1)(RowOp(kind=LOAD, dst=0, n=4, ld=4, p0=None))
    assert lib.racf_row_program_forward(bad, 1, 8, 8, 1, 4, None) == -1
    bad = (RowOp * 1)(RowOp(kind=14, dst=0, n=4))
    assert lib.racf_row_program_forward(bad, 1, 8, 8, 1, 4, None) == -6         # unknown operator kind
    bad = (RowOp * 1)(RowOp(kind=11, dst=0, n=4, ld=4, p2=None))
    assert lib.racf_row_program_forward(bad, 1, 8, 8, 1, 4, None) == -1         # LAYERNORM_BWD without the saved input


def _rel_err(a, ref64):
    return float((a.double() - ref64).abs().max() / ref64.abs().max().clamp_min(1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize("rows,rows_per_cta", [(900, 8), (13, 8), (1, 4), (2440, 4), (900, 7), (2440, 6), (33, 5), (900, None)])
def test_mlp_chain_matches_torch_and_fp64(rows, rows_per_cta):
    from racformer_b200 import rowops
    dev = torch.device("cuda", 0)
    torch.manual_seed(rows)
    E = 256
    mlp = nn.Sequential(nn.Linear(3, E), nn.LayerNorm(E), nn.ReLU(), nn.Linear(E, 2 * E), nn.ReLU(), nn.Linear(2 * E, E),
                        nn.LayerNorm(E), nn.Linear(E, 10)).to(dev)
    with torch.no_grad():
        for m in mlp:
            if isinstance(m, nn.LayerNorm):
                m.weight.uniform_(0.5, 1.5)
                m.bias.uniform_(-0.5, 0.5)
    box = torch.rand(rows, 10, device=dev)
    resid = torch.randn(rows, E, device=dev)
    with torch.no_grad():
        h = mlp[2](mlp[1](mlp[0](box[:, :3])))
        h = mlp[5](mlp[4](mlp[3](h)))
        feat_ref = mlp[6](h + resid)
        out_ref = mlp[7](feat_ref)
        m64 = nn.Sequential(*[type(m)(*([m.in_features, m.out_features] if isinstance(m, nn.Linear) else
                                        [m.normalized_shape] if isinstance(m, nn.LayerNorm) else [])) for m in mlp]).to(dev).double()
        m64.load_state_dict(mlp.state_dict())
        h64 = m64[2](m64[1](m64[0](box[:, :3].double())))
        h64 = m64[5](m64[4](m64[3](h64)))
        feat64 = m64[6](h64 + resid.double())
        out64 = m64[7](feat64)
    p = rowops.RowProgram(rows, width=2 * E, num_bufs=3, rows_per_cta=rows_per_cta)
    p.load(0, box, n=3)
    p.linear(1, 0, mlp[0])
    p.layernorm(1, mlp[1], relu=True)
    p.linear(2, 1, mlp[3], relu=True)
    p.linear(0, 2, mlp[5])
    p.load(1, resid)
    p.add(0, 1, E)
    p.layernorm(0, mlp[6])
    feat = p.store(0, E)
    p.linear(1, 0, mlp[7])
    out = p.store(1, 10)
    p.run()
    torch.cuda.synchronize()
    assert torch.allclose(feat, feat_ref, rtol=1e-4, atol=2e-5), float((feat - feat_ref).abs().max())
    assert torch.allclose(out, out_ref, rtol=1e-4, atol=2e-5), float((out - out_ref).abs().max())
    # no further from fp64 than the PyTorch chain (plus slack for the different summation order)
    assert _rel_err(feat, feat64) <= max(2 * _rel_err(feat_ref, feat64), 2e-6)
    assert _rel_err(out, out64) <= max(2 * _rel_err(out_ref, out64), 2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("batch,with_logits", [(1, True), (2, True), (1, False)])
def test_queue_fusion_and_column_offsets(batch, with_logits):
    """LOAD_QUEUE + Linear into a column range + residual + LayerNorm on that range = BEVSelfAttention.attend's tail
    (models/bev_self_attention.py:206-225) followed by the branch norm and the concatenation."""
    from racformer_b200 import rowops
    dev = torch.device("cuda", 0)
    torch.manual_seed(3 + batch)
    Q, T, E = 77, 8, 256
    proj, norm, fusion = nn.Linear(E, E).to(dev), nn.LayerNorm(E).to(dev), nn.Linear(2 * E, E).to(dev)
    msda = torch.randn(batch * T, Q, E, device=dev)
    logits = torch.randn(batch, Q, T, device=dev) if with_logits else None
    query = torch.randn(batch, Q, E, device=dev)
    first = torch.randn(batch, Q, E, device=dev)
    with torch.no_grad():
        out = msda.permute(1, 2, 0).reshape(Q, E, batch, T)
        if with_logits:
            qw = torch.softmax(logits.permute(1, 0, 2).reshape(Q, 1, batch, T), dim=-1)
            out = torch.sum(out * qw, dim=-1)
        else:
            out = torch.sum(out, dim=-1) / T
        branch = norm(proj(out.permute(2, 0, 1)) + query)
        ref = fusion(torch.cat((first, branch), dim=-1))
    p = rowops.RowProgram(batch * Q, width=2 * E, num_bufs=3)
    p.load(0, first)
    p.load(1, query)
    p.load_queue(2, msda, logits, Q, T)
    p.linear(0, 2, proj, dst_col=E)
    p.add(0, 1, E, dst_col=E)
    p.layernorm(0, norm, col=E)
    p.linear(1, 0, fusion)
    got = p.store(1, E).view(batch, Q, E)
    p.run()
    torch.cuda.synchronize()
    assert torch.allclose(got, ref, rtol=1e-4, atol=2e-5), float((got - ref).abs().max())


@pytest.mark.gpu
def test_decoder_layer_row_programs_match_the_pytorch_chain():
    """The decoder with the row programs on vs off (same kernels everywhere else)."""
    from racformer_b200.decoder import RaCFormerTransformer
    from racformer_b200.synthetic import fill_parameters_by_name
    from tests.decoder_cases import SMALL, small_inputs
    dev = torch.device("cuda", 0)
    model = RaCFormerTransformer(**SMALL).to(dev).eval()
    fill_parameters_by_name(model)
    inp = small_inputs(batch=1)
    args = [inp["query_bbox"].to(dev), inp["query_feat"].to(dev), [f.to(dev) for f in inp["mlvl_feats"]],
            inp["lss_bev"].to(dev), inp["radar_bev"].to(dev), None, inp["img_metas"]]
    layer = model.decoder.decoder_layer
    with torch.no_grad():
        layer.row_programs = True
        cls_a, box_a = model(*args)
        layer.row_programs = False
        cls_b, box_b = model(*args)
    assert torch.allclose(cls_a, cls_b, rtol=1e-4, atol=2e-4), float((cls_a - cls_b).abs().max())
    assert torch.allclose(box_a, box_b, rtol=1e-4, atol=2e-4), float((box_a - box_b).abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("batch,frames", [(1, 8), (2, 2), (1, 1)])
def test_refine_bbox_matches_the_pytorch_chain(batch, frames):
    """racf_refine_bbox_forward vs refine_bbox + velocity scaling + theta_d2xy_coods (racformer_transformer.py:255-279)."""
    from racformer_b200 import points
    from racformer_b200.decoder import RaCFormerTransformerDecoderLayer, theta_d2xy_coods
    from racformer_b200.synthetic import PC_RANGE
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(11 + batch)
    Q = 333
    proposal = torch.rand(batch, Q, 10, device=dev, generator=g)
    proposal[0, :4, 1] = torch.tensor([0.0, 1.0, 1e-7, 1 - 1e-7], device=dev)      # inverse_sigmoid clamps
    delta = torch.randn(batch, Q, 10, device=dev, generator=g) * 2
    time_diff = torch.arange(frames, device=dev, dtype=torch.float32).repeat(batch, 1) * 0.5
    if batch > 1:
        time_diff[1, 1:] = 0.0                                                      # < 1e-5 -> 1
    layer = RaCFormerTransformerDecoderLayer(256, num_frames=frames, pc_range=PC_RANGE, num_ray=150)
    with torch.no_grad():
        ref = layer.refine_bbox(proposal, delta)
        if frames > 1:
            td = torch.where(time_diff < 1e-5, torch.ones_like(time_diff), time_diff)
            ref = torch.cat([ref[..., :8], ref[..., 8:] / td[:, 1:2, None]], dim=-1)
        ref_xy = theta_d2xy_coods(ref)
    pred, pred_xy = points.refine_bbox(proposal, delta, time_diff, 150)
    torch.cuda.synchronize()
    assert torch.allclose(pred, ref, rtol=1e-6, atol=1e-6), float((pred - ref).abs().max())
    assert torch.allclose(pred_xy, ref_xy, rtol=1e-6, atol=1e-6), float((pred_xy - ref_xy).abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("batch,num_query", [(1, 900), (2, 77), (1, 5)])
def test_sasa_attention_core_matches_the_masked_attention(batch, num_query):
    """racf_sasa_attention_forward vs the reference construction: dist * tau as additive mask of scaled dot-product
    attention (models/racformer_transformer.py:308-336)."""
    from racformer_b200 import points
    from racformer_b200.decoder import decode_bbox, theta_d2xy_coods
    from racformer_b200.synthetic import PC_RANGE
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(num_query)
    H, D = 8, 32
    E = H * D
    qkv = torch.randn(batch, num_query, 3 * E, device=dev, generator=g)
    tau = torch.rand(batch, num_query, H, device=dev, generator=g) * 2
    ray = torch.rand(batch, num_query, 10, device=dev, generator=g)
    centres = decode_bbox(theta_d2xy_coods(ray), PC_RANGE)[..., :2]
    dist = -torch.norm(centres[:, :, None, :] - centres[:, None, :, :], dim=-1)
    mask = dist[:, None] * tau.permute(0, 2, 1)[..., None]                                   # [B,H,Q,Q]
    q, k, v = (t.view(batch, num_query, H, D).transpose(1, 2) for t in qkv.split(E, dim=-1))
    ref = F.scaled_dot_product_attention(q.double(), k.double(), v.double(), attn_mask=mask.double())
    ref = ref.transpose(1, 2).reshape(batch, num_query, E)
    ref32 = F.scaled_dot_product_attention(q, k, v, attn_mask=mask).transpose(1, 2).reshape(batch, num_query, E)
    got = points.sasa_attention(qkv, tau, ray, PC_RANGE, H)
    torch.cuda.synchronize()
    err, err32 = float((got.double() - ref).abs().max()), float((ref32.double() - ref).abs().max())
    # the kernel recomputes the bias (up to ~200 in magnitude here, ulp 1.5e-5) instead of reading the fp32 mask the fp64
    # evaluation was given: a last-ulp difference in a distance moves a softmax weight by ~1e-5
    print("sasa err vs fp64", err, "torch fp32 SDPA err", err32)
    assert err <= max(2 * err32, 3e-5), (err, err32)
    assert torch.allclose(got, ref32, rtol=1e-4, atol=5e-5), float((got - ref32).abs().max())


@pytest.mark.gpu
def test_decoder_layer_fused_self_attention_matches_the_pytorch_chain():
    from racformer_b200.decoder import RaCFormerTransformer
    from racformer_b200.synthetic import fill_parameters_by_name
    from tests.decoder_cases import SMALL, small_inputs
    dev = torch.device("cuda", 0)
    model = RaCFormerTransformer(**SMALL).to(dev).eval()
    fill_parameters_by_name(model)
    inp = small_inputs(batch=2)
    args = [inp["query_bbox"].to(dev), inp["query_feat"].to(dev), [f.to(dev) for f in inp["mlvl_feats"]],
            inp["lss_bev"].to(dev), inp["radar_bev"].to(dev), None, inp["img_metas"]]
    layer = model.decoder.decoder_layer
    with torch.no_grad():
        layer.fused_self_attention = True
        cls_a, box_a = model(*args)
        layer.fused_self_attention = False
        cls_b, box_b = model(*args)
    assert torch.allclose(cls_a, cls_b, rtol=1e-4, atol=2e-4), float((cls_a - cls_b).abs().max())
    assert torch.allclose(box_a, box_b, rtol=1e-4, atol=2e-4), float((box_a - box_b).abs().max())


def test_chunked_transpose_layout_matches_the_header():
    """rowops.chunked_transpose: element [c, k, j] = W[c * 256 + j, k], zero padded (include/racformer_ops.h, LINEAR)."""
    from racformer_b200.rowops import CHUNK_COLS, chunked_transpose
    torch.manual_seed(0)
    for n, k in ((10, 7), (256, 32), (300, 5), (768, 256)):
        w = torch.randn(n, k)
        t = chunked_transpose(w)
        chunks = (n + CHUNK_COLS - 1) // CHUNK_COLS
        assert t.shape == (chunks, k, CHUNK_COLS) and t.is_contiguous()
        for c in range(chunks):
            cols = min(CHUNK_COLS, n - c * CHUNK_COLS)
            assert torch.equal(t[c, :, :cols], w[c * CHUNK_COLS:c * CHUNK_COLS + cols].t())
            assert not t[c, :, cols:].any()


def test_row_program_builder_rejects_cpu_tensors_and_bad_shapes():
    """There is no CPU path: the builder raises instead of falling back."""
    from racformer_b200.rowops import MAX_OPS, RowProgram
    p = RowProgram(4, width=8)
    with pytest.raises(RuntimeError, match="fp32 CUDA"):
        p.load(0, torch.zeros(4, 8))
    with pytest.raises(RuntimeError, match="empty row program"):
        p.run()
    with pytest.raises(RuntimeError, match="queue values"):
        p.load_queue(0, torch.zeros(8, 4, 8), None, 4, 2)
    for _ in range(MAX_OPS):
        p.add(0, 1, 4)
    with pytest.raises(RuntimeError, match="too long"):
        p.add(0, 1, 4)


def test_rows_per_cta_choice_minimises_waves_times_rows():
    """Host logic of rowops.choose_rows_per_cta (no GPU: 148 SMs assumed): a launch costs waves x rows per CTA; ties go to the
    larger tile; the row buffers must fit the shared memory the weight-tile ring leaves."""
    from racformer_b200.rowops import choose_rows_per_cta
    assert choose_rows_per_cta(900, 3, 768) == 7        # 129 CTAs, one wave (8 rows: 113 CTAs x 8; 6 rows: two waves)
    assert choose_rows_per_cta(2440, 3, 768) == 6       # 407 CTAs = three waves of 6 (8 rows: 305 CTAs = three waves of 8)
    assert choose_rows_per_cta(148 * 8, 2, 256) == 8    # exactly one wave of full tiles
    assert choose_rows_per_cta(100, 2, 256) == 4        # few rows: the smallest tile, one wave either way
    assert choose_rows_per_cta(900, 3, 1536) == 4       # wide buffers: only 4 rows fit
    for rows in (1, 37, 900, 1220, 2440, 7200, 50000):
        for nb, w in ((2, 256), (3, 768), (2, 1024)):
            r = choose_rows_per_cta(rows, nb, w)
            assert 4 <= r <= 8 and r * (nb * w * 4 + 3 * 256 * 4) <= 98 * 1024


def test_weight_cache_epoch_invalidates_keys():
    from racformer_b200 import caches
    e0 = caches.cache_epoch()
    caches.invalidate_weight_caches()
    assert caches.cache_epoch() == e0 + 1
    from racformer_b200.decoder import RaCFormerTransformer
    from tests.decoder_cases import SMALL
    m = RaCFormerTransformer(**SMALL)
    e1 = caches.cache_epoch()
    m.eval()                                            # mode change: EMA-style .data updates may have happened in between
    assert caches.cache_epoch() == e1 + 1
    m.eval()                                            # no change of mode: captured graphs stay valid
    assert caches.cache_epoch() == e1 + 1
    m.load_state_dict(m.state_dict())
    assert caches.cache_epoch() == e1 + 2
