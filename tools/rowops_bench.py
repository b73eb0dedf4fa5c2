"""Row programs at the f8 decoder shapes: time of the position-encoder and tail programs (csrc/rowops.cu) against the
PyTorch operator chains they replace, and the per-Linear cost. Usage: python tools/rowops_bench.py [rows_per_cta]"""
import json
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from racformer_b200 import rowops  # noqa: E402
from racformer_b200.decoder import RaCFormerTransformerDecoderLayer  # noqa: E402
from racformer_b200.synthetic import PC_RANGE  # noqa: E402


def timed(fn, iters=50, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    B, Q, E, T = 1, 900, 256, 8
    layer = RaCFormerTransformerDecoderLayer(E, num_frames=T, pc_range=PC_RANGE).to(dev).eval()
    layer.init_weights()
    mixed = torch.randn(B, Q, E, device=dev)
    qf = torch.randn(B, Q, E, device=dev)
    qb = torch.rand(B, Q, 10, device=dev)
    radar = (torch.randn(B * T, Q, E, device=dev), torch.randn(B, Q, T, device=dev))
    lss = (torch.randn(B * T, Q, E, device=dev), torch.randn(B, Q, T, device=dev))
    res = {}
    with torch.no_grad():
        def torch_tail():
            outs = []
            for branch, norm, (values, logits) in ((layer.sampling_radar_bev, layer.norm_radar_bev, radar),
                                                   (layer.sampling_lss_bev, layer.norm_lss_bev, lss)):
                o = values.permute(1, 2, 0).reshape(Q, E, B, T)
                qw = torch.softmax(logits.permute(1, 0, 2).reshape(Q, 1, B, T), dim=-1)
                o = torch.sum(o * qw, dim=-1)
                outs.append(norm(branch.attention.output_proj(o.permute(2, 0, 1)) + qf))
            x = layer.norm2(mixed)
            x = layer.norm_fusion(layer.fusion(torch.cat((x, outs[0], outs[1]), dim=-1)))
            x = layer.norm3(layer.ffn(x))
            return x, layer.cls_branch(x), layer.reg_branch(x)

        def torch_pos():
            return qf + layer.position_encoder(qb[..., :3])

        ref = torch_tail()
        got = layer._tail_rows(mixed, qf, radar, lss)
        res["tail_max_abs_diff"] = [float((a - b).abs().max()) for a, b in zip(got, ref)]
        res["pos_max_abs_diff"] = float((layer._pos_encode_rows(qb, qf) - torch_pos()).abs().max())
        res["tail_rows_us"] = timed(lambda: layer._tail_rows(mixed, qf, radar, lss))
        res["tail_torch_us"] = timed(torch_tail)
        res["pos_rows_us"] = timed(lambda: layer._pos_encode_rows(qb, qf))
        res["pos_torch_us"] = timed(torch_pos)
        # per-Linear cost: chains of n 256x256 Linear layers in one program
        lin = nn.Linear(E, E).to(dev)
        for rpc in (8, 4):
            for n in (1, 9, 17):
                def chain():
                    p = rowops.RowProgram(B * Q, width=E, num_bufs=2, rows_per_cta=rpc)
                    p.load(0, qf)
                    for i in range(n):
                        p.linear((i + 1) & 1, i & 1, lin)
                    p.store(n & 1, E)
                    p.run()
                res[f"chain_{n}_linears_rows{rpc}_us"] = timed(chain)
        # under a CUDA graph (launch overhead of the Python builder removed)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            layer._tail_rows(mixed, qf, radar, lss)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                out = layer._tail_rows(mixed, qf, radar, lss)
        res["tail_rows_graph_us"] = timed(g.replay)
        # scale-adaptive self-attention: fused path (2 row programs + csrc/sasa.cu) vs the PyTorch chain, both under a graph
        from racformer_b200 import points

        def graph_time(fn):
            g2, s2 = torch.cuda.CUDAGraph(), torch.cuda.Stream()
            with torch.cuda.stream(s2):
                fn()
                torch.cuda.synchronize()
                with torch.cuda.graph(g2, stream=s2):
                    keep = fn()
            return timed(g2.replay), keep

        qkv = torch.randn(B, Q, 3 * E, device=dev)
        tau = torch.rand(B, Q, 8, device=dev)
        res["sasa_core_graph_us"], _ = graph_time(lambda: points.sasa_attention(qkv, tau, qb, PC_RANGE, 8))
        res["self_attn_fused_graph_us"], _ = graph_time(lambda: layer._self_attn_rows(qb, qf))
        res["self_attn_torch_graph_us"], _ = graph_time(
            lambda: layer.norm1(layer.self_attn(qb, qf + layer.position_encoder(qb[..., :3]), None)))
        res["tail_torch_graph_us"], _ = graph_time(torch_tail)
        # rows per CTA (the default is rowops.choose_rows_per_cta: waves x rows)
        auto = rowops.choose_rows_per_cta
        res["tail_rows_per_cta_auto"] = auto(B * Q, 3, 3 * E, dev)
        for rpc in (8, 7, 6, 5):
            rowops.choose_rows_per_cta = lambda *a, _r=rpc, **k: _r
            res[f"tail_rows_graph_rows{rpc}_us"], _ = graph_time(lambda: layer._tail_rows(mixed, qf, radar, lss))
            res[f"self_attn_fused_graph_rows{rpc}_us"], _ = graph_time(lambda: layer._self_attn_rows(qb, qf))
        rowops.choose_rows_per_cta = auto
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
