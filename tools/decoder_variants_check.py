"""Run the decoder forward at the f8 shapes in its switchable configurations and compare each with the default one
(row programs + fused self-attention + fused points + tensor-core Linear layers): max |difference| of the outputs."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_workloads  # noqa: E402

dev = torch.device("cuda", 0)
wl = bench_workloads.build("decoder_forward_f8", dev)
wl.use_graph = False
layer = wl.model.decoder.decoder_layer


def run():
    with torch.no_grad():
        cls, box = wl.step()
    torch.cuda.synchronize()
    return cls.clone(), box.clone()


base = run()
res = {}


def compare(name):
    cls, box = run()
    res[name] = {"cls_max_abs_diff": float((cls - base[0]).abs().max()), "box_max_abs_diff": float((box - base[1]).abs().max()),
                 "rows_differing_1e-3": int(((cls - base[0]).abs().amax(-1) > 1e-3).sum())}


layer.row_programs = False
compare("row_programs_off")
layer.row_programs = True
layer.fused_self_attention = False
compare("fused_self_attention_off")
layer.fused_self_attention = True
wl.model.set_fused_points(False)
compare("fused_points_off")
wl.model.set_fused_points(True)
wl.model.set_mixing_precision("fp32")
compare("mixing_fp32_sgemm")
wl.model.set_mixing_precision("bf16x6")
enc = layer.sampling_radar_bev.temporal_encoder
enc.channels_last = False
compare("temporal_encoder_nchw")
enc.channels_last = True
compare("default_again")
print(json.dumps(res, indent=1))
