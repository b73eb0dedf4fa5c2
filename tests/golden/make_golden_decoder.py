"""Generate tests/golden/decoder_small.npz by RUNNING THE UNCHANGED REFERENCE DECODER on CPU (build container only).

    python tests/golden/make_golden_decoder.py

The reference's models/racformer_transformer.py (+ sparsebev_sampling.py, bev_self_attention.py, bbox/utils.py,
utils.py, csrc/wrapper.py) is imported from /root/reference through tests/reference_shim.py; weights are filled
deterministically from (seed, parameter name) so no checkpoint has to be stored; the fixture keeps inputs + outputs.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from racformer_b200.synthetic import fill_parameters_by_name  # noqa: E402
from tests import reference_shim  # noqa: E402
from tests.decoder_cases import SMALL, small_inputs  # noqa: E402


def main():
    torch.set_num_threads(1)
    rt = reference_shim.load_reference_transformer_module()
    model = rt.RaCFormerTransformer(**SMALL)
    model.init_weights()
    fill_parameters_by_name(model, seed=3)
    model.eval()
    d = small_inputs(seed=5)
    with torch.no_grad():
        cls, box = model(d["query_bbox"].clone(), d["query_feat"], [f.clone() for f in d["mlvl_feats"]], d["lss_bev"],
                         d["radar_bev"], None, [dict(m) for m in d["img_metas"]])
    out = {"cls_scores": cls.numpy(), "bbox_preds": box.numpy(), "query_bbox": d["query_bbox"].numpy(),
           "query_feat": d["query_feat"].numpy(), "lss_bev": d["lss_bev"].numpy(), "radar_bev": d["radar_bev"].numpy()}
    for i, f in enumerate(d["mlvl_feats"]):
        out[f"feat{i}"] = f.numpy()
    np.savez_compressed(os.path.join(HERE, "decoder_small.npz"), **out)
    print("decoder_small: cls", cls.shape, "abs mean", float(cls.abs().mean()), "box abs mean", float(box.abs().mean()),
          "params %.2fM" % (sum(p.numel() for p in model.parameters()) / 1e6))


if __name__ == "__main__":
    main()
