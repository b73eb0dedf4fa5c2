"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE (needs /root/reference; build container only).

    python tests/golden/make_golden.py

MSMV: imports the reference's own models/csrc/wrapper.py standalone (its CUDA import fails -> MSMV_CUDA False) and
records `msmv_sampling_pytorch` / `msmv_sampling_pytorch_v2` outputs plus autograd gradients on seeded inputs.
MSDA: mmcv-full 1.6.0 is absent (third-party, no network); the fixture is produced with the arithmetic-identical
pure-PyTorch `multi_scale_deformable_attention` from transformers.models.mask2former (the same published
Deformable-DETR fallback the reference calls at models/bev_self_attention.py:202-204).

Features are stored channel-LAST ([B',N,H,W,C], the CUDA op's layout); the reference function is fed the
channel-first permutation. The fixtures are what pins the oracle (tests/test_oracle.py) and, through it and
directly, the CUDA kernels (tests/test_gpu_parity.py).
"""
import contextlib
import importlib.util
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RACFORMER_REFERENCE", "/root/reference")


def load_reference_wrapper():
    pkg = types.ModuleType("refcsrc")
    pkg.__path__ = [os.path.join(REF, "models", "csrc")]
    sys.modules["refcsrc"] = pkg
    spec = importlib.util.spec_from_file_location("refcsrc.wrapper", os.path.join(REF, "models", "csrc", "wrapper.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["refcsrc.wrapper"] = mod
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    assert mod.MSMV_CUDA is False
    return mod


def msmv_case(ref, name, seed, Bp, N, C, Q, P, shapes, lo, hi, specials=True):
    g = torch.Generator().manual_seed(seed)
    L = len(shapes)
    feats = [torch.randn(Bp, N, h, w, C, generator=g) for h, w in shapes]
    xy = torch.rand(Bp, Q, P, 2, generator=g) * (hi - lo) + lo
    if specials:  # boundary coordinates: pixel -1, -0.5, 0, W-1, W-0.5, W on the first level, and far outside
        h0, w0 = shapes[0]
        sx = torch.tensor([-1.0, -0.5, 0.0, w0 - 1.0, w0 - 0.5, float(w0), -3.0, 2.0 * w0]) / max(w0 - 1, 1)
        sy = torch.tensor([-1.0, -0.5, 0.0, h0 - 1.0, h0 - 0.5, float(h0), 0.3 * h0, 0.6 * h0]) / max(h0 - 1, 1)
        k = min(sx.numel(), Q * P)
        flat = xy.view(Bp, Q * P, 2)
        flat[0, :k, 0] = sx[:k]
        flat[0, :k, 1] = sy[:k]
        flat[-1, :k, 0] = sx[:k].flip(0)
        flat[-1, :k, 1] = sy[:k]
    view = torch.randint(0, N, (Bp, Q, P, 1), generator=g).float() / (N - 1)
    loc = torch.cat([xy, view], -1).contiguous()
    w = torch.softmax(torch.randn(Bp, Q, P, L, generator=g), -1).contiguous()
    grad_out = torch.randn(Bp, Q, C, P, generator=g)

    feats_cf = [f.permute(0, 4, 1, 2, 3).contiguous().requires_grad_() for f in feats]
    loc_g, w_g = loc.clone().requires_grad_(), w.clone().requires_grad_()
    out = ref.msmv_sampling_pytorch(feats_cf, loc_g, w_g)
    out.backward(grad_out)
    out_v2 = ref.msmv_sampling_pytorch_v2([f.detach() for f in feats_cf], loc, w)
    data = {"loc": loc, "w": w, "grad_out": grad_out, "out": out.detach(), "out_v2": out_v2,
            "grad_loc": loc_g.grad, "grad_w": w_g.grad, "num_views": torch.tensor(N)}
    for i, (f, fc) in enumerate(zip(feats, feats_cf)):
        data[f"feat{i}"] = f
        data[f"grad_feat{i}"] = fc.grad.permute(0, 2, 3, 4, 1).contiguous()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: v.numpy() for k, v in data.items()})
    print(name, "out", tuple(out.shape), "abs mean", float(out.abs().mean()))


def msda_case(name, seed, B, M, D, Q, P, shapes, lo, hi):
    from transformers.models.mask2former.modeling_mask2former import multi_scale_deformable_attention as hf_msda
    g = torch.Generator().manual_seed(seed)
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    value = torch.randn(B, S, M, D, generator=g).requires_grad_()
    loc = torch.rand(B, Q, M, L, P, 2, generator=g) * (hi - lo) + lo
    # boundary coordinates on level 0: pixel centres/edges -0.5, 0, W-0.5 ... in align_corners=False units
    h0, w0 = shapes[0]
    sx = torch.tensor([-0.5, 0.0, 0.5, w0 - 0.5, w0 - 1.0, w0 + 0.5, -2.0, 0.25 * w0]) / w0
    sy = torch.tensor([-0.5, 0.0, 0.5, h0 - 0.5, h0 - 1.0, 0.5 * h0, 0.5 * h0, h0 + 2.0]) / h0
    k = min(sx.numel(), P)
    loc[0, 0, 0, 0, :k, 0] = sx[:k]
    loc[0, 0, 0, 0, :k, 1] = sy[:k]
    loc = loc.contiguous().requires_grad_()
    aw = torch.softmax(torch.randn(B, Q, M, L * P, generator=g), -1).view(B, Q, M, L, P).contiguous().requires_grad_()
    grad_out = torch.randn(B, Q, M * D, generator=g)
    out = hf_msda(value, [tuple(s) for s in shapes], loc, aw)
    out.backward(grad_out)
    lsi = np.cumsum([0] + [h * w for h, w in shapes[:-1]]).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), value=value.detach().numpy(),
                        spatial_shapes=np.asarray(shapes, dtype=np.int64), level_start_index=lsi,
                        loc=loc.detach().numpy(), aw=aw.detach().numpy(), grad_out=grad_out.numpy(),
                        out=out.detach().numpy(), grad_value=value.grad.numpy(), grad_loc=loc.grad.numpy(),
                        grad_aw=aw.grad.numpy())
    print(name, "out", tuple(out.shape), "abs mean", float(out.abs().mean()))


def main():
    torch.set_num_threads(1)
    ref = load_reference_wrapper()
    # generic-path case (C not 64), all four pyramid levels incl. a 1-row level, boundary coordinates
    msmv_case(ref, "msmv_l4_c8", 1, Bp=2, N=3, C=8, Q=5, P=3, shapes=[(8, 11), (4, 6), (2, 3), (1, 2)], lo=-0.2, hi=1.2)
    # fast-path cases (C = 64): c2345 with the RaCFormer point count, c45 and c23456, ragged P
    msmv_case(ref, "msmv_l4_c64", 2, Bp=2, N=2, C=64, Q=7, P=12, shapes=[(8, 12), (4, 6), (2, 3), (1, 2)], lo=-0.1, hi=1.1)
    msmv_case(ref, "msmv_l2_c64", 3, Bp=1, N=6, C=64, Q=4, P=5, shapes=[(4, 6), (2, 3)], lo=-0.1, hi=1.1)
    msmv_case(ref, "msmv_l5_c64", 4, Bp=2, N=3, C=64, Q=3, P=8, shapes=[(8, 12), (4, 6), (2, 3), (2, 2), (1, 1)], lo=0.0, hi=1.0, specials=False)
    # MSDA: RaCFormer-like single level with 20 points (fast path, D = 64), multi-level, and a generic head_dim
    msda_case("msda_l1_d64", 5, B=2, M=4, D=64, Q=6, P=20, shapes=[(16, 16)], lo=-0.05, hi=1.05)
    msda_case("msda_l3_d64", 6, B=2, M=2, D=64, Q=5, P=4, shapes=[(8, 10), (4, 5), (2, 3)], lo=-0.1, hi=1.1)
    msda_case("msda_l2_d32", 7, B=1, M=8, D=32, Q=4, P=3, shapes=[(6, 7), (3, 4)], lo=-0.1, hi=1.1)


if __name__ == "__main__":
    main()
