"""CPU: the product package must not reach the oracle or any CPU fallback."""
import ast
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "racformer_b200")


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(dirpath, f)).read())
            for node in ast.walk(tree):
                mods = []
                if isinstance(node, ast.Import):
                    mods = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    mods = [node.module or ""]
                for m in mods:
                    assert not m.split(".")[0] == "oracle", f"{f} imports {m}"
    for f in os.listdir(os.path.join(PKG, "csrc")):
        src = open(os.path.join(PKG, "csrc", f)).read()
        assert "racf_oracle_" not in src, f
        for line in src.splitlines():
            if line.lstrip().startswith("#include"):
                assert "oracle" not in line, f"{f}: {line}"


def test_cpu_tensors_are_rejected_not_silently_computed():
    from racformer_b200 import wrapper
    from racformer_b200.multi_scale_deformable_attn_function import ext_module
    feats = [torch.zeros(1, 2, 4, 4, 64), torch.zeros(1, 2, 2, 2, 64)]
    loc = torch.zeros(1, 1, 1, 3)
    w = torch.zeros(1, 1, 1, 2)
    assert wrapper.MSMV_CUDA is True
    with pytest.raises(RuntimeError, match="value must be a CUDA tensor"):
        wrapper.msmv_sampling(feats, loc, w)
    with pytest.raises(RuntimeError, match="value must be a CUDA tensor"):
        ext_module.ms_deform_attn_forward(torch.zeros(1, 16, 1, 64), torch.tensor([[4, 4]]), torch.tensor([0]),
                                          torch.zeros(1, 1, 1, 1, 1, 2), torch.zeros(1, 1, 1, 1, 1), im2col_step=64)


def test_reference_error_messages():
    from racformer_b200 import wrapper
    f = torch.zeros(1, 2, 4, 4, 64)
    with pytest.raises(RuntimeError, match="value tensor has to be contiguous"):
        wrapper.msmv_forward([f.permute(0, 1, 3, 2, 4), f], torch.zeros(1, 1, 1, 3), torch.zeros(1, 1, 1, 2))


def test_missing_library_fails_loudly(monkeypatch):
    from racformer_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "lib_path", lambda: "/nonexistent/libracformer_ops.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()
