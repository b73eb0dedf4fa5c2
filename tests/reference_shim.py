"""TEST INFRASTRUCTURE (build container only): import the UNCHANGED reference decoder call-site files from
/root/reference although mmcv / mmdet / mmdet3d are not installed.

Minimal stand-ins for the third-party symbols the files import (list verified in SURVEY.md 7.3-3) are placed in
`sys.modules`, and dummy parent packages make `models/__init__.py` / `models/bbox/__init__.py` (which import the
whole mm* stack) be skipped. Nothing here is used by the product package; GPU-box tests never import this module
because /root/reference does not exist there.
"""
import contextlib
import importlib
import io
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("RACFORMER_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "models"))


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], child, m)
    return m


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        for child in self.children():
            if hasattr(child, "init_weights"):
                child.init_weights()


def _decorator_factory(*_a, **_k):
    return lambda fn: fn


def xavier_init(module, gain=1, bias=0, distribution="normal"):
    (nn.init.xavier_uniform_ if distribution == "uniform" else nn.init.xavier_normal_)(module.weight, gain=gain)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)


def bias_init_with_prob(p):
    return float(-np.log((1 - p) / p))


class MultiheadAttention(BaseModule):
    def __init__(self, embed_dims, num_heads, attn_drop=0.0, proj_drop=0.0, dropout_layer=None, init_cfg=None,
                 batch_first=False, **kwargs):
        super().__init__(init_cfg)
        self.batch_first = batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = nn.Identity()

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None, attn_mask=None,
                key_padding_mask=None, **kwargs):
        key = query if key is None else key
        value = key if value is None else value
        identity = query if identity is None else identity
        if self.batch_first:
            query, key, value = (t.transpose(0, 1) for t in (query, key, value))
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask, key_padding_mask=key_padding_mask)[0]
        if self.batch_first:
            out = out.transpose(0, 1)
        return identity + self.dropout_layer(self.proj_drop(out))


class FFN(BaseModule):
    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, ffn_drop=0.0, add_identity=True, **kwargs):
        super().__init__()
        assert num_fcs == 2
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(ffn_drop)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop))
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return out
        return (x if identity is None else identity) + out


class LearnedPositionalEncoding(BaseModule):
    def __init__(self, num_feats, row_num_embed=50, col_num_embed=50, **kwargs):
        super().__init__()
        self.row_embed = nn.Embedding(row_num_embed, num_feats)
        self.col_embed = nn.Embedding(col_num_embed, num_feats)

    def forward(self, mask):
        h, w = mask.shape[-2:]
        x_embed = self.col_embed(torch.arange(w, device=mask.device))
        y_embed = self.row_embed(torch.arange(h, device=mask.device))
        pos = torch.cat((x_embed.unsqueeze(0).repeat(h, 1, 1), y_embed.unsqueeze(1).repeat(1, w, 1)), dim=-1)
        return pos.permute(2, 0, 1).unsqueeze(0).repeat(mask.shape[0], 1, 1, 1)


def build_positional_encoding(cfg):
    cfg = dict(cfg)
    assert cfg.pop("type") == "LearnedPositionalEncoding"
    return LearnedPositionalEncoding(**cfg)


class _Registry:
    def register_module(self, *_a, **_k):
        return lambda cls: cls


class _NoExt:
    def __getattr__(self, name):
        raise RuntimeError("mmcv _ext is not available: " + name)


def _msda_pytorch(value, shapes, loc, aw):
    from transformers.models.mask2former.modeling_mask2former import multi_scale_deformable_attention
    return multi_scale_deformable_attention(value, [tuple(int(v) for v in s) for s in shapes], loc, aw)


_loaded = None


def load_reference_transformer_module():
    """Returns the reference's `models.racformer_transformer` module (MSMV_CUDA is False: PyTorch fallbacks)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    assert available(), "reference tree not found"
    _module("mmcv")
    _module("mmcv.runner", BaseModule=BaseModule, auto_fp16=_decorator_factory, force_fp32=_decorator_factory)
    _module("mmcv.runner.base_module", BaseModule=BaseModule)
    _module("mmcv.cnn", xavier_init=xavier_init, bias_init_with_prob=bias_init_with_prob)
    _module("mmcv.cnn.bricks")
    _module("mmcv.cnn.bricks.transformer", MultiheadAttention=MultiheadAttention, FFN=FFN,
            build_positional_encoding=build_positional_encoding)
    _module("mmcv.utils", ext_loader=types.SimpleNamespace(load_ext=lambda *a, **k: _NoExt()))
    _module("mmcv.ops")
    _module("mmcv.ops.multi_scale_deform_attn", multi_scale_deformable_attn_pytorch=_msda_pytorch)
    _module("mmdet")
    _module("mmdet.models")
    _module("mmdet.models.utils")
    _module("mmdet.models.utils.builder", TRANSFORMER=_Registry())
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REF, "models")]
    sys.modules["models"] = pkg
    for sub in ("bbox", "csrc"):
        sp = types.ModuleType("models." + sub)
        sp.__path__ = [os.path.join(REF, "models", sub)]
        sys.modules["models." + sub] = sp
        setattr(pkg, sub, sp)
    with contextlib.redirect_stdout(io.StringIO()):
        _loaded = importlib.import_module("models.racformer_transformer")
    assert _loaded.MSMV_CUDA is False
    return _loaded
