"""fp32-grade Linear layers on the tcgen05 tensor cores (racformer_b200/csrc/linear.cu).

`y = F.linear(x, weight, bias)` for AdaptiveMixing's parameter_generator and out_proj
(models/racformer_transformer.py:560-566) with every fp32 operand split exactly into three bf16 pieces; all nine (or the
six largest) piece products are accumulated in fp32 in tensor memory. CUDA only -- there is no CPU or PyTorch fallback in
this module. `SplitLinear` is the inference stand-in for an nn.Linear; `TrainableSplitLinear` adds autograd: the input
and weight gradients are two more GEMMs of the same kernel on transposed operand splits.
"""
import ctypes

import torch

from . import _lib

_lib.load()

ALL_TERMS = 4      # max_order: products a_i * w_j with i + j <= max_order; 4 keeps all nine
SIX_TERMS = 2


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def split_bf16x3(x):
    """fp32 CUDA tensor -> bf16 tensor [3, *x.shape] with x == pieces.float().sum(0) exactly."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
        raise RuntimeError("split_bf16x3 needs a contiguous fp32 CUDA tensor")
    if x.numel() % 4 != 0 or x.numel() == 0:
        raise RuntimeError("split_bf16x3: element count must be a positive multiple of 4")
    out = torch.empty((3,) + tuple(x.shape), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().racf_split_bf16x3(x.data_ptr(), x.numel(), out.data_ptr(), _stream(x.device))
    _lib.check(rc, "racf_split_bf16x3")
    return out


class TiledOperand:
    """A [rows, K] fp32 matrix as its three bf16 pieces in the kernel's pre-tiled format (csrc/linear_tiled.cuh): an
    opaque device buffer whose pipeline stages are contiguous 24 KB blocks."""
    __slots__ = ("buf", "rows", "K")

    def __init__(self, buf, rows, K):
        self.buf, self.rows, self.K = buf, int(rows), int(K)

    @property
    def device(self):
        return self.buf.device


def tiled_bytes(rows, K):
    return int(_lib.load().racf_linear_tiled_bytes(int(rows), int(K)))


def empty_tiled(rows, K, device):
    return TiledOperand(torch.empty(tiled_bytes(rows, K) // 2, dtype=torch.bfloat16, device=device), rows, K)


from .caches import cache_epoch, invalidate_weight_caches  # noqa: E402,F401  (re-exported)


def split_tiled(x, addend=None):
    """fp32 CUDA matrix [rows, K] (+ addend [rows_a, K], broadcast with period rows_a over the rows) -> TiledOperand
    (the K tail of the last 32-wide block is zero-filled)."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2 and x.numel() > 0):
        raise RuntimeError("split_tiled needs a contiguous fp32 CUDA matrix")
    if addend is not None and not (addend.is_cuda and addend.dtype == torch.float32 and addend.is_contiguous()
                                   and addend.dim() == 2 and addend.shape[1] == x.shape[1] and addend.device == x.device
                                   and x.shape[0] % addend.shape[0] == 0):
        raise RuntimeError("split_tiled: addend must be a contiguous fp32 CUDA matrix [rows / n, K] on the same device")
    out = empty_tiled(x.shape[0], x.shape[1], x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().racf_split_bf16x3_tiled_add(x.data_ptr(), x.shape[0], x.shape[1],
                                                     addend.data_ptr() if addend is not None else None,
                                                     addend.shape[0] if addend is not None else 1, out.buf.data_ptr(),
                                                     _stream(x.device))
    _lib.check(rc, "racf_split_bf16x3_tiled_add")
    return out


def untile(op):
    """TiledOperand -> plain pieces [3, rows, K] bf16 (index arithmetic of csrc/linear_tiled.cuh in torch; used by tests)."""
    dev = op.buf.device
    kblocks = (op.K + 31) // 32
    row = torch.arange(op.rows, device=dev)[:, None]
    k = torch.arange(op.K, device=dev)[None, :]
    rr, kk = row & 127, k & 31
    chunk = (kk >> 3) ^ ((rr >> 1) & 3)
    base = ((row >> 7) * kblocks + (k >> 5)) * 3
    inner = rr * 32 + chunk * 8 + (kk & 7)
    return torch.stack([op.buf[(base + p) * 4096 + inner] for p in range(3)])


def split_bf16x3_chw_to_hwc(x, pos=None, tiled=False):
    """x [BT, C, S] fp32 (+ pos [C, S], broadcast over BT) -> bf16 pieces [3, BT * S, C] of the channel-last sum (or the
    same matrix as a TiledOperand): the add, permute and copy in front of BEVSelfAttention.value_proj fused with the
    operand split."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 3):
        raise RuntimeError("split_bf16x3_chw_to_hwc needs a contiguous fp32 CUDA tensor [BT, C, S]")
    BT, C, S = x.shape
    if pos is not None and not (pos.is_cuda and pos.dtype == torch.float32 and pos.is_contiguous()
                                and pos.numel() == C * S and pos.device == x.device):
        raise RuntimeError("split_bf16x3_chw_to_hwc: pos must be a contiguous fp32 CUDA tensor of C * S elements")
    if C % 8 != 0:
        raise RuntimeError("split_bf16x3_chw_to_hwc: channels must be a multiple of 8")
    out = empty_tiled(BT * S, C, x.device) if tiled else torch.empty((3, BT * S, C), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().racf_split_bf16x3_chw_to_hwc(x.data_ptr(), pos.data_ptr() if pos is not None else None, BT, C, S,
                                                      1 if tiled else 0, (out.buf if tiled else out).data_ptr(),
                                                      _stream(x.device))
    _lib.check(rc, "racf_split_bf16x3_chw_to_hwc")
    return out


def plan(M, N, K):
    """-> (split_k, workspace_bytes) the library wants for this problem."""
    s, w = ctypes.c_int(0), ctypes.c_longlong(0)
    _lib.check(_lib.load().racf_linear_bf16x3_plan(M, N, K, ctypes.byref(s), ctypes.byref(w)), "racf_linear_bf16x3_plan")
    return s.value, w.value


WIDE_TILES = True     # TiledOperands: 128 x 256 tiles on persistent CTAs (csrc/linear_wide.cu) instead of 128 x 128 tiles


def linear_bf16x3(a3, w3, bias=None, max_order=ALL_TERMS, split_k=None, variant=0):
    """a3 [3, M, K], w3 [3, N, K] (bf16 pieces from split_bf16x3) or two TiledOperands, bias [N] fp32 or None
    -> [M, N] fp32. TiledOperands run as variant 3 (wide tiles; bit-identical to variant 2) when N > 128 unless the module
    switch WIDE_TILES is off; `variant` only selects between the two plain-piece kernels (0 / 1)."""
    if isinstance(a3, TiledOperand) or isinstance(w3, TiledOperand):
        if not (isinstance(a3, TiledOperand) and isinstance(w3, TiledOperand)) or a3.K != w3.K or a3.device != w3.device:
            raise RuntimeError("linear_bf16x3: both operands must be TiledOperands with the same K on one device")
        M, K, N = a3.rows, a3.K, w3.rows
        variant = 3 if (WIDE_TILES and N > 128) else 2
        a_ptr, w_ptr, dev = a3.buf.data_ptr(), w3.buf.data_ptr(), a3.device
    else:
        if not (a3.is_cuda and w3.is_cuda and a3.dtype == torch.bfloat16 and w3.dtype == torch.bfloat16
                and a3.is_contiguous() and w3.is_contiguous() and a3.device == w3.device):
            raise RuntimeError("linear_bf16x3 needs contiguous bf16 CUDA piece tensors on one device")
        if a3.dim() != 3 or w3.dim() != 3 or a3.shape[0] != 3 or w3.shape[0] != 3 or a3.shape[2] != w3.shape[2]:
            raise RuntimeError("linear_bf16x3: a3 must be [3, M, K] and w3 [3, N, K]")
        M, K = a3.shape[1], a3.shape[2]
        N = w3.shape[1]
        if K % 8 != 0:
            raise RuntimeError("linear_bf16x3: K must be a multiple of 8")
        if variant not in (0, 1):
            raise RuntimeError("linear_bf16x3: plain pieces run as variant 0 or 1")
        a_ptr, w_ptr, dev = a3.data_ptr(), w3.data_ptr(), a3.device
    if bias is not None and not (bias.is_cuda and bias.dtype == torch.float32 and bias.is_contiguous()
                                 and bias.numel() == N and bias.device == dev):
        raise RuntimeError("linear_bf16x3: bias must be a contiguous fp32 CUDA tensor of N elements")
    if split_k is None:
        split_k, _ = plan(M, N, K)
    out = torch.empty((M, N), dtype=torch.float32, device=dev)
    ws = torch.empty((split_k, M, N), dtype=torch.float32, device=dev) if split_k > 1 else None
    with torch.cuda.device(dev):
        rc = _lib.load().racf_linear_bf16x3_forward(
            a_ptr, w_ptr, bias.data_ptr() if bias is not None else None, M, N, K, int(max_order),
            int(split_k), int(variant), ws.data_ptr() if ws is not None else None, out.data_ptr(), _stream(dev))
    _lib.check(rc, "racf_linear_bf16x3_forward")
    return out


class SplitLinear:
    """Inference-time stand-in for an nn.Linear: caches the bf16 pieces of the weight (re-split when the parameter
    changes) and runs x @ W^T + b through racf_linear_bf16x3_forward."""

    def __init__(self, linear, max_order=ALL_TERMS, variant=2):
        """variant 2 (default): operands in the pre-tiled format (bulk copies); 0 / 1: plain pieces through tensor maps."""
        self.linear, self.max_order, self.variant = linear, max_order, variant
        self._key, self._w3 = None, None

    def _split(self, x2d):
        return split_tiled(x2d) if self.variant == 2 else split_bf16x3(x2d)

    def weight_pieces(self):
        w = self.linear.weight
        key = (w.data_ptr(), w._version, w.device, cache_epoch())
        if self._key != key:
            self._w3 = self._split(w.detach().contiguous())
            self._key = key
        return self._w3

    def __call__(self, x=None, x3=None, lead=None):
        """x [..., K] fp32, or its pieces x3 (TiledOperand, or [3, rows, K] for variants 0 / 1) with the leading shape
        `lead` of the result -> [..., N] fp32."""
        if x3 is None:
            lead = x.shape[:-1]
            x3 = self._split(x.reshape(-1, x.shape[-1]).contiguous())
        elif lead is None:
            lead = (x3.rows if isinstance(x3, TiledOperand) else x3.shape[1],)
        bias = self.linear.bias.detach() if self.linear.bias is not None else None
        y = linear_bf16x3(x3, self.weight_pieces(), bias, self.max_order, variant=self.variant)
        return y.reshape(*lead, -1)


def split_tiled_transposed(x):
    """fp32 CUDA matrix [R, C] -> TiledOperand of its transpose [C rows, K = R] (R % 8 == 0): the transposing operand
    split (racf_split_bf16x3_chw_to_hwc with one batch element), used for the backward GEMMs."""
    if x.dim() != 2:
        raise RuntimeError("split_tiled_transposed needs a matrix")
    return split_bf16x3_chw_to_hwc(x.contiguous().unsqueeze(0), None, tiled=True)


class _SplitLinearFunction(torch.autograd.Function):
    """y = x @ W^T + b on the tcgen05 Linear kernel with autograd (fp32-grade in all three products):
        grad_x [M,K] = g [M,N] . (W^T) [K,N]^T      grad_W [N,K] = g^T [N,M] . (x^T) [K,M]^T      grad_b = sum_m g
    Replaces F.linear for AdaptiveMixing's parameter_generator / out_proj while training
    (models/racformer_transformer.py:560-566, 592, 606); the reference runs three cuBLAS SGEMMs per layer call."""

    @staticmethod
    def forward(ctx, x, weight, bias, owner):
        lead, K = x.shape[:-1], x.shape[-1]
        x2 = x.reshape(-1, K).contiguous()
        y = linear_bf16x3(split_tiled(x2), owner.weight_pieces(), bias.detach() if bias is not None else None,
                          owner.max_order, variant=2)
        ctx.save_for_backward(x2, weight)
        ctx.owner, ctx.has_bias, ctx.lead = owner, bias is not None, lead
        return y.reshape(*lead, weight.shape[0])

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_y):
        x2, weight = ctx.saved_tensors
        owner = ctx.owner
        g2 = grad_y.reshape(-1, weight.shape[0]).contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = linear_bf16x3(split_tiled(g2), owner.weight_t_pieces(), None, owner.max_order, variant=2)
            gx = gx.reshape(*ctx.lead, weight.shape[1])
        if ctx.needs_input_grad[1]:
            gw = linear_bf16x3(split_tiled_transposed(g2), split_tiled_transposed(x2), None, owner.max_order, variant=2)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g2.sum(0)
        return gx, gw, gb, None


class TrainableSplitLinear:
    """nn.Linear stand-in with autograd on the tcgen05 kernel. Caches the tiled pieces of W (forward) and of W^T (input
    gradient); both are re-split when the parameter changes (optimizer step). The row count of the input must be a
    multiple of 8 when a weight gradient is needed (the transposing split works on 16-byte chunks)."""

    def __init__(self, linear, max_order=SIX_TERMS):
        self.linear, self.max_order = linear, max_order
        self._key = self._w3 = self._keyt = self._wt3 = None

    def _version_key(self):
        w = self.linear.weight
        return (w.data_ptr(), w._version, w.device, cache_epoch())

    def weight_pieces(self):
        key = self._version_key()
        if self._key != key:
            self._w3, self._key = split_tiled(self.linear.weight.detach().contiguous()), key
        return self._w3

    def weight_t_pieces(self):
        key = self._version_key()
        if self._keyt != key:
            self._wt3, self._keyt = split_tiled_transposed(self.linear.weight.detach()), key
        return self._wt3

    def supports(self, x):
        rows = x.numel() // x.shape[-1]
        return (x.is_cuda and x.dtype == torch.float32 and rows % 8 == 0 and self.linear.out_features % 8 == 0
                and self.linear.in_features % 8 == 0)

    def __call__(self, x):
        return _SplitLinearFunction.apply(x, self.linear.weight, self.linear.bias, self)


class MultiSplitLinear:
    """Several nn.Linear layers applied to the SAME input in one tcgen05 launch (racf_linear_bf16x3_multi_forward):
    their weights are stacked along N (each padded with zero rows to a multiple of 128) and split once; every layer gets
    its own dense output. Inference only."""

    MAX_SEGMENTS = 16

    def __init__(self, linears, max_order=SIX_TERMS):
        self.linears, self.max_order = list(linears), max_order
        if not 0 < len(self.linears) <= self.MAX_SEGMENTS:
            raise RuntimeError("MultiSplitLinear: 1..16 layers")
        k = {lin.in_features for lin in self.linears}
        if len(k) != 1 or next(iter(k)) > 512:
            raise RuntimeError("MultiSplitLinear: the layers must share in_features (at most 512)")
        self._key, self._w3 = None, None

    def weight_pieces(self):
        key = tuple((lin.weight.data_ptr(), lin.weight._version) for lin in self.linears) + (cache_epoch(),)
        if self._key != key:
            rows = []
            for lin in self.linears:
                w = lin.weight.detach()
                pad = (-w.shape[0]) % 128
                rows.append(torch.cat([w, w.new_zeros(pad, w.shape[1])]) if pad else w)
            self._w3 = split_tiled(torch.cat(rows).contiguous())
            self._key = key
        return self._w3

    def __call__(self, x=None, x3=None):
        """x [..., K] (or its TiledOperand x3) -> list of [rows, N_i] fp32 tensors, one per layer."""
        if x3 is None:
            x3 = split_tiled(x.reshape(-1, x.shape[-1]).contiguous())
        w3 = self.weight_pieces()
        M, K = x3.rows, x3.K
        n = len(self.linears)
        outs = [torch.empty((M, lin.out_features), dtype=torch.float32, device=x3.device) for lin in self.linears]
        seg_n = (ctypes.c_int * n)(*[lin.out_features for lin in self.linears])
        seg_bias = (ctypes.c_void_p * n)(*[lin.bias.data_ptr() if lin.bias is not None else None for lin in self.linears])
        seg_out = (ctypes.c_void_p * n)(*[o.data_ptr() for o in outs])
        with torch.cuda.device(x3.device):
            rc = _lib.load().racf_linear_bf16x3_multi_forward(x3.buf.data_ptr(), w3.buf.data_ptr(), M, K, n, seg_n, seg_bias,
                                                              seg_out, int(self.max_order), 1, _stream(x3.device))
        _lib.check(rc, "racf_linear_bf16x3_multi_forward")
        return outs
