"""Row programs with autograd: the row-wise operator chains of a decoder iteration for TRAINING (csrc/rowops.cu).

`racformer_b200/rowops.py` runs such a chain (models/racformer_transformer.py:204-262, models/bev_self_attention.py:206-225)
as one launch for inference. Here the same chain is recorded once and executed as

  forward  : the inference program plus STORE operators that save what the backward needs (every Linear's input, the
             output of a Linear with a fused ReLU, every LayerNorm's input) -- one launch;
  backward : ANOTHER row program, generated from the recorded chain by reverse-mode rules over the shared-memory buffers
             (gradient buffers G[b] mirror the forward buffers; walking the chain backwards, STORE -> G += grad_out,
             ADD -> G[src] += G[dst], LayerNorm -> LAYERNORM_BWD in place, Linear -> [RELU_MASK,] STORE_COLSUM of the
             output gradient (+ bias gradient), G[src] += G[dst] W (a Linear with the un-transposed weight), ZERO G[dst];
             LOAD -> STORE of the input gradient; LOAD_QUEUE -> QUEUE_BWD) -- one launch;
  weights  : one tcgen05 GEMM per Linear, grad_W = (saved output gradients)^T (saved inputs) (racformer_b200/linear.py).

The reference differentiates these chains with autograd over ~25 PyTorch modules per iteration (cuBLAS SGEMMs, layer-norm,
bias and ReLU kernels: ~1000 launches per training step). CUDA only, fp32; no CPU fallback in this module.
"""
import ctypes

import torch

from . import _lib, linear as tc_linear, rowops
from .caches import cache_epoch as _cache_epoch
from .rowops import (ACCUM, ADD, DROPOUT, LAYERNORM, LAYERNORM_BWD, LINEAR, LINEAR_NARROW, LOAD, LOAD_QUEUE, QUEUE_BWD, RELU,
                     RELU_MASK, STORE, STORE_COLSUM, ZERO, RowOp, chunked_transpose)


def _pad8(n):
    return (n + 7) // 8 * 8


class RowChain:
    """Recorder with the operator vocabulary of rowops.RowProgram. Tensors that take part in autograd (inputs, queue
    values / logits, parameters) are registered in `self.inputs`; `run()` executes the chain through RowChainFunction."""

    def __init__(self, rows, width, num_bufs=3, device=None):
        self.rows, self.width, self.num_bufs, self.device = int(rows), int(width), int(num_bufs), device
        self.ir, self.inputs, self._index, self.num_outputs = [], [], {}, 0

    def _reg(self, t):
        """Index of tensor t among the autograd inputs of the chain (registered once per tensor object)."""
        if t is None:
            return None
        key = id(t)
        if key not in self._index:
            self._index[key] = len(self.inputs)
            self.inputs.append(t)
        return self._index[key]

    def load(self, dst, t, n=None, dst_col=0):
        n = t.shape[-1] if n is None else n
        self.ir.append(dict(op="load", dst=dst, t=self._reg(t), n=n, dst_col=dst_col))

    def load_queue(self, dst, values, logits, rows_per_batch, queue, dst_col=0):
        self.ir.append(dict(op="load_queue", dst=dst, values=self._reg(values), logits=self._reg(logits), k=rows_per_batch,
                            queue=queue, dst_col=dst_col, n=values.shape[2]))

    def linear(self, dst, src, lin, relu=False, dst_col=0, src_col=0, weight=None, bias=None):
        weight = lin.weight if weight is None else weight
        bias = (lin.bias if hasattr(lin, "bias") else None) if bias is None else bias
        self.ir.append(dict(op="linear", dst=dst, src=src, owner=lin, w=self._reg(weight), b=self._reg(bias), relu=relu,
                            dst_col=dst_col, src_col=src_col, n=weight.shape[0], k=weight.shape[1]))

    def layernorm(self, buf, ln, col=0, relu=False):
        self.ir.append(dict(op="layernorm", buf=buf, g=self._reg(ln.weight), b=self._reg(ln.bias), col=col, relu=relu,
                            n=ln.normalized_shape[-1], eps=float(ln.eps)))

    def add(self, dst, src, n, dst_col=0, src_col=0):
        self.ir.append(dict(op="add", dst=dst, src=src, n=n, dst_col=dst_col, src_col=src_col))

    def dropout(self, buf, n, p, seed, col=0):
        """In-place inverted dropout with drop probability p (no-op when p == 0): counter-based mask from (seed, row, column)."""
        if p > 0.0:
            self.ir.append(dict(op="dropout", buf=buf, n=n, col=col, p=float(p), seed=int(seed) & 0x7fffffff))

    def store(self, src, n, src_col=0):
        """-> index of the output (RowChain.run returns the outputs in this order)."""
        self.ir.append(dict(op="store", src=src, n=n, src_col=src_col, out=self.num_outputs))
        self.num_outputs += 1
        return self.num_outputs - 1

    def run(self):
        outs = RowChainFunction.apply(self, *self.inputs)
        return outs if isinstance(outs, tuple) else (outs,)


def _weight_fwd(e, w):
    """Forward weight operand of a Linear record: chunked transpose (wide) or the weight itself (narrow)."""
    N, K = w.shape
    narrow = N < 32 and N * K <= 8192 and (N * K) % 4 == 0
    if narrow:
        return w.detach().contiguous(), True
    return rowops.weight_t(e["owner"], w), False


_dgrad_cache = {}


def _weight_bwd(e, w):
    """Operand of the input-gradient Linear gx = gy @ W: the chunked transpose of W^T, i.e. W's rows in 256-column chunks."""
    key = (id(e["owner"]), w.data_ptr(), w._version, _cache_epoch())
    hit = _dgrad_cache.get(id(e["owner"]))
    if hit is None or hit[0] != key:
        hit = _dgrad_cache[id(e["owner"])] = (key, chunked_transpose(w.detach().t()))
    return hit[1]


def _launch(ops, rows, num_bufs, width, device):
    if len(ops) > rowops.MAX_OPS:
        raise _lib.Unsupported(f"row program of {len(ops)} operators (limit {rowops.MAX_OPS})")
    rows_per_cta = rowops.choose_rows_per_cta(rows, num_bufs, width, device)
    arr = (RowOp * len(ops))(*ops)
    with torch.cuda.device(device):
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        rc = _lib.load().racf_row_program_forward(arr, len(ops), rows, rows_per_cta, num_bufs, width, stream)
    _lib.check(rc, "racf_row_program_forward")


class RowChainFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, chain, *tensors):
        dev = next(t.device for t in tensors if t is not None)
        rows = chain.rows
        ops, keep, tape, outs = [], [], {}, [None] * chain.num_outputs
        f32 = dict(dtype=torch.float32, device=dev)

        def mat(t, n):
            t2 = t.detach()
            t2 = t2.reshape(-1, t2.shape[-1]) if t2.dim() != 2 else t2
            if t2.shape[0] != rows or t2.shape[1] < n or t2.dtype != torch.float32 or not t2.is_cuda:
                raise RuntimeError(f"row chain: expected an fp32 CUDA [{rows}, >= {n}] matrix, got {tuple(t2.shape)}")
            if t2.stride(1) != 1:
                t2 = t2.contiguous()
            keep.append(t2)
            return t2

        def param(i):
            if i is None:
                return None
            p = tensors[i].detach()
            if not p.is_contiguous():
                p = p.contiguous()
            keep.append(p)
            return p.data_ptr()

        version = [0] * chain.num_bufs      # bumped whenever an operator writes a buffer
        saved_x = {}                         # (buffer, column, n, version) -> tensor: several Linears reading the same rows
                                             # (the eleven sampling heads) share ONE saved input -- and one weight-gradient GEMM

        def save(idx, what, buf, col, n):
            key = (buf, col, n, version[buf])
            if what == "x" and key in saved_x:
                tape[(idx, what)] = saved_x[key]
                return
            t = torch.empty((rows, n), **f32)
            tape[(idx, what)] = t
            if what == "x":
                saved_x[key] = t
            ops.append(RowOp(kind=STORE, src=buf, src_col=col, n=n, ld=n, out=t.data_ptr()))

        for idx, e in enumerate(chain.ir):
            k = e["op"]
            written = e.get("dst", e.get("buf"))
            if k == "load":
                t2 = mat(tensors[e["t"]], e["n"])
                ops.append(RowOp(kind=LOAD, dst=e["dst"], dst_col=e["dst_col"], n=e["n"], ld=t2.stride(0), p0=t2.data_ptr()))
            elif k == "load_queue":
                v = tensors[e["values"]].detach().contiguous()
                keep.append(v)
                lg = None
                if e["logits"] is not None:
                    lg = mat(tensors[e["logits"]], e["queue"])
                    if lg.stride(0) != e["queue"]:
                        lg = lg.contiguous()
                        keep.append(lg)
                ops.append(RowOp(kind=LOAD_QUEUE, dst=e["dst"], dst_col=e["dst_col"], n=e["n"], k=e["k"], aux=e["queue"],
                                 ld=v.shape[2], p0=v.data_ptr(), p1=lg.data_ptr() if lg is not None else None))
            elif k == "linear":
                save(idx, "x", e["src"], e["src_col"], e["k"])
                w, narrow = _weight_fwd(e, tensors[e["w"]])
                keep.append(w)
                ops.append(RowOp(kind=LINEAR_NARROW if narrow else LINEAR, dst=e["dst"], dst_col=e["dst_col"], src=e["src"],
                                 src_col=e["src_col"], n=e["n"], k=e["k"], flags=RELU if e["relu"] else 0, p0=w.data_ptr(),
                                 p1=param(e["b"])))
                if e["relu"]:
                    save(idx, "y", e["dst"], e["dst_col"], e["n"])
            elif k == "layernorm":
                save(idx, "x", e["buf"], e["col"], e["n"])
                ops.append(RowOp(kind=LAYERNORM, dst=e["buf"], dst_col=e["col"], n=e["n"], eps=e["eps"],
                                 flags=RELU if e["relu"] else 0, p0=param(e["g"]), p1=param(e["b"])))
            elif k == "add":
                ops.append(RowOp(kind=ADD, dst=e["dst"], dst_col=e["dst_col"], src=e["src"], src_col=e["src_col"], n=e["n"]))
            elif k == "dropout":
                ops.append(RowOp(kind=DROPOUT, dst=e["buf"], dst_col=e["col"], n=e["n"], k=e["col"], aux=e["seed"], eps=e["p"]))
            elif k == "store":
                o = torch.empty((rows, e["n"]), **f32)
                outs[e["out"]] = o
                ops.append(RowOp(kind=STORE, src=e["src"], src_col=e["src_col"], n=e["n"], ld=e["n"], out=o.data_ptr()))
            if k != "store" and written is not None:
                version[written] += 1
        _launch(ops, rows, chain.num_bufs, chain.width, dev)
        ctx.chain, ctx.tape, ctx.dev = chain, tape, dev
        ctx.save_for_backward(*[t for t in tensors if t is not None])
        ctx.present = [t is not None for t in tensors]
        return tuple(outs)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *grad_outs):
        chain, tape, dev = ctx.chain, ctx.tape, ctx.dev
        saved = iter(ctx.saved_tensors)
        tensors = [next(saved) if p else None for p in ctx.present]
        rows, f32 = chain.rows, dict(dtype=torch.float32, device=dev)
        grads = [None] * len(tensors)
        needs = list(ctx.needs_input_grad[1:])                     # needs_input_grad[0] is the chain object
        ops, keep, loaded = [], [], []

        def acc(i, shape):
            """Zero-initialised fp32 accumulator for input i (atomics / several contributions)."""
            if grads[i] is None:
                grads[i] = torch.zeros(shape, **f32)
            return grads[i]

        # does buffer b carry a gradient that anyone needs, at each point of the chain (conservative: never reset)
        carries, flags = [], [False] * chain.num_bufs
        for e in chain.ir:
            carries.append(list(flags))
            k = e["op"]
            if k == "load":
                flags[e["dst"]] |= bool(needs[e["t"]])
            elif k == "load_queue":
                flags[e["dst"]] |= bool(needs[e["values"]]) or (e["logits"] is not None and bool(needs[e["logits"]]))
            elif k == "linear":
                flags[e["dst"]] = True
            elif k == "add":
                flags[e["dst"]] |= flags[e["src"]]
        # Linears that read the same saved input share one weight-gradient GEMM: their output gradients are stored side by
        # side in one [rows, sum of padded N] matrix (grad_W of all of them = that matrix^T @ x, sliced by rows afterwards)
        stacks, slot = {}, {}
        for idx, e in enumerate(chain.ir):
            if e["op"] == "linear" and needs[e["w"]]:
                st = stacks.setdefault(id(tape[(idx, "x")]), {"members": [], "width": 0, "x": tape[(idx, "x")]})
                slot[idx] = (st, st["width"])
                st["members"].append((e, st["width"]))
                st["width"] += _pad8(e["n"])
        for st in stacks.values():
            st["gy"] = torch.zeros((rows, st["width"]), **f32)                      # pad columns stay 0 for the GEMM
        for b in range(chain.num_bufs):
            ops.append(RowOp(kind=ZERO, dst=b, dst_col=0, n=chain.width))
        for idx in range(len(chain.ir) - 1, -1, -1):
            e = chain.ir[idx]
            k = e["op"]
            if k == "store":
                g = grad_outs[e["out"]]
                if g is not None:
                    g = g.reshape(rows, e["n"]).contiguous()
                    keep.append(g)
                    ops.append(RowOp(kind=LOAD, flags=ACCUM, dst=e["src"], dst_col=e["src_col"], n=e["n"], ld=e["n"], p0=g.data_ptr()))
            elif k == "add":
                if carries[idx][e["src"]]:
                    ops.append(RowOp(kind=ADD, dst=e["src"], dst_col=e["src_col"], src=e["dst"], src_col=e["dst_col"], n=e["n"]))
            elif k == "dropout":
                ops.append(RowOp(kind=DROPOUT, dst=e["buf"], dst_col=e["col"], n=e["n"], k=e["col"], aux=e["seed"], eps=e["p"]))
            elif k == "layernorm":
                x = tape[(idx, "x")]
                gam = tensors[e["g"]].detach() if e["g"] is not None else None
                bet = tensors[e["b"]].detach() if e["b"] is not None else None
                ops.append(RowOp(kind=LAYERNORM_BWD, dst=e["buf"], dst_col=e["col"], n=e["n"], eps=e["eps"], ld=e["n"],
                                 flags=RELU if e["relu"] else 0, p0=gam.data_ptr() if gam is not None else None,
                                 p1=bet.data_ptr() if bet is not None else None, p2=x.data_ptr(),
                                 out=acc(e["g"], gam.shape).data_ptr() if e["g"] is not None and needs[e["g"]] else None,
                                 out2=acc(e["b"], bet.shape).data_ptr() if e["b"] is not None and needs[e["b"]] else None))
            elif k == "linear":
                N, K = e["n"], e["k"]
                if e["relu"]:
                    y = tape[(idx, "y")]
                    ops.append(RowOp(kind=RELU_MASK, dst=e["dst"], dst_col=e["dst_col"], n=N, ld=N, p0=y.data_ptr()))
                need_w = needs[e["w"]]
                gy_ptr, gy_ld = None, _pad8(N)
                if need_w:
                    st, off = slot[idx]
                    gy_ptr, gy_ld = st["gy"].data_ptr() + 4 * off, st["width"]
                gb = acc(e["b"], tensors[e["b"]].shape) if e["b"] is not None and needs[e["b"]] else None
                if gy_ptr is not None or gb is not None:
                    ops.append(RowOp(kind=STORE_COLSUM, src=e["dst"], src_col=e["dst_col"], n=N, ld=gy_ld,
                                     out=gy_ptr, out2=gb.data_ptr() if gb is not None else None))
                if carries[idx][e["src"]]:
                    wt = _weight_bwd(e, tensors[e["w"]])
                    keep.append(wt)
                    ops.append(RowOp(kind=LINEAR, flags=ACCUM, dst=e["src"], dst_col=e["src_col"], src=e["dst"],
                                     src_col=e["dst_col"], n=K, k=N, p0=wt.data_ptr()))
                ops.append(RowOp(kind=ZERO, dst=e["dst"], dst_col=e["dst_col"], n=N))
            elif k == "load":
                if needs[e["t"]]:
                    t = tensors[e["t"]]
                    g = torch.zeros((rows, t.shape[-1]), **f32) if t.shape[-1] != e["n"] else torch.empty((rows, e["n"]), **f32)
                    ops.append(RowOp(kind=STORE, src=e["dst"], src_col=e["dst_col"], n=e["n"], ld=g.shape[1], out=g.data_ptr()))
                    loaded.append((e["t"], g.reshape(t.shape)))        # summed after the launch (a tensor may be loaded twice)
                ops.append(RowOp(kind=ZERO, dst=e["dst"], dst_col=e["dst_col"], n=e["n"]))
            elif k == "load_queue":
                v = tensors[e["values"]].detach().contiguous()
                keep.append(v)
                lg = tensors[e["logits"]].detach().reshape(rows, e["queue"]).contiguous() if e["logits"] is not None else None
                keep.append(lg)
                gv = torch.empty_like(v) if needs[e["values"]] else None
                gl = torch.empty((rows, e["queue"]), **f32) if lg is not None and needs[e["logits"]] else None
                if gv is not None or gl is not None:
                    ops.append(RowOp(kind=QUEUE_BWD, src=e["dst"], src_col=e["dst_col"], n=e["n"], k=e["k"], aux=e["queue"],
                                     ld=v.shape[2], p0=v.data_ptr(), p1=lg.data_ptr() if lg is not None else None,
                                     out=gv.data_ptr() if gv is not None else None, out2=gl.data_ptr() if gl is not None else None))
                if gv is not None:
                    grads[e["values"]] = gv
                if gl is not None:
                    grads[e["logits"]] = gl.reshape(tensors[e["logits"]].shape)
                ops.append(RowOp(kind=ZERO, dst=e["dst"], dst_col=e["dst_col"], n=e["n"]))
        _launch(ops, rows, chain.num_bufs, chain.width, dev)
        for i, g in loaded:
            grads[i] = g if grads[i] is None else grads[i] + g
        # weight gradients: grad_W [N, K] = gy^T [N, rows] x [rows, K], one tcgen05 GEMM per Linear (tiny shapes on cuBLAS)
        for st in stacks.values():
            gy, x = st["gy"], st["x"]
            if rows % 8 == 0 and x.shape[1] % 8 == 0:
                gw_all = tc_linear.linear_bf16x3(tc_linear.split_tiled_transposed(gy), tc_linear.split_tiled_transposed(x), None,
                                                 tc_linear.SIX_TERMS, variant=2)
            else:
                gw_all = gy.t() @ x
            for e, off in st["members"]:
                gw = gw_all[off:off + e["n"]]
                grads[e["w"]] = gw if grads[e["w"]] is None else grads[e["w"]] + gw
        for i, (t, g) in enumerate(zip(tensors, grads)):
            if g is not None and t is not None and g.shape != t.shape:
                grads[i] = g.reshape(t.shape)
        return (None, *grads)
