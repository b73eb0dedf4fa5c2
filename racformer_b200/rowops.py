"""Row programs: the row-wise operators of a decoder iteration in one launch (racformer_b200/csrc/rowops.cu).

Host-side builder for `racf_row_program_forward` (include/racformer_ops.h). A program is a list of operators on a few
shared-memory row buffers; `RowProgram` fills the C records from tensors and `nn.Linear` / `nn.LayerNorm` modules and
launches once. Replaces, for inference, the PyTorch op chains of models/racformer_transformer.py:204-262
(position_encoder, norm1..3, fusion, FFN, cls / reg branches) and models/bev_self_attention.py:206-225 (queue fusion,
output_proj). CUDA only, fp32 only, forward only -- there is no CPU or PyTorch fallback in this module.
"""
import ctypes
import weakref

import torch

from . import _lib
from .caches import cache_epoch as _cache_epoch

_lib.load()

LOAD, LOAD_QUEUE, STORE, ADD, LINEAR, LINEAR_NARROW, LAYERNORM = 1, 2, 3, 4, 5, 6, 7
ZERO, RELU_MASK, DROPOUT, LAYERNORM_BWD, STORE_COLSUM, QUEUE_BWD = 8, 9, 10, 11, 12, 13
RELU, ACCUM = 1, 2
MAX_OPS = 128
MAX_QUEUE = 16


class RowOp(ctypes.Structure):
    """racf_row_op_t"""
    _fields_ = [("kind", ctypes.c_int), ("dst", ctypes.c_int), ("dst_col", ctypes.c_int), ("src", ctypes.c_int),
                ("src_col", ctypes.c_int), ("n", ctypes.c_int), ("k", ctypes.c_int), ("flags", ctypes.c_int),
                ("ld", ctypes.c_int), ("aux", ctypes.c_int), ("eps", ctypes.c_float), ("p0", ctypes.c_void_p),
                ("p1", ctypes.c_void_p), ("out", ctypes.c_void_p), ("p2", ctypes.c_void_p), ("out2", ctypes.c_void_p)]


_transposed = weakref.WeakKeyDictionary()     # module -> (key, W^T); dies with the module (ids are reused, modules are not)


CHUNK_COLS = 256
ROW_BUFFER_BYTES = 74 * 1024      # csrc/rowops.cu: 226 KB - 128 KB weight-tile ring - 24 KB reduction scratch


def chunked_transpose(w):
    """W [N, K] -> [ceil(N / 256), K, 256] with element [c, k, j] = W[c * 256 + j, k] (zero padded)."""
    N, K = w.shape
    chunks = (N + CHUNK_COLS - 1) // CHUNK_COLS
    out = w.new_zeros(chunks * CHUNK_COLS, K)
    out[:N] = w
    return out.view(chunks, CHUNK_COLS, K).transpose(1, 2).contiguous()


def weight_t(owner, w=None):
    """Weight [N, K] of `owner` (an nn.Linear, or any module holding the parameter `w`) -> its cached chunked transpose
    (rebuilt when the parameter changes)."""
    w = owner.weight if w is None else w
    key = (w.data_ptr(), w._version, w.device, _cache_epoch())
    hit = _transposed.get(owner)
    if hit is None or hit[0] != key:
        hit = _transposed[owner] = (key, chunked_transpose(w.detach()))
    return hit[1]


def _ptr(t):
    return None if t is None else t.data_ptr()


_sm_count = {}


def choose_rows_per_cta(rows, num_bufs, width, device=None):
    """Rows a CTA carries through the chain (4..8). A CTA's time grows with its rows (the FMA work; every CTA streams the same
    weights), and CTAs run one per SM in waves, so the cost of a launch is waves x rows_per_cta: 900 rows -> 7 (129 CTAs, one
    wave on 148 SMs, instead of 113 CTAs of 8), 2440 rows -> 6 (407 CTAs, three waves of 6, instead of 305 CTAs = two full
    waves + 9 CTAs of 8). Ties go to the larger tile (fewer weight re-reads). The row buffers and the k-group scratch share
    the 98 KB of shared memory the weight-tile ring leaves."""
    per_row = int(num_bufs) * int(width) * 4 + 3 * CHUNK_COLS * 4
    fits = [r for r in range(4, 9) if r * per_row <= ROW_BUFFER_BYTES + 3 * CHUNK_COLS * 4 * 8]
    if not fits:
        return 4
    if not torch.cuda.is_available():        # building a program on a machine without a GPU (tests of the host logic)
        sms = 148
    else:
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        sms = _sm_count.get(dev)
        if sms is None:
            sms = _sm_count[dev] = torch.cuda.get_device_properties(dev).multi_processor_count
    cost = lambda r: (-(-(-(-int(rows) // r)) // sms)) * r
    return min(fits, key=lambda r: (cost(r), -r))


class RowProgram:
    """Builds and launches one row program over `rows` rows. Buffers are numbered 0..num_bufs-1, `width` floats each."""

    def __init__(self, rows, width, num_bufs=3, rows_per_cta=None, device=None):
        if rows_per_cta is None:
            rows_per_cta = choose_rows_per_cta(rows, num_bufs, width, device)
        self.rows, self.width, self.num_bufs, self.rows_per_cta = int(rows), int(width), int(num_bufs), int(rows_per_cta)
        self.device = device
        self.ops, self._keep = [], []

    # ---- helpers
    def _rows2d(self, t, n):
        """-> (tensor, ld) of a [rows, >= n] fp32 CUDA matrix whose rows are unit-stride."""
        if not (t.is_cuda and t.dtype == torch.float32):
            raise RuntimeError("row program tensors must be fp32 CUDA tensors")
        t2 = t.reshape(-1, t.shape[-1]) if t.dim() != 2 else t
        if t2.shape[0] != self.rows or t2.shape[1] < n:
            raise RuntimeError(f"row program: expected a [{self.rows}, >= {n}] matrix, got {tuple(t2.shape)}")
        if t2.stride(1) != 1:
            t2 = t2.contiguous()
        if self.device is None:
            self.device = t2.device
        elif t2.device != self.device:
            raise RuntimeError("row program: tensors on different devices")
        self._keep.append(t2)
        return t2, t2.stride(0)

    def _param(self, p):
        if p is None:
            return None
        p = p.detach()
        if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
            raise RuntimeError("row program parameters must be contiguous fp32 CUDA tensors")
        self._keep.append(p)
        return p.data_ptr()

    def _push(self, **kw):
        if len(self.ops) >= MAX_OPS:
            raise RuntimeError("row program too long")
        self.ops.append(RowOp(**kw))

    # ---- operators
    def load(self, dst, t, n=None, dst_col=0):
        n = t.shape[-1] if n is None else n
        t2, ld = self._rows2d(t, n)
        self._push(kind=LOAD, dst=dst, dst_col=dst_col, n=n, ld=ld, p0=t2.data_ptr())

    def load_queue(self, dst, values, logits, rows_per_batch, queue, dst_col=0):
        """values [B*queue, rows_per_batch, C] (MSDA output), logits [rows, queue] or None -> softmax-weighted queue sum."""
        if not (values.is_cuda and values.dtype == torch.float32 and values.is_contiguous() and values.dim() == 3
                and values.shape[1] == rows_per_batch and values.shape[0] * rows_per_batch == self.rows * queue):
            raise RuntimeError("row program: queue values must be a contiguous [B * queue, rows_per_batch, C] CUDA tensor")
        if queue > MAX_QUEUE:
            raise RuntimeError("row program: queue too long")
        self._keep.append(values)
        lg = None
        if logits is not None:
            lg, ld = self._rows2d(logits, queue)
            if ld != queue:
                lg = lg.contiguous()
                self._keep.append(lg)
        self._push(kind=LOAD_QUEUE, dst=dst, dst_col=dst_col, n=values.shape[2], k=rows_per_batch, aux=queue,
                   ld=values.shape[2], p0=values.data_ptr(), p1=_ptr(lg))

    def store(self, src, n, src_col=0, out=None):
        """-> the [rows, n] output tensor (allocated here unless given)."""
        if out is None:
            out = torch.empty((self.rows, n), dtype=torch.float32, device=self.device)
        o2, ld = self._rows2d(out, n)
        if o2.data_ptr() != out.data_ptr():
            raise RuntimeError("row program: output rows must be unit-stride")
        self._push(kind=STORE, src=src, src_col=src_col, n=n, ld=ld, out=o2.data_ptr())
        return out

    def add(self, dst, src, n, dst_col=0, src_col=0):
        self._push(kind=ADD, dst=dst, dst_col=dst_col, src=src, src_col=src_col, n=n)

    def linear(self, dst, src, lin, relu=False, dst_col=0, src_col=0, weight=None, bias=None):
        """buf[dst] = act(lin(buf[src])); wide outputs read the cached chunked transpose of the weight, narrow ones (< 32
        columns) the weight itself. `weight` / `bias`: parameters held by `lin` under other names (e.g. the in_proj of an
        nn.MultiheadAttention); `lin` is then only the cache key."""
        weight = lin.weight if weight is None else weight
        bias = (lin.bias if hasattr(lin, "bias") else None) if bias is None else bias
        N, K = weight.shape
        narrow = N < 32 and N * K <= 8192 and (N * K) % 4 == 0
        w = self._param(weight if narrow else weight_t(lin, weight))
        self._push(kind=LINEAR_NARROW if narrow else LINEAR, dst=dst, dst_col=dst_col, src=src, src_col=src_col, n=N, k=K,
                   flags=RELU if relu else 0, p0=w, p1=self._param(bias))

    def layernorm(self, buf, ln, col=0, relu=False):
        n = ln.normalized_shape[-1]
        if len(ln.normalized_shape) != 1:
            raise RuntimeError("row program: LayerNorm over the last dimension only")
        self._push(kind=LAYERNORM, dst=buf, dst_col=col, n=n, eps=float(ln.eps), flags=RELU if relu else 0,
                   p0=self._param(ln.weight), p1=self._param(ln.bias))

    # ---- launch
    def run(self):
        if not self.ops:
            raise RuntimeError("empty row program")
        arr = (RowOp * len(self.ops))(*self.ops)
        with torch.cuda.device(self.device):
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            rc = _lib.load().racf_row_program_forward(arr, len(self.ops), self.rows, self.rows_per_cta, self.num_bufs,
                                                      self.width, stream)
        _lib.check(rc, "racf_row_program_forward")
        self._keep.clear()
