"""Multi-GPU plumbing: one process per GPU (torchrun env), torch.distributed over NCCL (gloo on CPU for tests).

The sampling path has no exchange step: every sample -- with its 6 cameras x 8 frames -- is
independent, so inference shards samples across ranks with no data-path collective. The only collective is the
parameter-gradient all-reduce of the training configuration (reference: MMDistributedDataParallel, train.py:139-140).
"""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs NVML reports as local to the GPU, BEFORE any pinned host buffer is allocated, so the
    staging memory of the end-to-end path is first-touched on the GPU's NUMA node and eight ranks do not funnel their
    host-to-device traffic through one socket. Best effort: returns the CPU list or None (no NVML, containers that hide
    the topology, affinity not permitted)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = torch.cuda.get_device_properties(device_index).uuid
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
        except Exception:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = sorted(c for c in cpus if c in allowed)
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
        return cpus or None
    except Exception:
        return None


def init_distributed(backend=None):
    """Initialise the default process group from the torchrun environment. Returns (rank, world, device)."""
    rank, world, local_rank = env_rank_world()
    use_cuda = torch.cuda.is_available()
    device = torch.device("cuda", local_rank) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)
        if world > 1 and os.environ.get("RACF_NUMA_BIND", "1") != "0":
            bind_to_gpu_numa_node(local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if use_cuda else "gloo")
        if backend == "nccl":
            dist.init_process_group(backend, device_id=device)
        else:
            dist.init_process_group(backend)
    return rank, world, device


def shard_range(num_samples, rank, world):
    """Contiguous, balanced [start, stop) of the samples owned by `rank` (first num_samples % world ranks get one more)."""
    base, extra = divmod(num_samples, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value, device):
    """Max of a python float over all ranks (device-timed milliseconds -> job time)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(device):
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if device.type == "cuda":
        torch.cuda.synchronize(device)


class GradientAllReducer:
    """Bucketed average of parameter gradients across ranks (what DDP does for the reference: train.py:139-140,
    MMDistributedDataParallel; C2 in SURVEY 2.4) over NCCL / NVLink.

    Parameters are packed, in reverse order (~ the order in which backward finishes them), into flat fp32 buckets of
    `bucket_bytes` (25 MB, DDP's default).

    * `prepare()` + backward + `finish()` -- the overlapped mode: every parameter's `.grad` is a view into its bucket
      (no pack / unpack copies), autograd accumulates straight into the bucket, and a post-accumulate hook launches the
      bucket's asynchronous all-reduce the moment its last gradient has been produced, while backward is still running.
      `finish()` launches what is left (buckets holding parameters that received no gradient), waits, and averages.
    * `all_reduce()` -- the plain mode for gradients that already exist: pack, reduce, unpack after backward.
    """

    def __init__(self, params, bucket_bytes=25 << 20):
        self.params = [p for p in params if p.requires_grad]
        self.buckets, cur, size = [], [], 0
        for p in reversed(self.params):   # reverse order ~ order in which backward produces gradients
            nbytes = p.numel() * p.element_size()
            if cur and size + nbytes > bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
            cur.append(p)
            size += nbytes
        if cur:
            self.buckets.append(cur)
        self._flat = None
        self._views = None
        self._bucket_of = {id(p): b for b, bucket in enumerate(self.buckets) for p in bucket}
        self._hooks, self._armed = [], False
        self._pending, self._works = [], {}
        self.launched_in_backward = 0

    # ---- shared pieces
    def _buffers(self):
        if self._flat is None:
            self._flat = [torch.zeros(sum(p.numel() for p in b), dtype=b[0].dtype, device=b[0].device) for b in self.buckets]
            self._views = []
            for flat, bucket in zip(self._flat, self.buckets):
                off, views = 0, []
                for p in bucket:
                    views.append(flat[off:off + p.numel()].view(p.shape))
                    off += p.numel()
                self._views.append(views)
        return self._flat

    @staticmethod
    def _distributed():
        return dist.is_initialized() and dist.get_world_size() > 1

    def _launch(self, b):
        flat = self._flat[b]
        if dist.get_backend() == "nccl":
            self._works[b] = (dist.all_reduce(flat, op=dist.ReduceOp.AVG, async_op=True), False)
        else:
            self._works[b] = (dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True), True)

    @property
    def nbytes(self):
        return sum(p.numel() * p.element_size() for p in self.params)

    # ---- overlapped mode
    def prepare(self):
        """Call before backward: zero the buckets (one memset each), point every `.grad` at its slice, arm the hooks."""
        flats = self._buffers()
        for flat in flats:
            flat.zero_()
        for bucket, views in zip(self.buckets, self._views):
            for p, v in zip(bucket, views):
                p.grad = v
        if not self._hooks:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self._pending = [len(b) for b in self.buckets]
        self._works, self._armed, self.launched_in_backward = {}, True, 0

    def _on_grad(self, p):
        if not self._armed:
            return
        b = self._bucket_of[id(p)]
        self._pending[b] -= 1
        if self._pending[b] == 0 and self._distributed():
            self._launch(b)
            self.launched_in_backward += 1

    def finish(self):
        """Call after backward: reduce the buckets whose hooks did not all fire, wait for everything, average.
        Returns the bytes all-reduced (0 on a single rank)."""
        self._armed = False
        if not self._distributed():
            return 0
        world = dist.get_world_size()
        for b in range(len(self.buckets)):
            if b not in self._works:
                self._launch(b)
        for b, (work, divide) in self._works.items():
            work.wait()
            if divide:
                self._flat[b].div_(world)
        return self.nbytes

    # ---- plain mode
    def all_reduce(self):
        if not self._distributed():
            return 0
        world = dist.get_world_size()
        flats = self._buffers()
        self._works = {}
        for b, (flat, bucket, views) in enumerate(zip(flats, self.buckets, self._views)):
            for p, v in zip(bucket, views):
                if p.grad is None:
                    v.zero_()
                elif p.grad.data_ptr() != v.data_ptr():
                    v.copy_(p.grad)
            self._launch(b)
        for b, (work, divide) in self._works.items():
            work.wait()
            if divide:
                flats[b].div_(world)
            for p, v in zip(self.buckets[b], self._views[b]):
                if p.grad is None:
                    p.grad = v.clone()
                elif p.grad.data_ptr() != v.data_ptr():
                    p.grad.copy_(v)
        return self.nbytes
