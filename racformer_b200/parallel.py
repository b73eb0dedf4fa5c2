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
    """Bucketed average of parameter gradients across ranks (what DDP does for the reference, C2 in SURVEY 2.4).

    Gradients are packed into flat fp32 buckets of `bucket_bytes` (default 25 MB, DDP's default) so the NCCL
    all-reduce is launch-latency-efficient over NVLink/NVSwitch; one async all-reduce per bucket, then unpack.
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

    def all_reduce(self):
        if not (dist.is_initialized() and dist.get_world_size() > 1):
            return 0
        world = dist.get_world_size()
        if self._flat is None:
            self._flat = [torch.empty(sum(p.numel() for p in b), dtype=b[0].dtype, device=b[0].device) for b in self.buckets]
        works = []
        for flat, bucket in zip(self._flat, self.buckets):
            off = 0
            for p in bucket:
                n = p.numel()
                if p.grad is None:
                    flat[off:off + n].zero_()
                else:
                    flat[off:off + n].copy_(p.grad.reshape(-1))
                off += n
            works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True))
        nbytes = 0
        for work, flat, bucket in zip(works, self._flat, self.buckets):
            work.wait()
            flat.div_(world)
            off = 0
            for p in bucket:
                n = p.numel()
                if p.grad is None:
                    p.grad = flat[off:off + n].reshape(p.shape).clone()
                else:
                    p.grad.copy_(flat[off:off + n].reshape(p.shape))
                off += n
            nbytes += flat.numel() * flat.element_size()
        return nbytes
