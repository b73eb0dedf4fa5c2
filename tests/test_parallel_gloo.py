"""CPU: the N>1 host logic (sample sharding, max-over-ranks timing, bucketed gradient all-reduce) with gloo, world_size 2."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from racformer_b200 import parallel


def test_shard_range_covers_every_sample_once():
    for n in (0, 1, 7, 8, 9, 64):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                a, b = parallel.shard_range(n, r, world)
                assert 0 <= a <= b <= n and b - a in (n // world, n // world + 1)
                seen += list(range(a, b))
            assert seen == list(range(n))


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w, dev = parallel.init_distributed(backend="gloo")
    assert (r, w) == (rank, world) and dev.type == "cpu"
    parallel.barrier(dev)
    # device-time reduction: the job time is the slowest rank's
    assert parallel.max_over_ranks(10.0 + rank, dev) == 10.0 + world - 1
    # gradient averaging over small buckets (forces several buckets) incl. a parameter without gradient
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 4))
    extra = torch.nn.Parameter(torch.zeros(5))
    x = torch.full((3, 16), float(rank + 1))
    model(x).sum().backward()
    local = [p.grad.clone() for p in model.parameters()]
    reducer = parallel.GradientAllReducer(list(model.parameters()) + [extra], bucket_bytes=256)
    assert len(reducer.buckets) > 2
    moved = reducer.all_reduce()
    assert moved == sum(p.numel() * 4 for p in list(model.parameters()) + [extra])
    gathered = [[torch.zeros_like(g) for _ in range(world)] for g in local]
    for g, lst in zip(local, gathered):
        dist.all_gather(lst, g)
    for p, lst in zip(model.parameters(), gathered):
        torch.testing.assert_close(p.grad, sum(lst) / world)
    assert torch.equal(extra.grad, torch.zeros(5))
    # overlapped mode: grads are views into the buckets, hooks launch each bucket's all-reduce during backward
    torch.manual_seed(0)
    model2 = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 4))
    extra2 = torch.nn.Parameter(torch.zeros(5))
    red2 = parallel.GradientAllReducer(list(model2.parameters()) + [extra2], bucket_bytes=256)
    for step in range(2):                                  # second step: buckets are re-zeroed, hooks re-armed
        red2.prepare()
        (model2(x).sum() * (step + 1)).backward()
        assert red2.launched_in_backward >= 1              # at least one bucket went out before backward returned
        assert red2.finish() == moved
        for p, q in zip(model2.parameters(), model.parameters()):
            torch.testing.assert_close(p.grad, q.grad * (step + 1))
        assert torch.equal(extra2.grad, torch.zeros(5))
        assert all(p.grad.data_ptr() == v.data_ptr() for b, vs in zip(red2.buckets, red2._views) for p, v in zip(b, vs))
    # sharding: each rank works on its own samples, results gathered only for the check
    a, b = parallel.shard_range(5, rank, world)
    mine = torch.arange(a, b, dtype=torch.float32) * 2
    sizes = [parallel.shard_range(5, r, world) for r in range(world)]
    bufs = [torch.zeros(e - s) for s, e in sizes]
    dist.all_gather_object(out_list := [None] * world, mine.tolist())
    assert sum(out_list, []) == [0.0, 2.0, 4.0, 6.0, 8.0]
    dist.destroy_process_group()


def test_world_size_two_gloo():
    port = 29650 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, None), nprocs=2, join=True)
