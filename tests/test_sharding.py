"""Multi-GPU host logic on the CPU: contiguous shards, optional gather, max-over-ranks timing,
with a real world_size-2 gloo process group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gridcodegenerator_b200.sharding import gather_to_rank0, max_over_ranks, shard_range


@pytest.mark.parametrize("N,W", [(65536, 8), (128, 8), (7, 2), (5, 8), (0, 4), (65537, 4)])
def test_shards_partition_the_batch(N, W):
    spans = [shard_range(N, r, W) for r in range(W)]
    assert spans[0][0] == 0 and spans[-1][1] == N
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def test_bad_rank_rejected():
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, N, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(N, rank, world)
    # stand-in for the per-state kernel: any per-row function of the global state index
    rows = torch.arange(a, b, dtype=torch.float32).unsqueeze(1) * torch.tensor([[1.0, 2.0, 3.0]])
    full = gather_to_rank0(rows, N, rank, world)
    slow = max_over_ranks(10.0 + rank, world)
    dist.barrier()
    if rank == 0:
        q.put((full.numpy(), slow))
    dist.destroy_process_group()


def test_world_size_2_gloo_gather_and_timing():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    N, W = 101, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, W, port, N, q)) for r in range(W)]
    for p in procs:
        p.start()
    full, slow = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.arange(N, dtype=np.float32)[:, None] * np.array([[1.0, 2.0, 3.0]], dtype=np.float32)
    assert np.array_equal(full, expect)
    assert slow == 11.0
