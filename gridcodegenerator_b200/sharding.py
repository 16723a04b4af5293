"""Batch sharding for multi-GPU runs: contiguous slices of the state batch, no collective
(SURVEY.md section 8e).  Pure host logic, tested with world_size-2 gloo in tests/test_sharding.py."""
from typing import Tuple


def shard_range(num_states: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[first, last) of the states owned by `rank`; sizes differ by at most one."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(num_states, world_size)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def gather_to_rank0(local, num_states: int, rank: int, world_size: int, dist=None):
    """Optional gather of per-rank outputs (kept OFF the timed path: for Atlas N=65536 the gather
    would cost ~7x the compute).  `local` is a torch tensor of this rank's rows; shards may differ
    by one row, so rows are padded to the largest shard for the collective."""
    import torch
    if world_size == 1:
        return local
    dist = dist or torch.distributed
    sizes = [b - a for a, b in (shard_range(num_states, r, world_size) for r in range(world_size))]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    bufs = [torch.empty_like(buf) for _ in range(world_size)]
    dist.all_gather(bufs, buf)
    return torch.cat([b[:k] for b, k in zip(bufs, sizes)], 0) if rank == 0 else None


def max_over_ranks(value: float, world_size: int, device="cpu", dist=None) -> float:
    """Timing reduction used by bench.py: the slowest rank defines the step time."""
    import torch
    if world_size == 1:
        return float(value)
    dist = dist or torch.distributed
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
