"""Multi-GPU plumbing for the hot path: streams are independent units, so the only
"parallelism" is a contiguous partition of the stream index space over ranks -- no data-path collective.
torch.distributed is used for the barrier and for reducing the timing / unit counters (NCCL on GPUs,
gloo on CPU in the tests)."""
from __future__ import annotations


def shard_range(total_streams: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block of streams owned by `rank` (reference plan: stream s -> GPU s // (S/world));
    remainders go to the lowest ranks."""
    base, rem = divmod(total_streams, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def aggregate_throughput(local_units: float, local_ms: float, device=None):
    """(sum of units over ranks, max of elapsed ms over ranks, units per second).  Works without an
    initialised process group (single process)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local_units, local_ms, local_units / (local_ms * 1e-3)
    u = torch.tensor([float(local_units)], dtype=torch.float64, device=device)
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(u.item()), float(t.item()), float(u.item()) / (float(t.item()) * 1e-3)
