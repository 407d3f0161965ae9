"""Sharding of independent sequences over ranks: no data-path collective, only the launch barrier and
the reduction of the timing (max over ranks) / unit counts (sum over ranks) that bench.py reports.

One process per GPU; rank r of a world of N owns sequences r*S .. r*S + S-1 (S sequences per GPU, weak scaling)."""
from __future__ import annotations


def sequences_of_rank(rank: int, world: int, sequences_per_gpu: int = 1) -> list[int]:
    if not (0 <= rank < world) or sequences_per_gpu < 1:
        raise ValueError("bad rank/world/sequences_per_gpu")
    return [rank * sequences_per_gpu + k for k in range(sequences_per_gpu)]


def aggregate_rate(region_ms: float, units: float, device=None):
    """Whole-job rate: units summed over ranks / the slowest rank's region time.  Returns (rate_per_s, max_region_s, units).

    Works with any initialised torch.distributed backend (nccl on the GPUs, gloo in the CPU tests) and without one."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([region_ms / 1e3], dtype=torch.float64, device=device)
    n = torch.tensor([float(units)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
    sec, cnt = float(t.item()), float(n.item())
    return (cnt / sec if sec > 0 else 0.0), sec, cnt
