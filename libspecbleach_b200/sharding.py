"""Host-side sharding of independent units (stream, channel) across ranks.

libspecbleach handles are mono and share nothing (reference:
src/processors/specbleach_2d_denoiser.c:30-38), so multi-GPU is pure data
parallelism: every rank owns a contiguous block of units, no collective touches
the data path; only the timing is reduced (max over ranks)."""
from __future__ import annotations


def shard_bounds(n_units: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced partition: the first ``n_units % world`` ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_units(units, world: int, rank: int):
    lo, hi = shard_bounds(len(units), world, rank)
    return units[lo:hi]


def job_throughput(local_units_seconds: float, local_ms: float, dist=None, device=None):
    """Whole-job audio-seconds per second: sum of audio over ranks / max time over ranks."""
    if dist is None or not dist.is_initialized():
        return local_units_seconds / (local_ms / 1e3), local_ms
    import torch
    t = torch.tensor([local_ms], dtype=torch.float64, device=device)
    a = torch.tensor([local_units_seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(a, op=dist.ReduceOp.SUM)
    return float(a.item()) / (float(t.item()) / 1e3), float(t.item())
