"""How independent pairings shard over devices / ranks: contiguous slices, no data-path
collective.  The same split is used inside libzkpair.so across the devices of one context
(csrc/kernels.cu slice_of) and by bench.py across torchrun ranks; the only collective is the MAX of
the per-rank device times the measurement contract asks for."""
from __future__ import annotations


def slice_bounds(n: int, parts: int, idx: int) -> tuple[int, int]:
    """Elements [lo, hi) of a batch of n that part `idx` of `parts` owns (mirrors slice_of in kernels.cu)."""
    if parts < 1 or not 0 <= idx < parts:
        raise ValueError("bad partition")
    return n * idx // parts, n * (idx + 1) // parts


def synthetic_first_index(rank: int, per_rank: int) -> int:
    """Weak scaling: rank r generates / owns the seeded pairs [r * per_rank, (r + 1) * per_rank)."""
    return rank * per_rank


def max_over_ranks(value: float, world: int, device=None) -> float:
    """MAX over ranks of a per-rank time (torch.distributed must be initialised when world > 1)."""
    if world <= 1:
        return float(value)
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_rate(units_per_rank: int, steps: int, world: int, max_ms: float) -> float:
    """Units all ranks processed divided by the slowest rank's time."""
    return world * units_per_rank * steps / (max_ms * 1e-3)
