"""How independent pairings shard over devices / ranks: contiguous slices, no data-path
collective.  The same split is used inside libzkpair.so across the devices of one context
(csrc/kernels.cu slice_of) and by bench.py across torchrun ranks; the collectives are the MAX of
the per-rank device times the measurement contract asks for and -- for ONE product over all pairs -- the gather of
one 576-byte Fp12 partial per rank."""
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


def gather_partials(partial, world: int):
    """The one exchange step of a sharded multi-pairing product: every rank contributes its Fp12
    partial product (72 little-endian u64 limbs as an int64 tensor, on the device the process group's backend
    expects) and receives all of them as a (world, 72) tensor, rank-major.  Fp12 multiplication is exact,
    commutative and associative, so the product of the rows is bit-identical to the single-device product."""
    import torch
    import torch.distributed as dist
    partial = partial.reshape(72).contiguous()
    if world <= 1:
        return partial.reshape(1, 72).clone()
    out = torch.empty((world, 72), dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(out.view(-1), partial)
    return out
