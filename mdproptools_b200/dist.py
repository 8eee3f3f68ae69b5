"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests).

The hot path shards naturally: frames for RDF / CN / clusters / hydration / flux, atoms for MSD,
channels for the ACFs, central atoms for residence time.  Each API call does ONE small collective after all
local work: integer histograms are summed with an int64 all-reduce (order independent -> bit exact), fp64
partial sums with an fp64 all-reduce.  Without an initialised process group everything degenerates to a
single rank.
"""
from __future__ import annotations

import torch


def _dist():
    try:
        import torch.distributed as d
    except Exception:  # pragma: no cover
        return None
    return d if d.is_available() and d.is_initialized() else None


def rank() -> int:
    d = _dist()
    return d.get_rank() if d else 0


def world_size() -> int:
    d = _dist()
    return d.get_world_size() if d else 1


def shard_range(n: int, r: int | None = None, w: int | None = None):
    """Contiguous block [lo, hi) of ``n`` units owned by rank r of w (blocks differ by at most one unit)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    base, rem = divmod(n, w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def owner_of(i: int, n: int, w: int | None = None) -> int:
    w = world_size() if w is None else w
    base, rem = divmod(n, w)
    cut = rem * (base + 1)
    return i // (base + 1) if i < cut else rem + (i - cut) // max(base, 1)


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place SUM all-reduce (int64 or float64); no-op on a single rank."""
    d = _dist()
    if d and d.get_world_size() > 1:
        d.all_reduce(t, op=d.ReduceOp.SUM)
    return t


def barrier():
    d = _dist()
    if d and d.get_world_size() > 1:
        d.barrier()
