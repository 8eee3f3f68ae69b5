"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests).

The hot path shards naturally: frames for RDF / CN / clusters / hydration / flux, atoms for MSD,
channels for the ACFs; residence time shards its neighbour search by frames and its survival correlation by
central atoms.  Each API call does ONE small collective after all local work: integer histograms are summed with
an int64 all-reduce (order independent -> bit exact), fp64 partial sums with an fp64 all-reduce.  Residence time
is the one path with a real exchange step: the (frame, central, neighbour) entries found on a rank's frames are
routed to the rank that owns the central atom (``exchange_rows``: all-to-all over NVLink with NCCL, an all-gather
based equivalent with gloo).  Without an initialised process group everything degenerates to a single rank.
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


def owner_of_rows(idx: torch.Tensor, n: int, w: int | None = None) -> torch.Tensor:
    """Vectorised ``owner_of``: rank owning each unit index of ``idx`` under ``shard_range``'s block partition."""
    w = world_size() if w is None else w
    base, rem = divmod(n, w)
    cut = rem * (base + 1)
    if idx.dtype != torch.int32 or n >= (1 << 31):           # (int32 indices stay int32: half the bytes per pass)
        idx = idx.to(torch.int64)
    if rem == 0:
        return torch.div(idx, max(base, 1), rounding_mode="floor")
    return torch.where(idx < cut, torch.div(idx, base + 1, rounding_mode="floor"),
                       rem + torch.div(idx - cut, max(base, 1), rounding_mode="floor"))


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """In-place SUM all-reduce (int64 or float64); no-op on a single rank."""
    d = _dist()
    if d and d.get_world_size() > 1:
        d.all_reduce(t, op=d.ReduceOp.SUM)
    return t


def all_gather_blocks(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """Inverse of ``shard_range``: every rank holds rows [lo, hi) of an [n_total, ...] array (``local`` = those rows) and
    gets the whole array.  One ``all_gather_into_tensor`` of blocks padded to the longest shard (at most one row of
    padding per rank) -- the bytes on the wire are the array itself, not world_size zero-padded copies of it."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return local
    w = d.get_world_size()
    longest = (n_total + w - 1) // w
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((w * longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    d.all_gather_into_tensor(out, pad)
    if n_total == w * longest:
        return out
    parts = []
    for r in range(w):
        lo, hi = shard_range(n_total, r, w)
        parts.append(out[r * longest: r * longest + (hi - lo)])
    return torch.cat(parts, dim=0)


def barrier():
    d = _dist()
    if d and d.get_world_size() > 1:
        d.barrier()


def all_reduce_max_(t: torch.Tensor) -> torch.Tensor:
    d = _dist()
    if d and d.get_world_size() > 1:
        d.all_reduce(t, op=d.ReduceOp.MAX)
    return t


def exchange_rows(rows: torch.Tensor, dest: torch.Tensor) -> torch.Tensor:
    """Route row i of ``rows`` [M, C] to rank ``dest[i]``; returns the rows addressed to this rank, grouped by source
    rank and in their original order within a source.  Every rank must call it (M may be 0).

    NCCL: one all_to_all_single of the per-destination counts and one of the rows (variable splits).  Backends
    without all-to-all (gloo, CPU tests): every rank all-gathers the destination-sorted rows, padded to the longest
    contribution, and keeps its own slices.
    """
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return rows
    w, me = d.get_world_size(), d.get_rank()
    rows = rows.contiguous()
    # stable sort by destination: 8-bit keys (one radix pass instead of the eight of an int64 sort)
    key = dest.to(torch.uint8) if w <= 255 else dest.to(torch.int64)
    order = torch.argsort(key, stable=True)
    srt = rows.index_select(0, order)
    send = torch.bincount(key if w <= 255 else dest.to(torch.int64), minlength=w).to(torch.int64)
    if d.get_backend() == "nccl":
        recv = torch.empty_like(send)
        d.all_to_all_single(recv, send)
        send_l, recv_l = send.tolist(), recv.tolist()
        out = torch.empty((sum(recv_l),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
        d.all_to_all_single(out, srt, output_split_sizes=recv_l, input_split_sizes=send_l)
        return out
    counts = [torch.empty_like(send) for _ in range(w)]
    d.all_gather(counts, send)
    counts = torch.stack(counts).cpu()                      # [source, destination]
    longest = int(counts.sum(dim=1).max().item())
    pad = torch.zeros((longest,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    pad[: srt.shape[0]] = srt
    gathered = [torch.empty_like(pad) for _ in range(w)]
    d.all_gather(gathered, pad)
    parts = []
    for s_ in range(w):
        off = int(counts[s_, :me].sum().item())
        parts.append(gathered[s_][off: off + int(counts[s_, me].item())])
    return torch.cat(parts, dim=0)
