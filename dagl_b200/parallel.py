"""Multi-GPU host logic: one process per GPU, image/tile-parallel replicas.

The graph block is independent per image (the reference loops over the batch,
dagl.py:245) and per chop tile (model/__init__.py:201-214), so ranks shard the
*units* (images or tiles) with no data-path collective (SURVEY.md §8e schemes 1
and 2).  torch.distributed is used only for the rendezvous, for the
max-over-ranks timing the bench contract asks for, and to gather results when a
caller wants the whole batch back on one rank.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def partition(n_units: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of ``n_units`` owned by ``rank``; sizes differ by at most 1."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_units, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device: torch.device) -> float:
    """All-reduce(MAX) of a scalar (used for step time: the job is as slow as its slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: torch.device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def forward_sharded(fn, batch: torch.Tensor, gather: bool = True) -> torch.Tensor:
    """Run ``fn`` on this rank's slice of ``batch`` (dim 0) and, if ``gather``,
    all-gather the per-rank outputs back into batch order on every rank.

    ``fn`` maps [b, C, H, W] -> [b, C', H, W].  Ranks whose slice is empty
    contribute nothing.  The only collective is the final (optional) gather of
    outputs; the graph block itself never communicates."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return fn(batch)
    world, rank = dist.get_world_size(), dist.get_rank()
    b0, b1 = partition(batch.shape[0], world, rank)
    local = fn(batch[b0:b1]) if b1 > b0 else None
    if not gather:
        return local
    # shapes: every rank's slice output is [n_r, C', H, W]; pad to the max slice for all_gather
    nmax = (batch.shape[0] + world - 1) // world
    if local is None:
        probe = fn(batch[:1])
        local_pad = torch.zeros((nmax,) + tuple(probe.shape[1:]), dtype=probe.dtype, device=probe.device)
    else:
        local_pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local_pad[: local.shape[0]] = local
    outs: List[torch.Tensor] = [torch.empty_like(local_pad) for _ in range(world)]
    dist.all_gather(outs, local_pad)
    parts = []
    for r in range(world):
        r0, r1 = partition(batch.shape[0], world, r)
        parts.append(outs[r][: r1 - r0])
    return torch.cat(parts, dim=0)


def gather_query_rows(rows: torch.Tensor, rows_per_rank: int, n_rows: int, group=None) -> torch.Tensor:
    """All-gather of per-rank query rows for single-image query sharding.

    ``rows`` is [B, world*rows_per_rank (>= n_rows), D] on every rank with only this rank's slice
    [rank*rows_per_rank, (rank+1)*rows_per_rank) filled.  Returns the assembled [B, n_rows, D] tensor (the same
    on every rank).  One collective: ``all_gather_into_tensor`` of the owned slices."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    B, _, D = rows.shape
    if rows.shape[1] < world * rows_per_rank:
        pad = torch.zeros(B, world * rows_per_rank - rows.shape[1], D, dtype=rows.dtype, device=rows.device)
        rows = torch.cat([rows, pad], dim=1)
    mine = rows[:, rank * rows_per_rank:(rank + 1) * rows_per_rank].contiguous()          # [B, rpr, D]
    out = torch.empty((world * B, rows_per_rank, D), dtype=rows.dtype, device=rows.device)   # ranks concatenated on dim 0
    dist.all_gather_into_tensor(out, mine, group=group)
    full = out.view(world, B, rows_per_rank, D).permute(1, 0, 2, 3).reshape(B, world * rows_per_rank, D)
    return full[:, :n_rows].contiguous()
