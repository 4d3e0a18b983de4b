"""Multi-GPU plumbing: ensemble members are independent, so they are dealt round-robin to the
ranks (one process per GPU) and trained with NO data-path collective.  The only exchange step
is an all-gather of fixed-size per-member records (per-ROI normative statistics, per-ROI AUC,
per-subject deviations, fold AUC) so that every rank can average over the modalities / seeds of
a fold, which live on different GPUs (group analysis :212-215).  NCCL over NVLink on GPUs,
gloo on CPU (tests)."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_members(n_members: int, rank: int, world_size: int, cost: Sequence[float] = None) -> List[int]:
    """Indices owned by `rank`.  Members are sorted by cost class (e.g. FLOPs ~ input width) and
    dealt round-robin so every rank gets the same mix."""
    order = list(range(n_members))
    if cost is not None:
        order.sort(key=lambda i: (-float(cost[i]), i))
    return sorted(order[rank::world_size])


def gather_member_tables(local: torch.Tensor, owned: Sequence[int], n_members: int,
                         collective: bool = True) -> torch.Tensor:
    """local: [len(owned), W] records of this rank's members -> [n_members, W] on every rank.
    Shards may differ in length by one; they are padded to the longest for the collective.
    collective=False: the caller owns every member (no exchange even inside a distributed run)."""
    rank, ws = world()
    if ws == 1 or not collective:
        out = torch.empty((n_members, local.shape[1]), dtype=local.dtype, device=local.device)
        out[torch.as_tensor(list(owned), device=local.device, dtype=torch.long)] = local
        return out
    width = local.shape[1]
    longest = -(-n_members // ws)
    pad = torch.zeros((longest, width + 1), dtype=local.dtype, device=local.device)
    pad[: len(owned), :width] = local
    pad[: len(owned), width] = torch.as_tensor(list(owned), dtype=local.dtype, device=local.device)
    pad[len(owned):, width] = -1
    buf = torch.empty((ws * longest, width + 1), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad)
    idx = buf[:, width].round().long()
    keep = idx >= 0
    out = torch.empty((n_members, width), dtype=local.dtype, device=local.device)
    out[idx[keep]] = buf[keep, :width]
    return out
