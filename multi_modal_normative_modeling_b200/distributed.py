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


_IDX_CACHE = {}


def _owned_index(owned: Sequence[int], device, dtype) -> torch.Tensor:
    """Device copy of the member ids of this rank, made once: building it per call is a pageable host-to-device copy that
    stalls the launch pipeline of every scoring pass."""
    key = (tuple(owned), str(device), dtype)
    t = _IDX_CACHE.get(key)
    if t is None:
        if len(_IDX_CACHE) > 64:
            _IDX_CACHE.clear()
        t = _IDX_CACHE[key] = torch.as_tensor(list(owned), dtype=dtype).to(device)
    return t


def gather_member_tables(local: torch.Tensor, owned: Sequence[int], n_members: int,
                         collective: bool = True) -> torch.Tensor:
    """local: [len(owned), W] records of this rank's members -> [n_members, W] on every rank.
    Shards may differ in length by one; they are padded to the longest for the collective.
    collective=False: the caller owns every member (no exchange even inside a distributed run).
    No host synchronisation on any path (no boolean-mask indexing, cached index tensors)."""
    rank, ws = world()
    if ws == 1 or not collective:
        if len(owned) == n_members and all(o == i for i, o in enumerate(owned)):
            return local                                     # every member, already in member order
        out = torch.empty((n_members, local.shape[1]), dtype=local.dtype, device=local.device)
        out.index_copy_(0, _owned_index(owned, local.device, torch.long), local)
        return out
    width = local.shape[1]
    longest = -(-n_members // ws)
    pad = torch.zeros((longest, width + 1), dtype=local.dtype, device=local.device)
    pad[: len(owned), :width] = local
    pad[: len(owned), width] = _owned_index(owned, local.device, local.dtype)
    pad[len(owned):, width] = -1
    buf = torch.empty((ws * longest, width + 1), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad)
    idx = buf[:, width].round().long()
    # padding rows (id -1) land in a scratch row behind the table instead of being masked out (a boolean mask would need
    # the number of kept rows on the host: a device -> host synchronisation per scoring pass)
    out = torch.empty((n_members + 1, width), dtype=local.dtype, device=local.device)
    out.index_copy_(0, torch.where(idx < 0, torch.full_like(idx, n_members), idx), buf[:, :width])
    return out[:n_members]
