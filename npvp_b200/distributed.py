"""Multi-GPU inference: clips are independent, so the batch is sharded contiguously across ranks
(one process per GPU, weights replicated) and the only exchange is one all-gather of the predicted
frames (NCCL over NVLink/NVSwitch on the GPU box; gloo in the CPU tests).  No all-reduce exists on this path.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of rank ``rank``; the first ``n_clips % world`` ranks hold one extra clip."""
    base, extra = divmod(n_clips, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_frames(local: torch.Tensor, n_clips: int, group=None) -> torch.Tensor:
    """All-gather per-rank predicted frames (n_local, T, C, H, W) into the global batch (n_clips, T, C, H, W).
    Uneven shards are padded to the largest shard for the collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return local
    sizes = [shard_bounds(n_clips, r, world) for r in range(world)]
    n_max = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    assert local.shape[0] == hi - lo, f"rank {rank} holds {local.shape[0]} clips, expected {hi - lo}"
    if local.shape[0] < n_max:
        pad = torch.zeros((n_max - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    local = local.contiguous()
    out = torch.empty((world * n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, local, group=group)
    else:
        dist.all_gather(list(out.split(n_max, dim=0)), local, group=group)
    if all(h - l == n_max for l, h in sizes):
        return out
    return torch.cat([out[r * n_max: r * n_max + (h - l)] for r, (l, h) in enumerate(sizes)], dim=0)


def predict_sharded(fn, past_frames_global: torch.Tensor, group=None, **kw) -> torch.Tensor:
    """Run ``fn`` (e.g. ``model.predict`` / ``model.rollout``) on this rank's shard of the global batch, gather all."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = past_frames_global.shape[0]
    lo, hi = shard_bounds(n, rank, world)
    assert n >= world, f"predict_sharded: {n} clips cannot be spread over {world} ranks (every rank needs at least one clip)"
    local = fn(past_frames_global[lo:hi], **kw)
    return gather_frames(local, n, group)
