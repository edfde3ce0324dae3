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


class BlockGather:
    """The one exchange of the multi-GPU path, issued block by block: every rank submits the frames of an autoregressive block
    (n_local, take, C, H, W) as soon as they exist; the collective runs asynchronously (NCCL: on its own stream, overlapping
    the next block's kernels) and the receiving side assembles the rank-major global batch (world * n_local, num_future, C, H, W).

    ``dst``: group rank that receives (default 0: ``dist.gather`` = grouped NCCL send / recv, the other ranks only send and
    get ``None`` from :meth:`result`), or ``None`` for an all-gather to every rank (r01 behaviour: each rank then receives
    ``world - 1`` shards per block, which at 8 GPUs and fp32 frames cost 7 x 352 MB of NVLink + HBM traffic per rank and step).
    The payload type is the caller's business (fp32 model-space frames, half, or uint8 pixel frames: 4x smaller)."""

    def __init__(self, group, dst, num_future: int):
        self.group, self.dst, self.num_future = group, dst, int(num_future)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.receives = dst is None or self.rank == dst
        self.full, self.pending, self.stream = None, [], None

    def submit(self, block: torch.Tensor, done: int):
        """block: contiguous (n_local, take, ...) frames of target positions [done, done + take)."""
        assert block.is_contiguous()
        n, take = block.shape[0], block.shape[1]
        allb, nccl = None, dist.get_backend(self.group) == "nccl"
        if self.receives:
            allb = torch.empty((self.world,) + tuple(block.shape), dtype=block.dtype, device=block.device)
        if self.dst is None:
            if nccl:
                work = dist.all_gather_into_tensor(allb, block, group=self.group, async_op=True)
            else:
                work = dist.all_gather(list(allb.unbind(0)), block, group=self.group, async_op=True)
        else:
            dst_global = self.dst if self.group is None else dist.get_global_rank(self.group, self.dst)
            work = dist.gather(block, list(allb.unbind(0)) if self.receives else None, dst=dst_global, group=self.group, async_op=True)
        if self.receives:
            if self.full is None:
                self.full = torch.empty((self.world * n, self.num_future) + tuple(block.shape[2:]), dtype=block.dtype, device=block.device)
                if block.is_cuda:
                    self.stream = torch.cuda.Stream(device=block.device)
                    self.stream.wait_stream(torch.cuda.current_stream(block.device))   # `full` may reuse memory of earlier work
                    self.full.record_stream(self.stream)
            dstv = self.full.view(self.world, n, self.num_future, *block.shape[2:])[:, :, done:done + take]
            if block.is_cuda:                                   # assembly on a side stream, under the next block's kernels
                with torch.cuda.stream(self.stream):
                    work.wait()
                    dstv.copy_(allb)
                allb.record_stream(self.stream)
            else:
                work.wait()
                dstv.copy_(allb)
        self.pending.append((work, allb, block))                # keep the buffers alive until the collective has run

    def result(self):
        """The assembled global batch on receiving ranks (the caller's stream waits for the last assembly), else ``None``."""
        if not self.receives:
            for work, _, blk in self.pending:                   # senders: the payloads must outlive their sends
                if not blk.is_cuda:
                    work.wait()
            return None
        if self.stream is not None:
            torch.cuda.current_stream(self.full.device).wait_stream(self.stream)
        return self.full
