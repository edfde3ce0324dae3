"""CPU, world_size 2 over gloo: batch sharding + the frame all-gather reproduce the single-process result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from npvp_b200.distributed import gather_frames, predict_sharded, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_predict(x):
    """Stand-in for model.rollout: per-clip function (no cross-clip mixing), like the real path."""
    return torch.stack([x[:, 0] * 2 + 1, x[:, 1] - x[:, 0], x.sum(1)], dim=1)


def _worker(rank, world, port, n_clips, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        x = torch.rand((n_clips, 2, 3, 8, 8), generator=g)
        full = _fake_predict(x)
        out = predict_sharded(_fake_predict, x)
        lo, hi = shard_bounds(n_clips, rank, world)
        again = gather_frames(_fake_predict(x[lo:hi]), n_clips)
        ret[rank] = bool(torch.equal(out, full) and torch.equal(again, full))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [4, 5, 2])
def test_sharded_predict_matches_single_process(n_clips):
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, n_clips, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)


def _block_worker(rank, world, port, dst, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from npvp_b200.distributed import BlockGather
        g = torch.Generator().manual_seed(1)
        n, nf = 3, 7                                            # clips per rank, frames per clip: blocks of 3 + 3 + 1
        full = torch.rand((world * n, nf, 2, 4, 4), generator=g)
        ok = True
        for dtype in (torch.float32, torch.uint8):
            ref = (full * 255).to(torch.uint8) if dtype == torch.uint8 else full
            bg = BlockGather(None, dst, nf)
            for done in (0, 3, 6):
                take = min(3, nf - done)
                bg.submit(ref[rank * n:(rank + 1) * n, done:done + take].contiguous(), done)
            res = bg.result()
            if dst is None or rank == dst:
                ok = ok and res is not None and res.dtype == dtype and torch.equal(res, ref)
            else:
                ok = ok and res is None
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dst", [0, 1, None])
def test_block_gather_assembles_rank_major_batch(dst):
    """distributed.BlockGather (the per-block exchange of NPVPInference.rollout): gather to one rank or to all, any payload type."""
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_block_worker, args=(world, port, dst, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)


def test_shard_bounds_cover_batch():
    for n in (1, 7, 8, 64, 513):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
