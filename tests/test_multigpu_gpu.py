"""GPU, >= 2 devices: batch-sharded prediction + NCCL all-gather equals the single-GPU result bit for bit."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        torch.set_grad_enabled(False)
        from npvp_b200.distributed import gather_frames, shard_bounds
        from npvp_b200.pipeline import build_from_config
        model = build_from_config("KITTI_VFP_NPVP-S", device=f"cuda:{rank}", seed=0)
        g = torch.Generator().manual_seed(3)
        n = 5                                                     # uneven shards on purpose
        x = (torch.rand((n, 4, 3, 128, 128), generator=g) * 2 - 1).to(rank)
        eps = torch.randn((n, 512, 8, 8), generator=g).to(rank)
        lo, hi = shard_bounds(n, rank, world)
        local = model.predict(x[lo:hi], eps[lo:hi])
        full = gather_frames(local, n)
        if rank == 0:
            single = model.predict(x, eps)
            ret["equal"] = bool(torch.equal(full, single))
            ret["shape"] = tuple(full.shape)
        # block-autoregressive rollout with the per-block overlapped all-gather == rollout + one gather at the end
        x4 = (torch.rand((4, 4, 3, 128, 128), generator=g) * 2 - 1).to(rank)
        eps4 = [torch.randn((4, 512, 8, 8), generator=g).to(rank) for _ in range(2)]
        l4, h4 = shard_bounds(4, rank, world)
        el = [e[l4:h4] for e in eps4]
        local4 = model.rollout(x4[l4:h4], 7, eps_list=el)
        b = gather_frames(local4, 4)
        a = model.rollout(x4[l4:h4], 7, eps_list=el, gather_group=True)                      # default: gathered on rank 0 only
        torch.cuda.synchronize()
        if rank == 0:
            ret["rollout_equal"] = bool(torch.equal(a, b)) and tuple(a.shape) == (4, 7, 3, 128, 128)
        else:
            ret["sender_keeps_shard"] = bool(torch.equal(a, local4))
        every = model.rollout(x4[l4:h4], 7, eps_list=el, gather_group=True, gather_dst=None)  # all-gather: every rank
        u8 = model.rollout(x4[l4:h4], 7, eps_list=el, gather_group=True, gather_dtype=torch.uint8)
        f16 = model.rollout(x4[l4:h4], 7, eps_list=el, gather_group=True, gather_dst=1, gather_dtype=torch.float16)
        torch.cuda.synchronize()
        ret[f"all_equal_{rank}"] = bool(torch.equal(every, b))
        if rank == 0:
            ret["u8_equal"] = u8.dtype == torch.uint8 and bool(torch.equal(u8, model.to_pixels(b, uint8=True)))
        if rank == 1:
            ret["f16_equal"] = f16.dtype == torch.float16 and bool(torch.equal(f16, b.to(torch.float16)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_prediction_bitwise_equal_to_single_gpu():
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret["shape"] == (5, 5, 3, 128, 128)
    assert ret["equal"], "gathered multi-GPU frames differ from the single-GPU batch"
    assert ret["rollout_equal"], "per-block overlapped gather differs from rollout + gather_frames"
    assert ret["sender_keeps_shard"] and ret["all_equal_0"] and ret["all_equal_1"], dict(ret)
    assert ret["u8_equal"] and ret["f16_equal"], dict(ret)
