"""Throughput of every BASELINE.json configuration on 1 / 2 / 4 / 8 GPUs (VERDICT r01 row g), one command per GPU count:

    python tools/bench_configs.py [--cpu] [--out profiles/r02_configs_1gpu]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_configs.py --out profiles/r02_configs_Ngpu

Weak scaling like bench.py: every rank runs `clips` clips of the config (batch-sharded, weights replicated), the predicted frames are
gathered on rank 0 as uint8 pixel frames (the path's one exchange), CUDA-event timed per step, L2 flushed between steps, max over
ranks.  Rank 0 writes <out>.json (one record per row) and <out>.md.  `--cpu`: rank 0 also times the CPU oracle (the reference's
path on the host cores) on 8 clips of each config; `--faithful`: LitPredictor.forward-faithful call next to the throughput path.
Synthetic clips, random-init weights.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from npvp_b200.pipeline import build_from_config  # noqa: E402

torch.set_grad_enabled(False)

# (label, preset, clips per GPU, mode, frames predicted per clip)
#   predict  : one forward of the YAML's (to, tp)
#   rollout  : block-autoregressive 2 -> 28 (10 + 10 + 8, the third block queries its 8 timestamps): configs 3 / 4 as the YAMLs allow
#   oneshot  : the non-YAML max_T = 30 variant, 28 target timestamps in one call
#   samples8 : config 3's "8 stochastic samples per clip" through predict_samples (encoder / EVT_Former / prior once per clip)
ROWS = [
    ("C1 SMMNIST NPVP-D 5->10 (YAML)", "SMMNIST_VFP_NPVP-D", 8, "predict", 10),
    ("C1 SMMNIST NPVP-D 5->10 (YAML)", "SMMNIST_VFP_NPVP-D", 64, "predict", 10),
    ("C1b SMMNIST NPVP-D 10->10 (BASELINE.json wording)", "SMMNIST_VFP_NPVP-D_10to10", 64, "predict", 10),
    ("C2 KTH unified NPVP-S VFP 10->10", "KTH_Unified_NPVP-S", 8, "predict", 10),
    ("C2 KTH unified NPVP-S VFP 10->10", "KTH_Unified_NPVP-S", 64, "predict", 10),
    ("C3 BAIR NPVP-S 2->28 block-AR", "BAIR_VFP_NPVP-S", 64, "rollout", 28),
    ("C3 BAIR NPVP-S 2->10, 8 samples per clip", "BAIR_VFP_NPVP-S", 16, "samples8", 80),
    ("C3b BAIR NPVP-S 2->28 one-shot (max_T 30)", "BAIR_VFP_NPVP-S_oneshot28", 32, "predict", 28),
    ("C4 Cityscapes NPVP-D 2->28 block-AR", "Cityscapes_VFP_NPVP-D", 64, "rollout", 28),
    ("C4b Cityscapes NPVP-D 2->28 one-shot (max_T 30)", "Cityscapes_VFP_NPVP-D_oneshot28", 32, "predict", 28),
    ("headline Cityscapes NPVP-S 2->28 block-AR", "Cityscapes_VFP_NPVP-S", 64, "rollout", 28),
    ("C5 KITTI NPVP-S 4->5", "KITTI_VFP_NPVP-S", 1, "predict", 5),
    ("C5 KITTI NPVP-S 4->5", "KITTI_VFP_NPVP-S", 8, "predict", 5),
    ("C5 KITTI NPVP-S 4->5", "KITTI_VFP_NPVP-S", 64, "predict", 5),
    ("C5 KITTI NPVP-S 4->5", "KITTI_VFP_NPVP-S", 512, "predict", 5),
]
KITTI_SWEEP = (1, 2, 4, 8, 16, 32, 64, 128, 256, 512)


def dist_env():
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    return world, rank, local


def timed_fps(step, frames_per_step, world, dev, flush, iters):
    import torch.distributed as dist
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    evs = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return frames_per_step * world * iters / (float(ms.item()) * 1e-3), float(ms.item()) / iters


def make_step(model, mode, x, world):
    """The timed call: prediction on this rank's clips + (N > 1) the gather of the uint8 pixel frames on rank 0."""
    from npvp_b200.distributed import BlockGather
    if mode == "rollout":
        return lambda: model.rollout(x, 28, last_block="query", gather_group=True if world > 1 else None, gather_dst=0, gather_dtype=torch.uint8)

    def gathered(frames_u8):
        if world > 1:
            bg = BlockGather(None, 0, frames_u8.shape[1])
            bg.submit(frames_u8.clone(), 0)
            return bg.result()
        return frames_u8
    if mode == "samples8":
        def step():
            smp = model.predict_samples(x, 8)
            return gathered(model.to_pixels(smp.flatten(0, 1), uint8=True)) if world > 1 else smp
        return step
    if world > 1:
        return lambda: gathered(model.predict(x, pixels_u8=True)[1])
    return lambda: model.predict(x)


def cpu_fps(preset, n, to, tp_n):
    """CPU baseline of one forward of a config: bench.py's cpu_baseline leg (the only code outside tests/ that runs oracle/)."""
    import bench
    return bench.cpu_baseline_forward(preset, n, to, tp_n)


def faithful():
    """the LitPredictor.forward-faithful call (also decodes the context reconstructions, Predictor.py:72-86) next to
    the throughput path, one 2 -> 10 block of the headline config, eager launches for both."""
    model = build_from_config("Cityscapes_VFP_NPVP-S", device="cuda", seed=0)
    x = torch.rand(64, 2, 3, 128, 128, device="cuda") * 2 - 1

    def run(fn, iters=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters
    t_pred, t_fwd = run(lambda: model.predict(x)), run(lambda: model(x))
    print("| call (Cityscapes NPVP-S, 64 clips, 2 -> 10, eager) | ms | predicted frames/s |\n|---|---:|---:|")
    print(f"| `model.predict(past)` (Enc(context) -> Predictor -> Dec(predictions), channels-last hand-offs) | {t_pred:.2f} | {640 / t_pred * 1e3:,.0f} |")
    print(f"| `model(past)` = LitPredictor.forward triple (+ Dec(context) reconstructions, NCHW fp32 between the modules) | {t_fwd:.2f} | {640 / t_fwd * 1e3:,.0f} |")


def main():
    if "--faithful" in sys.argv:
        return faithful()
    import torch.distributed as dist
    world, rank, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    do_cpu = "--cpu" in sys.argv and rank == 0
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    quick = "--quick" in sys.argv
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    records, cpu_cache, model, model_name = [], {}, None, None
    rows = list(ROWS)
    if world == 1 and not quick:                                   # single GPU: the full KITTI batch sweep
        rows = [r for r in rows if not r[0].startswith("C5")] + [("C5 KITTI NPVP-S 4->5", "KITTI_VFP_NPVP-S", n, "predict", 5) for n in KITTI_SWEEP]
    for label, preset, clips, mode, nf in rows:
        if model_name != preset:
            del model
            torch.cuda.empty_cache()
            model = build_from_config(preset, device=dev, seed=0).use_cuda_graphs(True)
            model_name = preset
        c = model.cfg
        if c.Predictor.rand_context:
            model.predictor.reset_pos_coor(model.to_list, model.tp_list)
        g = torch.Generator().manual_seed(1234 + rank)
        x = torch.rand((clips, c.Dataset.num_past_frames, c.Dataset.img_channels, c.Dataset.img_size, c.Dataset.img_size), generator=g)
        x = (x * 2 - 1 if c.AE.out_layer == "Tanh" else x).to(dev)
        iters = 5 if clips * nf >= 640 else 20
        fps, ms = timed_fps(make_step(model, mode, x, world), clips * nf, world, dev, flush, iters)
        rec = {"config": label, "preset": preset, "mode": mode, "n_gpus": world, "clips_per_gpu": clips, "global_clips": clips * world,
               "frames_per_clip": nf, "frames_per_s": fps, "ms_per_step": ms, "cuda_graphs": True,
               "exchange": "uint8 pixel frames gathered on rank 0 (NCCL)" if world > 1 else None}
        if do_cpu and mode == "predict":
            key = (preset, min(clips, 8))
            if key not in cpu_cache:
                cpu_cache[key] = cpu_fps(preset, min(clips, 8), c.Dataset.num_past_frames, c.Dataset.num_future_frames)
            rec["cpu_oracle_frames_per_s"], rec["cpu_clips"], rec["cpu_cores"] = cpu_cache[key], key[1], os.cpu_count()
        records.append(rec)
        if rank == 0:
            print(json.dumps(rec), flush=True)
    if rank == 0 and out:
        os.makedirs(os.path.dirname(os.path.abspath(out)) or ".", exist_ok=True)
        with open(out + ".json", "w") as f:
            json.dump(records, f, indent=1)
        with open(out + ".md", "w") as f:
            f.write(f"# Per-config throughput, {world} x B200 (tools/bench_configs.py; weak scaling, CUDA graphs, CUDA events, L2 flushed between steps, max over ranks)\n\n")
            f.write("| config | clips / GPU | global clips | predicted frames/s | ms / step | CPU oracle frames/s (clips, cores) |\n|---|---:|---:|---:|---:|---|\n")
            for r in records:
                cpu = f"{r['cpu_oracle_frames_per_s']:.1f} ({r['cpu_clips']}, {r['cpu_cores']})" if "cpu_oracle_frames_per_s" in r else ""
                f.write(f"| {r['config']} [{r['mode']}] | {r['clips_per_gpu']} | {r['global_clips']} | {r['frames_per_s']:,.0f} | {r['ms_per_step']:.2f} | {cpu} |\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
