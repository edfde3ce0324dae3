"""Throughput of every BASELINE.json configuration + the KITTI batch sweep (config 5), one GPU.
usage: python tools/bench_configs.py [--cpu]   ->  markdown tables on stdout"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_grad_enabled(False)
from npvp_b200.pipeline import build_from_config  # noqa: E402


def gpu_fps(model, x, n_future, iters=5, rollout=False):
    fn = (lambda: model.rollout(x, n_future, last_block="query")) if rollout else (lambda: model.predict(x))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return x.shape[0] * n_future * iters / (e0.elapsed_time(e1) * 1e-3)


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
    do_cpu = "--cpu" in sys.argv
    print("| config | clips | frames/s (B200, CUDA graphs) | CPU oracle frames/s |\n|---|---:|---:|---:|")
    rows = [("SMMNIST_VFP_NPVP-D", 8, 10, False), ("SMMNIST_VFP_NPVP-D", 64, 10, False), ("SMMNIST_VFP_NPVP-D_10to10", 8, 10, False), ("KTH_Unified_NPVP-S", 8, 10, False), ("KTH_Unified_NPVP-S", 64, 10, False),
            ("BAIR_VFP_NPVP-S", 64, 28, True), ("Cityscapes_VFP_NPVP-D", 64, 28, True), ("Cityscapes_VFP_NPVP-S", 64, 28, True)]
    for preset, n, nf, roll in rows:
        model = build_from_config(preset, device="cuda", seed=0).use_cuda_graphs(True)
        c = model.cfg
        x = torch.rand(n, c.Dataset.num_past_frames, c.Dataset.img_channels, c.Dataset.img_size, c.Dataset.img_size, device="cuda") * 2 - 1
        fps = gpu_fps(model, x, nf, rollout=roll)
        cpu = f"{cpu_fps(preset, min(n, 8), c.Dataset.num_past_frames, c.Dataset.num_future_frames):.1f}" if do_cpu and n == 8 else ""
        print(f"| {preset} ({c.Dataset.num_past_frames}->{nf}{' block-AR' if roll else ''}) | {n} | {fps:,.0f} | {cpu} |", flush=True)
        del model
        torch.cuda.empty_cache()
    # BASELINE config 3: 8 stochastic samples per clip (frame encoder, EVT_Former and prior run once per clip)
    model = build_from_config("BAIR_VFP_NPVP-S", device="cuda", seed=0)
    x = torch.rand(16, 2, 3, 64, 64, device="cuda") * 2 - 1
    for _ in range(2):
        model.predict_samples(x, 8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        model.predict_samples(x, 8)
    e1.record()
    torch.cuda.synchronize()
    print(f"| BAIR_VFP_NPVP-S (2->10, 8 samples per clip, predict_samples, eager) | 16 x 8 | {16 * 8 * 10 * 5 / (e0.elapsed_time(e1) * 1e-3):,.0f} | |", flush=True)
    del model
    torch.cuda.empty_cache()
    print("\nKITTI VFP NPVP-S 4->5, 128x128 RGB, batch sweep (one B200):\n\n| clips | frames/s | ms / forward |\n|---:|---:|---:|")
    model = build_from_config("KITTI_VFP_NPVP-S", device="cuda", seed=0).use_cuda_graphs(True)
    for n in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512):
        x = torch.rand(n, 4, 3, 128, 128, device="cuda") * 2 - 1
        fps = gpu_fps(model, x, 5, iters=5 if n >= 64 else 20)
        print(f"| {n} | {fps:,.0f} | {1e3 * n * 5 / fps:.2f} |", flush=True)
    if do_cpu:
        print(f"\nCPU oracle on this host ({os.cpu_count()} threads), KITTI 4->5: N=1 {cpu_fps('KITTI_VFP_NPVP-S', 1, 4, 5):.1f} frames/s, N=8 {cpu_fps('KITTI_VFP_NPVP-S', 8, 4, 5):.1f} frames/s")


if __name__ == "__main__":
    main()
