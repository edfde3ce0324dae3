"""Experiment: does running two half-batches on two streams (independent clips; tails of one lane's kernels filled by the other
lane) beat one full batch?  python tools/exp_two_lanes.py [clips]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from npvp_b200.pipeline import build_from_config  # noqa: E402

PRESET, N_FUTURE = "Cityscapes_VFP_NPVP-S", 28


def main():
    clips = int(sys.argv[1]) if len(sys.argv) > 1 else 74
    torch.set_grad_enabled(False)
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    lanes = [build_from_config(PRESET, device=dev, seed=0) for _ in range(2)]
    for m in lanes:
        m.use_cuda_graphs(True)
    full = lanes[0]
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(clips, 2, 3, 128, 128, generator=g) * 2 - 1).to(dev)
    halves = [x[: clips // 2].contiguous(), x[clips // 2:].contiguous()]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def one():
        full.rollout(x, N_FUTURE, last_block="query")

    def two():
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        for m, h, s in zip(lanes, halves, streams):
            with torch.cuda.stream(s):
                m.rollout(h, N_FUTURE, last_block="query")
        for s in streams:
            cur.wait_stream(s)

    for name, fn in (("one lane ", one), ("two lanes", two), ("one lane ", one), ("two lanes", two)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        t = ts[len(ts) // 2]
        print(f"{name}: {t:.2f} ms per {clips} clips -> {clips * N_FUTURE / t * 1e3:.0f} frames/s", flush=True)


if __name__ == "__main__":
    main()
