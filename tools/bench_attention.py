"""Attention cores at the bench shapes (64 clips): CUDA-event time and achieved GB/s.  python tools/bench_attention.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from npvp_b200 import _lib  # noqa: E402


def main():
    op, dev = _lib.ops(), "cuda"
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, mode, Tq, Tk in (("temporal self 10x10", 1, 10, 10), ("temporal cross 10x2", 1, 10, 2), ("spatial windows T=10", 0, 10, 10),
                               ("temporal self 2x2", 1, 2, 2), ("temporal self 28x28", 1, 28, 28)):
        Mq, Mk = n * Tq * 64, n * Tk * 64
        qk = torch.randn(Mq, 1024, generator=g, device=dev).to(torch.bfloat16)
        k = qk[:, 512:] if Tq == Tk else torch.randn(Mk, 512, generator=g, device=dev).to(torch.bfloat16)
        v = torch.randn(Mk, 512, generator=g, device=dev).to(torch.bfloat16)
        o = torch.empty(Mq, 512, device=dev, dtype=torch.bfloat16)
        ts = []
        for i in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            op.attention(qk[:, :512], k, v, o, mode, n, Tq, Tk, False)
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(ts)[len(ts) // 2]
        byts = (2 * Mq + 2 * Mk) * 512 * 2
        print(f"{name:24s} {t:7.1f} us   {byts / t * 1e-3:6.0f} GB/s algorithmic (q + k + v + out = {byts / 1e6:.0f} MB)")


if __name__ == "__main__":
    main()
