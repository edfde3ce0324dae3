"""7x7 stem convolution: tcgen05 row-streaming kernel (head_tc.cu) against the mma.sync tile kernel, same inputs.
python tools/bench_stem.py            # 148 frames of 128 x 128 x 3 -> 32 channels (the bench step's encoder input)
STEM_SHAPE=Cin,Cout,H,W STEM_FRAMES=n python tools/bench_stem.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from npvp_b200 import _lib  # noqa: E402


def main():
    op = _lib.ops()
    frames = int(os.environ.get("STEM_FRAMES", "148"))
    Cin, Cout, H, W = (int(v) for v in os.environ.get("STEM_SHAPE", "3,32,128,128").split(","))
    x = torch.randn(frames, Cin, H, W, device="cuda")
    w, sh = torch.randn(49 * Cin, Cout, device="cuda") * 0.1, torch.randn(Cout, device="cuda") * 0.2
    out = torch.empty(frames * H * W, Cout, device="cuda", dtype=torch.float16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for tc in (0, 1, 0, 1):
        op.lib.npvp_set_option(b"stem_tc", tc)
        ts = []
        for _ in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            op.conv7x7_stem(x, w, sh, out, Cin, Cout, H, W)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(ts[1:])[2]
        gb = (x.numel() * 4 + out.numel() * 2) / 1e9
        print(f"stem_tc={tc}: {t:.1f} us per {frames} frames of {H}x{W}, Cin {Cin} Cout {Cout}  ({gb / t * 1e6:.0f} GB/s algorithmic)", flush=True)
    op.lib.npvp_set_option(b"stem_tc", 1)


if __name__ == "__main__":
    main()
