"""7x7 head convolution: tcgen05 row-streaming kernel (head_tc.cu) against the mma.sync tile kernel, same inputs.

python tools/bench_head.py            # equality over shapes + timing at the Cityscapes decoder shape (640 frames, 128 x 128, Cin 32)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from npvp_b200 import _lib  # noqa: E402
from npvp_b200._lib import pack_head_weights  # noqa: E402

DEV = "cuda"


def run(op, x, w, b, frames, Cin, Cout, H, W, act, tc, u8=False):
    op.lib.npvp_set_option(b"head_tc", int(tc))
    out = torch.empty(frames, Cout, H, W, device=DEV)
    o8 = torch.empty(frames, Cout, H, W, device=DEV, dtype=torch.uint8) if u8 else None
    ren = ([0.5] * Cout, [0.5] * Cout) if u8 else None
    op.conv7x7_head(x, w, b, out, Cin, Cout, H, W, False, act, out_u8=o8, renorm=ren)
    torch.cuda.synchronize()
    return out, o8


def main():
    op = _lib.ops()
    if "--time-only" in sys.argv:
        return timing(op)
    g = torch.Generator(device="cpu").manual_seed(0)
    worst = 0.0
    for dt in (torch.float16, torch.bfloat16):
        for frames, Cin, Cout, H, W, act in [(2, 32, 3, 48, 48, 3), (2, 32, 2, 36, 36, 0), (2, 64, 1, 76, 76, 4), (3, 32, 3, 128, 128, 3),
                                             (5, 64, 3, 64, 64, 3), (40, 32, 3, 128, 128, 3), (7, 64, 1, 64, 64, 4), (3, 32, 3, 20, 36, 3),
                                             (300, 64, 1, 64, 64, 4), (2, 32, 3, 128, 160, 3), (1, 64, 3, 8, 8, 3)]:
            x = (torch.randn(frames * H * W, Cin, generator=g)).to(DEV).to(dt)
            w = pack_head_weights((torch.randn(49 * Cin, Cout, generator=g) * 0.03).to(DEV), dt)
            b = (torch.randn(Cout, generator=g) * 0.2).to(DEV)
            o_ref, u_ref = run(op, x, w, b, frames, Cin, Cout, H, W, act, False, u8=True)
            o_tc, u_tc = run(op, x, w, b, frames, Cin, Cout, H, W, act, True, u8=True)
            err = float((o_ref - o_tc).abs().max())
            du8 = int((u_ref.int() - u_tc.int()).abs().max())
            worst = max(worst, err)
            print(f"{str(dt)[6:]:9s} frames {frames:4d} Cin {Cin} Cout {Cout} {H}x{W} act {act}: max |tc - mma| {err:.2e}  u8 diff {du8}", flush=True)
    print("worst", worst)
    timing(op)


def timing(op):
    # timing, L2 flushed between launches
    frames = int(os.environ.get('HEAD_FRAMES', '640'))
    Cin, Cout, H, W = (int(v) for v in os.environ.get('HEAD_SHAPE', '32,3,128,128').split(','))
    x = torch.randn(frames * H * W, Cin, device=DEV).half()
    w = pack_head_weights(torch.randn(49 * Cin, Cout, device=DEV) * 0.03, torch.float16)
    b = torch.randn(Cout, device=DEV) * 0.2
    out = torch.empty(frames, Cout, H, W, device=DEV)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for tc in (0, 1, 0, 1):
        op.lib.npvp_set_option(b"head_tc", tc)
        wm = int(os.environ.get('HEAD_WAIT', '1'))
        op.lib.npvp_set_option(b"head_tc_wait", wm)
        ts = []
        for it in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            op.conv7x7_head(x, w, b, out, Cin, Cout, H, W, False, 3)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        t = sorted(ts[1:])[len(ts[1:]) // 2]
        gb = (x.numel() * 2 + out.numel() * 4) / 1e9
        print(f"head_tc={tc} wait_mode={wm}: {t:.1f} us per {frames} frames of {H}x{W}, Cin {Cin} Cout {Cout}  ({gb / t * 1e6:.0f} GB/s algorithmic)")
    op.lib.npvp_set_option(b"head_tc", 1)


if __name__ == "__main__":
    main()
