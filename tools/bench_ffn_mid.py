"""Conv-FFN middle: the single-pass half-precision kernel (npvp_ffn_mid16) vs the r01 fp32 two-kernel path (npvp_ffn_stats_finalize +
npvp_ffn_dwconv + npvp_ffn_norm2).  CUDA-event timing per launch, 256 MiB L2 flush between launches, inputs larger than L2
(640 frames = 168 MB per 16-bit buffer).  Algorithmic bytes of the middle: one 16-bit frame in + one out = 0.524 MB per frame.

    python tools/bench_ffn_mid.py [frames]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from npvp_b200 import _lib  # noqa: E402

DEV = "cuda"


def rn(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def timeit(fn, flush, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 640
    Ch = 2048
    op = _lib.ops()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    a = rn(frames * 64, 512, seed=11, dtype=torch.bfloat16)
    w1, b1 = rn(Ch, 512, seed=12, scale=0.06, dtype=torch.bfloat16), rn(Ch, seed=13, scale=0.3)
    n1w, n1b = rn(64, Ch, seed=2) * 0.3 + 1, rn(64, Ch, seed=3) * 0.3
    n2w, n2b = rn(64, Ch, seed=4) * 0.3 + 1, rn(64, Ch, seed=5) * 0.3
    dw_w, dw_b = rn(9, Ch, seed=6, scale=0.4), rn(Ch, seed=7, scale=0.2)
    part1 = torch.empty(frames, 32, 2, device=DEV)
    # half path
    h1 = torch.empty(frames * 64, Ch, dtype=torch.float16, device=DEV)
    op.gemm(a, w1, bias=b1, out_bf16=h1, frame_stats=part1)
    pair = lambda w, b: torch.stack([w.view(64, Ch // 2, 2), b.view(64, Ch // 2, 2)], dim=2)
    ln_wb = torch.stack([pair(n1w, n1b), pair(n2w, n2b)], 0).to(torch.float16).contiguous()
    xch, cnt = _lib.ffn_mid16_scratch(frames, DEV)
    out16 = torch.empty_like(h1)
    t_half = timeit(lambda: op.ffn_mid16(h1, part1, ln_wb, dw_w.half(), dw_b.half(), out16, xch, cnt), flush)
    # fp32 two-kernel path
    h1b = torch.empty(frames * 64, Ch, dtype=torch.bfloat16, device=DEV)
    op.gemm(a, w1, bias=b1, out_bf16=h1b, frame_stats=part1)
    st1, pt2 = torch.empty(frames, 2, device=DEV), torch.empty(frames, Ch // _lib.FFN_CHUNK, 2, device=DEV)
    y2, outb = torch.empty_like(h1b), torch.empty_like(h1b)
    t_fin = timeit(lambda: op.ffn_stats_finalize(part1, st1, 64 * Ch), flush)
    t_dw = timeit(lambda: op.ffn_dwconv(h1b, st1, n1w, n1b, dw_w, dw_b, y2, pt2), flush)
    t_n2 = timeit(lambda: op.ffn_norm2(y2, pt2, n2w, n2b, outb), flush)
    alg = frames * 64 * Ch * 2 * 2
    d = (out16.float() - outb.float()).abs()
    print(f"frames {frames}: algorithmic bytes {alg / 1e6:.1f} MB")
    print(f"| path | us | GB/s algorithmic | frac of 6457 GB/s |\n|---|---:|---:|---:|")
    print(f"| ffn_mid16 (one pass, half2) | {t_half:.1f} | {alg / t_half * 1e-3:.0f} | {alg / t_half * 1e-3 / 6457.4:.3f} |")
    t_old = t_fin + t_dw + t_n2
    print(f"| finalize {t_fin:.1f} + dwconv2 {t_dw:.1f} + norm2 {t_n2:.1f} (fp32, two passes) | {t_old:.1f} | {alg / t_old * 1e-3:.0f} | {alg / t_old * 1e-3 / 6457.4:.3f} |")
    print(f"half vs fp32 path outputs: max abs diff {float(d.max()):.3e}, mean abs {float(d.mean()):.3e}, max |out| {float(outb.float().abs().max()):.2f}")


if __name__ == "__main__":
    main()
