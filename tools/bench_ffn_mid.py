"""Conv-FFN middle: fused single-pass kernel (npvp_ffn_mid) vs the two-kernel path (npvp_ffn_dwconv + npvp_ffn_norm2).
CUDA-event timing, inputs larger than L2 (640 frames = 168 MB per bf16 buffer).  python tools/bench_ffn_mid.py [frames]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from npvp_b200 import _lib  # noqa: E402
from npvp_b200._lib import FFN_CHUNK  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 640
    Ch, dev = 2048, "cuda"
    op = _lib.ops()
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g, device=dev)
    h = r(frames * 64, Ch).to(torch.bfloat16)
    n1w, n1b, n2w, n2b = r(64, Ch) * 0.3 + 1, r(64, Ch) * 0.3, r(64, Ch) * 0.3 + 1, r(64, Ch) * 0.3
    dw_w, dw_b = r(9, Ch) * 0.4, r(Ch) * 0.2
    st = torch.empty(frames, 2, device=dev)
    op.ffn_frame_stats(h, st)
    y, out, out2 = torch.empty_like(h), torch.empty_like(h), torch.empty_like(h)
    pt = torch.empty(frames, Ch // FFN_CHUNK, 2, device=dev)
    print("clusters:", op.ffn_mid_clusters())

    quick = os.environ.get("NPVP_BENCH_QUICK") == "1"          # one launch per kernel (for ncu captures)

    def timeit(fn, reps=20):
        reps = 1 if quick else reps
        for _ in range(0 if quick else 3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3

    op.lib.npvp_set_option(b"ffn_scalar", 1)
    t_dw0 = timeit(lambda: op.ffn_dwconv(h, st, n1w, n1b, dw_w, dw_b, y, pt))
    t_n20 = timeit(lambda: op.ffn_norm2(y, pt, n2w, n2b, out2))
    ref_y, ref_o = y.clone(), out2.clone()
    op.lib.npvp_set_option(b"ffn_scalar", 0)
    t_dw = timeit(lambda: op.ffn_dwconv(h, st, n1w, n1b, dw_w, dw_b, y, pt))
    t_n2 = timeit(lambda: op.ffn_norm2(y, pt, n2w, n2b, out2))
    print(f"scalar kernels: dwconv {t_dw0:.1f} us + norm2 {t_n20:.1f} us = {t_dw0 + t_n20:.1f} us;  packed vs scalar max |d|: "
          f"y {float((y.float() - ref_y.float()).abs().max()):.3e}, out {float((out2.float() - ref_o.float()).abs().max()):.3e}")
    t_c = timeit(lambda: op.ffn_mid(h, st, n1w, n1b, dw_w, dw_b, n2w, n2b, out))
    t_f = timeit(lambda: op.ffn_mid(h, st, n1w, n1b, dw_w, dw_b, n2w, n2b, out, xch=pt))
    print(f"fused, 16-block clusters (DSMEM exchange): {t_c:.1f} us;  lanes {op.ffn_mid_lanes()}")
    alg = 2 * frames * 64 * Ch * 2                     # bf16 in + bf16 out
    d = (out.float() - out2.float()).abs()
    print(f"frames {frames}: dwconv {t_dw:.1f} us + norm2 {t_n2:.1f} us = {t_dw + t_n2:.1f} us;  fused {t_f:.1f} us "
          f"({alg / t_f * 1e-3:.0f} GB/s algorithmic);  fused vs split max |d| {float(d.max()):.3e}")


if __name__ == "__main__":
    main()
