"""Micro-benchmark of the GEMM back-ends on the predictor's shapes (run on the GPU box)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from npvp_b200 import _lib

def main():
    op = _lib.Ops()
    dev = "cuda"
    M = int(os.environ.get("M", 40960))
    shapes = [(M, 2048, 512, "bf16"), (M, 512, 2048, "bf16"), (M, 512, 512, "bf16"), (M, 1024, 512, "bf16"), (M, 1024, 512, "gelu"),
              (M, 512, 1024, "bf16"), (M // 5, 512, 512, "bf16"), (M // 5, 2048, 512, "bf16"), (8192, 256, 4608, "bf16"),
              (M * 8, 128, 256, "bf16"), (M * 8, 64, 576, "bf16"), (M, 512, 512, "res"), (M, 512, 2048, "f32")]
    if os.environ.get("SHAPES"):
        shapes = [shapes[int(i)] for i in os.environ["SHAPES"].split(",")]
    backends = [b for b in (("v2", 1), ("2cta", 4)) if os.environ.get("ONLY", b[0]) == b[0]]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for (m, n, k, kind) in shapes:
        a = torch.randn(m, k, device=dev).to(torch.bfloat16)
        w = (torch.randn(n, k, device=dev) * k ** -0.5).to(torch.bfloat16)
        bias = torch.randn(n, device=dev)
        res = torch.randn(m, n, device=dev)
        of = torch.empty(m, n, device=dev)
        ob = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
        kw = dict(bias=bias)
        if os.environ.get("POST_RELU"): kw["post_relu"] = int(os.environ["POST_RELU"])
        if kind == "res": kw.update(res1=res, out_f32=of)
        elif kind == "f32": kw.update(out_f32=of)
        elif kind == "gelu": kw.update(act=2, out_bf16=ob)
        else: kw.update(out_bf16=ob)
        line = f"M={m:7d} N={n:5d} K={k:5d} {kind:5s}"
        ref = None
        for name, be in backends:
            for _ in range(3):
                op.gemm(a, w, backend=be, **kw)
            ts = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); op.gemm(a, w, backend=be, **kw); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = sorted(ts)[len(ts) // 2]
            out = (of if "out_f32" in kw else ob).float().clone()
            if ref is None: ref = out
            line += f" | {name}: {t*1e3:8.1f} us {2*m*n*k/t/1e9:7.1f} TFLOP/s"
            if name != backends[0][0]: line += f" (maxdiff vs {backends[0][0]} {float((out-ref).abs().max()):.1e})"
        if os.environ.get("CUBLAS", "1") == "1":          # library yardstick on the same shape (never used by the product)
            import torch.nn.functional as F
            bb = bias.to(torch.bfloat16)
            for _ in range(3):
                F.linear(a, w, bb)
            ts = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); F.linear(a, w, bb); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = sorted(ts)[len(ts) // 2]
            line += f" | cuBLASLt: {t*1e3:8.1f} us {2*m*n*k/t/1e9:7.1f} TFLOP/s"
        print(line, flush=True)

if __name__ == "__main__":
    main()
