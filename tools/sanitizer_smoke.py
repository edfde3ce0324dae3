"""One small launch of every hand-written kernel family, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitizer_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitizer_smoke.py
    compute-sanitizer --tool synccheck python tools/sanitizer_smoke.py

Shapes are small (the tools slow kernels down 10-100x) but cover: ragged M / N / K tails of the tcgen05 GEMMs (1-CTA persistent,
2-CTA cta_group::2, implicit-GEMM conv gather, frame-statistics epilogue), the 4-block-cluster frame kernels (DSMEM st.async +
mbarrier), the single-pass conv-FFN middle (cp.async tiles, L2 statistics exchange, several frames per stream), the attention
cores, and the autoencoder's stem / head / non-local kernels.  Every result is compared with the torch specification, so a run
that is clean under the sanitizer is also a correct one.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from npvp_b200 import _lib  # noqa: E402
from kernel_specs import SpecOps  # noqa: E402

torch.set_grad_enabled(False)
DEV = "cuda"


def rn(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def close(a, b, rel, what):
    a, b = a.float(), b.float()
    err, ref = float((a - b).abs().max()), float(b.abs().max())
    assert err <= rel * max(ref, 1e-6), f"{what}: {err:.3e} vs {ref:.3e}"
    print(f"ok  {what}: max err {err:.3e} (ref max {ref:.3e})", flush=True)


def main():
    op, spec = _lib.ops(), SpecOps()
    # ---- GEMMs: v2 persistent (BN 256 / 128 / 64), 2-CTA, ragged shapes, residual epilogues
    for M, N, K, backend in [(300, 512, 512, 1), (130, 64, 64, 1), (257, 384, 136, 1), (520, 512, 1024, 4), (384, 256, 2048, 0)]:
        a, w, bias = rn(M, K, seed=1, dtype=torch.bfloat16), rn(N, K, seed=2, scale=K ** -0.5, dtype=torch.bfloat16), rn(N, seed=3)
        res = rn(M, N, seed=4)
        o1, o2 = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
        op.gemm(a, w, bias=bias, res1=res, out_f32=o1, backend=backend)
        spec.gemm(a, w, bias=bias, res1=res, out_f32=o2)
        close(o1, o2, 3e-3, f"gemm M={M} N={N} K={K} backend={backend}")
    # ---- fc1 with half output + frame statistics -> single-pass conv-FFN middle (3 frame streams deep) -> fc2 half operands
    frames, Ch = 75, 2048
    a = rn(frames * 64, 512, seed=11, dtype=torch.bfloat16)
    w1, b1 = rn(Ch, 512, seed=12, scale=0.06, dtype=torch.bfloat16), rn(Ch, seed=13, scale=0.3)
    h1, p1 = torch.empty(frames * 64, Ch, dtype=torch.float16, device=DEV), torch.empty(frames, 32, 2, device=DEV)
    h1r, p1r = torch.empty_like(h1), torch.empty_like(p1)
    op.gemm(a, w1, bias=b1, out_bf16=h1, frame_stats=p1)
    spec.gemm(a, w1, bias=b1, out_bf16=h1r, frame_stats=p1r)
    close(h1, h1r, 2e-3, "fc1 half output + frame statistics")
    pair = lambda w, b: torch.stack([w.view(64, Ch // 2, 2), b.view(64, Ch // 2, 2)], dim=2)
    ln_wb = torch.stack([pair(rn(64, Ch, seed=2) * 0.3 + 1, rn(64, Ch, seed=3) * 0.3), pair(rn(64, Ch, seed=4) * 0.3 + 1, rn(64, Ch, seed=5) * 0.3)], 0)
    ln_wb = ln_wb.to(torch.float16).contiguous()
    dw_w, dw_b = rn(9, Ch, seed=6, scale=0.4).half(), rn(Ch, seed=7, scale=0.2).half()
    xch, cnt = _lib.ffn_mid16_scratch(frames, DEV)
    out, ref = torch.empty_like(h1), torch.empty_like(h1)
    for _ in range(2):
        op.ffn_mid16(h1, p1, ln_wb, dw_w, dw_b, out, xch, cnt)
    spec.ffn_mid16(h1, p1, ln_wb, dw_w, dw_b, ref, None, None)
    close(out, ref, 1e-2, "ffn_mid16 (twice on one exchange scratch)")
    assert bool((xch.view(torch.int32) == -1).all()) and bool((cnt == 0).all())
    # ---- 4-block-cluster frame kernels
    n, T = 2, 3
    x = rn(n * T * 64, 512, seed=21)
    lw, lb, qe = rn(512, seed=22) * 0.2 + 1, rn(512, seed=23) * 0.2, rn(n * 64, 512, seed=24)
    beta, gamma = rn(T * 64, 512, seed=25), rn(T * 64, 512, seed=26) * 0.3
    delta = rn(n * T * 64, 512, seed=27, dtype=torch.bfloat16)
    outs = []
    for o in (op, spec):
        xx = x.clone()
        a1, f1 = torch.empty(n * T * 64, 512, dtype=torch.bfloat16, device=DEV), torch.empty(n * T * 64, 512, dtype=torch.bfloat16, device=DEV)
        o.add_ln_posfuse(xx, delta, lw, lb, qe, beta, gamma, a1, f1, n, T)
        h3 = rn(n * T * 64, 512, seed=28, dtype=torch.bfloat16)
        a2, f2 = torch.empty_like(a1), torch.empty_like(f1)
        o.frame_ln_gelu_residual_posfuse(h3, rn(64, 512, seed=29) * 0.3 + 1, rn(64, 512, seed=30) * 0.3, xx, lw, lb, qe, beta, gamma, a2, f2, n, T)
        outs.append((xx, a1, f1, a2, f2))
    for i, name in enumerate(("stream", "ln", "fused", "tail ln", "tail fused")):
        close(outs[0][i], outs[1][i], 2e-2, f"frame cluster kernels: {name}")
    # ---- attention cores
    for mode, Tq, Tk, mask in [(0, 3, 3, False), (1, 5, 5, True), (1, 7, 2, False)]:
        q, k, v = (rn(2 * t * 64, 512, seed=31 + i, dtype=torch.bfloat16) for i, t in enumerate((Tq, Tk, Tk)))
        o1, o2 = torch.empty_like(q), torch.empty_like(q)
        op.attention(q, k, v, o1, mode, 2, Tq, Tk, mask)
        spec.attention(q, k, v, o2, mode, 2, Tq, Tk, mask)
        close(o1, o2, 2e-2, f"attention mode={mode} Tq={Tq} Tk={Tk} mask={mask}")
    # ---- autoencoder: stem (fp32 and uint8 input), implicit-GEMM conv / transposed conv, non-local attention, head (+ uint8 epilogue)
    dt = torch.float16
    xs = rn(2, 3, 32, 32, seed=41)
    ws, sh = rn(147, 32, seed=42, scale=0.1), rn(32, seed=43, scale=0.1)
    s1, s2 = torch.empty(2 * 32 * 32, 32, dtype=dt, device=DEV), torch.empty(2 * 32 * 32, 32, dtype=dt, device=DEV)
    op.conv7x7_stem(xs, ws, sh, s1, 3, 32, 32, 32)
    spec.conv7x7_stem(xs, ws, sh, s2, 3, 32, 32, 32)
    close(s1, s2, 5e-3, "conv7x7 stem")
    u8 = torch.randint(0, 256, (2, 3, 32, 32), dtype=torch.uint8, device=DEV)
    norm = ((0.3, 0.35, 0.31), (1.2, 1.3, 1.2))
    op.conv7x7_stem(u8, ws, sh, s1, 3, 32, 32, 32, norm=norm)
    spec.conv7x7_stem(u8, ws, sh, s2, 3, 32, 32, 32, norm=norm)
    close(s1, s2, 5e-3, "conv7x7 stem, uint8 ingest")
    op.lib.npvp_set_option(b"stem_tc", 0)            # the calls above ran the tcgen05 row-streaming stem; this one the mma.sync tile kernel
    op.conv7x7_stem(xs, ws, sh, s1, 3, 32, 32, 32)
    op.lib.npvp_set_option(b"stem_tc", 1)
    spec.conv7x7_stem(xs, ws, sh, s2, 3, 32, 32, 32)
    close(s1, s2, 5e-3, "conv7x7 stem (mma.sync tile kernel)")
    wc = rn(64, 9 * 32, seed=44, scale=0.06, dtype=dt)
    c1, c2 = torch.empty(2 * 16 * 16, 64, dtype=dt, device=DEV), torch.empty(2 * 16 * 16, 64, dtype=dt, device=DEV)
    op.conv_gemm(s2, wc, 2, 32, 32, 32, 3, 3, 2, 1, 0, 16, 16, bias=rn(64, seed=45), act=1, out_bf16=c1)
    spec.conv_gemm(s2, wc, 2, 32, 32, 32, 3, 3, 2, 1, 0, 16, 16, bias=rn(64, seed=45), act=1, out_bf16=c2)
    close(c1, c2, 5e-3, "implicit-GEMM 3x3 stride-2 conv")
    wt = rn(4 * 32, 4 * 64, seed=46, scale=0.06, dtype=dt)
    t1, t2 = torch.empty(2 * 16 * 16, 128, dtype=dt, device=DEV), torch.empty(2 * 16 * 16, 128, dtype=dt, device=DEV)
    op.conv_gemm(c2, wt, 2, 16, 16, 64, 2, 2, 1, 0, 0, 16, 16, False, bias=rn(128, seed=47), act=1, out_bf16=t1)
    spec.conv_gemm(c2, wt, 2, 16, 16, 64, 2, 2, 1, 0, 0, 16, 16, False, bias=rn(128, seed=47), act=1, out_bf16=t2)
    close(t1, t2, 5e-3, "transposed conv as a 2x2-neighbourhood GEMM")
    hw = _lib.pack_head_weights(rn(49 * 32, 3, seed=48, scale=0.05), dt)
    hb = rn(3, seed=49, scale=0.1)
    f1, f2 = torch.empty(2, 3, 32, 32, device=DEV), torch.empty(2, 3, 32, 32, device=DEV)
    g1, g2 = torch.empty(2, 3, 32, 32, dtype=torch.uint8, device=DEV), torch.empty(2, 3, 32, 32, dtype=torch.uint8, device=DEV)
    op.conv7x7_head(t2, hw, hb, f1, 32, 3, 32, 32, True, 3, out_u8=g1, renorm=norm)
    spec.conv7x7_head(t2, hw, hb, f2, 32, 3, 32, 32, True, 3, out_u8=g2, renorm=norm)
    close(f1, f2, 5e-3, "conv7x7 head (phase-major input)")
    assert int((g1.int() - g2.int()).abs().max()) <= 1
    # tcgen05 row-streaming head (plain NHWC input): TMA row ring with wrap-around copies, generic-proxy padding writes,
    # two MMA issuers, 12 epilogue warps; 5 frames x 40 rows over 12 CTAs = several frame segments per CTA
    for cin_h in (32, 64):
        t3 = rn(5 * 40 * 48, cin_h, seed=53, dtype=dt)
        hw3 = _lib.pack_head_weights(rn(49 * cin_h, 3, seed=54, scale=0.05), dt)
        f3, f4 = torch.empty(5, 3, 40, 48, device=DEV), torch.empty(5, 3, 40, 48, device=DEV)
        g3, g4 = torch.empty(5, 3, 40, 48, dtype=torch.uint8, device=DEV), torch.empty(5, 3, 40, 48, dtype=torch.uint8, device=DEV)
        op.conv7x7_head(t3, hw3, hb, f3, cin_h, 3, 40, 48, False, 3, out_u8=g3, renorm=norm)
        spec.conv7x7_head(t3, hw3, hb, f4, cin_h, 3, 40, 48, False, 3, out_u8=g4, renorm=norm)
        close(f3, f4, 5e-3, f"conv7x7 head on tcgen05 (NHWC input, Cin {cin_h})")
        assert int((g3.int() - g4.int()).abs().max()) <= 1
    q, kv = rn(2 * 256, 8, seed=51, dtype=dt), rn(2 * 64, 40, seed=52, dtype=dt)
    n1, n2 = torch.empty(2 * 256, 32, dtype=dt, device=DEV), torch.empty(2 * 256, 32, dtype=dt, device=DEV)
    op.nonlocal_attention(q, kv, n1, 2, 256, 64, 8, 32)
    spec.nonlocal_attention(q, kv, n2, 2, 256, 64, 8, 32)
    close(n1, n2, 1e-2, "non-local attention")
    torch.cuda.synchronize()
    print("sanitizer smoke: all kernels ran and matched their specifications")


if __name__ == "__main__":
    main()
