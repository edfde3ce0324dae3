"""GPU: every C-ABI kernel against its torch specification (tests/kernel_specs.py) on seeded inputs."""
import pytest
import torch
from npvp_b200._lib import FFN_CHUNK

from kernel_specs import SpecOps

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda"


@pytest.fixture(scope="module")
def op():
    from npvp_b200 import _lib
    return _lib.Ops()


@pytest.fixture(scope="module")
def spec():
    return SpecOps()


def rn(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def close(a, b, rel, what=""):
    a, b = a.float(), b.float()
    err = float((a - b).abs().max())
    ref = float(b.abs().max())
    assert err <= rel * max(ref, 1e-6), f"{what}: max err {err:.4e} vs ref max {ref:.4e} (rel {err / max(ref, 1e-6):.3e} > {rel})"


GEMM_SHAPES = [(40000, 512, 512), (20480, 2048, 512), (256, 512, 512), (192, 1024, 512), (128, 2048, 512), (1024, 512, 2048), (4096, 256, 4608),
               (130, 64, 64), (1000, 48, 64), (512, 64, 32), (64, 256, 40), (640, 384, 512), (300, 512, 128), (64, 1024, 256), (4096, 128, 288)]


H16 = [torch.bfloat16, torch.float16]


@pytest.mark.parametrize("dt", H16)
@pytest.mark.parametrize("backend", [2, 1, 4, 0])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_plain(op, spec, backend, M, N, K, dt):
    a, w = rn(M, K, seed=1, dtype=dt), rn(N, K, seed=2, scale=K ** -0.5, dtype=dt)
    bias = rn(N, seed=3)
    o1, o2 = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    b1, b2 = torch.empty(M, N, device=DEV, dtype=dt), torch.empty(M, N, device=DEV, dtype=dt)
    op.gemm(a, w, bias=bias, out_f32=o1, out_bf16=b1, backend=backend)
    spec.gemm(a, w, bias=bias, out_f32=o2, out_bf16=b2)
    torch.cuda.synchronize()
    close(o1, o2, 2e-3, f"gemm f32 backend={backend}")
    close(b1, b2, 1e-2 if dt == torch.bfloat16 else 2e-3, f"gemm 16-bit backend={backend}")


@pytest.mark.parametrize("dt", H16)
@pytest.mark.parametrize("backend", [2, 1, 4])
def test_gemm_epilogues(op, spec, backend, dt):
    M, N, K = 384, 512, 1024
    a, w = rn(M, K, seed=1, dtype=dt), rn(N, K, seed=2, scale=K ** -0.5, dtype=dt)
    bias, r1, r2 = rn(N, seed=3), rn(M, N, seed=4), rn(M, N, seed=5, dtype=dt)
    for kw in (dict(act=2), dict(act=1, alpha=0.6, res1=r1, res2=r2), dict(res1=r2, post_relu=True), dict(act=1, res1=r1)):
        o1, o2 = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
        op.gemm(a, w, bias=bias, out_f32=o1, backend=backend, **kw)
        spec.gemm(a, w, bias=bias, out_f32=o2, **kw)
        close(o1, o2, 2e-3, f"epilogue {kw.keys()} backend={backend}")
    # in-place residual stream update (out aliases res1), strided A view
    x = rn(M, N, seed=6)
    x2 = x.clone()
    big = rn(M, 2 * K, seed=7, dtype=dt)
    op.gemm(big[:, K:], w, bias=bias, res1=x, out_f32=x, backend=backend)
    spec.gemm(big[:, K:], w, bias=bias, res1=x2, out_f32=x2)
    close(x, x2, 2e-3, "in-place residual")


@pytest.mark.parametrize("dt", H16)
@pytest.mark.parametrize("backend", [1, 4])
@pytest.mark.parametrize("M,N,K,kind", [(1000, 512, 512, "act0"), (384, 1024, 512, "gelu"), (300, 128, 256, "relu"), (4096, 64, 576, "relu"),
                                        (1000, 512, 512, "res_f32_inplace"), (700, 512, 1024, "f32"), (900, 256, 512, "relu_res16"),
                                        (640, 512, 128, "res16x2"), (520, 512, 512, "both")])
def test_gemm_direct_epilogue_matches_staged(op, M, N, K, kind, dt, backend):
    """The direct (accumulator-layout) epilogue - the default - must be bit-identical to the staged one for every epilogue family."""
    a, w = rn(M, K, seed=1, dtype=dt), rn(N, K, seed=2, scale=K ** -0.5, dtype=dt)
    bias = rn(N, seed=3)
    r16a, r16b, r32 = rn(M, N, seed=4, dtype=dt), rn(M, N, seed=5, dtype=dt), rn(M, N, seed=6)
    outs = []
    try:
        for direct in (0, 1):
            op.lib.npvp_set_option(b"gemm_epi_direct", direct)
            o16 = torch.zeros(M, N, device=DEV, dtype=dt)
            o32 = torch.zeros(M, N, device=DEV)
            if kind == "act0":
                op.gemm(a, w, bias=bias, out_bf16=o16, backend=backend)
            elif kind == "gelu":
                op.gemm(a, w, bias=bias, act=2, out_bf16=o16, backend=backend)
            elif kind == "relu":
                op.gemm(a, w, bias=bias, act=1, out_bf16=o16, backend=backend)
            elif kind == "res_f32_inplace":
                o32.copy_(r32)
                op.gemm(a, w, bias=bias, res1=o32, out_f32=o32, backend=backend)
            elif kind == "f32":
                op.gemm(a, w, bias=bias, out_f32=o32, backend=backend)
            elif kind == "relu_res16":
                op.gemm(a, w, bias=bias, act=1, res1=r16a, out_bf16=o16, backend=backend)
            elif kind == "res16x2":
                op.gemm(a, w, bias=bias, act=1, res1=r16a, res2=r16b, out_bf16=o16, backend=backend)
            else:
                op.gemm(a, w, bias=bias, out_f32=o32, out_bf16=o16, backend=backend)
            torch.cuda.synchronize()
            outs.append((o16, o32))
    finally:
        op.lib.npvp_set_option(b"gemm_epi_direct", 1)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert float(outs[1][0].float().abs().max()) + float(outs[1][1].abs().max()) > 0


@pytest.mark.parametrize("M,N,K", [(640, 2048, 512), (128, 256, 64), (1216, 512, 192)])
def test_gemm_frame_stats(op, spec, M, N, K):
    """fc1 epilogue leaves per-frame partial (sum, sumsq); finalised they must match a separate pass over the output."""
    a, w = rn(M, K, seed=1, dtype=torch.bfloat16), rn(N, K, seed=2, scale=K ** -0.5, dtype=torch.bfloat16)
    bias = rn(N, seed=3) + 0.3
    res = []
    for o in (op, spec):
        out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
        pt = torch.full((M // 64, 4 * N // 256, 2), float("nan"), device=DEV)
        st = torch.empty(M // 64, 2, device=DEV)
        o.gemm(a, w, bias=bias, out_bf16=out, frame_stats=pt)
        o.ffn_stats_finalize(pt, st, 64 * N)
        res.append((out, pt, st))
    close(res[0][0], res[1][0], 1e-2, "gemm out")
    close(res[0][1], res[1][1], 2e-3, "partials")
    close(res[0][2], res[1][2], 1e-4, "finalised stats")
    ref = torch.empty(M // 64, 2, device=DEV)
    spec.ffn_frame_stats(res[0][0].reshape(M // 64, 64, N), ref)       # statistics of the rounded output, second pass
    close(res[0][2], ref, 2e-3, "stats vs second pass")
    with pytest.raises(RuntimeError):
        op.gemm(a, w, bias=bias, act=2, out_bf16=res[0][0], frame_stats=res[0][1])


def test_gemm_f32_and_fourier(op, spec):
    a, w, b = rn(700, 512, seed=1), rn(256, 512, seed=2, scale=0.05), rn(256, seed=3)
    o1, o2 = torch.empty(700, 256, device=DEV), torch.empty(700, 256, device=DEV)
    op.gemm_f32(a, w, b, 1, o1)
    spec.gemm_f32(a, w, b, 1, o2)
    close(o1, o2, 1e-5, "gemm_f32")
    coor = torch.rand(320, 3, device=DEV)
    B = rn(256, 3, seed=5, scale=10.0)
    f1, f2 = torch.empty(320, 512, device=DEV), torch.empty(320, 512, device=DEV)
    op.fourier_features(coor, B, f1)
    spec.fourier_features(coor, B, f2)
    assert float((f1 - f2).abs().max()) < 2e-4       # fp32 rounding of a ~240 rad argument


@pytest.mark.parametrize("n,T,use_ln,use_qe,use_gamma", [(2, 3, True, True, False), (1, 5, True, False, True), (3, 2, False, False, False)])
def test_ln_posfuse(op, spec, n, T, use_ln, use_qe, use_gamma):
    x = rn(n * T * 64, 512, seed=1, scale=2.0) + 0.5
    lw, lb = (rn(512, seed=2) * 0.3 + 1, rn(512, seed=3) * 0.3) if use_ln else (None, None)
    qe = rn(n * 64, 512, seed=4) if use_qe else None
    beta = rn(T * 64, 512, seed=5)
    gamma = rn(T * 64, 512, seed=6, scale=0.2) if use_gamma else None
    outs = [[torch.empty(n * T * 64, 512, device=DEV, dtype=torch.bfloat16) for _ in range(2)] for _ in range(2)]
    op.ln_posfuse(x, lw, lb, qe, beta, gamma, outs[0][0], outs[0][1], n, T)
    spec.ln_posfuse(x, lw, lb, qe, beta, gamma, outs[1][0], outs[1][1], n, T)
    close(outs[0][0], outs[1][0], 1e-2, "ln out")
    close(outs[0][1], outs[1][1], 1e-2, "fused out")


def test_layernorm_rows_and_frame_ln(op, spec):
    x = rn(1000, 512, seed=1, scale=3.0) - 1.0
    w, b = rn(512, seed=2) * 0.3 + 1, rn(512, seed=3) * 0.3
    for relu in (False, True):
        o1, o2 = torch.empty_like(x), torch.empty_like(x)
        b1, b2 = torch.empty_like(x, dtype=torch.bfloat16), torch.empty_like(x, dtype=torch.bfloat16)
        op.layernorm_rows(x, w, b, o1, b1, relu)
        spec.layernorm_rows(x, w, b, o2, b2, relu)
        close(o1, o2, 1e-5, "layernorm f32")
        close(b1, b2, 1e-2, "layernorm bf16")
    h = rn(5 * 64, 512, seed=4, scale=2.0) + 0.3
    aw, ab = rn(64, 512, seed=5) * 0.3 + 1, rn(64, 512, seed=6) * 0.3
    y1 = rn(5 * 64, 512, seed=7)
    y2 = y1.clone()
    op.frame_ln_gelu_residual(h, aw, ab, y1)
    spec.frame_ln_gelu_residual(h, aw, ab, y2)
    close(y1, y2, 1e-5, "frame_ln_gelu_residual")
    hb = h.to(torch.bfloat16)
    y1, y2 = rn(5 * 64, 512, seed=7), rn(5 * 64, 512, seed=7)
    op.frame_ln_gelu_residual(hb, aw, ab, y1)
    spec.frame_ln_gelu_residual(hb, aw, ab, y2)
    close(y1, y2, 1e-5, "frame_ln_gelu_residual (bf16 h)")
    mem = rn(3 * 4 * 64, 512, seed=8)
    e1, e2 = torch.empty(3 * 64, 512, device=DEV), torch.empty(3 * 64, 512, device=DEV)
    op.temporal_mean(mem, e1, 3, 4)
    spec.temporal_mean(mem, e2, 3, 4)
    close(e1, e2, 1e-6, "temporal_mean")


def test_conv_ffn_middle(op, spec):
    frames, Ch = 3, 2048
    h = (rn(frames * 64, Ch, seed=1, scale=1.5) + 0.2).to(torch.bfloat16)
    n1w, n1b = rn(64, Ch, seed=2) * 0.3 + 1, rn(64, Ch, seed=3) * 0.3
    n2w, n2b = rn(64, Ch, seed=4) * 0.3 + 1, rn(64, Ch, seed=5) * 0.3
    dw_w, dw_b = rn(9, Ch, seed=6, scale=0.4), rn(Ch, seed=7, scale=0.2)
    res = []
    for o in (op, spec):
        st = torch.empty(frames, 2, device=DEV)
        y = torch.empty_like(h)
        pt = torch.empty(frames, Ch // FFN_CHUNK, 2, device=DEV)
        g = torch.empty_like(h)
        o.ffn_frame_stats(h, st)
        o.ffn_dwconv(h, st, n1w, n1b, dw_w, dw_b, y, pt)
        o.ffn_norm2(y, pt, n2w, n2b, g)
        res.append((st, y, pt, g))
    close(res[0][0], res[1][0], 1e-4, "ffn stats")
    close(res[0][1], res[1][1], 1e-2, "ffn dwconv")
    close(res[0][2], res[1][2], 2e-2, "ffn partial stats")
    close(res[0][3], res[1][3], 2e-2, "ffn norm2")


@pytest.mark.parametrize("frames", [1, 3, 37, 130, 640])
def test_conv_ffn_middle_half(op, spec, frames):
    """npvp_ffn_mid16 (one pass, half2 arithmetic, statistics exchange through L2) fed by the fc1 GEMM (half output + LN1
    partial statistics from its epilogue) against the fp32 specification; run twice on one persistent exchange scratch (the
    kernel must leave it at rest), and the scratch is checked to be back at rest."""
    from npvp_b200._lib import ffn_mid16_scratch
    Ch = 2048
    assert op.ffn_mid16_lanes() > 0
    a = rn(frames * 64, 512, seed=11, dtype=torch.bfloat16)
    w1, b1 = rn(Ch, 512, seed=12, scale=0.06, dtype=torch.bfloat16), rn(Ch, seed=13, scale=0.3)
    h1, part1 = torch.empty(frames * 64, Ch, dtype=torch.float16, device=DEV), torch.empty(frames, 32, 2, device=DEV)
    op.gemm(a, w1, bias=b1, out_bf16=h1, frame_stats=part1)                       # bf16 operands -> half output (out16 = 1)
    h1_ref, part1_ref = torch.empty_like(h1), torch.empty_like(part1)
    spec.gemm(a, w1, bias=b1, out_bf16=h1_ref, frame_stats=part1_ref)
    close(h1, h1_ref, 2e-3, "fc1 half output")
    close(part1, part1_ref, 1e-3, "fc1 frame statistics")
    n1w, n1b = rn(64, Ch, seed=2) * 0.3 + 1, rn(64, Ch, seed=3) * 0.3
    n2w, n2b = rn(64, Ch, seed=4) * 0.3 + 1, rn(64, Ch, seed=5) * 0.3
    pair = lambda w, b: torch.stack([w.view(64, Ch // 2, 2), b.view(64, Ch // 2, 2)], dim=2)
    ln_wb = torch.stack([pair(n1w, n1b), pair(n2w, n2b)], 0).to(torch.float16).contiguous()
    dw_w, dw_b = rn(9, Ch, seed=6, scale=0.4).to(torch.float16), rn(Ch, seed=7, scale=0.2).to(torch.float16)
    xch, cnt = ffn_mid16_scratch(frames, DEV)
    out, ref = torch.empty_like(h1), torch.empty_like(h1)
    for _ in range(2):
        out.fill_(float("nan"))
        op.ffn_mid16(h1, part1, ln_wb, dw_w, dw_b, out, xch, cnt)
    torch.cuda.synchronize()
    assert bool((xch.view(torch.int32) == -1).all()) and bool((cnt == 0).all()), "exchange scratch not back at rest"
    spec.ffn_mid16(h1, part1, ln_wb, dw_w, dw_b, ref, None, None)
    d = (out.float() - ref.float()).abs()
    print(f"ffn_mid16 frames={frames}: max abs {float(d.max()):.3e} (max |ref| {float(ref.float().abs().max()):.2f}), mean abs {float(d.mean()):.3e}")
    assert bool(torch.isfinite(out.float()).all())
    close(out, ref, 1e-2, "ffn_mid16 vs spec")
    assert float(d.mean()) < 1.5e-3
    # fc2 on the half activations with half weights, bf16 output (out16 = 2)
    w2, b2 = rn(512, Ch, seed=14, scale=0.03).to(torch.float16), rn(512, seed=15, scale=0.3)
    h3, h3_ref = torch.empty(frames * 64, 512, dtype=torch.bfloat16, device=DEV), torch.empty(frames * 64, 512, dtype=torch.bfloat16, device=DEV)
    op.gemm(out, w2, bias=b2, out_bf16=h3)
    spec.gemm(out, w2, bias=b2, out_bf16=h3_ref)
    close(h3, h3_ref, 1e-2, "fc2 half operands -> bf16 output")


@pytest.mark.parametrize("mode,n,Tq,Tk,mask", [(0, 2, 3, 3, False), (1, 2, 5, 5, True), (1, 2, 5, 5, False), (1, 1, 7, 3, False),
                                               (1, 1, 2, 2, True), (1, 1, 28, 2, False), (1, 1, 23, 10, False), (1, 1, 1, 1, True)])
def test_attention(op, spec, mode, n, Tq, Tk, mask):
    qk = rn(n * Tq * 64, 1024, seed=1, scale=1.5, dtype=torch.bfloat16)
    kk = qk[:, 512:] if Tq == Tk else rn(n * Tk * 64, 512, seed=4, scale=1.5, dtype=torch.bfloat16)
    v = rn(n * Tk * 64, 512, seed=2, dtype=torch.bfloat16)
    o1 = torch.empty(n * Tq * 64, 512, device=DEV, dtype=torch.bfloat16)
    o2 = torch.empty_like(o1)
    op.attention(qk[:, :512], kk, v, o1, mode, n, Tq, Tk, mask)
    spec.attention(qk[:, :512], kk, v, o2, mode, n, Tq, Tk, mask)
    close(o1, o2, 1.5e-2, f"attention mode={mode}")


def test_event_encoder_pieces(op, spec):
    n, Cc = 3, 512
    x = rn(n * 64, Cc, seed=1)
    w, sh = rn(9, Cc, seed=2, scale=0.4), rn(Cc, seed=3, scale=0.2)
    o1 = torch.empty(n * 64, Cc, device=DEV, dtype=torch.bfloat16)
    o2 = torch.empty_like(o1)
    op.dwconv3x3_tokens(x, w, sh, o1, True)
    spec.dwconv3x3_tokens(x, w, sh, o2, True)
    close(o1, o2, 1e-2, "dwconv3x3_tokens")
    mulv = rn(n * 64, 2 * Cc, seed=4)
    eps = rn(n, Cc, 8, 8, seed=5)
    for e in (eps, None):
        z1, z2 = torch.empty(n * 64, Cc, device=DEV), torch.empty(n * 64, Cc, device=DEV)
        op.latent_reparam(mulv, e, z1, n, Cc)
        spec.latent_reparam(mulv, e, z2, n, Cc)
        close(z1, z2, 1e-5, "latent_reparam")


@pytest.mark.parametrize("dt", H16)
def test_layout_kernels(op, spec, dt):
    x = rn(5, 512, 64, seed=1)
    t1, t2 = torch.empty(5, 64, 512, device=DEV), torch.empty(5, 64, 512, device=DEV)
    b1 = torch.empty(5, 64, 512, device=DEV, dtype=dt)
    op.nchw_to_tokens(x, t1, b1)
    spec.nchw_to_tokens(x, t2, None)
    assert torch.equal(t1, t2) and torch.equal(b1, t2.to(dt))
    l1, l2 = torch.empty(320, 512, device=DEV, dtype=dt), torch.empty(320, 512, device=DEV, dtype=dt)
    op.layernorm_rows(x.reshape(320, 512), torch.ones(512, device=DEV), torch.zeros(512, device=DEV), None, l1, True)
    spec.layernorm_rows(x.reshape(320, 512), torch.ones(512, device=DEV), torch.zeros(512, device=DEV), None, l2, True)
    close(l1, l2, 1e-2, "layernorm 16-bit out")
    back = torch.empty(5, 512, 64, device=DEV)
    op.tokens_to_nchw(t1, back)
    assert torch.equal(back, x)
    op.tokens_to_nchw(b1, back, relu=True)
    assert torch.equal(back, torch.relu(b1.float()).permute(0, 2, 1))


@pytest.mark.parametrize("dt", H16)
@pytest.mark.parametrize("Cin,Cout,HW", [(3, 32, 128), (1, 64, 64), (3, 64, 40)])
def test_conv7x7_stem(op, spec, Cin, Cout, HW, dt):
    x = rn(3, Cin, HW, HW, seed=1)
    w, sh = rn(49 * Cin, Cout, seed=2, scale=0.1), rn(Cout, seed=3, scale=0.2)
    o1 = torch.empty(3 * HW * HW, Cout, device=DEV, dtype=dt)
    o2 = torch.empty_like(o1)
    op.conv7x7_stem(x, w, sh, o1, Cin, Cout, HW, HW)
    spec.conv7x7_stem(x, w, sh, o2, Cin, Cout, HW, HW)
    close(o1, o2, 1e-2, "conv7x7_stem")


@pytest.mark.parametrize("dt", H16)
@pytest.mark.parametrize("frames,Cin,Cout,H,W,u8", [(3, 3, 32, 128, 128, False), (5, 1, 64, 64, 64, False), (40, 3, 32, 128, 128, True), (3, 3, 64, 20, 36, False),
                                                    (300, 1, 64, 64, 64, True), (2, 3, 32, 128, 160, False), (1, 1, 32, 8, 8, False), (9, 3, 32, 4, 250, True)])
def test_conv7x7_stem_tcgen05_rows(op, spec, frames, Cin, Cout, H, W, u8, dt):
    """The row-streaming tcgen05 stem (head_tc.cu: im2col row blocks built once per input row, Cout 32 / 64) against the kernel
    specification and the mma.sync tile kernel: fp32 and uint8 input, frame segments, ring wrap-around, rectangular / tiny frames."""
    g = torch.Generator().manual_seed(5)
    if u8:
        x = torch.randint(0, 256, (frames, Cin, H, W), generator=g, dtype=torch.uint8).to(DEV)
        norm = ([0.4 + 0.05 * c for c in range(Cin)], [0.25 + 0.02 * c for c in range(Cin)])
    else:
        x, norm = rn(frames, Cin, H, W, seed=1), None
    w, sh = rn(49 * Cin, Cout, seed=2, scale=0.1), rn(Cout, seed=3, scale=0.2)
    outs = []
    for tc in (1, 0):
        op.lib.npvp_set_option(b"stem_tc", tc)
        o = torch.empty(frames * H * W, Cout, device=DEV, dtype=dt)
        op.conv7x7_stem(x, w, sh, o, Cin, Cout, H, W, norm=norm)
        outs.append(o)
    op.lib.npvp_set_option(b"stem_tc", 1)
    o2 = torch.empty_like(outs[0])
    spec.conv7x7_stem(x, w, sh, o2, Cin, Cout, H, W, norm=norm)
    close(outs[0], o2, 1e-2, "conv7x7_stem tcgen05")
    close(outs[0], outs[1], 4e-3, "conv7x7_stem tcgen05 vs mma.sync")     # one 16-bit rounding of the output apart at most


@pytest.mark.parametrize("dt", H16)
@pytest.mark.parametrize("Cin,Cout,HW,phase,act", [(32, 3, 128, True, 3), (64, 1, 64, True, 4), (32, 3, 48, False, 3), (64, 3, 64, True, 3),
                                                  (32, 2, 36, False, 0), (64, 1, 76, False, 4)])
def test_conv7x7_head(op, spec, Cin, Cout, HW, phase, act, dt):
    from npvp_b200._lib import pack_head_weights
    x = rn(2 * HW * HW, Cin, seed=1, dtype=dt)
    w, b = pack_head_weights(rn(49 * Cin, Cout, seed=2, scale=0.03), dt), rn(Cout, seed=3, scale=0.2)
    o1, o2 = torch.empty(2, Cout, HW, HW, device=DEV), torch.empty(2, Cout, HW, HW, device=DEV)
    op.conv7x7_head(x, w, b, o1, Cin, Cout, HW, HW, phase, act)
    spec.conv7x7_head(x, w, b, o2, Cin, Cout, HW, HW, phase, act)
    assert float((o1 - o2).abs().max()) < 2e-4, float((o1 - o2).abs().max())


@pytest.mark.parametrize("dt", H16)
@pytest.mark.parametrize("frames,Cin,Cout,H,W,act", [(3, 32, 3, 128, 128, 3), (5, 64, 3, 64, 64, 3), (40, 32, 3, 128, 128, 3), (3, 32, 3, 20, 36, 3),
                                                     (300, 64, 1, 64, 64, 4), (2, 32, 2, 128, 160, 0), (1, 64, 3, 8, 8, 3), (9, 32, 1, 4, 250, 4)])
def test_conv7x7_head_tcgen05_rows(op, spec, frames, Cin, Cout, H, W, act, dt):
    """The row-streaming tcgen05 head (head_tc.cu; plain NHWC input) against the kernel specification and against the mma.sync
    tile kernel on the same input: frame segments, ring wrap-around, rectangular and tiny frames, fused uint8 pixels."""
    from npvp_b200._lib import pack_head_weights
    x = rn(frames * H * W, Cin, seed=1, dtype=dt)
    w, b = pack_head_weights(rn(49 * Cin, Cout, seed=2, scale=0.03), dt), rn(Cout, seed=3, scale=0.2)
    ren = ([0.5] * Cout, [0.25] * Cout)
    outs = []
    for tc in (1, 0):
        op.lib.npvp_set_option(b"head_tc", tc)
        o, u = torch.empty(frames, Cout, H, W, device=DEV), torch.empty(frames, Cout, H, W, device=DEV, dtype=torch.uint8)
        op.conv7x7_head(x, w, b, o, Cin, Cout, H, W, False, act, out_u8=u, renorm=ren)
        outs.append((o, u))
    op.lib.npvp_set_option(b"head_tc", 1)
    o2 = torch.empty(frames, Cout, H, W, device=DEV)
    spec.conv7x7_head(x, w, b, o2, Cin, Cout, H, W, False, act)
    assert float((outs[0][0] - o2).abs().max()) < 2e-4, float((outs[0][0] - o2).abs().max())
    assert float((outs[0][0] - outs[1][0]).abs().max()) < 1e-5
    assert int((outs[0][1].int() - outs[1][1].int()).abs().max()) <= 1      # a 1e-6 difference can cross a truncation boundary
    pix = ((outs[0][0] * torch.tensor(ren[1], device=DEV).view(1, -1, 1, 1) + torch.tensor(ren[0], device=DEV).view(1, -1, 1, 1)).clamp(0, 1) * 255).floor()
    assert float((pix - outs[0][1].float()).abs().max()) <= 1


def test_conv7x7_tcgen05_more_blocks_than_sms(op):
    """Batches whose row tables would overflow get more CTAs than SMs (whole waves): 1250 frames of 128 x 128 = 160 000 output rows.
    Head and stem, tcgen05 against the mma.sync tile kernels on the same input."""
    from npvp_b200._lib import pack_head_weights
    frames, H, W = 1250, 128, 128
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(frames * H * W, 32, generator=g) * 0.7).to(DEV).half()
    w, b = pack_head_weights(rn(49 * 32, 3, seed=2, scale=0.03), torch.float16), rn(3, seed=3, scale=0.2)
    o = []
    for tc in (1, 0):
        op.lib.npvp_set_option(b"head_tc", tc)
        o.append(torch.empty(frames, 3, H, W, device=DEV))
        op.conv7x7_head(x, w, b, o[-1], 32, 3, H, W, False, 3)
    op.lib.npvp_set_option(b"head_tc", 1)
    assert float((o[0] - o[1]).abs().max()) < 1e-5
    del x, o
    xs = torch.randint(0, 256, (frames, 3, H, W), generator=g, dtype=torch.uint8).to(DEV)
    ws, sh = rn(147, 32, seed=4, scale=0.1), rn(32, seed=5, scale=0.2)
    o = []
    for tc in (1, 0):
        op.lib.npvp_set_option(b"stem_tc", tc)
        o.append(torch.empty(frames * H * W, 32, device=DEV, dtype=torch.float16))
        op.conv7x7_stem(xs, ws, sh, o[-1], 3, 32, H, W, norm=([0.4, 0.45, 0.5], [0.25, 0.27, 0.29]))
    op.lib.npvp_set_option(b"stem_tc", 1)
    close(o[0], o[1], 4e-3, "stem, 1250 frames")


@pytest.mark.parametrize("dt", H16)
@pytest.mark.parametrize("C,H", [(64, 64), (128, 32), (256, 16), (512, 8)])
def test_nonlocal_pieces(op, spec, C, H, dt):
    frames, dq, dv = 2, C // 8, C // 2
    qkv = rn(frames * H * H, 2 * dq + dv, seed=1, scale=0.7, dtype=dt)
    kv1 = torch.empty(frames * H * H // 4, dq + dv, device=DEV, dtype=dt)
    kv2 = torch.empty_like(kv1)
    op.maxpool2x2_cols(qkv, dq, dq + dv, kv1, frames, H, H)
    spec.maxpool2x2_cols(qkv, dq, dq + dv, kv2, frames, H, H)
    assert torch.equal(kv1, kv2)
    o1 = torch.empty(frames * H * H, dv, device=DEV, dtype=dt)
    o2 = torch.empty_like(o1)
    op.nonlocal_attention(qkv[:, :dq], kv1, o1, frames, H * H, H * H // 4, dq, dv)
    spec.nonlocal_attention(qkv[:, :dq], kv1, o2, frames, H * H, H * H // 4, dq, dv)
    close(o1, o2, 1.5e-2, "nonlocal_attention")


@pytest.mark.parametrize("dt", H16)
@pytest.mark.parametrize("frames,H,C,N,KH,stride,pad,mode,phase,epi", [
    (3, 16, 64, 128, 3, 1, 1, 0, False, "relu_res"), (5, 16, 32, 64, 3, 2, 1, 0, False, "relu"), (9, 8, 512, 512, 3, 1, 1, 1, False, "res_f32"),
    (2, 8, 128, 64, 3, 1, 1, 2, False, "relu"), (3, 16, 64, 128, 2, 1, 0, 0, True, "relu"), (70, 8, 512, 1024, 2, 1, 0, 0, False, "relu"),
    (2, 64, 64, 64, 3, 1, 1, 0, False, "relu_res"), (3, 32, 32, 256, 2, 1, 0, 0, True, "relu")])
def test_conv_gemm_implicit(op, spec, dt, frames, H, C, N, KH, stride, pad, mode, phase, epi):
    x = rn(frames * H * H, C, seed=1, dtype=dt)
    K = KH * KH * C
    w = rn(N, K, seed=2, scale=K ** -0.5, dtype=dt)
    bias = rn(N, seed=3)
    Ho = H // stride
    M = frames * Ho * Ho
    kw = dict(bias=bias)
    outs = []
    for o in (op, spec):
        if epi == "relu":
            out = torch.empty(M, N, device=DEV, dtype=dt)
            o.conv_gemm(x, w, frames, H, H, C, KH, KH, stride, pad, mode, Ho, Ho, phase, act=1, out_bf16=out, **kw)
        elif epi == "relu_res":
            out = torch.empty(M, N, device=DEV, dtype=dt)
            o.conv_gemm(x, w, frames, H, H, C, KH, KH, stride, pad, mode, Ho, Ho, phase, act=1, res1=rn(M, N, seed=4, dtype=dt), out_bf16=out, **kw)
        else:
            out = torch.empty(M, N, device=DEV)
            o.conv_gemm(x, w, frames, H, H, C, KH, KH, stride, pad, mode, Ho, Ho, phase, res1=rn(M, N, seed=4, dtype=dt), post_relu=True, out_f32=out, **kw)
        outs.append(out)
    close(outs[0], outs[1], 1e-2 if dt == torch.bfloat16 else 2e-3, "conv_gemm")


@pytest.mark.parametrize("dt", H16)
@pytest.mark.parametrize("frames,H,C,N,KH,epi", [
    (3, 8, 128, 128, 3, "relu_res"),        # two 8x8 frames per 128-pixel tile, odd frame count (the last tile overhangs)
    (2, 32, 64, 64, 3, "relu"),             # four rows per tile
    (1, 128, 64, 64, 3, "relu_res"),        # one row per tile
    (5, 16, 256, 256, 3, "res_f32"),        # run-time epilogue flags, several k-blocks per tap
    (2, 16, 64, 512, 5, "relu"),            # 5x5 window: shifts of +-2
    (300, 16, 64, 64, 3, "relu_res")])      # resident weights, many tiles per CTA (stage ring wraps)
def test_conv_gemm_tma_window(op, spec, dt, frames, H, C, N, KH, epi):
    """stride-1 zero-padded convolutions load their A operand as shifted TMA windows; compared with the spec AND with the gather path"""
    pad = KH // 2
    x = rn(frames * H * H, C, seed=1, dtype=dt)
    K = KH * KH * C
    w = rn(N, K, seed=2, scale=K ** -0.5, dtype=dt)
    bias = rn(N, seed=3)
    M = frames * H * H
    res = rn(M, N, seed=4, dtype=dt)

    def run(o):
        if epi == "relu":
            out = torch.empty(M, N, device=DEV, dtype=dt)
            o.conv_gemm(x, w, frames, H, H, C, KH, KH, 1, pad, 0, H, H, False, act=1, out_bf16=out, bias=bias)
        elif epi == "relu_res":
            out = torch.empty(M, N, device=DEV, dtype=dt)
            o.conv_gemm(x, w, frames, H, H, C, KH, KH, 1, pad, 0, H, H, False, act=1, res1=res, out_bf16=out, bias=bias)
        else:
            out = torch.empty(M, N, device=DEV)
            o.conv_gemm(x, w, frames, H, H, C, KH, KH, 1, pad, 0, H, H, False, res1=res, post_relu=True, out_f32=out, bias=bias)
        return out
    o_tma, o_spec = run(op), run(spec)
    op.lib.npvp_set_option(b"conv_tma", 0)
    try:
        o_gather = run(op)
    finally:
        op.lib.npvp_set_option(b"conv_tma", 1)
    op.lib.npvp_set_option(b"conv_wres", 0)
    try:
        o_stream = run(op)
    finally:
        op.lib.npvp_set_option(b"conv_wres", 1)
    close(o_tma, o_spec, 1e-2 if dt == torch.bfloat16 else 2e-3, "conv_gemm tma")
    assert torch.equal(o_tma, o_gather), "TMA-window and gather paths must agree bitwise (same MMA order)"
    assert torch.equal(o_tma, o_stream), "resident and streamed weights must agree bitwise"


@pytest.mark.parametrize("dt", H16)
@pytest.mark.parametrize("frames,H,Cin,Cout", [(3, 8, 512, 256), (2, 16, 256, 128), (3, 32, 128, 64), (2, 64, 64, 32), (5, 8, 128, 64),
                                               (1, 16, 64, 96)])
def test_convt_gemm(op, spec, dt, frames, H, Cin, Cout):
    """transposed conv: live (phase, tap) blocks only, NHWC in / NHWC out; random dense weights in EVERY block of Wt so that a
    wrongly skipped or wrongly included block shows (the kernel must ignore the 7 dead blocks, the spec zeroes them)"""
    x = rn(frames * H * H, Cin, seed=1, dtype=dt)
    w = rn(4 * Cout, 4 * Cin, seed=2, scale=(2.25 * Cin) ** -0.5, dtype=dt)
    live = torch.zeros(4, 1, 4, 1, device=DEV)
    for q, (py, px) in enumerate(((0, 0), (0, 1), (1, 1), (1, 0))):
        for dy in range(py + 1):
            for dx in range(px + 1):
                live[q, 0, dy * 2 + dx, 0] = 1
    w_live = (w.float().reshape(4, Cout, 4, Cin) * live).reshape(4 * Cout, 4 * Cin).to(dt)
    bias = rn(4 * Cout, seed=3)
    o1 = torch.full((frames * 4 * H * H, Cout), float("nan"), device=DEV, dtype=dt)
    o2 = torch.empty_like(o1)
    op.convt_gemm(x, w, frames, H, H, Cin, Cout, bias=bias, act=1, out_bf16=o1)       # dead blocks hold garbage: must not be read
    spec.convt_gemm(x, w_live, frames, H, H, Cin, Cout, bias=bias, act=1, out_bf16=o2)
    assert not torch.isnan(o1.float()).any(), "convt_gemm left output pixels unwritten"
    close(o1, o2, 1e-2 if dt == torch.bfloat16 else 2e-3, "convt_gemm")
    o3 = torch.empty_like(o1)
    op.lib.npvp_set_option(b"conv_wres", 0)
    try:
        op.convt_gemm(x, w, frames, H, H, Cin, Cout, bias=bias, act=1, out_bf16=o3)
    finally:
        op.lib.npvp_set_option(b"conv_wres", 1)
    assert torch.equal(o1, o3), "resident and streamed weights must agree bitwise"
    # and against torch's own transposed convolution on a weight in the nn.ConvTranspose2d layout
    wt = rn(Cin, Cout, 3, 3, seed=5, scale=(2.25 * Cin) ** -0.5).to(dt).float()
    tap = {(0, 0): 1, (1, 0): 2, (1, 1): 0}
    B = torch.zeros(4, Cout, 2, 2, Cin, device=DEV)
    for q, (py, px) in enumerate(((0, 0), (0, 1), (1, 1), (1, 0))):
        for dy in (0, 1):
            for dx in (0, 1):
                if (py, dy) in tap and (px, dx) in tap:
                    B[q, :, dy, dx, :] = wt[:, :, tap[(py, dy)], tap[(px, dx)]].t()
    op.convt_gemm(x, B.reshape(4 * Cout, 4 * Cin).to(dt), frames, H, H, Cin, Cout, bias=None, act=0, out_bf16=o1)
    ref = torch.nn.functional.conv_transpose2d(x.float().reshape(frames, H, H, Cin).permute(0, 3, 1, 2), wt, stride=2, padding=1, output_padding=1)
    close(o1, ref.permute(0, 2, 3, 1).reshape(o1.shape).to(dt), 1e-2 if dt == torch.bfloat16 else 2e-3, "convt_gemm vs conv_transpose2d")


def test_deferred_residual_layernorms(op, spec):
    n, T = 2, 3
    rows = n * T * 64
    x1 = rn(rows, 512, seed=1, scale=2.0)
    x2 = x1.clone()
    delta = rn(rows, 512, seed=2, dtype=torch.bfloat16)
    w, b = rn(512, seed=3) * 0.3 + 1, rn(512, seed=4) * 0.3
    o1, o2 = torch.empty(rows, 512, device=DEV, dtype=torch.bfloat16), torch.empty(rows, 512, device=DEV, dtype=torch.bfloat16)
    op.add_layernorm_rows(x1, delta, w, b, None, o1, False)
    spec.add_layernorm_rows(x2, delta, w, b, None, o2, False)
    assert torch.equal(x1, x2)                         # the fp32 stream update is exact
    close(o1, o2, 1e-2, "add_layernorm_rows")
    qe, beta = rn(n * 64, 512, seed=5), rn(T * 64, 512, seed=6)
    a1, f1 = torch.empty_like(o1), torch.empty_like(o1)
    a2, f2 = torch.empty_like(o1), torch.empty_like(o1)
    op.add_ln_posfuse(x1, delta, w, b, qe, beta, None, a1, f1, n, T)
    spec.add_ln_posfuse(x2, delta, w, b, qe, beta, None, a2, f2, n, T)
    assert torch.equal(x1, x2)
    close(a1, a2, 1e-2, "add_ln_posfuse ln")
    close(f1, f2, 1e-2, "add_ln_posfuse fused")
