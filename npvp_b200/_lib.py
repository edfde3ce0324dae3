"""ctypes binding of ``libnpvp_b200.so`` (the C-ABI declared in ``include/npvp_b200.h``).

``ops()`` returns the singleton :class:`Ops`, whose methods take torch CUDA tensors, pass raw device
pointers + sizes + the current CUDA stream across the C boundary and raise ``RuntimeError`` with
``npvp_last_error()`` on a non-zero return code.  There is no fallback: if the library is missing or
the tensors are not on a CUDA device the call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

ACT_NONE, ACT_RELU, ACT_GELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
GEMM_AUTO, GEMM_TCGEN05, GEMM_SIMT, GEMM_TCGEN05_2CTA = 0, 1, 2, 4
PAD_ZERO, PAD_REFLECT, PAD_REPLICATE = 0, 1, 2
ATTN_SPATIAL, ATTN_TEMPORAL = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NPVP_B200_LIB") or os.path.join(_HERE, "libnpvp_b200.so")   # NPVP_B200_LIB: another build of the same C-ABI (same-box A/B runs)

_i64, _i32, _f32, _vp = C.c_int64, C.c_int, C.c_float, C.c_void_p


class Epilogue(C.Structure):
    _fields_ = [("bias", _vp), ("res1", _vp), ("res2", _vp), ("out_f32", _vp), ("out_bf16", _vp),
                ("alpha", _f32), ("act", C.c_int32), ("res1_bf16", C.c_int32), ("res2_bf16", C.c_int32),
                ("post_relu", C.c_int32), ("fp16", C.c_int32), ("out16", C.c_int32), ("ld_out", _i64), ("ld_res", _i64),
                ("frame_stats", _vp)]


# symbol -> argtypes; every function returns int.  Kept in one table so tests can check the export list.
SIGNATURES = {
    "npvp_gemm_bf16": [_vp, _i64, _vp, _i64, _i64, _i64, _i64, C.POINTER(Epilogue), _i32, _vp],
    "npvp_conv_gemm_bf16": [_vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i64, _i64,
                            C.POINTER(Epilogue), _vp],
    "npvp_convt_gemm_bf16": [_vp, _i64, _i32, _i32, _i32, _vp, _i64, _i32, C.POINTER(Epilogue), _vp],
    "npvp_gemm_f32": [_vp, _i64, _vp, _i64, _i64, _i64, _i64, _vp, _i32, _vp, _i64, _vp],
    "npvp_fourier_features": [_vp, _vp, _vp, _i64, _i32, _vp],
    "npvp_ln_posfuse": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    "npvp_add_layernorm_rows": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp],
    "npvp_add_ln_posfuse": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    "npvp_layernorm_rows": [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp],
    "npvp_frame_ln_gelu_residual": [_vp, _i32, _vp, _vp, _vp, _i64, _vp],
    "npvp_frame_ln_gelu_residual_posfuse": [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp],
    "npvp_temporal_mean": [_vp, _vp, _i64, _i64, _i64, _vp],
    "npvp_ffn_frame_stats": [_vp, _vp, _i64, _i64, _vp],
    "npvp_ffn_stats_finalize": [_vp, _i64, _vp, _i64, _i64, _vp],
    "npvp_frames_to_pixels": [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i64, _vp],
    "npvp_pixels_to_frames": [_vp, _vp, _vp, _vp, _i64, _i32, _i64, _vp],
    "npvp_psnr": [_vp, _vp, _vp, _i64, _i64, _f32, _vp],
    "npvp_ssim": [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp],
    "npvp_sample_scores": [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _f32, _vp],
    "npvp_best_of_k": [_vp, _vp, _i64, _i32, _i32, _i64, _vp, _vp, _vp, _vp],
    "npvp_ffn_dwconv": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp],
    "npvp_ffn_norm2": [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp],
    "npvp_ffn_mid16": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp],
    "npvp_ffn_mid16_lanes": [],
    "npvp_attention": [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _i64, _i32, _i32, _i32, _vp],
    "npvp_dwconv3x3_tokens": [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp],
    "npvp_latent_reparam": [_vp, _i64, _vp, _vp, _i64, _i64, _vp],
    "npvp_nchw_to_tokens": [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _vp],
    "npvp_tokens_to_nchw": [_vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp],
    "npvp_conv7x7_stem": [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp],
    "npvp_conv7x7_head": [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp],
    "npvp_maxpool2x2_cols": [_vp, _i64, _i32, _i32, _vp, _i64, _i32, _i32, _i32, _vp],
    "npvp_nonlocal_attention": [_vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp],
}
AUX_SYMBOLS = ["npvp_last_error", "npvp_version", "npvp_launch_count", "npvp_reset_launch_count", "npvp_set_option"]


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with `python -m npvp_b200.build` (nvcc, sm_100a). "
                           "npvp_b200 has no CPU or eager fallback.")
    lib = C.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = argtypes, C.c_int
    lib.npvp_last_error.restype = C.c_char_p
    lib.npvp_version.restype = C.c_int
    lib.npvp_launch_count.restype = C.c_int64
    lib.npvp_reset_launch_count.restype = None
    lib.npvp_set_option.argtypes, lib.npvp_set_option.restype = [C.c_char_p, C.c_int], C.c_int
    return lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk(t: Optional[torch.Tensor], dtype, name: str, contiguous: bool = True):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"npvp_b200: tensor '{name}' must live on a CUDA device (no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"npvp_b200: tensor '{name}' must be {dtype}, got {t.dtype}")
    if contiguous and not t.is_contiguous():
        raise ValueError(f"npvp_b200: tensor '{name}' must be contiguous")


H16 = (torch.bfloat16, torch.float16)


def _chk16(t: Optional[torch.Tensor], name: str, contiguous: bool = True, like: Optional[torch.Tensor] = None):
    """16-bit operand buffer: bfloat16 or float16 (all 16-bit operands of one call must agree)."""
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"npvp_b200: tensor '{name}' must live on a CUDA device (no CPU path)")
    if t.dtype not in H16:
        raise TypeError(f"npvp_b200: tensor '{name}' must be bfloat16 or float16, got {t.dtype}")
    if like is not None and t.dtype != like.dtype:
        raise TypeError(f"npvp_b200: tensor '{name}' is {t.dtype} but the call's operands are {like.dtype}")
    if contiguous and not t.is_contiguous():
        raise ValueError(f"npvp_b200: tensor '{name}' must be contiguous")


def _is_fp16(t: torch.Tensor) -> int:
    return int(t.dtype == torch.float16)


def _rowmajor(t: torch.Tensor, name: str):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"npvp_b200: '{name}' must be a 2-D view with unit inner stride")
    return t.stride(0)


class Ops:
    """Tensor-level wrappers over the C-ABI.  All outputs are preallocated by the caller."""

    def __init__(self, lib: Optional[C.CDLL] = None):
        self.lib = lib or load_library()
        self.gemm_backend = {"auto": GEMM_AUTO, "tcgen05": GEMM_TCGEN05, "simt": GEMM_SIMT, "tcgen05_2cta": GEMM_TCGEN05_2CTA}[os.environ.get("NPVP_B200_GEMM", "auto")]
        self.lib.npvp_set_option(b"gemm_2cta", int(os.environ.get("NPVP_B200_GEMM_2CTA", "-1")))
        self.lib.npvp_set_option(b"gemm_epi_direct", int(os.environ.get("NPVP_B200_GEMM_EPI_DIRECT", "1")))
        self.lib.npvp_set_option(b"ffn_mid16_mode", int(os.environ.get("NPVP_B200_FFN_MID16_MODE", "0")))
        self.lib.npvp_set_option(b"conv_tma", int(os.environ.get("NPVP_B200_CONV_TMA", "1")))
        self.lib.npvp_set_option(b"gemm_prefetch", int(os.environ.get("NPVP_B200_GEMM_PREFETCH", "1")))
        self.lib.npvp_set_option(b"head_tc", int(os.environ.get("NPVP_B200_HEAD_TC", "1")))
        self.lib.npvp_set_option(b"stem_tc", int(os.environ.get("NPVP_B200_STEM_TC", "1")))

    # -- plumbing -------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _call(self, name, *args):
        rc = getattr(self.lib, name)(*args)
        if rc != 0:
            raise RuntimeError(f"{name} failed ({rc}): {self.lib.npvp_last_error().decode()}")

    def launch_count(self) -> int:
        return int(self.lib.npvp_launch_count())

    def reset_launch_count(self):
        self.lib.npvp_reset_launch_count()

    # -- contractions ---------------------------------------------------------------------------
    def gemm(self, a, w, *, bias=None, act=ACT_NONE, alpha=1.0, res1=None, res2=None, out_f32=None, out_bf16=None,
             post_relu=False, backend=None, frame_stats=None):
        """``frame_stats``: fp32 [M/64, 4*N/256, 2] receiving partial (sum, sumsq) of the outputs per 64-row frame.
        ``out_bf16`` (and 16-bit residuals) may be of the other 16-bit type than the operands (npvp_epilogue_t.out16)."""
        _chk16(a, "a", False); _chk16(w, "w", False, like=a)
        _chk(bias, torch.float32, "bias"); _chk(out_f32, torch.float32, "out_f32", False)
        _chk16(out_bf16, "out_bf16", False)
        o16 = out_bf16.dtype if out_bf16 is not None else a.dtype
        lda, ldw = _rowmajor(a, "a"), _rowmajor(w, "w")
        M, K = a.shape
        N = w.shape[0]
        assert w.shape[1] == K, (a.shape, w.shape)
        outs = [o for o in (out_f32, out_bf16) if o is not None]
        assert outs and all(o.shape == (M, N) for o in outs), "gemm: output shape mismatch"
        ld_out = _rowmajor(outs[0], "out")
        assert all(_rowmajor(o, "out") == ld_out for o in outs)
        ld_res = 0
        for r in (res1, res2):
            if r is not None:
                assert r.is_cuda and r.shape == (M, N) and r.dtype in (torch.float32, o16)
                ld = _rowmajor(r, "res")
                assert ld_res in (0, ld), "gemm: residuals must share a row stride"
                ld_res = ld
        ep = Epilogue(_ptr(bias), _ptr(res1), _ptr(res2), _ptr(out_f32), _ptr(out_bf16), float(alpha), int(act),
                      int(res1 is not None and res1.dtype in H16),
                      int(res2 is not None and res2.dtype in H16), int(post_relu), _is_fp16(a),
                      0 if o16 == a.dtype else (1 if o16 == torch.float16 else 2), ld_out, ld_res, _ptr(frame_stats))
        if frame_stats is not None:
            _chk(frame_stats, torch.float32, "frame_stats")
            assert M % 64 == 0 and N % 256 == 0 and tuple(frame_stats.shape) == (M // 64, 4 * N // 256, 2)
        self._call("npvp_gemm_bf16", a.data_ptr(), lda, w.data_ptr(), ldw, M, N, K, C.byref(ep),
                   self.gemm_backend if backend is None else backend, self._stream())

    def conv_gemm(self, x, w, frames, H, W, Cc, KH, KW, stride, pad, pad_mode, Ho, Wo, phase_major=False, *, bias=None,
                  act=ACT_NONE, alpha=1.0, res1=None, res2=None, out_f32=None, out_bf16=None, post_relu=False):
        """Implicit-GEMM convolution: x 16-bit [frames*H*W, C] (NHWC or phase-major), w 16-bit [N, KH*KW*C]."""
        _chk16(x, "x"); _chk16(w, "w", False, like=x)
        _chk(bias, torch.float32, "bias"); _chk(out_f32, torch.float32, "out_f32", False); _chk16(out_bf16, "out_bf16", False, like=x)
        assert x.numel() == frames * H * W * Cc and w.shape[1] == KH * KW * Cc
        M, N = frames * Ho * Wo, w.shape[0]
        outs = [o for o in (out_f32, out_bf16) if o is not None]
        assert outs and all(o.shape == (M, N) for o in outs), "conv_gemm: output shape mismatch"
        ld_out = _rowmajor(outs[0], "out")
        ld_res = 0
        for r in (res1, res2):
            if r is not None:
                assert r.is_cuda and r.shape == (M, N) and r.dtype in (torch.float32, x.dtype)
                ld = _rowmajor(r, "res")
                assert ld_res in (0, ld)
                ld_res = ld
        ep = Epilogue(_ptr(bias), _ptr(res1), _ptr(res2), _ptr(out_f32), _ptr(out_bf16), float(alpha), int(act),
                      int(res1 is not None and res1.dtype in H16), int(res2 is not None and res2.dtype in H16),
                      int(bool(post_relu)), _is_fp16(x), 0, ld_out, ld_res)
        self._call("npvp_conv_gemm_bf16", x.data_ptr(), frames, H, W, Cc, KH, KW, stride, pad, pad_mode, Ho, Wo, int(phase_major),
                   w.data_ptr(), _rowmajor(w, "w"), N, C.byref(ep), self._stream())

    @staticmethod
    def convt_supported(H, W, Cin, Cout):
        """geometry npvp_convt_gemm_bf16 accepts (include/npvp_b200.h)"""
        HW = H * W
        return (Cin % 64 == 0 and Cout % 32 == 0 and W <= 128 and (W & (W - 1)) == 0 and 128 % W == 0
                and (HW % 128 == 0 or 128 % HW == 0))

    def convt_gemm(self, x, w, frames, H, W, Cin, Cout, *, bias=None, act=ACT_NONE, out_bf16=None):
        """ConvTranspose2d(3, s2, p1, op1): x 16-bit NHWC [frames*H*W, Cin], w 16-bit [(q,co), (dy,dx,ci)] with the phases
        q = (0,0), (0,1), (1,1), (1,0) -> out 16-bit NHWC [frames*2H*2W, Cout]."""
        _chk16(x, "x"); _chk16(w, "w", False, like=x); _chk(bias, torch.float32, "bias"); _chk16(out_bf16, "out_bf16", False, like=x)
        assert x.numel() == frames * H * W * Cin and tuple(w.shape) == (4 * Cout, 4 * Cin)
        assert out_bf16.is_contiguous() and out_bf16.numel() == frames * 4 * H * W * Cout
        assert bias is None or bias.numel() == 4 * Cout
        ep = Epilogue(_ptr(bias), None, None, None, _ptr(out_bf16), 1.0, int(act), 0, 0, 0, _is_fp16(x), 0, Cout, 0)
        self._call("npvp_convt_gemm_bf16", x.data_ptr(), frames, H, W, Cin, w.data_ptr(), _rowmajor(w, "w"), Cout, C.byref(ep),
                   self._stream())

    def gemm_f32(self, a, w, bias, act, out):
        for t, n in ((a, "a"), (w, "w"), (out, "out")):
            _chk(t, torch.float32, n, False)
        _chk(bias, torch.float32, "bias")
        M, K = a.shape
        N = w.shape[0]
        assert w.shape[1] == K and out.shape == (M, N)
        self._call("npvp_gemm_f32", a.data_ptr(), _rowmajor(a, "a"), w.data_ptr(), _rowmajor(w, "w"), M, N, K, _ptr(bias),
                   int(act), out.data_ptr(), _rowmajor(out, "out"), self._stream())

    # -- predictor ------------------------------------------------------------------------------
    def fourier_features(self, coor, B, out):
        _chk(coor, torch.float32, "coor"); _chk(B, torch.float32, "B"); _chk(out, torch.float32, "out")
        rows, half = coor.shape[0], B.shape[0]
        assert coor.shape[1] == 3 and B.shape[1] == 3 and out.shape == (rows, 2 * half)
        self._call("npvp_fourier_features", coor.data_ptr(), B.data_ptr(), out.data_ptr(), rows, half, self._stream())

    def ln_posfuse(self, x, ln_w, ln_b, qe, beta, gamma, out_ln, out_fused, n_clips, T):
        _chk(x, torch.float32, "x"); _chk(ln_w, torch.float32, "ln_w"); _chk(ln_b, torch.float32, "ln_b")
        _chk(qe, torch.float32, "qe"); _chk(beta, torch.float32, "beta"); _chk(gamma, torch.float32, "gamma")
        _chk(out_ln, torch.bfloat16, "out_ln"); _chk(out_fused, torch.bfloat16, "out_fused")
        assert x.numel() == n_clips * T * 64 * 512
        assert qe is None or qe.numel() == n_clips * 64 * 512
        self._call("npvp_ln_posfuse", x.data_ptr(), _ptr(ln_w), _ptr(ln_b), _ptr(qe), _ptr(beta), _ptr(gamma), _ptr(out_ln),
                   _ptr(out_fused), n_clips, T, _pos_frames(beta, gamma, n_clips, T), self._stream())

    def add_layernorm_rows(self, x, delta, w, b, out_f32=None, out_bf16=None, relu=False):
        """x += delta (deferred residual, written back), then LayerNorm(512) of the updated rows."""
        _chk(x, torch.float32, "x"); _chk(delta, torch.bfloat16, "delta"); _chk(w, torch.float32, "w"); _chk(b, torch.float32, "b")
        _chk(out_f32, torch.float32, "out_f32"); _chk16(out_bf16, "out_bf16")
        assert delta.numel() == x.numel()
        self._call("npvp_add_layernorm_rows", x.data_ptr(), delta.data_ptr(), w.data_ptr(), b.data_ptr(), _ptr(out_f32), _ptr(out_bf16),
                   x.numel() // 512, int(relu), 0 if out_bf16 is None else _is_fp16(out_bf16), self._stream())

    def add_ln_posfuse(self, x, delta, ln_w, ln_b, qe, beta, gamma, out_ln, out_fused, n_clips, T):
        """x += delta (deferred residual, written back), then ln_posfuse of the updated frames."""
        _chk(x, torch.float32, "x"); _chk(delta, torch.bfloat16, "delta"); _chk(ln_w, torch.float32, "ln_w"); _chk(ln_b, torch.float32, "ln_b")
        _chk(qe, torch.float32, "qe"); _chk(beta, torch.float32, "beta"); _chk(gamma, torch.float32, "gamma")
        _chk(out_ln, torch.bfloat16, "out_ln"); _chk(out_fused, torch.bfloat16, "out_fused")
        assert x.numel() == n_clips * T * 64 * 512 and delta.numel() == x.numel()
        self._call("npvp_add_ln_posfuse", x.data_ptr(), delta.data_ptr(), _ptr(ln_w), _ptr(ln_b), _ptr(qe), _ptr(beta), _ptr(gamma),
                   _ptr(out_ln), _ptr(out_fused), n_clips, T, _pos_frames(beta, gamma, n_clips, T), self._stream())

    def layernorm_rows(self, x, w, b, out_f32=None, out_bf16=None, relu=False):
        _chk(x, torch.float32, "x"); _chk(w, torch.float32, "w"); _chk(b, torch.float32, "b")
        _chk(out_f32, torch.float32, "out_f32"); _chk16(out_bf16, "out_bf16")
        rows = x.numel() // 512
        self._call("npvp_layernorm_rows", x.data_ptr(), w.data_ptr(), b.data_ptr(), _ptr(out_f32), _ptr(out_bf16), rows,
                   int(relu), 0 if out_bf16 is None else _is_fp16(out_bf16), self._stream())

    def frame_ln_gelu_residual(self, h, w_hwc, b_hwc, y):
        for t, n in ((w_hwc, "w"), (b_hwc, "b"), (y, "y")):
            _chk(t, torch.float32, n)
        assert h.is_cuda and h.is_contiguous() and h.dtype in (torch.float32, torch.bfloat16)
        frames = h.numel() // (64 * 512)
        assert y.numel() == h.numel() and w_hwc.numel() == 64 * 512
        self._call("npvp_frame_ln_gelu_residual", h.data_ptr(), int(h.dtype == torch.bfloat16), w_hwc.data_ptr(), b_hwc.data_ptr(),
                   y.data_ptr(), frames, self._stream())

    def frame_ln_gelu_residual_posfuse(self, h, w_hwc, b_hwc, y, ln_w, ln_b, qe, beta, gamma, out_ln, out_fused, n_clips, T):
        for t, n in ((w_hwc, "w"), (b_hwc, "b"), (y, "y"), (ln_w, "ln_w"), (ln_b, "ln_b"), (qe, "qe"), (beta, "beta"),
                     (gamma, "gamma")):
            _chk(t, torch.float32, n)
        assert h.is_cuda and h.is_contiguous() and h.dtype in (torch.float32, torch.bfloat16)
        _chk(out_ln, torch.bfloat16, "out_ln"); _chk(out_fused, torch.bfloat16, "out_fused")
        assert h.numel() == n_clips * T * 64 * 512 and y.numel() == h.numel()
        self._call("npvp_frame_ln_gelu_residual_posfuse", h.data_ptr(), int(h.dtype == torch.bfloat16), w_hwc.data_ptr(), b_hwc.data_ptr(),
                   y.data_ptr(), _ptr(ln_w),
                   _ptr(ln_b), _ptr(qe), _ptr(beta), _ptr(gamma), _ptr(out_ln), _ptr(out_fused), n_clips, T,
                   _pos_frames(beta, gamma, n_clips, T), self._stream())

    def temporal_mean(self, mem, evt, n_clips, T):
        _chk(mem, torch.float32, "mem"); _chk(evt, torch.float32, "evt")
        fe = mem.numel() // (n_clips * T)
        assert evt.numel() == n_clips * fe
        self._call("npvp_temporal_mean", mem.data_ptr(), evt.data_ptr(), n_clips, T, fe, self._stream())

    def ffn_frame_stats(self, h, stats):
        _chk(h, torch.bfloat16, "h"); _chk(stats, torch.float32, "stats")
        frames, Ch = stats.shape[0], h.shape[-1]
        assert h.numel() == frames * 64 * Ch
        self._call("npvp_ffn_frame_stats", h.data_ptr(), stats.data_ptr(), frames, Ch, self._stream())

    # -- pixel space / metrics ------------------------------------------------------------------
    @staticmethod
    def _host_f32(vals):
        return (C.c_float * len(vals))(*[float(v) for v in vals])

    def frames_to_pixels(self, frames, mean, std, out_f32=None, out_u8=None):
        """frames fp32 (..., C, H, W) model space -> clamp(x*std+mean, 0, 1) as fp32 and / or uint8 (reference op order)."""
        _chk(frames, torch.float32, "frames"); _chk(out_f32, torch.float32, "out_f32", False)
        Cc, H, W = frames.shape[-3:]
        assert len(mean) == Cc and len(std) == Cc and (out_f32 is not None or out_u8 is not None)
        for o, dt in ((out_f32, torch.float32), (out_u8, torch.uint8)):
            assert o is None or (o.is_cuda and o.is_contiguous() and o.dtype == dt and o.shape == frames.shape)
        n = frames.numel() // (Cc * H * W)
        inv_std = self._host_f32([1.0 / float(s) for s in std])
        inv_mean = self._host_f32([-float(m) for m in mean])
        self._call("npvp_frames_to_pixels", frames.data_ptr(), inv_std, inv_mean, _ptr(out_f32), _ptr(out_u8), n, Cc, H * W, self._stream())

    def pixels_to_frames(self, u8, mean, std, out):
        assert u8.is_cuda and u8.is_contiguous() and u8.dtype == torch.uint8
        _chk(out, torch.float32, "out")
        Cc, H, W = u8.shape[-3:]
        assert len(mean) == Cc and len(std) == Cc and out.shape == u8.shape
        self._call("npvp_pixels_to_frames", u8.data_ptr(), self._host_f32(mean), self._host_f32(std), out.data_ptr(),
                   u8.numel() // (Cc * H * W), Cc, H * W, self._stream())

    def psnr(self, x, y, out, data_range=1.0):
        _chk(x, torch.float32, "x"); _chk(y, torch.float32, "y"); _chk(out, torch.float32, "out")
        n = x.shape[0]
        assert x.shape == y.shape and out.shape == (n,)
        self._call("npvp_psnr", x.data_ptr(), y.data_ptr(), out.data_ptr(), n, x.numel() // n, float(data_range), self._stream())

    def ssim(self, x, y, window11, out):
        _chk(x, torch.float32, "x"); _chk(y, torch.float32, "y"); _chk(out, torch.float32, "out")
        n, Cc, H, W = x.shape
        assert x.shape == y.shape and out.shape == (n,) and len(window11) == 11
        self._call("npvp_ssim", x.data_ptr(), y.data_ptr(), self._host_f32(window11), out.data_ptr(), n, Cc, H, W, self._stream())

    def sample_scores(self, samples, gt, scores, window11=None, data_range=1.0):
        """samples fp32 (N, K, T, C, H, W) vs gt fp32 (N, T, C, H, W) -> scores (N, K, T): PSNR, or SSIM when ``window11`` is given."""
        _chk(samples, torch.float32, "samples"); _chk(gt, torch.float32, "gt"); _chk(scores, torch.float32, "scores")
        n, K, T, Cc, H, W = samples.shape
        assert tuple(gt.shape) == (n, T, Cc, H, W) and tuple(scores.shape) == (n, K, T) and (window11 is None or len(window11) == 11)
        self._call("npvp_sample_scores", samples.data_ptr(), gt.data_ptr(), None if window11 is None else self._host_f32(window11),
                   scores.data_ptr(), n, K, T, Cc, H, W, float(data_range), self._stream())

    def best_of_k(self, scores, samples, best_idx, mean_scores, best=None):
        """scores (N, K, T) -> mean_scores (N, K), best_idx int32 (N,), and (optional) the winner's frames of samples (N, K, ...)."""
        _chk(scores, torch.float32, "scores"); _chk(mean_scores, torch.float32, "mean_scores"); _chk(best_idx, torch.int32, "best_idx")
        _chk(samples, torch.float32, "samples"); _chk(best, torch.float32, "best")
        n, K, T = scores.shape
        clip_elems = 0 if samples is None else samples.numel() // (n * K)
        assert tuple(mean_scores.shape) == (n, K) and tuple(best_idx.shape) == (n,)
        assert best is None or (samples is not None and best.numel() == n * clip_elems)
        self._call("npvp_best_of_k", scores.data_ptr(), _ptr(samples), n, K, T, clip_elems, best_idx.data_ptr(), mean_scores.data_ptr(),
                   _ptr(best), self._stream())

    def ffn_stats_finalize(self, partial, stats, elems_per_frame):
        _chk(partial, torch.float32, "partial"); _chk(stats, torch.float32, "stats")
        frames, P = partial.shape[0], partial.shape[1]
        assert partial.shape == (frames, P, 2) and stats.shape == (frames, 2)
        self._call("npvp_ffn_stats_finalize", partial.data_ptr(), P, stats.data_ptr(), frames, int(elems_per_frame), self._stream())

    def ffn_dwconv(self, h, stats1, n1w, n1b, dw_w, dw_b, y, partial2):
        _chk(h, torch.bfloat16, "h"); _chk(y, torch.bfloat16, "y")
        for t, n in ((stats1, "stats1"), (n1w, "n1w"), (n1b, "n1b"), (dw_w, "dw_w"), (dw_b, "dw_b"), (partial2, "partial2")):
            _chk(t, torch.float32, n)
        frames, Ch = stats1.shape[0], h.shape[-1]
        assert partial2.shape == (frames, Ch // FFN_CHUNK, 2) and dw_w.shape == (9, Ch) and n1w.shape == (64, Ch)
        self._call("npvp_ffn_dwconv", h.data_ptr(), stats1.data_ptr(), n1w.data_ptr(), n1b.data_ptr(), dw_w.data_ptr(),
                   dw_b.data_ptr(), y.data_ptr(), partial2.data_ptr(), frames, Ch, self._stream())

    def ffn_norm2(self, y, partial2, n2w, n2b, out):
        _chk(y, torch.bfloat16, "y"); _chk(out, torch.bfloat16, "out")
        for t, n in ((partial2, "partial2"), (n2w, "n2w"), (n2b, "n2b")):
            _chk(t, torch.float32, n)
        frames, Ch = partial2.shape[0], y.shape[-1]
        self._call("npvp_ffn_norm2", y.data_ptr(), partial2.data_ptr(), n2w.data_ptr(), n2b.data_ptr(), out.data_ptr(), frames,
                   Ch, self._stream())

    def ffn_mid16_lanes(self):
        """Frame lanes (32 blocks each) of the single-pass conv-FFN middle resident at once (0: the kernel does not fit)."""
        return int(self.lib.npvp_ffn_mid16_lanes())

    def ffn_mid16(self, h1, part1, ln_wb, dw_w, dw_b, out, xch, cnt):
        """out = GELU(LN2(dw3x3(GELU(LN1(h1))) + b)), half in / half out, one pass (see include/npvp_b200.h).
        ``xch`` / ``cnt``: persistent exchange scratch from :func:`ffn_mid16_scratch`."""
        _chk(h1, torch.float16, "h1"); _chk(out, torch.float16, "out"); _chk(part1, torch.float32, "part1")
        _chk(ln_wb, torch.float16, "ln_wb"); _chk(dw_w, torch.float16, "dw_w"); _chk(dw_b, torch.float16, "dw_b")
        _chk(xch, torch.float32, "xch"); _chk(cnt, torch.int32, "cnt")
        frames, Ch = part1.shape[0], h1.shape[-1]
        assert h1.numel() == frames * 64 * Ch and out.numel() == h1.numel() and tuple(part1.shape) == (frames, 32, 2)
        assert tuple(ln_wb.shape) == (2, 64, Ch // 2, 2, 2) and tuple(dw_w.shape) == (9, Ch) and dw_b.numel() == Ch
        assert xch.numel() >= frames * 64 and cnt.numel() >= frames and out.data_ptr() != h1.data_ptr()
        self._call("npvp_ffn_mid16", h1.data_ptr(), part1.data_ptr(), ln_wb.data_ptr(), dw_w.data_ptr(), dw_b.data_ptr(), out.data_ptr(),
                   xch.data_ptr(), cnt.data_ptr(), frames, Ch, self._stream())

    def attention(self, q, k, v, out, mode, n_clips, Tq, Tk, mask_last=False):
        for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
            _chk(t, torch.bfloat16, n, False)
            assert t.shape[1] == 512
        assert q.shape[0] == n_clips * Tq * 64 and k.shape[0] == n_clips * Tk * 64 and v.shape[0] == k.shape[0]
        self._call("npvp_attention", q.data_ptr(), _rowmajor(q, "q"), k.data_ptr(), _rowmajor(k, "k"), v.data_ptr(),
                   _rowmajor(v, "v"), out.data_ptr(), _rowmajor(out, "out"), int(mode), n_clips, int(Tq), int(Tk),
                   int(bool(mask_last)), self._stream())

    def dwconv3x3_tokens(self, x, w, shift, out, relu=True):
        _chk(x, torch.float32, "x"); _chk(w, torch.float32, "w"); _chk(shift, torch.float32, "shift"); _chk(out, torch.bfloat16, "out")
        Cc = x.shape[-1]
        frames = x.numel() // (64 * Cc)
        assert w.shape == (9, Cc) and out.numel() == x.numel()
        self._call("npvp_dwconv3x3_tokens", x.data_ptr(), w.data_ptr(), shift.data_ptr(), out.data_ptr(), frames, Cc, int(relu),
                   self._stream())

    def latent_reparam(self, mulv, eps_nchw, z, n_clips, Cc):
        _chk(mulv, torch.float32, "mulv", False); _chk(eps_nchw, torch.float32, "eps"); _chk(z, torch.float32, "z")
        assert mulv.shape[0] == n_clips * 64 and z.numel() == n_clips * 64 * Cc
        assert eps_nchw is None or eps_nchw.numel() == z.numel()
        self._call("npvp_latent_reparam", mulv.data_ptr(), _rowmajor(mulv, "mulv"), _ptr(eps_nchw), z.data_ptr(), n_clips, Cc,
                   self._stream())

    # -- layouts --------------------------------------------------------------------------------
    def nchw_to_tokens(self, x, out_f32=None, out_bf16=None):
        """x: (frames, C, HW) fp32 -> (frames, HW, C)."""
        _chk(x, torch.float32, "x"); _chk(out_f32, torch.float32, "out_f32"); _chk16(out_bf16, "out_bf16")
        frames, Cc, HW = x.shape
        self._call("npvp_nchw_to_tokens", x.data_ptr(), _ptr(out_f32), _ptr(out_bf16), frames, Cc, HW,
                   0 if out_bf16 is None else _is_fp16(out_bf16), self._stream())

    def tokens_to_nchw(self, x, out, relu=False):
        """x: (frames, HW, C) fp32 or bf16 -> out (frames, C, HW) fp32."""
        assert x.is_cuda and x.is_contiguous() and x.dtype in (torch.float32,) + H16
        _chk(out, torch.float32, "out")
        frames, HW, Cc = x.shape
        xf, xb = (x.data_ptr(), None) if x.dtype == torch.float32 else (None, x.data_ptr())
        self._call("npvp_tokens_to_nchw", xf, xb, out.data_ptr(), frames, Cc, HW, int(relu), _is_fp16(x), self._stream())

    # -- autoencoder ----------------------------------------------------------------------------
    def conv7x7_stem(self, x, w, shift, out, Cin, Cout, H, W, norm=None):
        """x: fp32 frames (frames, Cin, H, W), or uint8 pixels with ``norm`` = (mean, std) of VidNormalize: the stem normalises."""
        _chk(w, torch.float32, "w"); _chk(shift, torch.float32, "shift"); _chk16(out, "out")
        u8 = x.dtype == torch.uint8
        if u8:
            assert x.is_cuda and x.is_contiguous() and norm is not None and len(norm[0]) == Cin and len(norm[1]) == Cin
            mean, std = self._host_f32(norm[0]), self._host_f32(norm[1])
        else:
            _chk(x, torch.float32, "x")
            mean = std = None
        esz = 1 if u8 else 4
        frames = x.numel() // (Cin * H * W)
        assert out.numel() == frames * H * W * Cout and w.shape == (49 * Cin, Cout)
        per = max(1, 65535 // (Cout // 32))
        for f0 in range(0, frames, per):      # grid.z limit
            n = min(per, frames - f0)
            ptr = x.data_ptr() + f0 * Cin * H * W * esz
            self._call("npvp_conv7x7_stem", None if u8 else ptr, w.data_ptr(), shift.data_ptr(),
                       out.data_ptr() + f0 * H * W * Cout * 2, n, Cin, Cout, H, W, _is_fp16(out), ptr if u8 else None, mean, std, self._stream())

    def conv7x7_head(self, x, w, bias, out, Cin, Cout, H, W, phase_major, act, out_u8=None, renorm=None):
        """w: 16-bit weights packed by :func:`pack_head_weights` ([Cin/32, 14, NT, 32, 4]).  ``out`` fp32 model-space frames and / or
        ``out_u8`` pixel-space uint8 frames (``renorm`` = (mean, std) of VidReNormalize; fused VidReNormalize + clamp + uint8)."""
        _chk16(x, "x"); _chk16(w, "w", like=x); _chk(bias, torch.float32, "bias"); _chk(out, torch.float32, "out")
        frames = x.numel() // (Cin * H * W)
        assert out is not None or out_u8 is not None
        assert out is None or out.numel() == frames * Cout * H * W
        assert tuple(w.shape) == (Cin // 32, 14, head_n_tiles(Cout), 32, 4)
        inv_std = inv_mean = None
        if out_u8 is not None:
            assert out_u8.is_cuda and out_u8.is_contiguous() and out_u8.dtype == torch.uint8 and out_u8.numel() == frames * Cout * H * W
            assert renorm is not None and len(renorm[0]) == Cout and len(renorm[1]) == Cout
            inv_std = self._host_f32([1.0 / float(v) for v in renorm[1]])
            inv_mean = self._host_f32([-float(v) for v in renorm[0]])
        for f0 in range(0, frames, 65535):
            n = min(65535, frames - f0)
            self._call("npvp_conv7x7_head", x.data_ptr() + f0 * Cin * H * W * 2, w.data_ptr(), bias.data_ptr(),
                       None if out is None else out.data_ptr() + f0 * Cout * H * W * 4, n, Cin, Cout, H, W, int(phase_major), int(act), _is_fp16(x),
                       None if out_u8 is None else out_u8.data_ptr() + f0 * Cout * H * W, inv_std, inv_mean, self._stream())

    def maxpool2x2_cols(self, x, col0, Cn, out, frames, H, W):
        _chk16(x, "x", False); _chk16(out, "out", like=x)
        assert x.shape[0] == frames * H * W and out.shape == (frames * (H // 2) * (W // 2), Cn)
        self._call("npvp_maxpool2x2_cols", x.data_ptr(), _rowmajor(x, "x"), col0, Cn, out.data_ptr(), frames, H, W, _is_fp16(x),
                   self._stream())

    def nonlocal_attention(self, q, kv, out, frames, HW, HWk, dq, dv):
        _chk16(q, "q", False); _chk16(kv, "kv", like=q); _chk16(out, "out", like=q)
        assert q.shape[0] == frames * HW and kv.shape == (frames * HWk, dq + dv) and out.shape == (frames * HW, dv)
        self._call("npvp_nonlocal_attention", q.data_ptr(), _rowmajor(q, "q"), kv.data_ptr(), out.data_ptr(), frames, HW, HWk, dq,
                   dv, _is_fp16(q), self._stream())


def _head_frag_index(device) -> torch.Tensor:
    """k indices held by thread-in-group ``tig`` of an m16n8k16 B fragment: b0 = (2t, 2t+1), b1 = (2t+8, 2t+9)."""
    t = torch.arange(4, device=device)
    return torch.stack([2 * t, 2 * t + 1, 2 * t + 8, 2 * t + 9], dim=1)


def head_n_tiles(Cout: int) -> int:
    return (7 * Cout + 7) // 8


def pack_head_weights(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """[(ky,kx,ci), Cout] fp32 -> 16-bit mma.sync B fragments [Cin/32 passes, 14 k-steps, NT n-tiles, 32 lanes, 4].

    The head kernel computes Z[x', (kx,co)] = sum_{ky,ci} X[oy+ky, x', ci] W[ky,kx,ci,co]: GEMM column n = kx*Cout + co
    (zero-padded to NT*8), k-step s = (ky, 16-channel half) within a 32-channel pass.  Lane (gid, tig) of n-tile j holds
    B[k, n = 8j + gid] for k in (2tig, 2tig+1, 2tig+8, 2tig+9)."""
    K, Cout = w.shape
    Cin = K // 49
    assert Cin % 32 == 0 and 1 <= Cout <= 3
    P, NT = Cin // 32, head_n_tiles(Cout)
    full = torch.zeros(7, Cin, NT * 8, dtype=torch.float32, device=w.device)
    full[:, :, :7 * Cout] = w.reshape(7, 7, Cin, Cout).permute(0, 2, 1, 3).reshape(7, Cin, 7 * Cout)
    b5 = full.reshape(7, P, 2, 16, NT * 8).permute(1, 0, 2, 3, 4).reshape(P, 14, 16, NT, 8)
    frag = b5[:, :, _head_frag_index(w.device), :, :]                       # [P, 14, tig, e, NT, gid]
    return frag.permute(0, 1, 4, 5, 2, 3).reshape(P, 14, NT, 32, 4).to(dtype).contiguous()


def unpack_head_weights(p: torch.Tensor, Cout: int) -> torch.Tensor:
    """Inverse of :func:`pack_head_weights` (used by the kernel specification)."""
    P, _, NT, _, _ = p.shape
    frag = p.float().reshape(P, 14, NT, 8, 4, 4).permute(0, 1, 4, 5, 2, 3)  # [P, 14, tig, e, NT, gid]
    b5 = torch.zeros(P, 14, 16, NT, 8, dtype=torch.float32, device=p.device)
    b5[:, :, _head_frag_index(p.device), :, :] = frag
    full = b5.reshape(P, 7, 2, 16, NT * 8).permute(1, 0, 2, 3, 4).reshape(7, P * 32, NT * 8)[:, :, :7 * Cout]
    return full.reshape(7, P * 32, 7, Cout).permute(0, 2, 1, 3).reshape(49 * P * 32, Cout).contiguous()


def _pos_frames(beta, gamma, n_clips, T):
    """Frames covered by the positional code: T (timestamps shared by the batch) or n_clips * T (per-clip timestamps)."""
    if beta is None:
        return 0
    frames = beta.numel() // (64 * 512)
    assert beta.numel() == frames * 64 * 512 and frames in (T, n_clips * T), \
        f"positional code covers {frames} frames; expected T = {T} or n_clips * T = {n_clips * T}"
    assert gamma is None or gamma.numel() == beta.numel()
    return frames


FFN_CHUNK = 128        # channels per partial-statistics chunk of the conv-FFN middle (kFfnChunk in predictor_kernels.cu)


def ffn_mid16_scratch(frames: int, device) -> tuple:
    """Persistent exchange scratch of :meth:`Ops.ffn_mid16` at rest: (xch fp32 [frames,32,2] of all-ones bit patterns, cnt int32
    [frames] zeros).  Every call returns it to this state, so one pair serves all launches of a stream (the launches are
    stream-ordered; concurrent streams need separate pairs)."""
    xch = torch.full((frames, 32, 2), -1, dtype=torch.int32, device=device).view(torch.float32)
    return xch, torch.zeros(frames, dtype=torch.int32, device=device)

_OPS: Optional[Ops] = None


def ops() -> Ops:
    """The process-wide kernel binding.  Tests may replace it with ``set_ops`` (kernel-spec emulation on CPU)."""
    global _OPS
    if _OPS is None:
        _OPS = Ops()
    return _OPS


def set_ops(obj) -> None:
    global _OPS
    _OPS = obj
