"""Executable specification of every C-ABI entry point, in plain torch (TEST INFRASTRUCTURE ONLY).

``SpecOps`` has the same method signatures as ``npvp_b200._lib.Ops`` and the same rounding points
(bf16 where the kernels store bf16, fp32 statistics).  It is used
  * on the GPU box as the per-kernel reference the CUDA kernels are compared against, and
  * on CPU to run the host-side engines end to end (``_lib.set_ops(SpecOps())``) so that weight
    packing / buffer plumbing is validated against the oracle without a GPU.
It is never imported by the package.
"""
from __future__ import annotations

import math

import torch
from npvp_b200._lib import FFN_CHUNK
import torch.nn.functional as F

ACT_NONE, ACT_RELU, ACT_GELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
EPS = 1e-5

# the specification is fp32: keep cuDNN / cuBLAS from silently using TF32 when the specs run on a GPU
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _act(v, act):
    if act == ACT_RELU:
        return torch.relu(v)
    if act == ACT_GELU:
        return 0.5 * v * (1.0 + torch.erf(v * (1.0 / math.sqrt(2.0))))
    if act == ACT_TANH:
        return torch.tanh(v)
    if act == ACT_SIGMOID:
        return torch.sigmoid(v)
    return v


def _reflect(i, n):
    i = i.abs()
    return torch.where(i >= n, 2 * n - 2 - i, i)


class SpecOps:
    def __init__(self):
        self.launches = 0

    def launch_count(self):
        return self.launches

    def reset_launch_count(self):
        self.launches = 0

    # -- contractions ---------------------------------------------------------------------------
    def gemm(self, a, w, *, bias=None, act=ACT_NONE, alpha=1.0, res1=None, res2=None, out_f32=None, out_bf16=None,
             post_relu=False, backend=None, frame_stats=None):
        self.launches += 1
        assert a.dtype in (torch.bfloat16, torch.float16) and w.dtype == a.dtype
        v = a.float() @ w.float().t()
        if bias is not None:
            v = v + bias
        v = _act(v, act) * alpha
        if res1 is not None:
            v = v + res1.float()
        if res2 is not None:
            v = v + res2.float()
        if post_relu:
            v = torch.relu(v)
        if out_f32 is not None:
            out_f32.copy_(v)
        if out_bf16 is not None:
            out_bf16.copy_(v.clamp(-65504, 65504).to(out_bf16.dtype))
        if frame_stats is not None:
            # same slot layout as the kernel: slot = ((n_tile * 2 + column half) * 2 + row half), blocks of 32 rows x 128 columns
            M, N = v.shape
            blk = v.reshape(M // 64, 2, 32, N // 256, 2, 128).permute(0, 3, 4, 1, 2, 5).reshape(M // 64, 4 * N // 256, 32 * 128)
            frame_stats[:, :, 0] = blk.sum(-1)
            frame_stats[:, :, 1] = (blk * blk).sum(-1)

    def conv_gemm(self, x, w, frames, H, W, Cc, KH, KW, stride, pad, pad_mode, Ho, Wo, phase_major=False, **epi):
        col = torch.empty(frames * Ho * Wo, KH * KW * Cc, dtype=x.dtype, device=x.device)
        self.im2col(x, col, frames, H, W, Cc, KH, KW, stride, pad, pad_mode, Ho, Wo, phase_major)
        self.launches -= 1
        self.gemm(col, w, **epi)

    @staticmethod
    def convt_supported(H, W, Cin, Cout):
        from npvp_b200._lib import Ops
        return Ops.convt_supported(H, W, Cin, Cout)

    def convt_gemm(self, x, w, frames, H, W, Cin, Cout, *, bias=None, act=ACT_NONE, out_bf16=None):
        self.launches += 1
        img = F.pad(x.float().reshape(frames, H, W, Cin), (0, 0, 0, 1, 0, 1))                  # zero row / column past the edge
        nb = torch.stack([img[:, dy:dy + H, dx:dx + W, :] for dy in (0, 1) for dx in (0, 1)], dim=3)   # (f,H,W,(dy,dx),ci)
        v = nb.reshape(frames * H * W, 4 * Cin) @ w.float().t()                                # columns (q, co)
        if bias is not None:
            v = v + bias
        v = _act(v, act).reshape(frames, H, W, 4, Cout)
        o = torch.empty(frames, H, 2, W, 2, Cout, dtype=torch.float32, device=x.device)
        for q, (py, px) in enumerate(((0, 0), (0, 1), (1, 1), (1, 0))):
            o[:, :, py, :, px, :] = v[:, :, :, q, :]
        out_bf16.copy_(o.reshape(out_bf16.shape).clamp(-65504, 65504).to(out_bf16.dtype))

    def gemm_f32(self, a, w, bias, act, out):
        self.launches += 1
        v = a @ w.t()
        if bias is not None:
            v = v + bias
        out.copy_(_act(v, act))

    # -- predictor ------------------------------------------------------------------------------
    def fourier_features(self, coor, B, out):
        self.launches += 1
        proj = (2.0 * float(math.pi) * coor) @ B.t()
        out.copy_(torch.cat([torch.cos(proj), torch.sin(proj)], dim=-1))

    def ln_posfuse(self, x, ln_w, ln_b, qe, beta, gamma, out_ln, out_fused, n_clips, T):
        self.launches += 1
        a = x.reshape(n_clips, T, 64, 512).float()
        if ln_w is not None:
            a = F.layer_norm(a, (512,), ln_w, ln_b, EPS)
        if out_ln is not None:
            out_ln.copy_(a.reshape(out_ln.shape).to(torch.bfloat16))
        if out_fused is None:
            return
        u = a
        if qe is not None:
            u = u + qe.reshape(n_clips, 1, 64, 512)
        flat = u.reshape(n_clips, T, -1)
        mu = flat.mean(-1, keepdim=True)
        var = ((flat - mu) ** 2).mean(-1, keepdim=True)
        nrm = ((flat - mu) * torch.rsqrt(var + EPS)).reshape(n_clips, T, 64, 512)
        nb = beta.numel() // (T * 64 * 512)          # 1: timestamps shared by the batch; n_clips: per-clip timestamps
        assert nb in (1, n_clips)
        g = 0.0 if gamma is None else gamma.reshape(nb, T, 64, 512)
        out_fused.copy_((nrm * (1.0 + g) + beta.reshape(nb, T, 64, 512)).reshape(out_fused.shape).to(torch.bfloat16))

    def add_layernorm_rows(self, x, delta, w, b, out_f32=None, out_bf16=None, relu=False):
        x.add_(delta.float().reshape(x.shape))
        self.layernorm_rows(x, w, b, out_f32, out_bf16, relu)

    def add_ln_posfuse(self, x, delta, ln_w, ln_b, qe, beta, gamma, out_ln, out_fused, n_clips, T):
        x.add_(delta.float().reshape(x.shape))
        self.ln_posfuse(x, ln_w, ln_b, qe, beta, gamma, out_ln, out_fused, n_clips, T)

    def layernorm_rows(self, x, w, b, out_f32=None, out_bf16=None, relu=False):
        self.launches += 1
        v = F.layer_norm(x.reshape(-1, 512).float(), (512,), w, b, EPS)
        if relu:
            v = torch.relu(v)
        if out_f32 is not None:
            out_f32.copy_(v.reshape(out_f32.shape))
        if out_bf16 is not None:
            out_bf16.copy_(v.reshape(out_bf16.shape).to(out_bf16.dtype))

    def frame_ln_gelu_residual(self, h, w_hwc, b_hwc, y):
        self.launches += 1
        hh = h.reshape(-1, 64 * 512).float()
        mu = hh.mean(-1, keepdim=True)
        var = ((hh - mu) ** 2).mean(-1, keepdim=True)
        v = (hh - mu) * torch.rsqrt(var + EPS) * w_hwc.reshape(1, -1) + b_hwc.reshape(1, -1)
        y.add_(_act(v, ACT_GELU).reshape(y.shape))

    def frame_ln_gelu_residual_posfuse(self, h, w_hwc, b_hwc, y, ln_w, ln_b, qe, beta, gamma, out_ln, out_fused, n_clips, T):
        self.frame_ln_gelu_residual(h, w_hwc, b_hwc, y)
        self.launches -= 1
        self.ln_posfuse(y, ln_w, ln_b, qe, beta, gamma, out_ln, out_fused, n_clips, T)

    def temporal_mean(self, mem, evt, n_clips, T):
        self.launches += 1
        evt.copy_(mem.reshape(n_clips, T, -1).sum(1).mul(1.0 / T).reshape(evt.shape))

    def ffn_frame_stats(self, h, stats):
        self.launches += 1
        frames = stats.shape[0]
        hh = h.reshape(frames, -1).double()
        mean = hh.mean(-1)
        var = (hh * hh).mean(-1) - mean * mean
        stats[:, 0] = mean.float()
        stats[:, 1] = (1.0 / torch.sqrt(var.clamp_min(0) + EPS)).float()

    def ffn_stats_finalize(self, partial, stats, elems_per_frame):
        self.launches += 1
        s = partial[:, :, 0].double().sum(-1) / elems_per_frame
        q = partial[:, :, 1].double().sum(-1) / elems_per_frame
        stats[:, 0] = s.float()
        stats[:, 1] = (1.0 / torch.sqrt((q - s * s).clamp_min(0) + EPS)).float()

    def ffn_dwconv(self, h, stats1, n1w, n1b, dw_w, dw_b, y, partial2):
        self.launches += 1
        frames, Ch = stats1.shape[0], h.shape[-1]
        hh = h.reshape(frames, 64, Ch).float()
        a = _act((hh - stats1[:, 0].reshape(-1, 1, 1)) * stats1[:, 1].reshape(-1, 1, 1) * n1w + n1b, ACT_GELU)
        img = a.reshape(frames, 8, 8, Ch).permute(0, 3, 1, 2)
        wt = dw_w.reshape(3, 3, Ch).permute(2, 0, 1).unsqueeze(1)
        o = F.conv2d(img, wt, dw_b, padding=1, groups=Ch).permute(0, 2, 3, 1).reshape(frames, 64, Ch)
        ob = o.to(torch.bfloat16)
        y.copy_(ob.reshape(y.shape))
        of = o.reshape(frames, 64, Ch // FFN_CHUNK, FFN_CHUNK)      # statistics of the fp32 conv output (before rounding to bf16)
        partial2[:, :, 0] = of.sum(dim=(1, 3))
        partial2[:, :, 1] = (of * of).sum(dim=(1, 3))

    def ffn_norm2(self, y, partial2, n2w, n2b, out):
        self.launches += 1
        frames, Ch = partial2.shape[0], y.shape[-1]
        n = 64.0 * Ch
        s = partial2[:, :, 0].double().sum(-1)
        q = partial2[:, :, 1].double().sum(-1)
        mean = s / n
        rstd = 1.0 / torch.sqrt((q / n - mean * mean).clamp_min(0) + EPS)
        yy = y.reshape(frames, 64, Ch).float()
        v = (yy - mean.float().reshape(-1, 1, 1)) * rstd.float().reshape(-1, 1, 1) * n2w + n2b
        out.copy_(_act(v, ACT_GELU).reshape(out.shape).to(torch.bfloat16))

    def ffn_mid16_lanes(self):
        return 9

    def ffn_mid16(self, h1, part1, ln_wb, dw_w, dw_b, out, xch, cnt):
        """Single-pass conv-FFN middle, fp32 restatement of the half-precision kernel: LN1 statistics from the fc1 partials,
        GELU(LN1) -> dw3x3 -> statistics of the fp32 conv output -> GELU(LN2) -> half.  (The kernel's element-wise math is
        half2; the GPU test compares with a tolerance that covers it.)"""
        self.launches += 1
        frames, Ch = part1.shape[0], h1.shape[-1]
        n = 64.0 * Ch
        s, q = part1[:, :, 0].double().sum(-1) / n, part1[:, :, 1].double().sum(-1) / n
        mean1, rstd1 = s.float(), (1.0 / torch.sqrt((q - s * s).clamp_min(0) + EPS)).float()
        wb = ln_wb.float()                                             # [2, 64, Ch/2, (w|b), 2]
        w1, b1, w2, b2 = (wb[i, :, :, j, :].reshape(64, Ch) for i in (0, 1) for j in (0, 1))
        hh = h1.reshape(frames, 64, Ch).float()
        a = _act((hh - mean1.reshape(-1, 1, 1)) * rstd1.reshape(-1, 1, 1) * w1 + b1, ACT_GELU)
        img = a.reshape(frames, 8, 8, Ch).permute(0, 3, 1, 2)
        wt = dw_w.float().reshape(3, 3, Ch).permute(2, 0, 1).unsqueeze(1)
        o = F.conv2d(img, wt, dw_b.float().reshape(-1), padding=1, groups=Ch).permute(0, 2, 3, 1).reshape(frames, 64, Ch)
        od = o.double()
        mean2 = od.mean(dim=(1, 2))
        rstd2 = 1.0 / torch.sqrt(((od * od).mean(dim=(1, 2)) - mean2 * mean2).clamp_min(0) + EPS)
        v = (o - mean2.float().reshape(-1, 1, 1)) * rstd2.float().reshape(-1, 1, 1) * w2 + b2
        out.copy_(_act(v, ACT_GELU).reshape(out.shape).to(torch.float16))

    def attention(self, q, k, v, out, mode, n_clips, Tq, Tk, mask_last=False):
        self.launches += 1
        H, D = 8, 64
        if mode == 0:
            Fr = n_clips * Tq

            def win(t):   # (Fr*64, 512) -> (Fr*4, heads, 16, D)
                t = t.float().reshape(Fr, 2, 4, 2, 4, H, D).permute(0, 1, 3, 5, 2, 4, 6)
                return t.reshape(Fr * 4, H, 16, D)
            Q, K, V = win(q), win(k), win(v)
            s = (Q * 0.125) @ K.transpose(-1, -2)
            o = torch.softmax(s, -1) @ V                                  # (Fr*4, H, 16, D)
            o = o.reshape(Fr, 2, 2, H, 4, 4, D).permute(0, 1, 4, 2, 5, 3, 6).reshape(Fr * 64, 512)
        else:
            def seq(t, T):  # (n*T*64, 512) -> (n*64, heads, T, D)
                return t.float().reshape(n_clips, T, 64, H, D).permute(0, 2, 3, 1, 4).reshape(n_clips * 64, H, T, D)
            Q, K, V = seq(q, Tq), seq(k, Tk), seq(v, Tk)
            s = (Q * 0.125) @ K.transpose(-1, -2)
            if mask_last:
                m = torch.zeros(Tq, Tk, dtype=torch.bool, device=s.device)
                m[0:Tq - 1, Tk - 1] = True
                s = s.masked_fill(m, float("-inf"))
            o = torch.softmax(s, -1) @ V
            o = o.reshape(n_clips, 64, H, Tq, D).permute(0, 3, 1, 2, 4).reshape(n_clips * Tq * 64, 512)
        out.copy_(o.to(torch.bfloat16))

    def dwconv3x3_tokens(self, x, w, shift, out, relu=True):
        self.launches += 1
        Cc = x.shape[-1]
        img = x.reshape(-1, 8, 8, Cc).permute(0, 3, 1, 2).float()
        wt = w.reshape(3, 3, Cc).permute(2, 0, 1).unsqueeze(1)
        o = F.conv2d(img, wt, shift, padding=1, groups=Cc).permute(0, 2, 3, 1)
        if relu:
            o = torch.relu(o)
        out.copy_(o.reshape(out.shape).to(torch.bfloat16))

    def latent_reparam(self, mulv, eps_nchw, z, n_clips, Cc):
        self.launches += 1
        mu = mulv[:, :Cc].reshape(n_clips, 64, Cc)
        if eps_nchw is None:
            z.copy_(mu.reshape(z.shape))
            return
        lv = mulv[:, Cc:2 * Cc].reshape(n_clips, 64, Cc)
        eps = eps_nchw.reshape(n_clips, Cc, 64).permute(0, 2, 1)
        z.copy_((mu + torch.exp(0.5 * lv) * eps).reshape(z.shape))

    # -- layouts --------------------------------------------------------------------------------
    def nchw_to_tokens(self, x, out_f32=None, out_bf16=None):
        self.launches += 1
        t = x.permute(0, 2, 1)
        if out_f32 is not None:
            out_f32.copy_(t.reshape(out_f32.shape))
        if out_bf16 is not None:
            out_bf16.copy_(t.reshape(out_bf16.shape).to(out_bf16.dtype))

    def tokens_to_nchw(self, x, out, relu=False):
        self.launches += 1
        t = x.float().permute(0, 2, 1)
        out.copy_((torch.relu(t) if relu else t).reshape(out.shape))

    # -- autoencoder ----------------------------------------------------------------------------
    @staticmethod
    def _unphase(x, frames, H, W, Cc, phase_major):
        if not phase_major:
            return x.reshape(frames, H, W, Cc)
        return x.reshape(frames, H // 2, W // 2, 2, 2, Cc).permute(0, 1, 3, 2, 4, 5).reshape(frames, H, W, Cc)

    def conv7x7_stem(self, x, w, shift, out, Cin, Cout, H, W, norm=None):
        self.launches += 1
        if x.dtype == torch.uint8:                 # VidToTensor + VidNormalize, reference operation order
            m = torch.tensor([float(v) for v in norm[0]], dtype=torch.float32, device=x.device).view(Cin, 1, 1)
            sdv = torch.tensor([float(v) for v in norm[1]], dtype=torch.float32, device=x.device).view(Cin, 1, 1)
            x = (x.reshape(-1, Cin, H, W).to(torch.float32).div(255) - m) / sdv
        img = x.reshape(-1, Cin, H, W)
        wt = w.reshape(7, 7, Cin, Cout).permute(3, 2, 0, 1)
        with torch.backends.cudnn.flags(enabled=False):     # see conv7x7_head
            o = F.conv2d(F.pad(img, (3, 3, 3, 3), mode="reflect"), wt, shift)
        out.copy_(torch.relu(o).permute(0, 2, 3, 1).reshape(out.shape).to(out.dtype))

    def conv7x7_head(self, x, w, bias, out, Cin, Cout, H, W, phase_major, act, out_u8=None, renorm=None):
        self.launches += 1
        frames = x.numel() // (Cin * H * W)
        img = self._unphase(x.float(), frames, H, W, Cin, phase_major).permute(0, 3, 1, 2)
        from npvp_b200._lib import unpack_head_weights
        wt = unpack_head_weights(w, Cout).reshape(7, 7, Cin, Cout).permute(3, 2, 0, 1)   # 16-bit packed weights (mma B fragments)
        with torch.backends.cudnn.flags(enabled=False):     # cuDNN finds no engine for some thin frames (H = 4, W = 250); the native kernel always works
            o = _act(F.conv2d(F.pad(img, (3, 3, 3, 3), mode="reflect"), wt, bias), act)
        if out is not None:
            out.copy_(o.reshape(out.shape))
        if out_u8 is not None:                      # VidReNormalize + clamp + ToPILImage (truncation), reference operation order
            inv_std = torch.tensor([1.0 / float(v) for v in renorm[1]], dtype=torch.float32, device=o.device).view(1, Cout, 1, 1)
            inv_mean = torch.tensor([-float(v) for v in renorm[0]], dtype=torch.float32, device=o.device).view(1, Cout, 1, 1)
            p = ((o / inv_std) - inv_mean).clamp(0.0, 1.0)
            out_u8.copy_(p.mul(255).to(torch.uint8).reshape(out_u8.shape))

    def im2col(self, x, out, frames, H, W, Cc, KH, KW, stride, pad, pad_mode, Ho, Wo, phase_major=False):
        self.launches += 1
        img = self._unphase(x, frames, H, W, Cc, phase_major)
        dev = x.device
        oy = torch.arange(Ho, device=dev).view(Ho, 1, 1, 1) * stride - pad + torch.arange(KH, device=dev).view(1, 1, KH, 1)
        ox = torch.arange(Wo, device=dev).view(1, Wo, 1, 1) * stride - pad + torch.arange(KW, device=dev).view(1, 1, 1, KW)
        oy, ox = oy.expand(Ho, Wo, KH, KW), ox.expand(Ho, Wo, KH, KW)
        if pad_mode == 1:
            iy, ix, valid = _reflect(oy, H), _reflect(ox, W), None
        elif pad_mode == 2:
            iy, ix, valid = oy.clamp(0, H - 1), ox.clamp(0, W - 1), None
        else:
            valid = (oy >= 0) & (oy < H) & (ox >= 0) & (ox < W)
            iy, ix = oy.clamp(0, H - 1), ox.clamp(0, W - 1)
        g = img[:, iy, ix, :]                                     # (frames,Ho,Wo,KH,KW,C)
        if valid is not None:
            g = g * valid.view(1, Ho, Wo, KH, KW, 1).to(g.dtype)
        out.copy_(g.reshape(out.shape))

    def maxpool2x2_cols(self, x, col0, Cn, out, frames, H, W):
        self.launches += 1
        t = x[:, col0:col0 + Cn].float().reshape(frames, H // 2, 2, W // 2, 2, Cn)
        out.copy_(t.amax(dim=(2, 4)).reshape(out.shape).to(out.dtype))

    def nonlocal_attention(self, q, kv, out, frames, HW, HWk, dq, dv):
        self.launches += 1
        Q = q[:, :dq].float().reshape(frames, HW, dq)
        K = kv[:, :dq].float().reshape(frames, HWk, dq)
        V = kv[:, dq:].float().reshape(frames, HWk, dv)
        a = torch.softmax(Q @ K.transpose(1, 2), dim=-1)
        out.copy_((a @ V).reshape(out.shape).to(out.dtype))
