"""Host-side drivers of the ResNet frame encoder / decoder (reference: models/ResNetAutoEncoder.py:51-261,
models/submodules.py:9-180).

Data layout in HBM: activations are 16-bit channels-last [frames, H, W, C]; every 3x3 / strided / transposed
convolution is an implicit GEMM (``npvp_conv_gemm_bf16``: the patch gather happens inside the tensor-core kernel) whose
epilogue applies the folded eval-mode BatchNorm, bias, ReLU, the non-local gamma and up to two residual adds.  A transposed conv
(3x3, stride 2, pad 1, output_pad 1) is computed as ONE GEMM over the 2x2 input neighbourhood with N = 4*Cout
enumerating the four output phases (9 live taps out of 16 blocks), leaving its output "phase-major"
[frames, H, W, (py,px), Cout]; the next layer's gather reads that layout directly, so no pixel shuffle runs.
The 7x7 stem / head convolutions (Cin or Cout in {1,3}) are direct kernels.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, PAD_REFLECT, PAD_REPLICATE, PAD_ZERO

from .engine_predictor import _bn_fold, _f
from .workspace import Workspace

_PAD = {"reflect": PAD_REFLECT, "replicate": PAD_REPLICATE, "zero": PAD_ZERO}


def ae_dtype() -> torch.dtype:
    """16-bit storage type of the autoencoder activations / weights.  IEEE half by default: the decoder's rounding
    errors land directly in pixels (x dataset std up to 2.76) and half has 8x the mantissa resolution of bfloat16
    (measured: decoder max error 6.9e-3 -> 9.1e-4 under stress init); activations are O(1..30) behind BatchNorm and
    stores saturate at +-65504.  ``NPVP_B200_AE_DTYPE=bf16`` switches back."""
    return torch.bfloat16 if os.environ.get("NPVP_B200_AE_DTYPE", "fp16") == "bf16" else torch.float16


def _h(t, dt):
    return t.detach().to(dt).contiguous()


def _pack_conv3x3(conv: nn.Conv2d, bn: nn.BatchNorm2d, dt):
    """[Cout,Cin,3,3] + BN -> bf16 [Cout, (ky,kx,ci)] with the BN scale folded, fp32 bias = shift (+ scale*conv bias)."""
    scale, shift = _bn_fold(bn)
    w = conv.weight.detach().float() * scale[:, None, None, None]
    if conv.bias is not None:
        shift = shift + scale * conv.bias.detach().float()
    return _h(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1), dt), shift.contiguous()


def _pack_conv7x7(conv: nn.Conv2d, scale=None):
    w = conv.weight.detach().float()
    if scale is not None:
        w = w * scale[:, None, None, None]
    return w.permute(2, 3, 1, 0).reshape(-1, w.shape[0]).contiguous()          # [(ky,kx,ci), Cout]


class _F3D:
    """Factorized3DConvAttn (learn_3d=False): 3x3 conv + BN + ReLU + skip, non-local attention, outer skip."""

    def __init__(self, m, dt):
        self.C = m.in_channels
        self.wc, self.bc = _pack_conv3x3(m.spatial_conv[0], m.spatial_conv[1], dt)
        a = m.attn2d
        self.dq, self.dv = a.attn_dim, a.value_dim
        self.wqkv = _h(torch.cat([a.Wq.weight, a.Wk.weight, a.Wv.weight], 0), dt)
        self.bqkv = _f(torch.cat([a.Wq.bias, a.Wk.bias, a.Wv.bias], 0))
        scale, shift = _bn_fold(a.norm_func)
        self.wo = _h(a.out_proj.weight.detach().float() * scale[:, None], dt)
        self.bo = (shift + scale * a.out_proj.bias.detach().float()).contiguous()
        self.gamma = float(a.gamma.detach()) if isinstance(a.gamma, torch.Tensor) else float(a.gamma)


class EncoderEngine:
    def __init__(self, mod):
        self.mod = mod
        self.device = next(mod.parameters()).device
        self.ws = Workspace(self.device)
        self.dt = dt = ae_dtype()
        self.cin = mod.input_nc
        scale, shift = _bn_fold(mod.block0[2])
        self.stem_w, self.stem_shift = _pack_conv7x7(mod.block0[1], scale), shift.contiguous()
        self.ngf = mod.block0[1].weight.shape[0]
        self.down = [_pack_conv3x3(mod.block1[0], mod.block1[1], dt)]
        self.f3d = []
        for i in range(1, mod.n_downsampling):
            self.f3d.append(_F3D(getattr(mod, f'block{i + 1}_3dConvAttn'), dt))
            seq = getattr(mod, f'block{i + 1}_conv')
            self.down.append(_pack_conv3x3(seq[0], seq[1], dt))
        self.res = []
        for i in range(mod.num_res_blocks):
            blk = getattr(mod, f'res_conv_{i}')
            (c1, n1), (c2, n2) = blk.convs()
            self.res.append((_F3D(getattr(mod, f'res_3dConvAttn_{i}'), dt), _pack_conv3x3(c1, n1, dt), _pack_conv3x3(c2, n2, dt),
                             _PAD[blk.padding_type]))

    # x: bf16 [frames*H*W, C]
    def _conv3x3(self, x, frames, H, W, C, wb, stride, pad_mode, tag, **epi):
        op, ws = _lib.ops(), self.ws
        Ho, Wo = H // stride, W // stride
        w, b = wb
        if "out_f32" not in epi:
            epi["out_bf16"] = ws.h16(tag, self.dt, frames * Ho * Wo, w.shape[0])
        op.conv_gemm(x, w, frames, H, W, C, 3, 3, stride, 1, pad_mode, Ho, Wo, bias=b, **epi)   # A gathered inside the GEMM
        return epi.get("out_bf16", epi.get("out_f32"))

    def _f3d(self, x, frames, H, W, p: _F3D, tag):
        op, ws = _lib.ops(), self.ws
        M, C = frames * H * W, p.C
        y = self._conv3x3(x, frames, H, W, C, (p.wc, p.bc), 1, PAD_ZERO, f"f3d_y_{tag}", act=ACT_RELU, res1=x)
        qkv = ws.h16(f"f3d_qkv_{tag}", self.dt, M, 2 * p.dq + p.dv)
        op.gemm(y, p.wqkv, bias=p.bqkv, out_bf16=qkv)
        kvp = ws.h16(f"f3d_kvp_{tag}", self.dt, M // 4, p.dq + p.dv)
        op.maxpool2x2_cols(qkv, p.dq, p.dq + p.dv, kvp, frames, H, W)
        att = ws.h16(f"f3d_att_{tag}", self.dt, M, p.dv)
        op.nonlocal_attention(qkv[:, :p.dq], kvp, att, frames, H * W, H * W // 4, p.dq, p.dv)
        out = ws.h16(f"f3d_out_{tag}", self.dt, M, C)
        op.gemm(att, p.wo, bias=p.bo, act=ACT_RELU, alpha=p.gamma, res1=y, res2=x, out_bf16=out)
        return out

    def run(self, x, channels_last=False, norm=None):
        """x: (N,T,Cin,H,W) fp32 -> (N,T,C,h,w) fp32, or channels-last (N,T,h,w,C) when ``channels_last``.
        uint8 ``x`` (pixels as a video decoder delivers them) with ``norm`` = (mean, std): VidToTensor + VidNormalize
        (utils/dataset.py:835-858) run inside the stem kernel - no fp32 copy of the input frames exists."""
        op, ws, mod = _lib.ops(), self.ws, self.mod
        N, T, Cin, H, W = x.shape
        assert Cin == self.cin, f"expected {self.cin} input channels, got {Cin}"
        frames = N * T
        if x.dtype == torch.uint8:
            assert norm is not None, "uint8 frames need the dataset's (mean, std)"
            x = x.detach().contiguous()
        else:
            x, norm = x.detach().to(torch.float32).contiguous(), None
        cur = ws.h16("stem", self.dt, frames * H * W, self.ngf)
        op.conv7x7_stem(x, self.stem_w, self.stem_shift, cur, Cin, self.ngf, H, W, norm=norm)
        C = self.ngf
        cur = self._conv3x3(cur, frames, H, W, C, self.down[0], 2, PAD_ZERO, "down0", act=ACT_RELU)
        H, W, C = H // 2, W // 2, C * 2
        for i, p in enumerate(self.f3d):
            cur = self._f3d(cur, frames, H, W, p, f"s{i}")
            cur = self._conv3x3(cur, frames, H, W, C, self.down[i + 1], 2, PAD_ZERO, f"down{i + 1}", act=ACT_RELU)
            H, W, C = H // 2, W // 2, C * 2
        out_tok = None
        for i, (p, c1, c2, pad_mode) in enumerate(self.res):
            cur = self._f3d(cur, frames, H, W, p, "res")
            t = self._conv3x3(cur, frames, H, W, C, c1, 1, pad_mode, "res_t", act=ACT_RELU)
            if i + 1 < len(self.res):
                cur = self._conv3x3(t, frames, H, W, C, c2, 1, pad_mode, f"res_o{i % 2}", res1=cur)
            else:   # last block: + skip, then the encoder's out_act ReLU; fp32 tokens for the predictor
                out_tok = torch.empty(frames * H * W, C, dtype=torch.float32, device=self.device)
                self._conv3x3(t, frames, H, W, C, c2, 1, pad_mode, "", res1=cur, post_relu=True, out_f32=out_tok)
        if out_tok is None:   # num_res_blocks == 0
            out_tok = torch.relu(cur.float())
        if channels_last:
            return out_tok.view(N, T, H, W, C)
        out = torch.empty(N, T, C, H, W, dtype=torch.float32, device=self.device)
        op.tokens_to_nchw(out_tok.view(frames, H * W, C), out.view(frames, C, H * W))
        return out


class DecoderEngine:
    def __init__(self, mod):
        self.mod = mod
        self.device = next(mod.parameters()).device
        self.ws = Workspace(self.device)
        self.dt = ae_dtype()
        self.ups = []
        self.ups_sig = tuple((mod.model[3 * i].weight.shape[0], mod.model[3 * i].weight.shape[1]) for i in range(mod.n_downsampling))
        self._ups_tma = None      # packed on first use: the layout depends on the feature-map size (see run)
        head = mod.model[3 * mod.n_downsampling + 1]
        self.head_w = _lib.pack_head_weights(_pack_conv7x7(head), self.dt)
        self.head_b = _f(head.bias)
        self.head_cin, self.cout = head.weight.shape[1], head.weight.shape[0]
        self.act = ACT_TANH if mod.out_layer == 'Tanh' else ACT_SIGMOID

    # output phases (py,px) along N.  The gather path keeps the natural order and leaves its output phase-major; the TMA path
    # (npvp_convt_gemm_bf16) orders them so that every tap feeds a contiguous column range (zero blocks skipped) and stores NHWC.
    PHASES_GATHER = ((0, 0), (0, 1), (1, 0), (1, 1))
    PHASES_TMA = ((0, 0), (0, 1), (1, 1), (1, 0))

    @staticmethod
    def _pack_convT(convt: nn.ConvTranspose2d, bn: nn.BatchNorm2d, dt, phases=PHASES_GATHER):
        """ConvTranspose2d(3,s2,p1,op1) weight [Cin,Cout,3,3] -> 16-bit [(q,co), (dy,dx,ci)] over the 2x2 input
        neighbourhood, q enumerating ``phases``: out[2a+py, 2b+px] = sum_{dy,dx} in[a+dy, b+dx] * w[ky(py,dy), kx(px,dx)], where
        (p=0,d=0)->k=1, (p=1,d=0)->k=2, (p=1,d=1)->k=0 and (p=0,d=1) is dead."""
        scale, shift = _bn_fold(bn)
        w = convt.weight.detach().float()                       # [Cin, Cout, 3, 3]
        Cin, Cout = w.shape[0], w.shape[1]
        tap = {(0, 0): 1, (1, 0): 2, (1, 1): 0}
        B = torch.zeros(4, Cout, 2, 2, Cin, dtype=torch.float32, device=w.device)
        for q, (py, px) in enumerate(phases):
            for dy in (0, 1):
                for dx in (0, 1):
                    if (py, dy) in tap and (px, dx) in tap:
                        B[q, :, dy, dx, :] = (w[:, :, tap[(py, dy)], tap[(px, dx)]] * scale[None, :]).t()
        bias = shift.repeat(4)
        if convt.bias is not None:
            bias = bias + (scale * convt.bias.detach().float()).repeat(4)
        return _h(B.reshape(4 * Cout, 4 * Cin), dt), bias.contiguous(), Cin, Cout

    def _packed_ups(self, H, W):
        """(tma, [packed layers]): the TMA form when every layer's geometry allows it, else the gather form for all
        (the two leave different activation layouts, so they are not mixed inside one chain)."""
        op, mod = _lib.ops(), self.mod
        tma, h, w = os.environ.get("NPVP_B200_CONVT", "tma") != "gather", H, W
        for cin, cout in self.ups_sig:
            tma = tma and op.convt_supported(h, w, cin, cout)
            h, w = 2 * h, 2 * w
        if self._ups_tma != tma:
            phases = self.PHASES_TMA if tma else self.PHASES_GATHER
            self.ups = [self._pack_convT(mod.model[3 * i], mod.model[3 * i + 1], self.dt, phases) for i in range(mod.n_downsampling)]
            self._ups_tma = tma
        return tma, self.ups

    def run(self, x, channels_last=False, renorm=None, want_f32=True):
        """x: (N,T,C,h,w) fp32, or channels-last (N,T,h,w,C) fp32/bf16 -> frames (N,T,Cimg,H,W) fp32.
        ``renorm`` = (mean, std): the head kernel ALSO writes the pixel-space uint8 frames (VidReNormalize + clamp + ToPILImage's
        truncation fused into its epilogue, utils/dataset.py:860-886, utils/train_summary.py:243-248) and the call returns
        (frames | None, frames_u8); ``want_f32=False`` skips the fp32 frames."""
        op, ws = _lib.ops(), self.ws
        N, T = x.shape[0], x.shape[1]
        frames = N * T
        x = x.detach().contiguous()
        if channels_last:
            H, W, C = x.shape[2], x.shape[3], x.shape[4]
            if x.dtype == self.dt:
                cur = x.view(frames * H * W, C)
            else:
                cur = ws.h16("in", self.dt, frames * H * W, C)
                cur.copy_(x.reshape(frames * H * W, C))
        else:
            C, H, W = x.shape[2], x.shape[3], x.shape[4]
            cur = ws.h16("in", self.dt, frames * H * W, C)
            op.nchw_to_tokens(x.to(torch.float32).view(frames, C, H * W), out_bf16=cur.view(frames, H * W, C))
        assert C == self.ups_sig[0][0], f"decoder expects {self.ups_sig[0][0]} feature channels, got {C}"
        tma, ups = self._packed_ups(H, W)
        phase = False
        for i, (w, b, Cin, Cout) in enumerate(ups):
            if tma:      # zero blocks skipped, A tiles by TMA, plain NHWC out
                nxt = ws.h16(f"up{i}", self.dt, frames * 4 * H * W, Cout)
                op.convt_gemm(cur, w, frames, H, W, Cin, Cout, bias=b, act=ACT_RELU, out_bf16=nxt)
                cur, H, W = nxt, 2 * H, 2 * W
            else:        # dense 2x2-neighbourhood GEMM with a cp.async gather, output left phase-major
                nxt = ws.h16(f"up{i}", self.dt, frames * H * W, 4 * Cout)
                op.conv_gemm(cur, w, frames, H, W, Cin, 2, 2, 1, 0, PAD_ZERO, H, W, phase, bias=b, act=ACT_RELU, out_bf16=nxt)
                cur, H, W, phase = nxt, 2 * H, 2 * W, True
        out = torch.empty(N, T, self.cout, H, W, dtype=torch.float32, device=self.device) if (want_f32 or renorm is None) else None
        out_u8 = torch.empty(N, T, self.cout, H, W, dtype=torch.uint8, device=self.device) if renorm is not None else None
        op.conv7x7_head(cur, self.head_w, self.head_b, out, self.head_cin, self.cout, H, W, phase, self.act, out_u8=out_u8, renorm=renorm)
        return out if renorm is None else (out, out_u8)
