"""Host-side driver of the NP predictor: packs the module's weights once and sequences the sm_100a kernels.

Data layout in HBM (per forward, N clips, T frames, 64 tokens/frame, C = 512):
  * residual streams  fp32  [N*T*64, 512]      (token-major, channels-last; never leaves fp32)
  * GEMM operands     bf16  [N*T*64, 512|1024|2048]
  * positional code   fp32  beta/gamma [T*64, 512]  (batch independent, from the NRMLP kernels)
  * per-frame stats   fp32  [frames, 2] / [frames, 8, 2]
Weights are packed at first use (and re-packed when any parameter's version changes):
bf16 [out, in] matrices for the tensor-core GEMMs, fp32 biases / LayerNorm affines, the (C,8,8)
LayerNorm affines transposed to token-major [64, C], depthwise taps as [9, C], BatchNorm folded.

Reference call graph being reproduced: models/Predictor.py:301-350 -> models/VidHRFormer.py:25-52,
79-116, 126-161, 198-245 (see SURVEY.md Appendix A for the math).
"""
from __future__ import annotations

import torch

import os

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, ATTN_SPATIAL, ATTN_TEMPORAL, PAD_ZERO
from .workspace import Workspace

C = 512
TOK = 64
# Deferred residual adds (default on): a residual branch's GEMM writes its bf16 output and the LayerNorm kernel that reads the
# stream next performs x += delta, instead of a read-modify-write of the fp32 stream in the GEMM epilogue.
_DEFER = os.environ.get("NPVP_B200_DEFER_RESIDUAL", "1") != "0"


def _bf(t):
    return t.detach().to(torch.bfloat16).contiguous()


def _f(t):
    return t.detach().float().contiguous()


def _bn_fold(bn):
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return scale, shift


def _chw_affine(ln):
    """LayerNorm((C,8,8)) elementwise affine -> token-major [64, C]."""
    w = ln.weight.detach().float().permute(1, 2, 0).reshape(TOK, -1).contiguous()
    b = ln.bias.detach().float().permute(1, 2, 0).reshape(TOK, -1).contiguous()
    return w, b


class _MHA:
    """Packed nn.MultiheadAttention weights: q|k fused [1024,512] (same source), v [512,512], out [512,512]."""

    def __init__(self, mha, split_q=False):
        W, b = mha.in_proj_weight.detach(), mha.in_proj_bias.detach()
        if split_q:      # encoder-decoder attention: q, k, v all read different sources
            self.wq, self.bq = _bf(W[:C]), _f(b[:C])
            self.wk, self.bk = _bf(W[C:2 * C]), _f(b[C:2 * C])
        else:
            self.wqk, self.bqk = _bf(W[:2 * C]), _f(b[:2 * C])
        self.wv, self.bv = _bf(W[2 * C:]), _f(b[2 * C:])
        self.wo, self.bo = _bf(mha.out_proj.weight), _f(mha.out_proj.bias)


class _ConvFFN:
    """MlpDWBN weights.  ``half_mid`` (hidden width 2048, the only one NPVP uses): the middle runs as ONE half-precision
    kernel (npvp_ffn_mid16) - both LayerNorm((hid,8,8)) affines pair-interleaved in IEEE half [2, 64, hid/2, (w|b), 2], half
    depthwise taps, and fc2 as a half GEMM.  Other widths keep the fp32 two-kernel path."""

    def __init__(self, m, half_mid):
        hid = m.fc1.weight.shape[0]
        self.hid = hid
        self.half_mid = bool(half_mid) and hid == 2048
        self.w1, self.b1 = _bf(m.fc1.weight.reshape(hid, C)), _f(m.fc1.bias)
        n1w, n1b = _chw_affine(m.norm1)
        n2w, n2b = _chw_affine(m.norm2)
        dw_w = _f(m.dw3x3.weight.reshape(hid, 9).t())
        if self.half_mid:
            pair = lambda w, b: torch.stack([w.view(TOK, hid // 2, 2), b.view(TOK, hid // 2, 2)], dim=2)
            self.ln_wb = torch.stack([pair(n1w, n1b), pair(n2w, n2b)], dim=0).to(torch.float16).contiguous()
            self.dw_w16 = dw_w.to(torch.float16).contiguous()
            self.dw_b16 = m.dw3x3.bias.detach().to(torch.float16).contiguous()
            self.w2 = m.fc2.weight.detach().reshape(-1, hid).to(torch.float16).contiguous()
        else:
            self.n1w, self.n1b, self.n2w, self.n2b = n1w, n1b, n2w, n2b
            self.dw_w, self.dw_b = dw_w, _f(m.dw3x3.bias)
            self.w2 = _bf(m.fc2.weight.reshape(-1, hid))
        self.b2 = _f(m.fc2.bias)
        self.n3w, self.n3b = _chw_affine(m.norm3)


class _LN:
    def __init__(self, ln):
        self.w, self.b = _f(ln.weight), _f(ln.bias)


class _EncLayer:
    def __init__(self, blk, half_mid=True):
        self.attn_s = _MHA(blk.SLMHSA.attn)
        self.ffn_s = _ConvFFN(blk.SpatialFFN, half_mid)
        self.attn_t = _MHA(blk.temporal_MHSA)
        self.n1, self.n2, self.n3, self.n4 = _LN(blk.norm1), _LN(blk.norm2), _LN(blk.norm3), _LN(blk.norm4)
        self.l1w, self.l1b = _bf(blk.linear1.weight), _f(blk.linear1.bias)
        self.l2w, self.l2b = _bf(blk.linear2.weight), _f(blk.linear2.bias)


class _DecLayer(_EncLayer):
    def __init__(self, blk, half_mid=True):
        super().__init__(blk, half_mid)
        self.attn_x = _MHA(blk.EncDecAttn, split_q=True)
        self.ffn_x = _ConvFFN(blk.SpatialFFN1, half_mid)
        self.n5, self.n6 = _LN(blk.norm5), _LN(blk.norm6)


class _EventEnc:
    def __init__(self, m, stochastic):
        s1, sh1 = _bn_fold(m.conv1[1])
        self.dw_w = (m.conv1[0].weight.detach().float().reshape(C, 9) * s1[:, None]).t().contiguous()   # [9, C]
        self.dw_shift = sh1.contiguous()
        s2, sh2 = _bn_fold(m.conv2[1])
        w2 = m.conv2[0].weight.detach().float() * s2[:, None, None, None]                                # [256,512,3,3]
        self.w2 = _bf(w2.permute(0, 2, 3, 1).reshape(w2.shape[0], -1))                                  # [(ky,kx,ci)]
        self.b2 = sh2.contiguous()
        s3, sh3 = _bn_fold(m.MLP_0[1])
        self.w3 = _bf(m.MLP_0[0].weight.detach().float().reshape(s3.numel(), -1) * s3[:, None])
        self.b3 = sh3.contiguous()
        ws, bs = [m.mu_net.weight.detach().float().reshape(C, -1)], [m.mu_net.bias.detach().float()]
        if stochastic:
            ws.append(m.logvar_net.weight.detach().float().reshape(C, -1))
            bs.append(m.logvar_net.bias.detach().float())
        self.w_head, self.b_head = _bf(torch.cat(ws, 0)), torch.cat(bs, 0).contiguous()
        self.hidden = self.w3.shape[0]
        self.stochastic = stochastic


class PredictorEngine:
    def __init__(self, mod):
        self.mod = mod
        dev = next(mod.parameters()).device
        self.device = dev
        self.ws = Workspace(dev)
        self.stochastic = bool(mod.stochastic)
        # conv-FFN middle: "half" (default) = one single-pass half-precision kernel (npvp_ffn_mid16);
        # "split" = the fp32 two-kernel path ffn_dwconv + ffn_norm2 (r01; kept for A/B runs and for hidden widths != 2048)
        mid = os.environ.get("NPVP_B200_FFN_MID", "half")
        half_mid = mid == "half" and _lib.ops().ffn_mid16_lanes() > 0
        self._mid16 = {}                                         # frames -> persistent (xch, cnt) exchange scratch
        self._pos_lru = {}                                       # coordinate tensor key -> (tensor, (beta, gamma), buffers)
        self.enc_layers = [_EncLayer(b, half_mid) for b in mod.EVT_Former.layers]
        self.dec_layers = [_DecLayer(b, half_mid) for b in mod.transformer.layers]
        # encoder-decoder attention: the K / V projections of the (layer-invariant) memory do not depend on the decoder state, so
        # the 8 layers' weights are stacked and each projection is ONE GEMM with N = 8 x 512 per forward instead of 8 GEMMs on
        # only 128 tiles each (M = clips x To x 64 is small); layer l reads columns [512 l, 512 l + 512) of the result
        self.wk_all = torch.cat([L.attn_x.wk for L in self.dec_layers], 0).contiguous()
        self.bk_all = torch.cat([L.attn_x.bk for L in self.dec_layers], 0).contiguous()
        self.wv_all = torch.cat([L.attn_x.wv for L in self.dec_layers], 0).contiguous()
        self.bv_all = torch.cat([L.attn_x.bv for L in self.dec_layers], 0).contiguous()
        self.norm_enc = _LN(mod.EVT_Former.norm)
        self.norm_dec = _LN(mod.transformer.norm)
        latent = mod.evt_prior if self.stochastic else mod.evt_posterior      # Predictor.py:310 / :330
        self.evt = _EventEnc(latent, self.stochastic)
        nr = mod.nrmlp
        self.nr_B = _f(nr.B)
        self.nr_mlp = [(_f(l.weight), _f(l.bias)) for l in nr.linears()]
        self.nr_beta = (_f(nr.mlp_beta.weight), _f(nr.mlp_beta.bias))
        self.nr_gamma = (_f(nr.mlp_gamma.weight), _f(nr.mlp_gamma.bias)) if nr.fuse_method == 'SPADE' else None

    # ------------------------------------------------------------------------------------------
    # positional code
    # ------------------------------------------------------------------------------------------
    def positional(self, coor, bufs):
        """NRMLP (submodules.py:299-327): coor fp32 [R,3] -> (beta [R,512], gamma [R,512] | None), written into ``bufs`` (a dict
        of scratch / result buffers owned by one table slot, allocated on first use)."""
        op = _lib.ops()
        coor = coor.detach().to(self.device, torch.float32).contiguous()
        R = coor.shape[0]
        half = self.nr_B.shape[0]

        def buf(name, cols):
            t = bufs.get(name)
            if t is None or t.shape != (R, cols):
                t = bufs[name] = torch.empty(R, cols, dtype=torch.float32, device=self.device)
            return t
        cur = buf("ff", 2 * half)
        op.fourier_features(coor, self.nr_B, cur)
        for i, (w, b) in enumerate(self.nr_mlp):
            nxt = buf(f"h{i}", w.shape[0])
            op.gemm_f32(cur, w, b, ACT_RELU, nxt)
            cur = nxt
        beta = buf("beta", C)
        op.gemm_f32(cur, self.nr_beta[0], self.nr_beta[1], ACT_NONE, beta)
        gamma = None
        if self.nr_gamma is not None:
            gamma = buf("gamma", C)
            op.gemm_f32(cur, self.nr_gamma[0], self.nr_gamma[1], ACT_NONE, gamma)
        return beta, gamma

    @staticmethod
    def _coor_key(coor):
        """Identity of a coordinate tensor's CONTENT as far as it can be told without reading it: storage address, shape and
        the tensor's in-place modification counter.  The table slot keeps a reference to the tensor, so the address cannot be
        recycled for other coordinates while the entry lives."""
        return (coor.data_ptr(), tuple(coor.shape), int(coor._version), coor.device)

    def positional_table(self, coor, slot=None):
        """Cached NRMLP code of ``coor``.  The code depends only on the coordinates and the NRMLP weights (0.6-1.5
        GFLOP of fp32 work per forward when recomputed, 24 launches), so it is computed once per coordinate tensor and kept in
        a small LRU table (the engine itself is rebuilt when any weight changes).  ``slot``: a dict owned by the caller
        (a captured CUDA graph) that pins the result buffers: the table is then recomputed IN PLACE when the coordinates
        change, so that pointers baked into the graph stay valid."""
        key = self._coor_key(coor)
        if slot is not None:
            if slot.get("key") != key:
                slot["key"], slot["coor"] = None, coor
                slot["val"] = self.positional(coor, slot.setdefault("bufs", {}))
                slot["key"] = key
            return slot["val"]
        lru = self._pos_lru
        hit = lru.pop(key, None)
        if hit is None:
            bufs = {}
            if len(lru) >= 8:                                   # recycle the buffers of the least recently used entry
                old_key = next(iter(lru))
                bufs = lru.pop(old_key)[2]
            hit = (coor, self.positional(coor, bufs), bufs)
        lru[key] = hit                                          # (re-)insert as most recently used
        return hit[1]

    def _positional_pair(self, oc, pc):
        slots = self.__dict__.get("_graph_slots")               # set by a captured forward (pipeline._GraphedPredict)
        if slots is not None:
            return self.positional_table(oc, slots[0]), self.positional_table(pc, slots[1])
        return self.positional_table(oc), self.positional_table(pc)

    # ------------------------------------------------------------------------------------------
    # building blocks (x: fp32 residual stream [M,512], updated in place)
    # ------------------------------------------------------------------------------------------
    def _self_attention(self, x, a_bf, f_bf, w: _MHA, mode, n, T, mask_last, tag):
        op, ws = _lib.ops(), self.ws
        M = x.shape[0]
        qk = ws.bf16(f"qk_{tag}", M, 2 * C)
        v = ws.bf16(f"v_{tag}", M, C)
        o = ws.bf16(f"o_{tag}", M, C)
        op.gemm(f_bf, w.wqk, bias=w.bqk, out_bf16=qk)
        op.gemm(a_bf, w.wv, bias=w.bv, out_bf16=v)
        op.attention(qk[:, :C], qk[:, C:], v, o, mode, n, T, T, mask_last)
        return self._residual_gemm(x, o, w.wo, w.bo, tag)

    def _residual_gemm(self, x, a_bf, w, b, tag):
        """x += a @ w^T + b.  Deferred mode returns the bf16 branch output for the next LayerNorm kernel to add."""
        op = _lib.ops()
        if not _DEFER:
            op.gemm(a_bf, w, bias=b, res1=x, out_f32=x)
            return None
        d = self.ws.bf16(f"delta_{tag}", x.shape[0], C)
        op.gemm(a_bf, w, bias=b, out_bf16=d)
        return d

    def _ln_rows(self, x, delta, ln, **kw):
        op = _lib.ops()
        if delta is None:
            op.layernorm_rows(x, ln.w, ln.b, **kw)
        else:
            op.add_layernorm_rows(x, delta, ln.w, ln.b, **kw)

    def _ln_fuse(self, x, delta, ln, qe, beta, gamma, out_ln, out_fused, n, T):
        op = _lib.ops()
        if delta is None:
            op.ln_posfuse(x, ln.w, ln.b, qe, beta, gamma, out_ln, out_fused, n, T)
        else:
            op.add_ln_posfuse(x, delta, ln.w, ln.b, qe, beta, gamma, out_ln, out_fused, n, T)

    def _conv_ffn(self, x, a_bf, w: _ConvFFN, frames, tag, tail=None):
        """x += MlpDWBN(a)  (VidHRFormer.py:374-392).  ``tail`` = (ln, qe, beta, gamma, out_ln, out_fused, n, T): the
        LayerNorm + positional fuse that consumes the updated stream next, fused into the last kernel."""
        op, ws = _lib.ops(), self.ws
        M = x.shape[0]
        h3 = ws.bf16(f"h3_{tag}", M, C)       # 16-bit is enough: h3 is re-normalised by LayerNorm((C,8,8)) immediately
        pt1 = ws.f32(f"pt1_{tag}", frames, 4 * w.hid // 256, 2)
        if w.half_mid:                        # fc1 (bf16 operands -> half h1 + LN1 partial statistics) -> one-pass middle -> fc2 (half)
            h1 = ws.h16(f"h1_{tag}", torch.float16, M, w.hid)
            h2 = ws.h16(f"h2_{tag}", torch.float16, M, w.hid)
            op.gemm(a_bf, w.w1, bias=w.b1, out_bf16=h1, frame_stats=pt1)
            scratch = self._mid16.get(frames)
            if scratch is None:
                scratch = self._mid16[frames] = _lib.ffn_mid16_scratch(frames, self.device)
            op.ffn_mid16(h1, pt1, w.ln_wb, w.dw_w16, w.dw_b16, h2, *scratch)
        else:
            h1 = ws.bf16(f"h1_{tag}", M, w.hid)
            y2 = ws.bf16(f"y2_{tag}", M, w.hid)
            st1 = ws.f32(f"st1_{tag}", frames, 2)
            pt2 = ws.f32(f"pt2_{tag}", frames, w.hid // _lib.FFN_CHUNK, 2)
            op.gemm(a_bf, w.w1, bias=w.b1, out_bf16=h1, frame_stats=pt1)     # LayerNorm((hid,8,8)) statistics from the fc1 epilogue
            op.ffn_stats_finalize(pt1, st1, 64 * w.hid)
            op.ffn_dwconv(h1, st1, w.n1w, w.n1b, w.dw_w, w.dw_b, y2, pt2)
            op.ffn_norm2(y2, pt2, w.n2w, w.n2b, h1)             # h1 is dead: reuse it for GELU(LN2(.))
            h2 = h1
        op.gemm(h2, w.w2, bias=w.b2, out_bf16=h3)
        if tail is None:
            op.frame_ln_gelu_residual(h3, w.n3w, w.n3b, x)
        else:
            ln, qe, beta, gamma, out_ln, out_fused, n, T = tail
            op.frame_ln_gelu_residual_posfuse(h3, w.n3w, w.n3b, x, ln.w, ln.b, qe, beta, gamma, out_ln, out_fused, n, T)

    def _mlp_ffn(self, x, a_bf, L, tag):
        op, ws = _lib.ops(), self.ws
        f1 = ws.bf16(f"f1_{tag}", x.shape[0], L.l1w.shape[0])
        op.gemm(a_bf, L.l1w, bias=L.l1b, act=ACT_GELU, out_bf16=f1)
        return self._residual_gemm(x, f1, L.l2w, L.l2b, tag)

    # ------------------------------------------------------------------------------------------
    # EVT_Former (VidHRFormer.py:25-52, 79-116)
    # ------------------------------------------------------------------------------------------
    def encode(self, x_tokens, beta, gamma, n, T, tag="enc"):
        """x_tokens fp32 [n*T*64, 512] (consumed in place) -> (memory fp32, memory bf16) [n*T*64,512].
        ``tag`` names the scratch buffers: the posterior pass over the ground-truth future ("encp") keeps its own set, so
        the context memory survives it and neither pass re-allocates when To != Tp."""
        op, ws = _lib.ops(), self.ws
        M = n * T * TOK
        x = x_tokens
        a = ws.bf16(f"a_{tag}", M, C)
        f = ws.bf16(f"f_{tag}", M, C)
        d = None                                                 # pending (deferred) residual of the previous branch
        for L in self.enc_layers:
            self._ln_fuse(x, d, L.n1, None, beta, gamma, a, f, n, T)
            d = self._self_attention(x, a, f, L.attn_s, ATTN_SPATIAL, n, T, False, tag)
            self._ln_rows(x, d, L.n2, out_bf16=a)
            self._conv_ffn(x, a, L.ffn_s, n * T, tag, tail=(L.n3, None, beta, gamma, a, f, n, T))   # + LN3 / fuse
            d = self._self_attention(x, a, f, L.attn_t, ATTN_TEMPORAL, n, T, True, tag)      # mask quirk :100-102
            self._ln_rows(x, d, L.n4, out_bf16=a)
            d = self._mlp_ffn(x, a, L, tag)
        sfx = "" if tag == "enc" else "_" + tag
        mem = ws.f32("mem_f32" + sfx, M, C)
        mem_bf = ws.bf16("mem_bf16" + sfx, M, C)
        self._ln_rows(x, d, self.norm_enc, out_f32=mem, out_bf16=mem_bf)
        return mem, mem_bf

    # ------------------------------------------------------------------------------------------
    # latent event code (submodules.py:388-410)
    # ------------------------------------------------------------------------------------------
    def latent(self, evt, n, eps, n_samples=1, E=None, sfx=""):
        """evt fp32 [n*64,512] token-major -> z fp32 [n*K*64,512]; keeps (mu|logvar) of the n clips in the workspace.
        K > 1: the prior (mu, logvar) is computed once per clip and re-parameterised with K noise tensors.
        ``E`` / ``sfx``: another packed EventEncoder (the posterior) with its own scratch buffers."""
        op, ws = _lib.ops(), self.ws
        E = self.evt if E is None else E
        M = n * TOK
        e1 = ws.bf16("evt_e1" + sfx, M, C)
        op.dwconv3x3_tokens(evt, E.dw_w, E.dw_shift, e1, relu=True)
        e2 = ws.bf16("evt_e2" + sfx, M, E.hidden)
        op.conv_gemm(e1, E.w2, n, 8, 8, C, 3, 3, 1, 1, PAD_ZERO, 8, 8, bias=E.b2, act=ACT_RELU, out_bf16=e2)
        e3 = ws.bf16("evt_e3" + sfx, M, E.hidden)
        op.gemm(e2, E.w3, bias=E.b3, act=ACT_RELU, out_bf16=e3)
        mulv = ws.f32("evt_mulv" + sfx, M, E.w_head.shape[0])
        op.gemm(e3, E.w_head, bias=E.b_head, out_f32=mulv)
        K = int(n_samples)
        src = mulv
        if K > 1:
            W2 = mulv.shape[1]
            src = ws.f32("evt_mulv_rep" + sfx, M * K, W2)
            src.view(n, K, TOK, W2).copy_(mulv.view(n, 1, TOK, W2).expand(n, K, TOK, W2))
        z = ws.f32("evt_z" + sfx, M * K, C)
        op.latent_reparam(src, eps if E.stochastic else None, z, n * K, C)
        return z, mulv

    # ------------------------------------------------------------------------------------------
    # NAR decoder (VidHRFormer.py:126-161, 198-245)
    # ------------------------------------------------------------------------------------------
    def decode(self, z, mem, mem_bf, beta_o, gamma_o, beta_p, gamma_p, n, To, Tp, relu_out=True, out16=torch.bfloat16):
        op, ws = _lib.ops(), self.ws
        M, Mo = n * Tp * TOK, n * To * TOK
        y = ws.f32("y_dec", M, C)
        y.zero_()                                               # tgt = zeros (VidHRFormer.py:139)
        a = ws.bf16("a_dec", M, C)
        f = ws.bf16("f_dec", M, C)
        keyf = ws.bf16("keyf", Mo, C)
        op.ln_posfuse(mem, None, None, None, beta_o, gamma_o, None, keyf, n, To)   # fuse(memory): layer invariant
        nl = len(self.dec_layers)
        kx_all = ws.bf16("kx_all", Mo, nl * C)
        vx_all = ws.bf16("vx_all", Mo, nl * C)
        op.gemm(keyf, self.wk_all, bias=self.bk_all, out_bf16=kx_all)
        op.gemm(mem_bf, self.wv_all, bias=self.bv_all, out_bf16=vx_all)
        qx = ws.bf16("qx", M, C)
        ox = ws.bf16("o_dec", M, C)
        for li, L in enumerate(self.dec_layers):
            if li == 0:      # later layers get this from the previous layer's last kernel
                op.ln_posfuse(y, L.n1.w, L.n1.b, z, beta_p, gamma_p, a, f, n, Tp)
            d = self._self_attention(y, a, f, L.attn_s, ATTN_SPATIAL, n, Tp, False, "dec")
            self._ln_rows(y, d, L.n2, out_bf16=a)
            self._conv_ffn(y, a, L.ffn_s, n * Tp, "dec", tail=(L.n3, None, beta_p, gamma_p, a, f, n, Tp))   # + LN3 / fuse
            d = self._self_attention(y, a, f, L.attn_t, ATTN_TEMPORAL, n, Tp, False, "dec")
            self._ln_rows(y, d, L.n4, out_bf16=a)
            d = self._mlp_ffn(y, a, L, "dec")
            # encoder-decoder attention over time
            self._ln_fuse(y, d, L.n5, z, beta_p, gamma_p, None, f, n, Tp)
            X = L.attn_x
            op.gemm(f, X.wq, bias=X.bq, out_bf16=qx)
            op.attention(qx, kx_all[:, li * C:(li + 1) * C], vx_all[:, li * C:(li + 1) * C], ox, ATTN_TEMPORAL, n, Tp, To, False)
            d = self._residual_gemm(y, ox, X.wo, X.bo, "dec")
            self._ln_rows(y, d, L.n6, out_bf16=a)
            nxt = self.dec_layers[li + 1] if li + 1 < len(self.dec_layers) else None
            self._conv_ffn(y, a, L.ffn_x, n * Tp, "dec",          # + next layer's LN1 / (+ query_evt) / fuse
                           tail=None if nxt is None else (nxt.n1, z, beta_p, gamma_p, a, f, n, Tp))
        out = ws.f32("dec_out", M, C)
        out_bf = ws.h16("dec_out_16", out16, M, C)
        op.layernorm_rows(y, self.norm_dec.w, self.norm_dec.b, out_f32=out, out_bf16=out_bf, relu=relu_out)
        return out, out_bf

    # ------------------------------------------------------------------------------------------
    # entry points used by the Predictor module
    # ------------------------------------------------------------------------------------------
    def _to_tokens(self, x, channels_last, name):
        """(N,T,C,H,W) or channels-last (N,T,H,W,C) fp32 -> private fp32 token matrix [N*T*64, 512]."""
        op, ws = _lib.ops(), self.ws
        n, T = x.shape[0], x.shape[1]
        tok = ws.f32(name, n * T * TOK, C)
        x = x.detach().to(torch.float32).contiguous()
        if channels_last:
            assert tuple(x.shape[2:]) == (8, 8, C), x.shape
            tok.copy_(x.reshape(-1, C))
        else:
            assert tuple(x.shape[2:]) == (C, 8, 8), x.shape
            op.nchw_to_tokens(x.reshape(n * T, C, TOK), out_f32=tok.view(n * T, TOK, C))
        return tok, n, T

    def _from_tokens(self, tok, n, T, channels_last):
        op = _lib.ops()
        if channels_last:
            return tok.view(n, T, 8, 8, C).clone()
        out = torch.empty(n, T, C, 8, 8, dtype=torch.float32, device=tok.device)
        op.tokens_to_nchw(tok.view(n * T, TOK, C), out.view(n * T, C, TOK))
        return out

    @staticmethod
    def _mu_logvar_nchw(mulv, n):
        """(mu | logvar) token-major fp32 [n*64, 1024] -> two (n, 512, 8, 8) tensors in the reference layout."""
        t = mulv.view(n, 8, 8, 2, C).permute(3, 0, 4, 1, 2).contiguous()
        return t[0], t[1]

    def posterior(self, gt, n, channels_last, beta_p, gamma_p, sample_noise, eps_p=None, want_z=False):
        """NPVP-S posterior on the ground-truth future features (Predictor.py:311-313): EVT_Former over the Tp target frames with
        the target positional code, temporal mean, ``evt_posterior`` heads -> (mu_p, logvar_p).  z_p itself is only used in
        training mode (:316-318), so no re-parameterisation kernel runs; with ``sample_noise`` one noise tensor is still drawn
        and dropped so that the global generator advances exactly like the reference's second ``torch.randn`` (submodules.py:409)."""
        post = self.__dict__.get("_evt_post")
        if post is None:
            post = self._evt_post = _EventEnc(self.mod.evt_posterior, True)
        tok, n2, Tp = self._to_tokens(gt, channels_last, "x_encp")
        assert n2 == n and beta_p.shape[0] in (Tp * TOK, n * Tp * TOK), "predict_features_gt must hold one frame per target timestamp"
        mem_p, _ = self.encode(tok, beta_p, gamma_p, n, Tp, tag="encp")
        evt_p = self.ws.f32("evt_p", n * TOK, C)
        _lib.ops().temporal_mean(mem_p, evt_p, n, Tp)
        if want_z:            # z_p = mu_p + exp(logvar_p / 2) eps_p drives the decoder (Predictor.py:315-318)
            if eps_p is None:
                eps_p = torch.randn((n, C, 8, 8), device=self.device)                     # the reference's second draw (submodules.py:409)
            eps_p = eps_p.detach().to(self.device, torch.float32).contiguous()
            assert tuple(eps_p.shape) == (n, C, 8, 8), f"posterior noise must be {(n, C, 8, 8)}, got {tuple(eps_p.shape)}"
            z_p, mulv_p = self.latent(evt_p, n, eps_p, E=post, sfx="_p")
            return self._mu_logvar_nchw(mulv_p, n) + (z_p,)
        if sample_noise:
            torch.randn((n, C, 8, 8), device=self.device)
        _, mulv_p = self.latent(evt_p, n, None, E=post, sfx="_p")
        return self._mu_logvar_nchw(mulv_p, n)

    def run(self, observed, channels_last=False, out16=None, n_samples=1, predict_gt=None, decode_posterior=False):
        """``n_samples`` = K > 1 (NPVP-S): K stochastic futures per clip from ONE pass of the EVT_Former and the prior (only the
        latent and the NAR decoder run per sample); the batch dimension of the result is clip-major (clip 0 sample 0..K-1, ...).
        ``predict_gt`` (NPVP-S, eval): ground-truth future features -> (out, mu_o, logvar_o, mu_p, logvar_p) like the reference
        (Predictor.py:324-327); the decoder is still queried with the prior sample.  The deterministic model ignores it (:328-335)."""
        mod = self.mod
        K = int(n_samples)
        assert K >= 1 and (K == 1 or self.stochastic), "several samples per clip only make sense for the stochastic model (NPVP-S)"
        x, n, To = self._to_tokens(observed, channels_last, "x_enc")
        oc, pc = mod.observed_coor, mod.predict_coor
        per_clip = int(getattr(mod, "_coor_clips", 0))      # > 0: every clip has its own timestamps (reset_pos_coor_per_clip)
        if per_clip:
            assert per_clip == n, f"per-clip timestamps were set for {per_clip} clips but the batch has {n}"
            assert K == 1, "several samples per clip with per-clip timestamps are not supported"
        assert oc.shape[0] == (per_clip or 1) * To * TOK, \
            f"observed_coor has {oc.shape[0] // TOK // (per_clip or 1)} timestamps but the input has {To} frames"
        Tp = pc.shape[0] // TOK // (per_clip or 1)
        assert Tp <= 32 and To <= 32, "temporal attention kernels hold at most 32 timestamps per sequence"
        (beta_o, gamma_o), (beta_p, gamma_p) = self._positional_pair(oc, pc)
        mem, mem_bf = self.encode(x, beta_o, gamma_o, n, To)
        evt = self.ws.f32("evt", n * TOK, C)
        _lib.ops().temporal_mean(mem, evt, n, To)
        eps = None
        if self.stochastic:
            eps = mod.injected_eps
            if eps is None:
                eps = torch.randn((n * K, C, 8, 8), device=self.device)       # submodules.py:409
            eps = eps.detach().to(self.device, torch.float32).contiguous()
            assert tuple(eps.shape) == (n * K, C, 8, 8), f"latent noise must be {(n * K, C, 8, 8)}, got {tuple(eps.shape)}"
        z, mulv = self.latent(evt, n, eps, K)
        mod.last_latent = (z, mulv)
        post = None
        if decode_posterior:                  # Predictor.py:315-318 (the reference's training-mode branch, forward only)
            assert self.stochastic and predict_gt is not None and K == 1, \
                "please input groundtruth predict features for storchastic model training/val"
            (_, _), (bp_, gp_) = self._positional_pair(oc, pc)
            mu_p, logvar_p, z = self.posterior(predict_gt, n, channels_last, bp_, gp_, False,
                                               eps_p=getattr(mod, "injected_eps_p", None), want_z=True)
            post = (mu_p, logvar_p)
        if K > 1:            # the decoder sees n*K clips: replicate the (small) encoder memory, clip-major
            rep = lambda t, name, dt: self.ws.get(name, (n * K * To * TOK, C), dt).view(n, K, To * TOK, C).copy_(
                t.view(n, 1, To * TOK, C).expand(n, K, To * TOK, C)).view(n * K * To * TOK, C)
            mem, mem_bf = rep(mem, "mem_rep", torch.float32), rep(mem_bf, "mem_bf_rep", torch.bfloat16)
        out, out_bf = self.decode(z, mem, mem_bf, beta_o, gamma_o, beta_p, gamma_p, n * K, To, Tp,
                                  out16=out16 or torch.bfloat16)
        if out16 is not None:   # engine-internal hand-off to the frame decoder (a view of the workspace, consumed at once)
            assert channels_last
            return out_bf.view(n * K, Tp, 8, 8, C)
        res = self._from_tokens(out, n * K, Tp, channels_last)
        if predict_gt is not None and self.stochastic:
            assert K == 1, "predict_features_gt and n_samples > 1 cannot be combined"
            mu_o, logvar_o = self._mu_logvar_nchw(mulv, n)
            if post is None:
                post = self.posterior(predict_gt, n, channels_last, beta_p, gamma_p, sample_noise=mod.injected_eps is None)
            return res, mu_o, logvar_o, post[0], post[1]
        return res

    def evt_coding(self, x, pos_beta, pos_gamma):
        """Predictor.evt_coding_forward (Predictor.py:337-350): returns (memory (N,T,C,H,W), evt_coding (N,C,H,W))."""
        tok, n, T = self._to_tokens(x, False, "x_enc")
        beta = pos_beta.detach().to(self.device, torch.float32).contiguous()
        gamma = pos_gamma.detach().to(self.device, torch.float32).contiguous() if pos_gamma is not None else None
        if gamma is not None and not bool((gamma != 0).any()):
            gamma = None
        mem, _ = self.encode(tok, beta, gamma, n, T)
        evt = self.ws.f32("evt", n * TOK, C)
        _lib.ops().temporal_mean(mem, evt, n, T)
        return self._from_tokens(mem, n, T, False), self._from_tokens(evt, n, 1, False)[:, 0]
