"""CPU oracle for the NPVP inference hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-tensor restatement (fp32, torch on CPU) of the reference
algorithm for the path  ResnetEncoder -> Predictor (NPVP-D / NPVP-S) -> ResnetDecoder.
It is *not* part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it.  The shipped
modules in ``npvp_b200/`` never touch it and raise if the CUDA library is missing.

Parity pinning: the reference repository ships no golden vectors or tests
(SURVEY.md section 4), so this oracle is pinned against outputs of the reference
itself, produced in the build container by ``tests/golden/make_golden.py`` (which
imports ``/root/reference`` unmodified through a ``sys.modules`` shim) and committed
under ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` replays them.

Every function works on a flat ``state_dict``-style mapping ``sd`` (name -> tensor)
using the reference's own key names, so it can consume either the reference's
checkpoints or ``npvp_b200`` module state_dicts.  Citations are into /root/reference.

The functions are device agnostic (plain torch ops on whatever device the state_dict and the
inputs live on): on CPU they are the parity oracle, and ``bench.py``'s ``gpu_eager_baseline``
leg runs the same code on ``cuda`` to time what the reference's eager PyTorch path costs on a B200
(the same ATen / cuDNN / cuBLAS kernels the reference modules dispatch to).
"""
from __future__ import annotations

import math
from typing import Mapping, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Mapping[str, Tensor]

EPS = 1e-5            # LayerNorm / GroupNorm / BatchNorm default eps used everywhere
NUM_HEADS = 8         # models/Predictor.py:270
WINDOW = 4            # models/Predictor.py:270


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def _bn_eval(sd: SD, p: str, x: Tensor) -> Tensor:
    """Eval-mode BatchNorm2d on NCHW: running stats, affine (torch default eps 1e-5)."""
    w, b = sd[p + "weight"], sd[p + "bias"]
    rm, rv = sd[p + "running_mean"], sd[p + "running_var"]
    scale = w / torch.sqrt(rv + EPS)
    shift = b - rm * scale
    return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


def _layer_norm_c(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """nn.LayerNorm(C) over the last dim, biased variance."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + EPS) * w + b


def _gelu(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form (models/VidHRFormer.py:73,337)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


# --------------------------------------------------------------------------------------
# coordinates + NRMLP positional INR
# --------------------------------------------------------------------------------------
def coor_generator(t_list: Tensor, h_list: Tensor, w_list: Tensor,
                   max_T: float, max_H: float, max_W: float) -> Tensor:
    """CoorGenerator.forward, models/submodules.py:339-366.
    Rows ordered t-major then h then w; columns (t/max_T, h/max_H, w/max_W)."""
    assert float(h_list.max()) <= max_H and float(h_list.min()) >= 0.0, "Invalid H coordinates"
    assert float(w_list.max()) <= max_W and float(w_list.min()) >= 0.0, "Invalid W coordinates"
    assert float(t_list.max()) <= max_T and float(t_list.min()) >= 0.0, "Invalid T coordinates"
    T, H, W = t_list.shape[0], h_list.shape[0], w_list.shape[0]
    out = torch.empty(T, H, W, 3, dtype=torch.float32)
    out[..., 0] = (t_list.float() / max_T).view(T, 1, 1)
    out[..., 1] = (h_list.float() / max_H).view(1, H, 1)
    out[..., 2] = (w_list.float() / max_W).view(1, 1, W)
    return out.reshape(T * H * W, 3)


def nrmlp(sd: SD, p: str, coor: Tensor, fuse_method: str = "Add") -> Tuple[Tensor, Tensor]:
    """NRMLP.forward + gaussian_mapping, models/submodules.py:299-327."""
    proj = (2.0 * float(math.pi) * coor) @ sd[p + "B"].t()
    x = torch.cat([torch.cos(proj), torch.sin(proj)], dim=-1)
    for i in (0, 2, 4):                                   # MLP = [Linear, ReLU] x 3 (submodules.py:272-291)
        x = torch.relu(x @ sd[f"{p}MLP.{i}.weight"].t() + sd[f"{p}MLP.{i}.bias"])
    beta = x @ sd[p + "mlp_beta.weight"].t() + sd[p + "mlp_beta.bias"]
    if fuse_method == "SPADE":
        gamma = x @ sd[p + "mlp_gamma.weight"].t() + sd[p + "mlp_gamma.bias"]
    else:
        gamma = torch.zeros_like(beta)                    # submodules.py:312
    return beta, gamma


def pos_fuse(x: Tensor, beta: Tensor, gamma: Tensor) -> Tensor:
    """PosFeatFuser.forward with param_free_norm_type='layer' (GroupNorm(1,C,affine=False)),
    models/submodules.py:432-454.  x: (N,T,H,W,C); beta/gamma: (T*H*W, C)."""
    N, T, H, W, C = x.shape
    flat = x.reshape(N, T, H * W * C)
    mu = flat.mean(dim=-1, keepdim=True)
    var = ((flat - mu) ** 2).mean(dim=-1, keepdim=True)
    normed = ((flat - mu) / torch.sqrt(var + EPS)).reshape(N, T, H, W, C)
    return normed * (1.0 + gamma.reshape(1, T, H, W, C)) + beta.reshape(1, T, H, W, C)


# --------------------------------------------------------------------------------------
# attention pieces
# --------------------------------------------------------------------------------------
def mha(sd: SD, p: str, q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor] = None,
        num_heads: int = NUM_HEADS) -> Tensor:
    """nn.MultiheadAttention forward (eval, batch_first=False): q (Lq,B,C), k/v (Lk,B,C).
    Used at models/VidHRFormer.py:104,221,239,298.  mask: bool (Lq,Lk), True = blocked."""
    Lq, B, C = q.shape
    Lk = k.shape[0]
    d = C // num_heads
    W, b = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
    Q = q @ W[:C].t() + b[:C]
    K = k @ W[C:2 * C].t() + b[C:2 * C]
    V = v @ W[2 * C:].t() + b[2 * C:]
    Q = Q.reshape(Lq, B, num_heads, d).permute(1, 2, 0, 3) * (1.0 / math.sqrt(d))
    K = K.reshape(Lk, B, num_heads, d).permute(1, 2, 0, 3)
    V = V.reshape(Lk, B, num_heads, d).permute(1, 2, 0, 3)
    s = Q @ K.transpose(-1, -2)                           # (B,h,Lq,Lk)
    if mask is not None:
        s = s.masked_fill(mask.view(1, 1, Lq, Lk), float("-inf"))
    a = torch.softmax(s, dim=-1)
    o = (a @ V).permute(2, 0, 1, 3).reshape(Lq, B, C)
    return o @ sd[p + "out_proj.weight"].t() + sd[p + "out_proj.bias"]


def window_partition(x: Tensor, ws: int = WINDOW) -> Tensor:
    """LocalPermuteModule.permute, models/VidHRFormer.py:447-462:
    (F,H,W,C) -> (ws*ws, F*(H/ws)*(W/ws), C); sequence index ph*ws+pw, batch index (f,qh,qw)."""
    Fr, H, W, C = x.shape
    x = x.reshape(Fr, H // ws, ws, W // ws, ws, C)        # f qh ph qw pw c
    return x.permute(2, 4, 0, 1, 3, 5).reshape(ws * ws, Fr * (H // ws) * (W // ws), C)


def window_reverse(x: Tensor, Fr: int, H: int, W: int, ws: int = WINDOW) -> Tensor:
    """LocalPermuteModule.rev_permute, models/VidHRFormer.py:464-475."""
    C = x.shape[-1]
    x = x.reshape(ws, ws, Fr, H // ws, W // ws, C)        # ph pw f qh qw c
    return x.permute(2, 3, 0, 4, 1, 5).reshape(Fr, H, W, C)


def slmhsa(sd: SD, p: str, x_qk: Tensor, value: Tensor) -> Tensor:
    """SpatialLocalMultiheadAttention.forward, models/VidHRFormer.py:274-307 (8x8 grid, ws 4: PadBlock no-op)."""
    N, T, H, W, C = x_qk.shape
    qk = window_partition(x_qk.reshape(N * T, H, W, C))
    vv = window_partition(value.reshape(N * T, H, W, C))
    out = mha(sd, p + "attn.", qk, qk, vv)
    return window_reverse(out, N * T, H, W).reshape(N, T, H, W, C)


def mlp_dwbn(sd: SD, p: str, x: Tensor) -> Tensor:
    """MlpDWBN.forward with AR_model=True (LayerNorm over (C,H,W)), models/VidHRFormer.py:374-392."""
    N, T, H, W, C = x.shape
    y = x.reshape(N * T, H, W, C).permute(0, 3, 1, 2)
    y = F.conv2d(y, sd[p + "fc1.weight"], sd[p + "fc1.bias"])
    y = _gelu(F.layer_norm(y, y.shape[1:], sd[p + "norm1.weight"], sd[p + "norm1.bias"], EPS))
    y = F.conv2d(y, sd[p + "dw3x3.weight"], sd[p + "dw3x3.bias"], padding=1, groups=y.shape[1])
    y = _gelu(F.layer_norm(y, y.shape[1:], sd[p + "norm2.weight"], sd[p + "norm2.bias"], EPS))
    y = F.conv2d(y, sd[p + "fc2.weight"], sd[p + "fc2.bias"])
    y = _gelu(F.layer_norm(y, y.shape[1:], sd[p + "norm3.weight"], sd[p + "norm3.bias"], EPS))
    return y.permute(0, 2, 3, 1).reshape(N, T, H, W, -1)


def _to_seq(x: Tensor) -> Tensor:
    """(N,T,H,W,C) -> (T, N*H*W, C), batch index (n*H+h)*W+w  (models/VidHRFormer.py:94)."""
    N, T, H, W, C = x.shape
    return x.permute(1, 0, 2, 3, 4).reshape(T, N * H * W, C)


def _from_seq(x: Tensor, N: int, H: int, W: int) -> Tensor:
    T, _, C = x.shape
    return x.reshape(T, N, H, W, C).permute(1, 0, 2, 3, 4)


# --------------------------------------------------------------------------------------
# transformer blocks
# --------------------------------------------------------------------------------------
def enc_block(sd: SD, p: str, x: Tensor, beta: Tensor, gamma: Tensor) -> Tensor:
    """VidHRFormerBlockEnc.forward, models/VidHRFormer.py:79-116 (dropout/drop_path identity in eval)."""
    N, T, H, W, C = x.shape
    x1 = _layer_norm_c(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    x = x + slmhsa(sd, p + "SLMHSA.", pos_fuse(x1, beta, gamma), x1)
    x = x + mlp_dwbn(sd, p + "SpatialFFN.", _layer_norm_c(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"]))
    x1 = _layer_norm_c(x, sd[p + "norm3.weight"], sd[p + "norm3.bias"])
    temp = pos_fuse(x1, beta, gamma)
    # mask quirk (VidHRFormer.py:100-102): frames 0..T-2 may not attend to the last frame
    mask = torch.zeros(T, T, dtype=torch.bool, device=x.device)
    mask[0:-1, -1] = True
    x = x + _from_seq(mha(sd, p + "temporal_MHSA.", _to_seq(temp), _to_seq(temp), _to_seq(x1), mask), N, H, W)
    x1 = _layer_norm_c(x, sd[p + "norm4.weight"], sd[p + "norm4.bias"])
    x1 = _gelu(x1 @ sd[p + "linear1.weight"].t() + sd[p + "linear1.bias"])
    x = x + (x1 @ sd[p + "linear2.weight"].t() + sd[p + "linear2.bias"])
    return x


def evt_former(sd: SD, p: str, src: Tensor, beta: Tensor, gamma: Tensor, num_layers: int) -> Tensor:
    """VidHRFormerEncoder.forward (evt_token False), models/VidHRFormer.py:25-52. src (N,T,C,H,W)."""
    x = src.permute(0, 1, 3, 4, 2)
    for i in range(num_layers):
        x = enc_block(sd, f"{p}layers.{i}.", x, beta, gamma)
    x = _layer_norm_c(x, sd[p + "norm.weight"], sd[p + "norm.bias"])
    return x.permute(0, 1, 4, 2, 3)


def dec_block(sd: SD, p: str, tgt: Tensor, qe: Tensor, memory: Tensor,
              mem_pos: Tuple[Tensor, Tensor], tgt_pos: Tuple[Tensor, Tensor]) -> Tensor:
    """VidHRFormerBlockDecNAR.forward, models/VidHRFormer.py:198-245. All (N,T,H,W,C)."""
    N, T2, H, W, C = tgt.shape
    a = _layer_norm_c(tgt, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    y = tgt + slmhsa(sd, p + "SLMHSA.", pos_fuse(a + qe, *tgt_pos), a)
    y = y + mlp_dwbn(sd, p + "SpatialFFN.", _layer_norm_c(y, sd[p + "norm2.weight"], sd[p + "norm2.bias"]))
    a = _layer_norm_c(y, sd[p + "norm3.weight"], sd[p + "norm3.bias"])
    s = pos_fuse(a, *tgt_pos)
    y = y + _from_seq(mha(sd, p + "temporal_MHSA.", _to_seq(s), _to_seq(s), _to_seq(a)), N, H, W)
    f = _layer_norm_c(y, sd[p + "norm4.weight"], sd[p + "norm4.bias"])
    f = _gelu(f @ sd[p + "linear1.weight"].t() + sd[p + "linear1.bias"])
    y = y + (f @ sd[p + "linear2.weight"].t() + sd[p + "linear2.bias"])
    a = _layer_norm_c(y, sd[p + "norm5.weight"], sd[p + "norm5.bias"])
    key = pos_fuse(memory, *mem_pos)
    query = pos_fuse(a + qe, *tgt_pos)
    y = y + _from_seq(mha(sd, p + "EncDecAttn.", _to_seq(query), _to_seq(key), _to_seq(memory)), N, H, W)
    y = y + mlp_dwbn(sd, p + "SpatialFFN1.", _layer_norm_c(y, sd[p + "norm6.weight"], sd[p + "norm6.bias"]))
    return y


def decoder_nar(sd: SD, p: str, query_evt: Tensor, memory: Tensor, mem_pos, tgt_pos, num_layers: int) -> Tensor:
    """VidHRformerDecoderNAR.forward, models/VidHRFormer.py:126-161. In/out (N,T,C,H,W)."""
    qe = query_evt.permute(0, 1, 3, 4, 2)
    mem = memory.permute(0, 1, 3, 4, 2)
    y = torch.zeros_like(qe)
    for i in range(num_layers):
        y = dec_block(sd, f"{p}layers.{i}.", y, qe, mem, mem_pos, tgt_pos)
    y = _layer_norm_c(y, sd[p + "norm.weight"], sd[p + "norm.bias"])
    return torch.relu(y.permute(0, 1, 4, 2, 3))


def event_encoder(sd: SD, p: str, x: Tensor, stochastic: bool, eps: Optional[Tensor] = None):
    """EventEncoder.forward / reparameterize, models/submodules.py:388-410 (n_layers=1). x (N,C,H,W)."""
    C = x.shape[1]
    y = torch.relu(_bn_eval(sd, p + "conv1.1.", F.conv2d(x, sd[p + "conv1.0.weight"], None, padding=1, groups=C)))
    y = torch.relu(_bn_eval(sd, p + "conv2.1.", F.conv2d(y, sd[p + "conv2.0.weight"], None, padding=1)))
    y = torch.relu(_bn_eval(sd, p + "MLP_0.1.", F.conv2d(y, sd[p + "MLP_0.0.weight"], None)))
    mu = F.conv2d(y, sd[p + "mu_net.weight"], sd[p + "mu_net.bias"])
    if not stochastic:
        return mu
    logvar = F.conv2d(y, sd[p + "logvar_net.weight"], sd[p + "logvar_net.bias"])
    if eps is None:
        eps = torch.randn(mu.shape, device=mu.device)                     # submodules.py:409
    return mu + torch.exp(0.5 * logvar) * eps, mu, logvar


def predictor_forward(sd: SD, observed_features: Tensor, observed_coor: Tensor, predict_coor: Tensor,
                      stochastic: bool, eps: Optional[Tensor] = None, fuse_method: str = "Add",
                      evt_layers: int = 4, dec_layers: int = 8, prefix: str = "",
                      return_latent: bool = False, predict_features_gt: Optional[Tensor] = None,
                      decode_with_posterior: bool = False, eps_p: Optional[Tensor] = None):
    """Predictor.forward in eval mode, models/Predictor.py:301-350.  observed_features (N,To,C,H,W) -> (N,Tp,C,H,W).
    With ``predict_features_gt`` (N,Tp,C,H,W) the stochastic model also runs the posterior on the ground-truth future
    (Predictor.py:311-313) and returns (out, mu_o, logvar_o, mu_p, logvar_p) (:324-327) - the KL / ELBO evaluation path;
    in eval mode the decoder is still queried with the PRIOR sample z_o (:320-322).  The deterministic model ignores it (:328-335).
    ``decode_with_posterior``: the branch the reference takes when ``self.training`` is set (:315-318, forward only here): the
    decoder is queried with the POSTERIOR sample z_p = mu_p + exp(logvar_p / 2) eps_p - the reconstruction term of the ELBO.
    ``eps_p``: the posterior's noise (the reference's second ``torch.randn`` draw; defaults to ``eps``)."""
    p = prefix
    Tp = predict_coor.shape[0] // (observed_features.shape[-1] * observed_features.shape[-2])
    op = nrmlp(sd, p + "nrmlp.", observed_coor, fuse_method)
    pp = nrmlp(sd, p + "nrmlp.", predict_coor, fuse_method)
    memory = evt_former(sd, p + "EVT_Former.", observed_features, op[0], op[1], evt_layers)
    evt = memory.mean(dim=1)                              # Predictor.py:346
    if stochastic:
        z, mu, logvar = event_encoder(sd, p + "evt_prior.", evt, True, eps)        # Predictor.py:310
    else:
        z = event_encoder(sd, p + "evt_posterior.", evt, False)                     # Predictor.py:330
        mu, logvar = z, None
    if stochastic and predict_features_gt is not None:
        memory_p = evt_former(sd, p + "EVT_Former.", predict_features_gt, pp[0], pp[1], evt_layers)     # Predictor.py:312
        z_p, mu_p, logvar_p = event_encoder(sd, p + "evt_posterior.", memory_p.mean(dim=1), True,
                                            eps if eps_p is None else eps_p)                            # :313
        if decode_with_posterior:                                                                       # :315-318
            z = z_p
    else:
        assert not decode_with_posterior, "please input groundtruth predict features for storchastic model training/val"   # :316
    query_evt = z.unsqueeze(1).repeat(1, Tp, 1, 1, 1)
    out = decoder_nar(sd, p + "transformer.", query_evt, memory, op, pp, dec_layers)
    if return_latent:
        return out, memory, evt, z, mu, logvar
    if stochastic and predict_features_gt is not None:
        return out, mu, logvar, mu_p, logvar_p
    return out


# --------------------------------------------------------------------------------------
# ResNet autoencoder
# --------------------------------------------------------------------------------------
def nonlocal_attn2d(sd: SD, p: str, x: Tensor) -> Tensor:
    """NonLocalAttenion2D.forward, models/submodules.py:138-168: unscaled softmax, 2x2 max-pooled k/v."""
    N, C, H, W = x.shape
    t = x.flatten(2, 3).permute(0, 2, 1)                  # (N,HW,C)
    q = t @ sd[p + "Wq.weight"].t() + sd[p + "Wq.bias"]
    k = (t @ sd[p + "Wk.weight"].t() + sd[p + "Wk.bias"]).reshape(N, H, W, -1).permute(0, 3, 1, 2)
    v = (t @ sd[p + "Wv.weight"].t() + sd[p + "Wv.bias"]).reshape(N, H, W, -1).permute(0, 3, 1, 2)
    k = F.max_pool2d(k, 2, 2).flatten(2, 3)               # (N,dq,HW/4)
    v = F.max_pool2d(v, 2, 2).flatten(2, 3).permute(0, 2, 1)
    a = torch.softmax(q @ k, dim=-1)
    o = (a @ v) @ sd[p + "out_proj.weight"].t() + sd[p + "out_proj.bias"]
    o = o.reshape(N, H, W, C).permute(0, 3, 1, 2)
    o = torch.relu(_bn_eval(sd, p + "norm_func.", o))
    return x + sd[p + "gamma"] * o


def f3d_conv_attn(sd: SD, p: str, x: Tensor) -> Tensor:
    """Factorized3DConvAttn.conv_forward with learn_3d=False, models/submodules.py:48-70."""
    y = F.conv2d(x, sd[p + "spatial_conv.0.weight"], sd[p + "spatial_conv.0.bias"], padding=1)
    y = torch.relu(_bn_eval(sd, p + "spatial_conv.1.", y)) + x
    y = nonlocal_attn2d(sd, p + "attn2d.", y)
    return y + x


def resnet_block(sd: SD, p: str, x: Tensor) -> Tensor:
    """ResnetBlock.forward with reflect padding, models/ResNetAutoEncoder.py:220-261."""
    y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), sd[p + "conv_block.1.weight"], None)
    y = torch.relu(_bn_eval(sd, p + "conv_block.2.", y))
    y = F.conv2d(F.pad(y, (1, 1, 1, 1), mode="reflect"), sd[p + "conv_block.5.weight"], None)
    return x + _bn_eval(sd, p + "conv_block.6.", y)


def resnet_encoder(sd: SD, x: Tensor, n_downsampling: int, num_res_blocks: int, prefix: str = "") -> Tensor:
    """ResnetEncoder.forward, models/ResNetAutoEncoder.py:120-146. x (N,T,Cin,H,W) -> (N,T,ngf*2^n,h,w)."""
    p = prefix
    N, T = x.shape[:2]
    y = x.flatten(0, 1)
    y = F.conv2d(F.pad(y, (3, 3, 3, 3), mode="reflect"), sd[p + "block0.1.weight"], None)
    y = torch.relu(_bn_eval(sd, p + "block0.2.", y))
    y = torch.relu(_bn_eval(sd, p + "block1.1.", F.conv2d(y, sd[p + "block1.0.weight"], None, stride=2, padding=1)))
    for i in range(1, n_downsampling):
        y = f3d_conv_attn(sd, f"{p}block{i + 1}_3dConvAttn.", y)
        y = F.conv2d(y, sd[f"{p}block{i + 1}_conv.0.weight"], None, stride=2, padding=1)
        y = torch.relu(_bn_eval(sd, f"{p}block{i + 1}_conv.1.", y))
    for i in range(num_res_blocks):
        y = f3d_conv_attn(sd, f"{p}res_3dConvAttn_{i}.", y)
        y = resnet_block(sd, f"{p}res_conv_{i}.", y)
    y = torch.relu(y)
    return y.reshape(N, T, *y.shape[1:])


def resnet_decoder(sd: SD, x: Tensor, n_downsampling: int, out_layer: str = "Tanh", prefix: str = "") -> Tensor:
    """ResnetDecoder.forward, models/ResNetAutoEncoder.py:149-204. x (N,T,C,h,w) -> (N,T,Cimg,H,W)."""
    p = prefix
    N, T = x.shape[:2]
    y = x.flatten(0, 1)
    for i in range(n_downsampling):
        y = F.conv_transpose2d(y, sd[f"{p}model.{3 * i}.weight"], None, stride=2, padding=1, output_padding=1)
        y = torch.relu(_bn_eval(sd, f"{p}model.{3 * i + 1}.", y))
    head = 3 * n_downsampling + 1
    y = F.conv2d(F.pad(y, (3, 3, 3, 3), mode="reflect"), sd[f"{p}model.{head}.weight"], sd[f"{p}model.{head}.bias"])
    if out_layer == "Tanh":
        y = torch.tanh(y)
    elif out_layer == "Sigmoid":
        y = torch.sigmoid(y)
    else:
        raise ValueError("Unsupported output layer")
    return y.reshape(N, T, *y.shape[1:])


# --------------------------------------------------------------------------------------
# whole path (what LitPredictor.forward does for the predicted frames, Predictor.py:72-86)
# --------------------------------------------------------------------------------------
def npvp_predict_frames(enc_sd: SD, pred_sd: SD, dec_sd: SD, past_frames: Tensor, cfg: dict,
                        observed_coor: Tensor, predict_coor: Tensor, eps: Optional[Tensor] = None) -> Tensor:
    """Enc(context) -> Predictor -> Dec(predictions).  cfg keys: n_downsampling, num_res_blocks,
    out_layer, stochastic, fuse_method, evt_layers, dec_layers."""
    feats = resnet_encoder(enc_sd, past_frames, cfg["n_downsampling"], cfg["num_res_blocks"])
    pred = predictor_forward(pred_sd, feats, observed_coor, predict_coor, cfg["stochastic"], eps,
                             cfg.get("fuse_method", "Add"), cfg.get("evt_layers", 4), cfg.get("dec_layers", 8))
    return resnet_decoder(dec_sd, pred, cfg["n_downsampling"], cfg["out_layer"])


def psnr(x: Tensor, y: Tensor) -> Tensor:
    """PSNR as utils/metrics.py:12-30: -10 log10(mse + 1e-8), mse per image over (C,H,W), mean over batch*time."""
    mse = ((x - y) ** 2).flatten(-3).mean(dim=-1)
    return (-10.0 * torch.log10(mse + 1e-8)).mean()
