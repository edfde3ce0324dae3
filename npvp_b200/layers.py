"""Parameter containers for the NPVP hot path.

These classes only *own* parameters under the reference's state_dict key names
 so that reference checkpoints load with ``strict=True`` and so
that constructing them under a seed draws the same random numbers, in the same order,
as the reference constructors do.  They contain no math: all compute is done by the
sm_100a kernels driven from ``engine_predictor.py`` / ``engine_autoencoder.py``.
Calling ``forward`` on a container is an error by design (there is no eager fallback).
"""
from __future__ import annotations

import copy
import math

import torch
import torch.nn as nn


class _Holder(nn.Module):
    """Base class: a bag of parameters.  Never executed."""

    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(f"{type(self).__name__} is a parameter container; compute runs in the CUDA engine")


def _slot() -> nn.Module:
    """Parameter-free placeholder that keeps nn.Sequential indices aligned with the reference keys."""
    return nn.Identity()


def _bn_only(norm_layer) -> bool:
    """Bias rule of the reference AE (ResNetAutoEncoder.py:64-67): conv bias only with InstanceNorm."""
    import functools
    base = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
    if base is not nn.BatchNorm2d:
        raise NotImplementedError("npvp_b200 folds eval-mode BatchNorm2d into the conv kernels; "
                                  f"norm_layer={base} is not supported")
    return False


# ------------------------------------------------------------------------------------
# autoencoder pieces  (reference: models/submodules.py:9-180, models/ResNetAutoEncoder.py:207-261)
# ------------------------------------------------------------------------------------
class NonLocalAttenion2D(_Holder):
    """Keys: Wq, Wk, Wv, out_proj (Linear), gamma (scalar), norm_func (BN)."""

    def __init__(self, in_channels, atten_channels_downsample_ratio=8, value_channels_downsample_ratio=2,
                 bias=True, learn_gamma=True, norm_func=None, activ_func=None):
        super().__init__()
        self.in_channels = in_channels
        self.attn_dim = in_channels // atten_channels_downsample_ratio
        self.value_dim = in_channels // value_channels_downsample_ratio
        self.bias = bias
        self.Wq = nn.Linear(in_channels, self.attn_dim, bias=bias)
        self.Wk = nn.Linear(in_channels, self.attn_dim, bias=bias)
        self.Wv = nn.Linear(in_channels, self.value_dim, bias=bias)
        self.out_proj = nn.Linear(self.value_dim, in_channels, bias=bias)
        self.learn_gamma = learn_gamma
        if learn_gamma:
            self.gamma = nn.Parameter(torch.zeros((), dtype=torch.float32))
        else:
            self.gamma = 1.0
        self.norm_func = norm_func if norm_func is not None else nn.Identity()
        for lin in (self.Wq, self.Wk, self.Wv, self.out_proj):
            if bias:
                nn.init.zeros_(lin.bias)
        for lin in (self.Wq, self.Wk, self.Wv, self.out_proj):
            nn.init.xavier_uniform_(lin.weight)


class Factorized3DConvAttn(_Holder):
    """Keys: spatial_conv.{0,1}, attn2d.*  (learn_3d=False only: every shipped YAML sets it False)."""

    def __init__(self, in_channels, atten_channels_downsample_ratio=8, value_channels_downsample_ratio=2,
                 use_bias=True, learn_gamma=True, norm_layer_2d=nn.BatchNorm2d, norm_layer_1d=nn.BatchNorm1d,
                 activ_func=None, conv_first=True, learn_3d=True):
        super().__init__()
        if learn_3d:
            raise NotImplementedError("learn_3d=True (temporal conv/attention branch of the AE) is outside the "
                                      "B200 hot path: all reference configs use learn_3d: False")
        if not conv_first:
            raise NotImplementedError("conv_first=False is never used by the reference encoders")
        self.in_channels = in_channels
        self.learn_3d = False
        self.spatial_conv = nn.Sequential(
            nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1, bias=use_bias),
            norm_layer_2d(in_channels), _slot())
        self.attn2d = NonLocalAttenion2D(in_channels, atten_channels_downsample_ratio,
                                         value_channels_downsample_ratio, True, learn_gamma,
                                         norm_layer_2d(in_channels))


class ResnetBlock(_Holder):
    """Keys: conv_block.{1,2,5,6} (indices shift by one after a Dropout slot, as in the reference)."""

    def __init__(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        super().__init__()
        if padding_type not in ("reflect", "replicate", "zero"):
            raise NotImplementedError('padding [%s] is not implemented' % padding_type)
        self.padding_type = padding_type
        seq = []
        for half in range(2):
            if padding_type != "zero":
                seq.append(_slot())
            seq.append(nn.Conv2d(dim, dim, kernel_size=3, padding=0 if padding_type != "zero" else 1, bias=use_bias))
            seq.append(norm_layer(dim))
            if half == 0:
                seq.append(_slot())
                if use_dropout:
                    seq.append(_slot())          # Dropout(0.5) is identity in eval
        self.conv_block = nn.Sequential(*seq)

    def convs(self):
        mods = [m for m in self.conv_block if isinstance(m, nn.Conv2d)]
        norms = [m for m in self.conv_block if isinstance(m, nn.BatchNorm2d)]
        return list(zip(mods, norms))


# ------------------------------------------------------------------------------------
# predictor pieces  (reference: models/submodules.py:258-454, models/VidHRFormer.py)
# ------------------------------------------------------------------------------------
class CoorGenerator(nn.Module):
    """Normalised (t,h,w) grid, rows t-major; reference models/submodules.py:329-366.
    Runs on the host exactly like the reference (it is a few hundred floats)."""

    def __init__(self, max_H, max_W, max_T):
        super().__init__()
        self.max_H, self.max_W, self.max_T = max_H, max_W, max_T

    def forward(self, t_list, h_list, w_list):
        assert torch.max(h_list) <= self.max_H and torch.min(h_list) >= 0., "Invalid H coordinates"
        assert torch.max(w_list) <= self.max_W and torch.min(w_list) >= 0., "Invalid W coordinates"
        assert torch.max(t_list) <= self.max_T and torch.min(t_list) >= 0., "Invalid T coordinates"
        T, H, W = t_list.shape[0], h_list.shape[0], w_list.shape[0]
        grid = torch.empty(T, H, W, 3, dtype=torch.result_type(t_list, torch.tensor(1.0)), device=t_list.device)
        grid[..., 0] = (t_list / self.max_T).reshape(T, 1, 1)
        grid[..., 1] = (h_list / self.max_H).reshape(1, H, 1).to(grid.device)
        grid[..., 2] = (w_list / self.max_W).reshape(1, 1, W).to(grid.device)
        return grid.reshape(T * H * W, 3)


class NRMLP(_Holder):
    """Fourier-feature MLP.  Keys: B, MLP.{0,2,4}, mlp_beta (+ mlp_gamma for 'SPADE')."""

    def __init__(self, out_channels, dim_x=3, d_model=256, MLP_layers=4, scale=10, fix_B=False, fuse_method='SPADE'):
        super().__init__()
        self.out_channels, self.dim_x, self.d_model = out_channels, dim_x, d_model
        self.MLP_layers, self.scale, self.fix_B, self.fuse_method = MLP_layers, scale, fix_B, fuse_method
        first = nn.Linear(2 * d_model, d_model)                      # drawn before B, like the reference
        B = torch.normal(mean=0, std=1.0, size=(d_model, dim_x)) * scale
        if fix_B:
            self.register_buffer('B', B)
        else:
            self.B = nn.Parameter(B, requires_grad=True)
        seq = [first, _slot()]
        for _ in range(MLP_layers - 2):
            seq += [nn.Linear(d_model, d_model), _slot()]
        self.MLP = nn.Sequential(*seq)
        self.mlp_beta = nn.Linear(d_model, out_channels)
        if fuse_method == 'SPADE':
            self.mlp_gamma = nn.Linear(d_model, out_channels)

    def linears(self):
        return [m for m in self.MLP if isinstance(m, nn.Linear)]


class PosFeatFuser(_Holder):
    """Parameter-free GroupNorm(1,C) + positional beta/gamma (models/submodules.py:412-454)."""

    def __init__(self, x_channels, param_free_norm_type='layer'):
        super().__init__()
        if param_free_norm_type not in ('instance', 'syncbatch', 'batch', 'layer'):
            raise ValueError('%s is not a recognized param-free norm type in SPADE' % param_free_norm_type)
        if param_free_norm_type != 'layer':
            raise NotImplementedError("only param_free_norm_type='layer' (used by every reference config) has a kernel")
        self.x_channels = x_channels
        self.param_free_norm_type = param_free_norm_type


class EventEncoder(_Holder):
    """Latent heads.  Keys: conv1.{0,1}, conv2.{0,1}, MLP_i.{0,1}, mu_net, logvar_net (stochastic)."""

    def __init__(self, in_channels, hidden_channels, n_layers, stochastic):
        super().__init__()
        self.stochastic, self.n_layers = stochastic, n_layers
        self.in_channels, self.hidden_channels = in_channels, hidden_channels
        self.conv1 = nn.Sequential(
            nn.Conv2d(in_channels, in_channels, 3, 1, 1, bias=False, groups=in_channels),
            nn.BatchNorm2d(in_channels), _slot())
        self.conv2 = nn.Sequential(
            nn.Conv2d(in_channels, hidden_channels, 3, 1, 1, bias=False),
            nn.BatchNorm2d(hidden_channels), _slot())
        for i in range(n_layers):
            setattr(self, f'MLP_{i}', nn.Sequential(
                nn.Conv2d(hidden_channels, hidden_channels, 1, 1, bias=False),
                nn.BatchNorm2d(hidden_channels), _slot()))
        self.mu_net = nn.Conv2d(hidden_channels, in_channels, 1, 1, bias=True)
        if stochastic:
            self.logvar_net = nn.Conv2d(hidden_channels, in_channels, 1, 1, bias=True)


class SpatialLocalMultiheadAttention(_Holder):
    def __init__(self, embed_dim, num_heads, window_size=7, dropout=0.):
        super().__init__()
        self.dim, self.num_heads, self.window_size, self.dropout = embed_dim, num_heads, window_size, dropout
        self.attn = nn.MultiheadAttention(embed_dim, num_heads, dropout=dropout)


class MlpDWBN(_Holder):
    """Conv-FFN container, AR_model=True layout: LayerNorm over (C, encH, encW) with elementwise affine."""

    def __init__(self, encH, encW, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU,
                 dw_act_layer=nn.GELU, drop=0.0, AR_model=True):
        super().__init__()
        if not AR_model:
            raise NotImplementedError("MlpDWBN with BatchNorm (AR_model=False) is never built by the reference predictor")
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Conv2d(in_features, hidden_features, kernel_size=1)
        self.norm1 = nn.LayerNorm((hidden_features, encH, encW))
        self.dw3x3 = nn.Conv2d(hidden_features, hidden_features, 3, 1, padding=1, groups=hidden_features)
        self.norm2 = nn.LayerNorm((hidden_features, encH, encW))
        self.fc2 = nn.Conv2d(hidden_features, out_features, kernel_size=1)
        self.norm3 = nn.LayerNorm((out_features, encH, encW))
        self.out_features, self.hidden_features = out_features, hidden_features


class VidHRFormerBlockEnc(_Holder):
    def __init__(self, encH, encW, embed_dim, num_heads, window_size=7, dropout=0., drop_path=0.,
                 Spatial_FFN_hidden_ratio=4, dim_feedforward=1024):
        super().__init__()
        self.embed_dim, self.num_heads, self.window_size = embed_dim, num_heads, window_size
        self.SLMHSA = SpatialLocalMultiheadAttention(embed_dim, num_heads, window_size, dropout)
        self.SpatialFFN = MlpDWBN(encH, encW, embed_dim, int(Spatial_FFN_hidden_ratio * embed_dim), embed_dim, drop=dropout)
        self.norm1 = nn.LayerNorm(embed_dim)
        self.norm2 = nn.LayerNorm(embed_dim)
        self.norm3 = nn.LayerNorm(embed_dim)
        self.temporal_MHSA = nn.MultiheadAttention(embed_dim, num_heads, dropout=dropout)
        self.linear1 = nn.Linear(embed_dim, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, embed_dim)
        self.norm4 = nn.LayerNorm(embed_dim)


class VidHRFormerBlockDecNAR(_Holder):
    def __init__(self, encH, encW, embed_dim, num_heads, window_size=7, dropout=0., drop_path=0.,
                 Spatial_FFN_hidden_ratio=4, dim_feedforward=1024):
        super().__init__()
        self.embed_dim, self.num_heads, self.window_size = embed_dim, num_heads, window_size
        hid = int(Spatial_FFN_hidden_ratio * embed_dim)
        self.SLMHSA = SpatialLocalMultiheadAttention(embed_dim, num_heads, window_size, dropout)
        self.SpatialFFN = MlpDWBN(encH, encW, embed_dim, hid, embed_dim, drop=dropout)
        self.norm1 = nn.LayerNorm(embed_dim)
        self.norm2 = nn.LayerNorm(embed_dim)
        self.norm3 = nn.LayerNorm(embed_dim)
        self.temporal_MHSA = nn.MultiheadAttention(embed_dim, num_heads, dropout=dropout)
        self.linear1 = nn.Linear(embed_dim, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, embed_dim)
        self.norm4 = nn.LayerNorm(embed_dim)
        self.EncDecAttn = nn.MultiheadAttention(embed_dim, num_heads, dropout=dropout)
        self.SpatialFFN1 = MlpDWBN(encH, encW, embed_dim, hid, embed_dim, drop=dropout)
        self.norm5 = nn.LayerNorm(embed_dim)
        self.norm6 = nn.LayerNorm(embed_dim)


def _replicate(block: nn.Module, n: int) -> nn.ModuleList:
    """All layers of a stack start as deep copies of one initialised block (VidHRFormer.py:545-546)."""
    return nn.ModuleList(copy.deepcopy(block) for _ in range(n))


class VidHRFormerEncoder(_Holder):
    def __init__(self, num_layers, enc_H, enc_W, d_model, num_heads, window_size=7, dropout=0., drop_path=0.,
                 Spatial_FFN_hidden_ratio=4, dim_feedforward=1024, norm=None, evt_token=False):
        super().__init__()
        if evt_token:
            raise NotImplementedError("learn_evt_token=True is hard-coded off by the reference (Predictor.py:46)")
        self.layers = _replicate(VidHRFormerBlockEnc(enc_H, enc_W, d_model, num_heads, window_size, dropout, drop_path,
                                                     Spatial_FFN_hidden_ratio, dim_feedforward), num_layers)
        self.num_layers, self.norm, self.evt_token = num_layers, norm, False


class VidHRformerDecoderNAR(_Holder):
    def __init__(self, num_layers, encH, encW, embed_dim, num_heads, window_size=7, dropout=0., drop_path=0.,
                 Spatial_FFN_hidden_ratio=4, dim_feedforward=1024, norm=None, return_intermediate=False):
        super().__init__()
        if return_intermediate:
            raise NotImplementedError("return_intermediate=True is never used on the inference path")
        self.layers = _replicate(VidHRFormerBlockDecNAR(encH, encW, embed_dim, num_heads, window_size, dropout,
                                                        drop_path, Spatial_FFN_hidden_ratio, dim_feedforward), num_layers)
        self.num_layers, self.norm, self.return_intermediate = num_layers, norm, False
