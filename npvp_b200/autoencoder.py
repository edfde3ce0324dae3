"""Drop-in ``ResnetEncoder`` / ``ResnetDecoder`` (reference: models/ResNetAutoEncoder.py:51-204).

Same constructor signatures, forward signatures and state_dict keys as the reference;
the forward pass runs on sm_100a kernels through the C-ABI library (no eager fallback).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .layers import Factorized3DConvAttn, ResnetBlock, _bn_only, _slot


class _EngineModule(nn.Module):
    """Shared plumbing: inference-only guard + lazily built packed-weight cache."""

    _engine_cls = None

    def _guard(self, x: torch.Tensor):
        if self.training:
            raise NotImplementedError(f"{type(self).__name__}: npvp_b200 implements the inference path only; call .eval()")
        if not x.is_cuda:
            raise NotImplementedError(f"{type(self).__name__}: input must be a CUDA tensor (there is no CPU fallback)")

    _NOT_WEIGHTS = ("observed_coor", "predict_coor", "all_coor")      # derived coordinate buffers: not packed into the engine

    def _weights_version(self):
        """Cheap fingerprint of the weights the engine packed: (sum of the tensors' in-place modification counters, tensor
        count, device).  The counters only grow, so the sum changes whenever any weight is written (load_state_dict, optimiser
        step, manual edits).  The tensor list is cached - walking the module tree (~800 tensors) on every call cost 0.3 ms per
        module and forward, visible at batch 1 - and dropped whenever ``.to()`` / ``.cuda()`` / ``.half()`` may have replaced
        tensors (``_apply``)."""
        ts = self.__dict__.get("_ver_tensors")
        if ts is None:
            ts = [t for _, t in self.named_parameters()] + [t for n, t in self.named_buffers() if n.rsplit(".", 1)[-1] not in self._NOT_WEIGHTS]
            self.__dict__["_ver_tensors"] = ts
        v = 0
        for t in ts:
            v += t._version
        return v, len(ts), ts[0].device

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_ver_tensors", None)
        self.__dict__.pop("_eng", None)
        return super()._apply(fn, *args, **kwargs)

    def _on_device(self, x: torch.Tensor):
        """Context that makes the input's GPU the current device: the C-ABI launches on ``torch.cuda.current_stream()``, i.e. in
        the CURRENT device's context - a model built on cuda:1 while cuda:0 is current would otherwise launch on the wrong GPU."""
        return torch.cuda.device(x.device)

    def _engine(self):
        key = self._weights_version()
        eng = self.__dict__.get("_eng")
        if eng is None or self.__dict__.get("_eng_key") != key:
            eng = self._build_engine()
            self.__dict__["_eng"], self.__dict__["_eng_key"] = eng, key
        return eng


class ResnetEncoder(_EngineModule):
    def __init__(self, input_nc, ngf=64, n_downsampling=3, num_res_blocks=2, norm_layer=nn.BatchNorm2d,
                 norm_layer1d=nn.BatchNorm1d, use_dropout=False, padding_type='reflect', learn_3d=True):
        super().__init__()
        use_bias = _bn_only(norm_layer)
        self.input_nc, self.ngf0 = input_nc, ngf
        self.n_downsampling, self.num_res_blocks = n_downsampling, num_res_blocks
        self.padding_type = padding_type
        self.block0 = nn.Sequential(_slot(), nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0, bias=use_bias),
                                    norm_layer(ngf), _slot())
        self.block1 = nn.Sequential(nn.Conv2d(ngf, ngf * 2, kernel_size=3, stride=2, padding=1, bias=use_bias),
                                    norm_layer(ngf * 2), _slot())
        ch = ngf * 2
        for i in range(1, n_downsampling):
            setattr(self, f'block{i + 1}_3dConvAttn',
                    Factorized3DConvAttn(in_channels=ch, norm_layer_2d=norm_layer, norm_layer_1d=norm_layer1d, learn_3d=learn_3d))
            setattr(self, f'block{i + 1}_conv',
                    nn.Sequential(nn.Conv2d(ch, ch * 2, kernel_size=3, stride=2, padding=1, bias=use_bias),
                                  norm_layer(ch * 2), _slot()))
            ch *= 2
        for i in range(num_res_blocks):
            setattr(self, f'res_3dConvAttn_{i}',
                    Factorized3DConvAttn(in_channels=ch, norm_layer_2d=norm_layer, norm_layer_1d=norm_layer1d, learn_3d=learn_3d))
            setattr(self, f'res_conv_{i}', ResnetBlock(ch, padding_type=padding_type, norm_layer=norm_layer,
                                                       use_dropout=use_dropout, use_bias=use_bias))
        self.out_channels = ch

    def _build_engine(self):
        from .engine_autoencoder import EncoderEngine
        return EncoderEngine(self)

    def forward(self, x):
        """x: (N, T, C, H, W) fp32 CUDA -> (N, T, ngf*2^n, H/2^n, W/2^n) fp32."""
        self._guard(x)
        with self._on_device(x):
            return self._engine().run(x)

    def forward_tokens(self, x, norm=None):
        """Engine-internal variant: returns channels-last features (N, T, h, w, C) without the NCHW copy.  ``x`` may be uint8
        pixels with ``norm`` = (mean, std): the stem kernel applies VidToTensor + VidNormalize itself."""
        self._guard(x)
        with self._on_device(x):
            return self._engine().run(x, channels_last=True, norm=norm)


class ResnetDecoder(_EngineModule):
    def __init__(self, output_nc, ngf=64, n_downsampling=2, norm_layer=nn.BatchNorm2d, use_dropout=False,
                 padding_type='reflect', out_layer='Tanh'):
        super().__init__()
        use_bias = _bn_only(norm_layer)
        seq = []
        for i in range(n_downsampling):
            mult = 2 ** (n_downsampling - i)
            seq += [nn.ConvTranspose2d(ngf * mult, int(ngf * mult / 2), kernel_size=3, stride=2, padding=1,
                                       output_padding=1, bias=use_bias),
                    norm_layer(int(ngf * mult / 2)), _slot()]
        seq += [_slot(), nn.Conv2d(ngf, output_nc, kernel_size=7, padding=0)]
        if out_layer not in ('Tanh', 'Sigmoid'):
            raise ValueError("Unsupported output layer")
        seq += [_slot()]
        self.model = nn.Sequential(*seq)
        self.output_nc, self.ngf0, self.n_downsampling, self.out_layer = output_nc, ngf, n_downsampling, out_layer

    def _build_engine(self):
        from .engine_autoencoder import DecoderEngine
        return DecoderEngine(self)

    def forward(self, x):
        """x: (N, T, C, h, w) fp32 CUDA -> (N, T, output_nc, H, W) fp32."""
        self._guard(x)
        with self._on_device(x):
            return self._engine().run(x)

    def forward_tokens(self, x_cl, renorm=None, want_f32=True):
        """Engine-internal variant: takes channels-last features (N, T, h, w, C).  ``renorm`` = (mean, std): also returns the
        pixel-space uint8 frames written by the head kernel's fused epilogue -> (frames | None, frames_u8)."""
        self._guard(x_cl)
        with self._on_device(x_cl):
            return self._engine().run(x_cl, channels_last=True, renorm=renorm, want_f32=want_f32)
