"""Drop-in ``Predictor`` (reference: models/Predictor.py:265-359).

NPVP-D (deterministic) and NPVP-S (stochastic) neural-process predictor: continuous-time
coordinates -> NRMLP positional codes, EVT_Former over the context features, latent event
code (mean, or mean + sigma * eps), non-autoregressive VidHRFormer decoder queried at arbitrary
target timestamps.  Constructor / forward signatures, attributes and state_dict keys follow the
reference; compute runs on sm_100a kernels through ``engine_predictor.PredictorEngine``.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .autoencoder import _EngineModule
from .layers import (CoorGenerator, EventEncoder, NRMLP, PosFeatFuser, VidHRFormerEncoder,
                     VidHRformerDecoderNAR)


class Predictor(_EngineModule):
    # NOTE: ``norm=nn.LayerNorm(512)`` is deliberately a shared default instance, as in the reference
    # (Predictor.py:270): EVT_Former.norm and transformer.norm are one module (two state_dict keys).
    def __init__(self, max_H, max_W, max_T, h_list, w_list, to_list, tp_list,
                 embed_dim=512,
                 fuse_method='SPADE', param_free_norm_type='layer',
                 evt_hidden_channels=256, evt_n_layers=1, stochastic=True,
                 transformer_layers=4, num_heads=8, window_size=4, dropout=0.1, drop_path=0.1,
                 Spatial_FFN_hidden_ratio=4, dim_feedforward=1024, norm=nn.LayerNorm(512), return_intermediate=False,
                 evt_former=True, learn_evt_token=False, evt_former_num_layers=4, rand_context=False):
        super().__init__()
        if not evt_former:
            raise NotImplementedError("evt_former=False (mean-pooled event coding without EVT_Former) has no kernel path; "
                                      "every reference config sets evt_former: True")
        if embed_dim != 512 or num_heads != 8 or window_size != 4 or max_H != 8 or max_W != 8:
            raise NotImplementedError("kernels are specialised for embed_dim 512, 8 heads, 4x4 windows on an 8x8 grid "
                                      "(the only geometry the reference configs use)")
        if evt_n_layers != 1:
            raise NotImplementedError("evt_n_layers is hard-coded to 1 by the reference (Predictor.py:45)")
        self.stochastic = stochastic
        self.evt_former = evt_former
        self.h_list, self.w_list = h_list, w_list
        self.max_H, self.max_W, self.max_T = max_H, max_W, max_T
        self.embed_dim, self.fuse_method = embed_dim, fuse_method
        self.rand_context = rand_context
        self.coor_generator = CoorGenerator(max_H, max_W, max_T)
        if not rand_context:
            self.register_buffer("observed_coor", self.coor_generator(to_list, h_list, w_list))
            self.register_buffer("predict_coor", self.coor_generator(tp_list, h_list, w_list))
        else:
            self.observed_coor = None
            self.predict_coor = None
            self.register_buffer('all_coor', self.coor_generator(torch.cat([to_list, tp_list]), h_list, w_list)
                                 .reshape(max_T, max_H, max_W, 3))
        self.nrmlp = NRMLP(out_channels=embed_dim, fuse_method=fuse_method)
        self.fuser = PosFeatFuser(x_channels=embed_dim, param_free_norm_type=param_free_norm_type)
        self.EVT_Former = VidHRFormerEncoder(evt_former_num_layers, max_H, max_W, embed_dim, num_heads, window_size,
                                             dropout, drop_path, Spatial_FFN_hidden_ratio, dim_feedforward, norm,
                                             learn_evt_token)
        self.evt_posterior = EventEncoder(embed_dim, evt_hidden_channels, evt_n_layers, stochastic)
        self.evt_prior = None
        if stochastic:
            self.evt_prior = EventEncoder(embed_dim, evt_hidden_channels, evt_n_layers, stochastic)
        self.TP = tp_list.shape[0]
        self.transformer = VidHRformerDecoderNAR(transformer_layers, max_H, max_W, embed_dim, num_heads, window_size,
                                                 dropout, drop_path, Spatial_FFN_hidden_ratio, dim_feedforward, norm,
                                                 return_intermediate)
        # Latent noise injection: set ``injected_eps`` to an (N,512,8,8) tensor to make the
        # stochastic path reproducible; ``None`` samples torch.randn on the input's device like the reference.
        self.injected_eps = None
        self.last_latent = None
        # Forward-only form of the reference's training-mode branch (Predictor.py:315-318): with ``posterior_decode = True``
        # and ground-truth future features, the decoder is queried with the POSTERIOR sample z_p (reconstruction term of the
        # ELBO) instead of the prior sample.  ``injected_eps_p``: the posterior's noise (the reference's second torch.randn).
        self.posterior_decode = False
        self.injected_eps_p = None
        self._coor_clips = 0          # > 0 after reset_pos_coor_per_clip: number of clips the coordinate buffers describe

    # -- reference API ------------------------------------------------------------------------------
    def reset_pos_coor(self, to_list, tp_list):
        """Re-target the module at new context/target timestamps (Predictor.py:352-359)."""
        try:
            device = self.observed_coor.device
        except AttributeError:
            device = self.all_coor.device
        self.predict_coor = self.coor_generator(tp_list, self.h_list, self.w_list).to(device)
        self.observed_coor = self.coor_generator(to_list, self.h_list, self.w_list).to(device)
        self.TP = tp_list.shape[0]
        self._coor_clips = 0

    def reset_pos_coor_per_clip(self, to_lists, tp_lists):
        """Extension of ``reset_pos_coor`` (the reference shares one timestamp set per module, Predictor.py:352-359, so a
        "mixed" batch is a sequence of calls there): ``to_lists`` (N, To) and ``tp_lists`` (N, Tp) give every clip of the next
        batches its own context / target timestamps - prediction, interpolation and arbitrary continuous-time queries in one
        batch.  Clip i computes exactly what ``reset_pos_coor(to_lists[i], tp_lists[i])`` followed by a forward of clip i alone
        computes (per-clip arithmetic is batch invariant).  ``reset_pos_coor`` switches back to shared timestamps."""
        to_lists, tp_lists = torch.as_tensor(to_lists, dtype=torch.float32), torch.as_tensor(tp_lists, dtype=torch.float32)
        assert to_lists.dim() == 2 and tp_lists.dim() == 2 and to_lists.shape[0] == tp_lists.shape[0], \
            "expected (N, To) and (N, Tp) timestamp tables"
        try:
            device = self.observed_coor.device
        except AttributeError:
            device = self.all_coor.device
        gen = lambda tl: torch.cat([self.coor_generator(t, self.h_list, self.w_list) for t in tl], 0).to(device)
        self.predict_coor = gen(tp_lists)
        self.observed_coor = gen(to_lists)
        self.TP = tp_lists.shape[1]
        self._coor_clips = int(to_lists.shape[0])

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        # observed_coor / predict_coor follow the *current* (to, tp): a checkpoint saved after reset_pos_coor
        # carries differently shaped buffers.  Keep ours (they are derived data), accept any shape.
        for name in ("observed_coor", "predict_coor"):
            key = prefix + name
            if key in state_dict and getattr(self, name, None) is not None \
                    and state_dict[key].shape != getattr(self, name).shape:
                state_dict = dict(state_dict)
                state_dict[key] = getattr(self, name).detach().clone()
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def _build_engine(self):
        from .engine_predictor import PredictorEngine
        return PredictorEngine(self)

    def _coords_ready(self):
        if self.observed_coor is None or self.predict_coor is None:
            raise RuntimeError("Predictor built with rand_context=True has no coordinates yet: call "
                               "reset_pos_coor(to_list, tp_list) before forward()")
        # with rand_context=True the coordinates are plain attributes (reference quirk vii): module.to(device) does not
        # move them.  Move them once here so no host->device copy happens on the hot path (or inside a CUDA-graph capture).
        dev = self.nrmlp.B.device
        if self.observed_coor.device != dev or self.observed_coor.dtype != torch.float32:
            self.observed_coor = self.observed_coor.to(dev, torch.float32)
        if self.predict_coor.device != dev or self.predict_coor.dtype != torch.float32:
            self.predict_coor = self.predict_coor.to(dev, torch.float32)

    def forward(self, observed_features, predict_features_gt=None):
        """observed_features: (N, To, C, H, W) fp32 CUDA -> predicted features (N, Tp, C, H, W).
        NPVP-S with ``predict_features_gt`` (N, Tp, C, H, W): also runs the posterior on the ground-truth future and returns
        ``(out, mu_o, logvar_o, mu_p, logvar_p)`` like the reference in eval mode (Predictor.py:311-313, 320-327) - the
        inputs of the KL term (criterion.py:341-354).  NPVP-D ignores it (:328-335).  ``training=True`` is rejected by ``_guard``
        (no backward pass exists here); the forward of the training-mode branch - decoder queried with the posterior sample,
        :315-318 - is available through the explicit flag ``posterior_decode``."""
        self._guard(observed_features)
        self._coords_ready()
        with self._on_device(observed_features):
            if self.posterior_decode and self.stochastic:
                assert predict_features_gt is not None, "please input groundtruth predict features for storchastic model training/val"   # :316
            if predict_features_gt is not None and self.stochastic:
                self._guard(predict_features_gt)
                return self._engine().run(observed_features, predict_gt=predict_features_gt, decode_posterior=bool(self.posterior_decode))
            return self._engine().run(observed_features)

    def forward_tokens(self, observed_tokens, out16=None, n_samples=1):
        """Engine-internal: channels-last (N,To,H,W,C) in -> channels-last (N*n_samples,Tp,H,W,C) out (fp32, or the 16-bit
        workspace view of dtype ``out16`` that the frame decoder consumes directly)."""
        self._guard(observed_tokens)
        self._coords_ready()
        with self._on_device(observed_tokens):
            return self._engine().run(observed_tokens, channels_last=True, out16=out16, n_samples=n_samples)

    def forward_samples(self, observed_features, n_samples: int):
        """NPVP-S: ``n_samples`` stochastic futures per clip, (N, To, C, H, W) -> (N, n_samples, Tp, C, H, W).  Equivalent to
        ``n_samples`` reference forwards on the same clips with different noise (``injected_eps``: (N*n_samples,512,8,8),
        clip-major), but the EVT_Former and the prior run once."""
        self._guard(observed_features)
        self._coords_ready()
        with self._on_device(observed_features):
            out = self._engine().run(observed_features, n_samples=n_samples)
        return out.view(observed_features.shape[0], n_samples, *out.shape[1:])

    def prefetch_positional(self):
        """Make sure the NRMLP positional codes of the current coordinates exist (they are cached per coordinate tensor and
        recomputed only when ``reset_pos_coor`` / a batch-process function installs new coordinates)."""
        self._coords_ready()
        with self._on_device(self.nrmlp.B):
            self._engine()._positional_pair(self.observed_coor, self.predict_coor)

    def evt_coding_forward(self, x, pos_beta, pos_gamma):
        """EVT_Former + temporal mean (Predictor.py:337-350).  x (N,T,C,H,W); pos_beta/gamma (T*H*W, C)."""
        self._guard(x)
        with self._on_device(x):
            return self._engine().evt_coding(x, pos_beta, pos_gamma)
