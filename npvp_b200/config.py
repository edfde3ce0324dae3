"""Reference config handling: the 21 hydra YAMLs of the reference (configs/*.yaml) are read with
``yaml.safe_load`` (hydra/omegaconf are not needed), exposed with attribute access like ``cfg.Predictor.max_T``.

Also holds the per-dataset pixel (re)normalisation constants of utils/dataset.py:33-58 (needed to express the
1e-2 pixel tolerance in [0,1] image space) and built-in presets equal to the values in the shipped YAMLs for
the BASELINE.json configurations, so benches and tests do not depend on the reference tree being present.
"""
from __future__ import annotations

import copy
from typing import Any, Dict

import yaml


class AttrDict(dict):
    """dict with attribute access (stand-in for the OmegaConf node the reference code indexes)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


_FLOAT_KEYS = {"AE_lr", "predictor_lr", "KL_beta", "scheduler_eta_min", "lam_PF_L1", "lam_gan", "max_grad_norm"}


def _wrap(node: Any, key: str = "") -> Any:
    if isinstance(node, dict):
        return AttrDict({k: _wrap(v, k) for k, v in node.items()})
    if isinstance(node, str) and key in _FLOAT_KEYS:      # PyYAML (YAML 1.1) reads '1e-4' as str; OmegaConf gave float
        try:
            return float(node)
        except ValueError:
            return node
    return node


def load_config(path: str) -> AttrDict:
    with open(path, "r") as f:
        return _wrap(yaml.safe_load(f))


# (mean, std) used by VidReNormalize: pixel = x * std + mean   (utils/dataset.py:33-58, 860-886)
RENORM = {
    "KTH": ((0.6013795,), (2.7570653,)),
    "KITTI": ((0.44811612, 0.47147346, 0.46771598), (1.5177081, 1.5897311, 1.5952978)),
    "SMMNIST": ((0.0,), (1.0,)),
    "BAIR": ((0.61749697, 0.6050092, 0.52180636), (2.1824553, 2.1553133, 1.9115673)),
    "CityScapes": ((0.31604213, 0.35114038, 0.3104223), (1.2172801, 1.3219808, 1.2082524)),
    "Cityscapes": ((0.31604213, 0.35114038, 0.3104223), (1.2172801, 1.3219808, 1.2082524)),
}


# VidNormalize constants (utils/dataset.py:34-58): identical to RENORM except for KITTI, where the reference normalises and
# re-normalises with slightly different statistics.
NORM = dict(RENORM)
NORM["KITTI"] = ((0.44812047, 0.47147775, 0.4677183), (1.5147436, 1.5871466, 1.5925455))


def _preset(name, ch, hw, past, future, ngf, nd, nr, out_layer, max_T, stochastic, rand_context=False, batch=8,
            test_future=None, vfi=False):
    return _wrap({
        "Dataset": {"name": name, "img_channels": ch, "img_size": hw, "num_past_frames": past, "num_future_frames": future,
                    "test_num_past_frames": past, "test_num_future_frames": test_future or future, "batch_size": batch},
        "AE": {"ngf": ngf, "n_downsampling": nd, "num_res_blocks": nr, "out_layer": out_layer, "learn_3d": False},
        "Predictor": {"rand_context": rand_context, "VFI": vfi, "max_H": 8, "max_W": 8, "max_T": max_T, "embed_dim": 512,
                      "fuse_method": "Add", "param_free_norm_type": "layer", "evt_former": True, "evt_former_num_layers": 4,
                      "evt_hidden_channels": 256, "stochastic": stochastic, "transformer_layers": 8},
    })


# Values transcribed from the reference YAMLs named in BASELINE.json (configs/config_<...>.yaml).
PRESETS: Dict[str, AttrDict] = {
    "SMMNIST_VFP_NPVP-D": _preset("SMMNIST", 1, 64, 5, 10, 64, 3, 2, "Sigmoid", 15, False),
    # BASELINE.json's text for config 1 says "10 context -> 10 future"; the YAML above is 5 -> 10.  Labelled variant:
    "SMMNIST_VFP_NPVP-D_10to10": _preset("SMMNIST", 1, 64, 10, 10, 64, 3, 2, "Sigmoid", 20, False),
    "KTH_Unified_NPVP-S": _preset("KTH", 1, 64, 10, 10, 64, 3, 2, "Tanh", 20, True, rand_context=True, test_future=20),
    "BAIR_VFP_NPVP-S": _preset("BAIR", 3, 64, 2, 10, 64, 3, 2, "Tanh", 12, True, test_future=28),
    "Cityscapes_VFP_NPVP-D": _preset("CityScapes", 3, 128, 2, 10, 32, 4, 3, "Tanh", 12, False, test_future=28),
    "Cityscapes_VFP_NPVP-S": _preset("CityScapes", 3, 128, 2, 10, 32, 4, 3, "Tanh", 12, True, test_future=28),
    "KITTI_VFP_NPVP-S": _preset("KITTI", 3, 128, 4, 5, 32, 4, 3, "Tanh", 9, True, batch=16),
    # NOT shipped YAMLs: the one-shot 2 -> 28 variants of BASELINE configs 3 / 4 (the YAMLs' max_T = 12 forbids
    # 28 target timestamps in one call; max_T = 30 is the smallest change that allows it - Predictor.py:41, submodules.py:351)
    "BAIR_VFP_NPVP-S_oneshot28": _preset("BAIR", 3, 64, 2, 28, 64, 3, 2, "Tanh", 30, True),
    "Cityscapes_VFP_NPVP-D_oneshot28": _preset("CityScapes", 3, 128, 2, 28, 32, 4, 3, "Tanh", 30, False),
    "Cityscapes_VFP_NPVP-S_oneshot28": _preset("CityScapes", 3, 128, 2, 28, 32, 4, 3, "Tanh", 30, True),
}


def preset(name: str) -> AttrDict:
    return copy.deepcopy(PRESETS[name])
