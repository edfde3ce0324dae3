"""Deterministic helpers shared by the golden generator and the parity tests.

``stress_init_`` perturbs a module's parameters *through its state_dict* so the same
procedure can be applied to the reference modules (when the goldens are generated) and
to the npvp_b200 modules (when tests run) and yield bit-identical weights.  It removes
the blind spots of default init listed in SURVEY.md section 4.4: zero non-local gamma,
identity BatchNorm/LayerNorm, identical cloned layers, tiny positional beta.
"""
from __future__ import annotations

import re
from typing import Dict

import torch

_LAYER_RE = re.compile(r"^(EVT_Former|transformer)\.layers\.(\d+)\.")


def fingerprint(sd: Dict[str, torch.Tensor]) -> Dict[str, float]:
    """Order-independent fingerprint of a state_dict (float64 sums)."""
    tot, tot_abs, n = 0.0, 0.0, 0
    for k in sorted(sd):
        v = sd[k]
        if not v.dtype.is_floating_point:
            continue
        d = v.detach().double()
        tot += float(d.sum())
        tot_abs += float(d.abs().sum())
        n += d.numel()
    return {"sum": tot, "abs_sum": tot_abs, "numel": float(n), "keys": float(len(sd))}


@torch.no_grad()
def stress_init_(module: torch.nn.Module, seed: int = 7) -> None:
    g = torch.Generator().manual_seed(seed)
    sd = module.state_dict()
    seen = set()

    def randn(t, std=1.0):
        return torch.randn(t.shape, generator=g, dtype=torch.float32) * std

    def rand(t, lo, hi):
        return torch.rand(t.shape, generator=g, dtype=torch.float32) * (hi - lo) + lo

    bn_prefixes = {k[: -len("running_mean")] for k in sd if k.endswith("running_mean")}
    for k in sorted(sd):
        v = sd[k]
        if not v.dtype.is_floating_point:
            continue
        if v.data_ptr() in seen:          # EVT_Former.norm.* aliases transformer.norm.*
            continue
        seen.add(v.data_ptr())
        base = k[: k.rfind(".") + 1]
        leaf = k[k.rfind(".") + 1:]
        if base in bn_prefixes:
            if leaf == "running_mean":
                v.copy_(randn(v, 0.3))
            elif leaf == "running_var":
                v.copy_(rand(v, 0.5, 2.0))
            elif leaf == "weight":
                v.copy_(rand(v, 0.5, 1.5))
            elif leaf == "bias":
                v.copy_(randn(v, 0.2))
            continue
        if leaf == "gamma" and v.dim() == 0:
            v.fill_(0.6)
            continue
        m = _LAYER_RE.match(k)
        if m and int(m.group(2)) > 0 and v.dim() >= 2 and "norm" not in k:
            # layers start as deep copies of layer 0: re-draw with the same scale
            v.copy_(randn(v, float(v.std())))
        if v.dim() == 1 or "norm" in k:
            v.add_(randn(v, 0.3))
    if "nrmlp.mlp_beta.weight" in sd:
        sd["nrmlp.mlp_beta.weight"].mul_(10.0)
    heads = [k for k in sd if re.match(r"^model\.\d+\.bias$", k) and sd[k].dim() == 1
             and k.replace("bias", "weight") in sd and sd[k.replace("bias", "weight")].dim() == 4
             and sd[k.replace("bias", "weight")].shape[-1] == 7]
    for k in heads:
        sd[k.replace("bias", "weight")].mul_(6.0)


def seeded_randn(shape, seed: int) -> torch.Tensor:
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


def seeded_rand(shape, seed: int) -> torch.Tensor:
    return torch.rand(shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32)


@torch.no_grad()
def reset_shared_norm(predictor_cls) -> None:
    """``Predictor.__init__`` has ``norm=nn.LayerNorm(512)`` as a *shared default instance* (reference quirk,
    models/Predictor.py:270): every Predictor built in the process aliases it.  Reset it to identity so a test
    case does not depend on which cases ran before."""
    import inspect
    norm = inspect.signature(predictor_cls.__init__).parameters["norm"].default
    norm.to("cpu")        # a previous .cuda() of any Predictor moved the shared instance too
    norm.weight.fill_(1.0)
    norm.bias.zero_()
