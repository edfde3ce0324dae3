"""Builders shared by CPU and GPU tests: re-create the exact modules/inputs behind tests/golden/*.npz."""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn

import npvp_b200
from util_init import fingerprint, reset_shared_norm, seeded_rand, seeded_randn, stress_init_

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PRED_CASES = ["pred_S_stress_realT", "pred_D_default", "pred_D_stress_vfi"]
PRED_SPADE_CASES = ["pred_S_stress_spade"]    # fuse_method='SPADE' (the constructor default; every shipped YAML uses 'Add')
PRED_GT_CASES = ["pred_S_stress_gt"]          # NPVP-S with ground-truth future features (posterior branch, Predictor.py:311-327)
PRED_ZP_CASES = ["pred_S_stress_zp"]          # decoder driven by the posterior sample z_p (Predictor.py:315-318), two noise tensors
LATENT_KEYS = ("mu_o", "logvar_o", "mu_p", "logvar_p")
AE_CASES = ["ae_famB_stress", "ae_famA_default", "ae_famA_rgb_stress"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = {k[5:]: z[k] for k in z.files if k.startswith("meta_")}
    return z, meta


def golden_sample(t: torch.Tensor, z) -> np.ndarray:
    return t.detach().float().cpu().reshape(-1)[:: int(z["stride"])].numpy()


def check_fingerprint(module, meta):
    fp = fingerprint(module.state_dict())
    assert fp["keys"] == float(meta["fp_keys"]) and fp["numel"] == float(meta["fp_numel"])
    assert abs(fp["sum"] - float(meta["fp_sum"])) <= 1e-6 * max(1.0, abs(float(meta["fp_sum"])))
    assert abs(fp["abs_sum"] - float(meta["fp_abs"])) <= 1e-9 * float(meta["fp_abs"])


def build_predictor_case(name):
    z, m = load_golden(name)
    seed, stoch = int(m["seed"]), bool(m["stochastic"])
    to = torch.tensor(m["to"], dtype=torch.float32)
    tp = torch.tensor(m["tp"], dtype=torch.float32)
    hl = torch.linspace(0, 7, 8)
    reset_shared_norm(npvp_b200.Predictor)
    torch.manual_seed(seed)
    fuse = str(m["fuse_method"]) if "fuse_method" in m else 'Add'
    mod = npvp_b200.Predictor(8, 8, int(m["max_T"]), hl, hl, to, tp, 512, fuse, 'layer', 256, 1, stoch, 8,
                              evt_former=True, learn_evt_token=False, evt_former_num_layers=4, rand_context=False).eval()
    if bool(m["stress"]):
        stress_init_(mod, seed)
    if "gamma_scale" in m:                                  # SPADE fixtures: enlarge gamma so that the (1 + gamma) factor matters
        mod.nrmlp.mlp_gamma.weight.data.mul_(float(m["gamma_scale"]))
    check_fingerprint(mod, m)
    x = torch.relu(seeded_randn((int(m["N"]), len(to), 512, 8, 8), seed + 100))
    eps = seeded_randn((int(m["N"]), 512, 8, 8), seed + 200)
    return mod, x, eps, stoch, z


def build_predictor_gt_case(name):
    """As build_predictor_case plus the ground-truth future features the fixture was generated with."""
    mod, x, eps, stoch, z = build_predictor_case(name)
    _, m = load_golden(name)
    gt = torch.relu(seeded_randn((int(m["N"]), len(m["tp"]), 512, 8, 8), int(m["seed"]) + 300))
    return mod, x, gt, eps, z


def build_predictor_zp_case(name):
    """As build_predictor_gt_case plus the posterior's own noise tensor (the reference's second torch.randn draw)."""
    mod, x, gt, eps, z = build_predictor_gt_case(name)
    _, m = load_golden(name)
    eps_p = seeded_randn((int(m["N"]), 512, 8, 8), int(m["seed"]) + 400)
    return mod, x, gt, eps, eps_p, z


def golden_latents(outs, z):
    """[(name, strided sample of ours, golden sample)] for the four latent tensors of a posterior-branch fixture."""
    res = []
    for key, t in zip(LATENT_KEYS, outs[1:]):
        ours = t.detach().float().cpu().reshape(-1)[:: int(z[key + "_stride"])].numpy()
        res.append((key, ours, z[key]))
    return res


def build_ae_case(name):
    ze, m = load_golden(name + "_enc")
    zd, md = load_golden(name + "_dec")
    seed, cimg, ngf, nd, nr = int(m["seed"]), int(m["cimg"]), int(m["ngf"]), int(m["n_down"]), int(m["n_res"])
    outl, hw, N, T = str(m["out_layer"]), int(m["hw"]), int(m["N"]), int(m["T"])
    torch.manual_seed(seed)
    enc = npvp_b200.ResnetEncoder(cimg, ngf=ngf, n_downsampling=nd, num_res_blocks=nr, norm_layer=nn.BatchNorm2d,
                                  norm_layer1d=nn.BatchNorm1d, learn_3d=False).eval()
    dec = npvp_b200.ResnetDecoder(cimg, ngf=ngf, n_downsampling=nd, out_layer=outl, norm_layer=nn.BatchNorm2d).eval()
    if bool(m["stress"]):
        stress_init_(enc, seed)
        stress_init_(dec, seed + 1)
    check_fingerprint(enc, m)
    check_fingerprint(dec, md)
    x = seeded_rand((N, T, cimg, hw, hw), seed + 100)
    if outl == "Tanh":
        x = x * 2 - 1
    f_in = torch.relu(seeded_randn((N, T, ngf * 2 ** nd, hw // 2 ** nd, hw // 2 ** nd), seed + 300))
    return enc, dec, x, f_in, dict(n_down=nd, n_res=nr, out_layer=outl), ze, zd
