"""Generate the golden fixtures in tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

For every case it
  1. imports the reference modules through a sys.modules shim (pytorch_lightning / timm are absent),
  2. builds them under a fixed seed, optionally applies ``tests/util_init.stress_init_``,
  3. runs the reference forward on seeded inputs (injected latent noise for NPVP-S),
  4. checks the npvp_b200 constructors draw *identical* weights under the same seed,
  5. checks the oracle restatement against the reference output,
  6. stores a strided sample of the output + statistics + the weight fingerprint.
The fixtures pin ``oracle/npvp_oracle.py``; tests never need /root/reference.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("NPVP_REFERENCE", "/root/reference")


def import_reference():
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        def log(self, *a, **k):
            pass

    class LightningDataModule:
        def __init__(self):
            pass

    pl.LightningModule, pl.LightningDataModule = LightningModule, LightningDataModule
    cb = types.ModuleType("pytorch_lightning.callbacks")
    cb.Callback = object
    cb.ModelCheckpoint = object
    ut = types.ModuleType("pytorch_lightning.utilities")
    ut.rank_zero_only = lambda f: f
    pl.callbacks, pl.utilities = cb, ut
    sys.modules.update({"pytorch_lightning": pl, "pytorch_lightning.callbacks": cb, "pytorch_lightning.utilities": ut})
    tl = types.ModuleType("timm.models.layers")
    tl.to_2tuple = lambda x: x if isinstance(x, (tuple, list)) else (x, x)
    sys.modules.update({"timm": types.ModuleType("timm"), "timm.models": types.ModuleType("timm.models"),
                        "timm.models.layers": tl})
    # our drop-in package is also called ``models``: import the reference first, then rename it
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    import models as ref_models  # noqa
    from models.Predictor import Predictor as RefPredictor
    from models.ResNetAutoEncoder import ResnetEncoder as RefEnc, ResnetDecoder as RefDec
    import models.submodules as ref_sub
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        sys.modules["ref_" + k] = sys.modules.pop(k)
    sys.path.remove(REF)
    return RefPredictor, RefEnc, RefDec, ref_sub


def sample(t: torch.Tensor, max_n: int = 30000):
    flat = t.detach().reshape(-1)
    stride = max(1, flat.numel() // max_n)
    return flat[::stride].numpy().astype(np.float32), stride


def main():
    import warnings
    warnings.filterwarnings("ignore")
    from util_init import fingerprint, stress_init_, seeded_randn, seeded_rand, reset_shared_norm
    RefPredictor, RefEnc, RefDec, ref_sub = import_reference()
    import npvp_b200
    from oracle import npvp_oracle as O

    torch.set_grad_enabled(False)
    report = []

    def check_same_weights(ref_mod, my_mod, name):
        a, b = ref_mod.state_dict(), my_mod.state_dict()
        assert list(a.keys()) == list(b.keys()), f"{name}: state_dict keys differ"
        for k in a:
            assert a[k].shape == b[k].shape, (name, k, a[k].shape, b[k].shape)
            assert torch.equal(a[k], b[k]), f"{name}: weight {k} differs under identical seed"
        # and reference checkpoints load strictly
        my_mod.load_state_dict(a, strict=True)

    def save(name, out, meta, extra=None):
        vals, stride = sample(out)
        d = dict(sample=vals, stride=np.int64(stride), shape=np.array(out.shape, dtype=np.int64),
                 mean=np.float64(out.double().mean()), std=np.float64(out.double().std()),
                 absmax=np.float64(out.abs().max()))
        for k, v in meta.items():
            d["meta_" + k] = np.array(v)
        if extra:
            d.update(extra)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)

    # ---------------------------------------------------------------- predictor cases
    pred_cases = [
        # name, stochastic, max_T, to, tp, N, stress, seed
        ("pred_S_stress_realT", True, 9, [0, 1, 4], [2, 2.5, 3, 5, 6.75], 1, True, 11),
        ("pred_D_default", False, 12, [0, 1], [2, 3, 4], 2, False, 12),
        ("pred_D_stress_vfi", False, 10, [0, 1, 8, 9], [3, 5.5], 2, True, 13),
    ]
    hl = torch.linspace(0, 7, 8)
    only_new = "--gt-only" in sys.argv or "--spade-only" in sys.argv or "--zp-only" in sys.argv
    for name, stoch, max_T, to, tp, N, stress, seed in ([] if only_new else pred_cases):
        to_t, tp_t = torch.tensor(to, dtype=torch.float32), torch.tensor(tp, dtype=torch.float32)
        args = (8, 8, max_T, hl, hl, to_t, tp_t, 512, 'Add', 'layer', 256, 1, stoch, 8)
        kw = dict(evt_former=True, learn_evt_token=False, evt_former_num_layers=4, rand_context=False)
        reset_shared_norm(RefPredictor)
        reset_shared_norm(npvp_b200.Predictor)
        torch.manual_seed(seed)
        ref = RefPredictor(*args, **kw).eval()
        torch.manual_seed(seed)
        mine = npvp_b200.Predictor(*args, **kw).eval()
        check_same_weights(ref, mine, name)
        if stress:
            stress_init_(ref, seed)
            stress_init_(mine, seed)
            check_same_weights(ref, mine, name + "+stress")
        x = torch.relu(seeded_randn((N, len(to), 512, 8, 8), seed + 100))
        eps = seeded_randn((N, 512, 8, 8), seed + 200)
        real = ref_sub.torch.randn
        try:
            ref_sub.torch.randn = lambda *a, **k: eps.clone()
            out_ref = ref(x)
        finally:
            ref_sub.torch.randn = real
        sd = {k: v.clone() for k, v in mine.state_dict().items()}
        out_or = O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], stoch, eps if stoch else None)
        err = float((out_ref - out_or).abs().max())
        report.append((name, tuple(out_ref.shape), err))
        assert err < 5e-5, (name, err)
        fp = fingerprint(ref.state_dict())
        save(name, out_ref, dict(seed=seed, stochastic=stoch, max_T=max_T, to=to, tp=tp, N=N, stress=stress,
                                 fp_sum=fp["sum"], fp_abs=fp["abs_sum"], fp_numel=fp["numel"], fp_keys=fp["keys"]))

    # ---------------------------------------------------------------- NPVP-S with ground-truth future features (posterior branch)
    # Predictor.forward(observed, predict_features_gt) in eval mode -> (out, mu_o, logvar_o, mu_p, logvar_p), Predictor.py:311-327
    for name, max_T, to, tp, N, seed in ([] if ("--spade-only" in sys.argv or "--zp-only" in sys.argv) else [("pred_S_stress_gt", 9, [0, 1.5, 3], [2, 4, 5.25, 8], 2, 14)]):
        to_t, tp_t = torch.tensor(to, dtype=torch.float32), torch.tensor(tp, dtype=torch.float32)
        args = (8, 8, max_T, hl, hl, to_t, tp_t, 512, 'Add', 'layer', 256, 1, True, 8)
        kw = dict(evt_former=True, learn_evt_token=False, evt_former_num_layers=4, rand_context=False)
        reset_shared_norm(RefPredictor)
        reset_shared_norm(npvp_b200.Predictor)
        torch.manual_seed(seed)
        ref = RefPredictor(*args, **kw).eval()
        torch.manual_seed(seed)
        mine = npvp_b200.Predictor(*args, **kw).eval()
        stress_init_(ref, seed)
        stress_init_(mine, seed)
        check_same_weights(ref, mine, name)
        x = torch.relu(seeded_randn((N, len(to), 512, 8, 8), seed + 100))
        gt = torch.relu(seeded_randn((N, len(tp), 512, 8, 8), seed + 300))
        eps = seeded_randn((N, 512, 8, 8), seed + 200)
        real = ref_sub.torch.randn
        try:
            ref_sub.torch.randn = lambda *a, **k: eps.clone()
            outs_ref = ref(x, gt)
        finally:
            ref_sub.torch.randn = real
        sd = {k: v.clone() for k, v in mine.state_dict().items()}
        outs_or = O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], True, eps, predict_features_gt=gt)
        assert len(outs_ref) == 5 and len(outs_or) == 5
        errs = [float((a - b).abs().max()) for a, b in zip(outs_ref, outs_or)]
        report.append((name, tuple(outs_ref[0].shape), max(errs)))
        assert max(errs) < 5e-5, (name, errs)
        fp = fingerprint(ref.state_dict())
        extra = {}
        for key, t in zip(("mu_o", "logvar_o", "mu_p", "logvar_p"), outs_ref[1:]):
            vals, stride = sample(t, 8000)
            extra[key] = vals
            extra[key + "_stride"] = np.int64(stride)
        save(name, outs_ref[0], dict(seed=seed, stochastic=True, max_T=max_T, to=to, tp=tp, N=N, stress=True,
                                     fp_sum=fp["sum"], fp_abs=fp["abs_sum"], fp_numel=fp["numel"], fp_keys=fp["keys"]), extra)
    # ---------------------------------------------------------------- NPVP-S decoder driven by the POSTERIOR sample z_p (Predictor.py:315-318):
    # the branch the reference takes when self.training is set.  Only the top-level flag is raised (ref.training = True, no
    # recursion), so dropout / drop-path stay in eval mode and the forward is deterministic; two different noise tensors are
    # injected for the prior's and the posterior's torch.randn draws, in the reference's order.
    for name, max_T, to, tp, N, seed in ([] if ("--spade-only" in sys.argv or "--gt-only" in sys.argv) else [("pred_S_stress_zp", 9, [0, 2, 3.5], [1, 4.5, 6, 8], 2, 16)]):
        to_t, tp_t = torch.tensor(to, dtype=torch.float32), torch.tensor(tp, dtype=torch.float32)
        args = (8, 8, max_T, hl, hl, to_t, tp_t, 512, 'Add', 'layer', 256, 1, True, 8)
        kw = dict(evt_former=True, learn_evt_token=False, evt_former_num_layers=4, rand_context=False)
        reset_shared_norm(RefPredictor)
        reset_shared_norm(npvp_b200.Predictor)
        torch.manual_seed(seed)
        ref = RefPredictor(*args, **kw).eval()
        torch.manual_seed(seed)
        mine = npvp_b200.Predictor(*args, **kw).eval()
        stress_init_(ref, seed)
        stress_init_(mine, seed)
        check_same_weights(ref, mine, name)
        x = torch.relu(seeded_randn((N, len(to), 512, 8, 8), seed + 100))
        gt = torch.relu(seeded_randn((N, len(tp), 512, 8, 8), seed + 300))
        eps = seeded_randn((N, 512, 8, 8), seed + 200)
        eps_p = seeded_randn((N, 512, 8, 8), seed + 400)
        draws = [eps, eps_p]
        real = ref_sub.torch.randn
        try:
            ref_sub.torch.randn = lambda *a, **k: draws.pop(0).clone()
            ref.training = True                       # top-level flag only: children stay in eval mode
            outs_ref = ref(x, gt)
        finally:
            ref_sub.torch.randn = real
            ref.training = False
        assert not draws
        sd = {k: v.clone() for k, v in mine.state_dict().items()}
        outs_or = O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], True, eps, predict_features_gt=gt,
                                      decode_with_posterior=True, eps_p=eps_p)
        errs = [float((a - b).abs().max()) for a, b in zip(outs_ref, outs_or)]
        prior_out = O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], True, eps)
        assert float((prior_out - outs_ref[0]).abs().max()) > 1e-2, "the posterior-driven output must differ from the prior-driven one"
        report.append((name, tuple(outs_ref[0].shape), max(errs)))
        assert max(errs) < 5e-5, (name, errs)
        fp = fingerprint(ref.state_dict())
        extra = {}
        for key, t in zip(("mu_o", "logvar_o", "mu_p", "logvar_p"), outs_ref[1:]):
            vals, stride = sample(t, 8000)
            extra[key] = vals
            extra[key + "_stride"] = np.int64(stride)
        save(name, outs_ref[0], dict(seed=seed, stochastic=True, max_T=max_T, to=to, tp=tp, N=N, stress=True,
                                     fp_sum=fp["sum"], fp_abs=fp["abs_sum"], fp_numel=fp["numel"], fp_keys=fp["keys"]), extra)
    # ---------------------------------------------------------------- fuse_method='SPADE' (the constructor default, Predictor.py:268;
    # every shipped YAML uses 'Add'): the NRMLP also emits gamma and the fuser scales by (1 + gamma) (submodules.py:296-297, 441-447)
    for name, max_T, to, tp, N, seed in ([] if "--zp-only" in sys.argv else [("pred_S_stress_spade", 10, [0, 2, 3.5], [1, 4, 6, 9.5], 2, 15)]):
        to_t, tp_t = torch.tensor(to, dtype=torch.float32), torch.tensor(tp, dtype=torch.float32)
        args = (8, 8, max_T, hl, hl, to_t, tp_t, 512, 'SPADE', 'layer', 256, 1, True, 8)
        kw = dict(evt_former=True, learn_evt_token=False, evt_former_num_layers=4, rand_context=False)
        reset_shared_norm(RefPredictor)
        reset_shared_norm(npvp_b200.Predictor)
        torch.manual_seed(seed)
        ref = RefPredictor(*args, **kw).eval()
        torch.manual_seed(seed)
        mine = npvp_b200.Predictor(*args, **kw).eval()
        stress_init_(ref, seed)
        stress_init_(mine, seed)
        for m in (ref, mine):                      # make gamma matter: the default init leaves |gamma| ~ 0.05
            m.nrmlp.mlp_gamma.weight.data.mul_(6.0)
        check_same_weights(ref, mine, name)
        x = torch.relu(seeded_randn((N, len(to), 512, 8, 8), seed + 100))
        eps = seeded_randn((N, 512, 8, 8), seed + 200)
        real = ref_sub.torch.randn
        try:
            ref_sub.torch.randn = lambda *a, **k: eps.clone()
            out_ref = ref(x)
        finally:
            ref_sub.torch.randn = real
        sd = {k: v.clone() for k, v in mine.state_dict().items()}
        out_or = O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], True, eps, fuse_method="SPADE")
        gam = O.nrmlp(sd, "nrmlp.", sd["predict_coor"], "SPADE")[1]
        err = float((out_ref - out_or).abs().max())
        report.append((name, tuple(out_ref.shape), err))
        assert err < 5e-5 and float(gam.abs().max()) > 0.2, (name, err, float(gam.abs().max()))
        fp = fingerprint(ref.state_dict())
        save(name, out_ref, dict(seed=seed, stochastic=True, max_T=max_T, to=to, tp=tp, N=N, stress=True, fuse_method="SPADE",
                                 gamma_scale=6.0, fp_sum=fp["sum"], fp_abs=fp["abs_sum"], fp_numel=fp["numel"], fp_keys=fp["keys"]))
    if "--spade-only" in sys.argv or "--zp-only" in sys.argv:
        print(report)
        with open(os.path.join(HERE, "REPORT.txt"), "a") as f:
            for r in report[-1:]:
                f.write(f"{r[0]}, {r[1]}, {r[2]:.3e}\n")
        return
    if "--gt-only" in sys.argv:
        print(report)
        with open(os.path.join(HERE, "REPORT.txt"), "a") as f:
            for r in report[-1:]:
                f.write(f"{r[0]}, {r[1]}, {r[2]:.3e}\n")
        return

    # ---------------------------------------------------------------- autoencoder cases
    ae_cases = [
        # name, Cimg, ngf, n_down, n_res, out_layer, HW, N, T, stress, seed
        ("ae_famB_stress", 3, 32, 4, 3, "Tanh", 128, 1, 2, True, 21),
        ("ae_famA_default", 1, 64, 3, 2, "Sigmoid", 64, 2, 2, False, 22),
        ("ae_famA_rgb_stress", 3, 64, 3, 2, "Tanh", 64, 1, 3, True, 23),
    ]
    for name, cimg, ngf, nd, nr, outl, hw, N, T, stress, seed in ae_cases:
        torch.manual_seed(seed)
        renc = RefEnc(cimg, ngf=ngf, n_downsampling=nd, num_res_blocks=nr, norm_layer=nn.BatchNorm2d,
                      norm_layer1d=nn.BatchNorm1d, learn_3d=False).eval()
        rdec = RefDec(cimg, ngf=ngf, n_downsampling=nd, out_layer=outl, norm_layer=nn.BatchNorm2d).eval()
        torch.manual_seed(seed)
        menc = npvp_b200.ResnetEncoder(cimg, ngf=ngf, n_downsampling=nd, num_res_blocks=nr, norm_layer=nn.BatchNorm2d,
                                       norm_layer1d=nn.BatchNorm1d, learn_3d=False).eval()
        mdec = npvp_b200.ResnetDecoder(cimg, ngf=ngf, n_downsampling=nd, out_layer=outl, norm_layer=nn.BatchNorm2d).eval()
        check_same_weights(renc, menc, name + ".enc")
        check_same_weights(rdec, mdec, name + ".dec")
        if stress:
            for m in (renc, menc):
                stress_init_(m, seed)
            for m in (rdec, mdec):
                stress_init_(m, seed + 1)
            check_same_weights(renc, menc, name + ".enc+stress")
            check_same_weights(rdec, mdec, name + ".dec+stress")
        x = seeded_rand((N, T, cimg, hw, hw), seed + 100)
        if outl == "Tanh":
            x = x * 2 - 1
        feats_ref = renc(x)
        f_in = torch.relu(seeded_randn((N, T, ngf * 2 ** nd, hw // 2 ** nd, hw // 2 ** nd), seed + 300))
        frames_ref = rdec(f_in)
        esd = {k: v.clone() for k, v in menc.state_dict().items()}
        dsd = {k: v.clone() for k, v in mdec.state_dict().items()}
        e1 = float((feats_ref - O.resnet_encoder(esd, x, nd, nr)).abs().max())
        e2 = float((frames_ref - O.resnet_decoder(dsd, f_in, nd, outl)).abs().max())
        report.append((name + ".enc", tuple(feats_ref.shape), e1))
        report.append((name + ".dec", tuple(frames_ref.shape), e2))
        assert e1 < 5e-4 * max(1.0, float(feats_ref.abs().max())) and e2 < 5e-5, (name, e1, e2)
        fpe, fpd = fingerprint(renc.state_dict()), fingerprint(rdec.state_dict())
        meta = dict(seed=seed, cimg=cimg, ngf=ngf, n_down=nd, n_res=nr, out_layer=outl, hw=hw, N=N, T=T, stress=stress)
        save(name + "_enc", feats_ref, dict(meta, fp_sum=fpe["sum"], fp_abs=fpe["abs_sum"], fp_numel=fpe["numel"], fp_keys=fpe["keys"]))
        save(name + "_dec", frames_ref, dict(meta, fp_sum=fpd["sum"], fp_abs=fpd["abs_sum"], fp_numel=fpd["numel"], fp_keys=fpd["keys"]))

    print("case, shape, max|oracle - reference|")
    for r in report:
        print(r)
    with open(os.path.join(HERE, "REPORT.txt"), "w") as f:
        f.write("golden fixtures generated from /root/reference (unmodified, CPU fp32, torch %s)\n" % torch.__version__)
        f.write("case, output shape, max|oracle - reference|\n")
        for r in report:
            f.write(f"{r[0]}, {r[1]}, {r[2]:.3e}\n")


if __name__ == "__main__":
    main()
