"""Freeze the reference's state_dict layout over the constructor matrix: tests/golden/state_dict_matrix.json.

Run in the build container only (needs /root/reference):  python tests/golden/make_state_dict_matrix.py
For every variant (stochastic x fuse_method x rand_context for the predictor, both autoencoder families) it builds the
UNMODIFIED reference module and ours under the same seed, checks key order, shapes, bit-identical values and strict loading,
and stores sha1(key:shape list) of the REFERENCE so that tests/test_state_dict_matrix.py can re-check ours without it."""
from __future__ import annotations

import hashlib
import json
import os
import sys
import warnings

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def layout_hash(sd) -> str:
    return hashlib.sha1("\n".join(f"{k}:{tuple(v.shape)}:{v.dtype}" for k, v in sd.items()).encode()).hexdigest()


def predictor_args(stoch, fuse, rc):
    hl = torch.linspace(0, 7, 8)
    return ((8, 8, 9, hl, hl, torch.arange(0., 4.), torch.arange(4., 9.), 512, fuse, 'layer', 256, 1, stoch, 8),
            dict(evt_former=True, learn_evt_token=False, evt_former_num_layers=4, rand_context=rc))


AE_VARIANTS = {"famA": (1, 64, 3, 2, "Sigmoid"), "famB": (3, 32, 4, 3, "Tanh")}


def main():
    warnings.filterwarnings("ignore")
    import make_golden as G
    RefP, RefE, RefD, _ = G.import_reference()
    import npvp_b200
    from util_init import reset_shared_norm
    out = {}

    def check(name, ref, mine):
        a, b = ref.state_dict(), mine.state_dict()
        assert list(a.keys()) == list(b.keys()), name
        assert all(a[k].shape == b[k].shape and torch.equal(a[k], b[k]) for k in a), name
        mine.load_state_dict(a, strict=True)
        out[name] = {"keys": len(a), "sha1": layout_hash(a)}

    for stoch in (True, False):
        for fuse in ("Add", "SPADE"):
            for rc in (False, True):
                args, kw = predictor_args(stoch, fuse, rc)
                reset_shared_norm(RefP)
                reset_shared_norm(npvp_b200.Predictor)
                torch.manual_seed(3)
                ref = RefP(*args, **kw).eval()
                torch.manual_seed(3)
                mine = npvp_b200.Predictor(*args, **kw).eval()
                check(f"predictor/stochastic={stoch}/fuse={fuse}/rand_context={rc}", ref, mine)
    for fam, (cimg, ngf, nd, nr, outl) in AE_VARIANTS.items():
        torch.manual_seed(4)
        re_ = RefE(cimg, ngf=ngf, n_downsampling=nd, num_res_blocks=nr, norm_layer=nn.BatchNorm2d, norm_layer1d=nn.BatchNorm1d, learn_3d=False)
        rd = RefD(cimg, ngf=ngf, n_downsampling=nd, out_layer=outl, norm_layer=nn.BatchNorm2d)
        torch.manual_seed(4)
        me = npvp_b200.ResnetEncoder(cimg, ngf=ngf, n_downsampling=nd, num_res_blocks=nr, norm_layer=nn.BatchNorm2d, norm_layer1d=nn.BatchNorm1d, learn_3d=False)
        md = npvp_b200.ResnetDecoder(cimg, ngf=ngf, n_downsampling=nd, out_layer=outl, norm_layer=nn.BatchNorm2d)
        check(f"encoder/{fam}", re_, me)
        check(f"decoder/{fam}", rd, md)
    with open(os.path.join(HERE, "state_dict_matrix.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"{len(out)} variants: key order, shapes, values under the same seed and strict loading all equal to the reference")


if __name__ == "__main__":
    main()
