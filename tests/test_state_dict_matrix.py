"""CPU: the modules' state_dict layout (key order, shapes, dtypes) equals the reference's over the constructor matrix.
The hashes in tests/golden/state_dict_matrix.json were taken from the unmodified reference modules by
tests/golden/make_state_dict_matrix.py, which also checked bit-identical weights under the same seed and strict loading."""
import json
import os
import sys

import pytest
import torch
import torch.nn as nn

import npvp_b200

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_state_dict_matrix import AE_VARIANTS, layout_hash, predictor_args  # noqa: E402

MATRIX = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "state_dict_matrix.json")))


@pytest.mark.parametrize("name", sorted(MATRIX))
def test_state_dict_layout_matches_reference(name):
    kind, *rest = name.split("/")
    if kind == "predictor":
        opts = dict(r.split("=") for r in rest)
        args, kw = predictor_args(opts["stochastic"] == "True", opts["fuse"], opts["rand_context"] == "True")
        mod = npvp_b200.Predictor(*args, **kw)
    else:
        cimg, ngf, nd, nr, outl = AE_VARIANTS[rest[0]]
        if kind == "encoder":
            mod = npvp_b200.ResnetEncoder(cimg, ngf=ngf, n_downsampling=nd, num_res_blocks=nr, norm_layer=nn.BatchNorm2d,
                                          norm_layer1d=nn.BatchNorm1d, learn_3d=False)
        else:
            mod = npvp_b200.ResnetDecoder(cimg, ngf=ngf, n_downsampling=nd, out_layer=outl, norm_layer=nn.BatchNorm2d)
    sd = mod.state_dict()
    assert len(sd) == MATRIX[name]["keys"]
    assert layout_hash(sd) == MATRIX[name]["sha1"]
