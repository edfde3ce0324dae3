"""CPU: oracle/post_oracle.py against the fixtures generated from the unmodified reference (tests/golden/post_*.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import post_oracle as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["post_cityscapes", "post_kth", "post_smmnist"]


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.mark.parametrize("name", CASES)
def test_post_oracle_matches_reference_fixture(name):
    g = load(name)
    mean, std = g["mean"].tolist(), g["std"].tolist()
    pix = P.renormalize_clamp(g["frames"], mean, std)
    assert torch.equal(pix, g["pix"])                                   # bit-exact: same fp32 operation order
    assert torch.equal(P.to_uint8(pix), g["u8"])
    assert float((P.normalize_u8(g["u8"], mean, std) - g["back"]).abs().max()) <= 1e-6
    n = pix.shape[0] * pix.shape[1]
    fp, fg = pix.reshape(n, *pix.shape[2:]), g["gt"].reshape(n, *pix.shape[2:])
    assert float((P.psnr(fp, fg) - g["psnr"]).abs().max()) <= 1e-5
    assert float((P.ssim(fp, fg) - g["ssim"]).abs().max()) <= 1e-6
