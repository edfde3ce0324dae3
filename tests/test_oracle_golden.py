"""CPU: the oracle restatement reproduces the reference outputs frozen in tests/golden (pins the oracle)."""
import numpy as np
import pytest
import torch

from cases import (AE_CASES, PRED_CASES, PRED_GT_CASES, PRED_SPADE_CASES, PRED_ZP_CASES, build_ae_case, build_predictor_case,
                   build_predictor_gt_case, build_predictor_zp_case, golden_latents, golden_sample)
from oracle import npvp_oracle as O

torch.set_grad_enabled(False)


@pytest.mark.parametrize("name", PRED_CASES)
def test_predictor_oracle_matches_reference(name):
    mod, x, eps, stoch, z = build_predictor_case(name)
    sd = mod.state_dict()
    out = O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], stoch, eps if stoch else None)
    assert list(out.shape) == list(z["shape"])
    np.testing.assert_allclose(golden_sample(out, z), z["sample"], atol=5e-5, rtol=0)
    assert abs(float(out.double().mean()) - float(z["mean"])) < 1e-5


@pytest.mark.parametrize("name", PRED_SPADE_CASES)
def test_predictor_spade_oracle_matches_reference(name):
    """fuse_method='SPADE': NRMLP emits gamma as well and the fuser multiplies by (1 + gamma) (submodules.py:296-297, 441-447)."""
    mod, x, eps, stoch, z = build_predictor_case(name)
    sd = mod.state_dict()
    assert "nrmlp.mlp_gamma.weight" in sd
    out = O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], stoch, eps, fuse_method="SPADE")
    np.testing.assert_allclose(golden_sample(out, z), z["sample"], atol=5e-5, rtol=0)
    plain = O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], stoch, eps, fuse_method="Add")
    assert float((plain - out).abs().max()) > 1e-2           # gamma really changes the result in this fixture


@pytest.mark.parametrize("name", PRED_GT_CASES)
def test_predictor_posterior_oracle_matches_reference(name):
    """Predictor.forward(observed, predict_features_gt) in eval mode: (out, mu_o, logvar_o, mu_p, logvar_p)."""
    mod, x, gt, eps, z = build_predictor_gt_case(name)
    sd = mod.state_dict()
    outs = O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], True, eps, predict_features_gt=gt)
    assert len(outs) == 5 and list(outs[0].shape) == list(z["shape"])
    np.testing.assert_allclose(golden_sample(outs[0], z), z["sample"], atol=5e-5, rtol=0)
    for key, ours, gold in golden_latents(outs, z):
        np.testing.assert_allclose(ours, gold, atol=5e-5, rtol=0, err_msg=key)
    # the decoder is queried with the PRIOR sample: the prediction does not depend on the ground truth
    plain = O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], True, eps)
    assert torch.equal(plain, outs[0])


@pytest.mark.parametrize("name", PRED_ZP_CASES)
def test_predictor_posterior_decode_oracle_matches_reference(name):
    """decode_with_posterior: the reference's training-mode branch (Predictor.py:315-318), fixture generated from the reference with
    its top-level training flag raised and two injected noise tensors."""
    mod, x, gt, eps, eps_p, z = build_predictor_zp_case(name)
    sd = mod.state_dict()
    outs = O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], True, eps, predict_features_gt=gt,
                               decode_with_posterior=True, eps_p=eps_p)
    assert len(outs) == 5 and list(outs[0].shape) == list(z["shape"])
    np.testing.assert_allclose(golden_sample(outs[0], z), z["sample"], atol=5e-5, rtol=0)
    for key, ours, gold in golden_latents(outs, z):
        np.testing.assert_allclose(ours, gold, atol=5e-5, rtol=0, err_msg=key)
    with pytest.raises(AssertionError):                                # Predictor.py:316
        O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], True, eps, decode_with_posterior=True)


@pytest.mark.parametrize("name", AE_CASES)
def test_autoencoder_oracle_matches_reference(name):
    enc, dec, x, f_in, cfg, ze, zd = build_ae_case(name)
    feats = O.resnet_encoder(enc.state_dict(), x, cfg["n_down"], cfg["n_res"])
    frames = O.resnet_decoder(dec.state_dict(), f_in, cfg["n_down"], cfg["out_layer"])
    assert list(feats.shape) == list(ze["shape"]) and list(frames.shape) == list(zd["shape"])
    np.testing.assert_allclose(golden_sample(feats, ze), ze["sample"], atol=5e-4 * max(1.0, float(ze["absmax"])), rtol=0)
    np.testing.assert_allclose(golden_sample(frames, zd), zd["sample"], atol=5e-5, rtol=0)


def test_coordinate_asserts():
    hl = torch.linspace(0, 7, 8)
    with pytest.raises(AssertionError):
        O.coor_generator(torch.tensor([0., 13.]), hl, hl, 12, 8, 8)
    c = O.coor_generator(torch.tensor([1.5]), hl, hl, 12, 8, 8)
    assert c.shape == (64, 3) and abs(float(c[9, 0]) - 0.125) < 1e-7 and float(c[9, 1]) == 0.125 and float(c[9, 2]) == 0.125
