"""CPU: run the host-side engines end to end with the kernel *specifications* standing in for the CUDA
kernels (tests/kernel_specs.SpecOps) and compare with the oracle.  This validates weight packing, buffer
plumbing and the kernel sequence without a GPU; the CUDA kernels themselves are checked against the same
specs in the -m gpu tests."""
import pytest
import torch

import npvp_b200._lib as _lib
from cases import (AE_CASES, PRED_CASES, PRED_GT_CASES, PRED_SPADE_CASES, PRED_ZP_CASES, build_ae_case, build_predictor_case,
                   build_predictor_gt_case, build_predictor_zp_case)
from kernel_specs import SpecOps
from oracle import npvp_oracle as O

torch.set_grad_enabled(False)


@pytest.fixture(autouse=True)
def spec_ops():
    old = _lib._OPS
    _lib.set_ops(SpecOps())
    yield
    _lib.set_ops(old)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


@pytest.mark.parametrize("name", PRED_CASES)
def test_predictor_engine_vs_oracle(name):
    from npvp_b200.engine_predictor import PredictorEngine
    mod, x, eps, stoch, _ = build_predictor_case(name)
    sd = mod.state_dict()
    ref = O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], stoch, eps if stoch else None)
    mod.injected_eps = eps if stoch else None
    out = PredictorEngine(mod).run(x)
    assert out.shape == ref.shape
    assert _rel(out, ref) < 3e-2, _rel(out, ref)
    out_cl = PredictorEngine(mod).run(x.permute(0, 1, 3, 4, 2).contiguous(), channels_last=True)
    assert torch.equal(out_cl.permute(0, 1, 4, 2, 3), out)


@pytest.mark.parametrize("name", PRED_SPADE_CASES)
def test_predictor_engine_spade(name):
    from npvp_b200.engine_predictor import PredictorEngine
    mod, x, eps, stoch, _ = build_predictor_case(name)
    sd = mod.state_dict()
    ref = O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], stoch, eps, fuse_method="SPADE")
    mod.injected_eps = eps
    out = PredictorEngine(mod).run(x)
    assert _rel(out, ref) < 3e-2, _rel(out, ref)


@pytest.mark.parametrize("name", PRED_GT_CASES)
def test_predictor_engine_posterior_branch(name):
    from npvp_b200.engine_predictor import PredictorEngine
    mod, x, gt, eps, _ = build_predictor_gt_case(name)
    sd = mod.state_dict()
    ref = O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], True, eps, predict_features_gt=gt)
    mod.injected_eps = eps
    eng = PredictorEngine(mod)
    outs = eng.run(x, predict_gt=gt)
    assert len(outs) == 5
    for a, b in zip(outs, ref):
        assert a.shape == b.shape
        assert _rel(a, b) < 3e-2, _rel(a, b)
    assert torch.equal(eng.run(x), outs[0])         # same prediction with and without the ground truth


@pytest.mark.parametrize("name", PRED_ZP_CASES)
def test_predictor_engine_posterior_decode(name):
    """Decoder driven by z_p (Predictor.py:315-318): engine vs oracle; the result differs from the prior-driven forward."""
    from npvp_b200.engine_predictor import PredictorEngine
    mod, x, gt, eps, eps_p, _ = build_predictor_zp_case(name)
    sd = mod.state_dict()
    ref = O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], True, eps, predict_features_gt=gt,
                              decode_with_posterior=True, eps_p=eps_p)
    mod.injected_eps, mod.injected_eps_p = eps, eps_p
    eng = PredictorEngine(mod)
    outs = eng.run(x, predict_gt=gt, decode_posterior=True)
    assert len(outs) == 5
    for a, b in zip(outs, ref):
        assert a.shape == b.shape
        assert _rel(a, b) < 3e-2, _rel(a, b)
    # the fixture tells the two latent samples apart: the prior-driven forward is 4x further from the reference than our error
    prior_ref = O.predictor_forward(sd, x, sd["observed_coor"], sd["predict_coor"], True, eps)
    gap = float((prior_ref - ref[0]).abs().mean())
    assert float((outs[0] - ref[0]).abs().mean()) < 0.25 * gap
    assert float((eng.run(x) - ref[0]).abs().mean()) > 0.75 * gap


def test_predictor_engine_per_clip_timestamps():
    """reset_pos_coor_per_clip: clip i of a mixed batch == the oracle run on clip i with its own timestamps."""
    from npvp_b200.engine_predictor import PredictorEngine
    mod, x, eps, _, _ = build_predictor_case("pred_S_stress_realT")
    x = torch.cat([x, x.flip(1) * 0.5 + 0.1], 0)                      # two clips, 3 context frames each
    eps = torch.cat([eps, eps.flip(1)], 0)
    to = torch.tensor([[0., 1., 4.], [0., 7., 8.]])                   # clip 0: prediction-like, clip 1: interpolation-like
    tp = torch.tensor([[2., 2.5, 3., 5.], [1., 3.25, 4., 6.5]])
    sd = mod.state_dict()
    hl = torch.linspace(0, 7, 8)
    refs = []
    for i in range(2):
        oc = O.coor_generator(to[i], hl, hl, mod.max_T, 8, 8)
        pc = O.coor_generator(tp[i], hl, hl, mod.max_T, 8, 8)
        refs.append(O.predictor_forward(sd, x[i:i + 1], oc, pc, True, eps[i:i + 1]))
    ref = torch.cat(refs, 0)
    mod.reset_pos_coor_per_clip(to, tp)
    mod.injected_eps = eps
    out = PredictorEngine(mod).run(x)
    assert out.shape == ref.shape == (2, 4, 512, 8, 8)
    assert _rel(out, ref) < 3e-2, _rel(out, ref)
    mod.reset_pos_coor(to[0], tp[0])                                  # back to shared timestamps
    mod.injected_eps = eps[:1]
    assert _rel(PredictorEngine(mod).run(x[:1]), refs[0]) < 3e-2


@pytest.mark.parametrize("name", AE_CASES)
def test_autoencoder_engine_vs_oracle(name):
    from npvp_b200.engine_autoencoder import DecoderEngine, EncoderEngine
    enc, dec, x, f_in, cfg, _, _ = build_ae_case(name)
    feats_ref = O.resnet_encoder(enc.state_dict(), x, cfg["n_down"], cfg["n_res"])
    frames_ref = O.resnet_decoder(dec.state_dict(), f_in, cfg["n_down"], cfg["out_layer"])
    feats = EncoderEngine(enc).run(x)
    frames = DecoderEngine(dec).run(f_in)
    assert feats.shape == feats_ref.shape and frames.shape == frames_ref.shape
    assert _rel(feats, feats_ref) < 3e-2, _rel(feats, feats_ref)
    assert float((frames - frames_ref).abs().max()) < 2e-2
