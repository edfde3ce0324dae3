"""GPU: pixel-space post-processing and metrics kernels (through the C-ABI) against the reference fixtures and the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["post_cityscapes", "post_kth", "post_smmnist"]


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="module")
def op():
    from npvp_b200 import _lib
    return _lib.ops()


@pytest.mark.parametrize("name", CASES)
def test_pixels_bit_exact_vs_reference(op, name):
    g = load(name)
    mean, std = g["mean"].tolist(), g["std"].tolist()
    x = g["frames"].to(DEV)
    f32, u8 = torch.empty_like(x), torch.empty(x.shape, dtype=torch.uint8, device=DEV)
    op.frames_to_pixels(x, mean, std, out_f32=f32, out_u8=u8)
    assert torch.equal(f32.cpu(), g["pix"])
    assert torch.equal(u8.cpu(), g["u8"])
    back = torch.empty_like(x)
    op.pixels_to_frames(u8, mean, std, back)
    assert float((back.cpu() - g["back"]).abs().max()) <= 1e-6


@pytest.mark.parametrize("hw", [(128, 128), (37, 53)])
def test_pixels_vector_and_scalar_paths_vs_oracle(op, hw):
    from oracle import post_oracle as P
    mean, std = (0.31604213, 0.35114038, 0.3104223), (1.2172801, 1.3219808, 1.2082524)
    x = torch.randn(3, 2, 3, *hw, generator=torch.Generator().manual_seed(3)) * 0.7
    f32, u8 = torch.empty_like(x, device=DEV), torch.empty(x.shape, dtype=torch.uint8, device=DEV)
    op.frames_to_pixels(x.to(DEV), mean, std, out_f32=f32, out_u8=u8)
    ref = P.renormalize_clamp(x, mean, std)
    assert torch.equal(f32.cpu(), ref) and torch.equal(u8.cpu(), P.to_uint8(ref))


@pytest.mark.parametrize("name", CASES)
def test_metrics_vs_reference(name):
    from npvp_b200.metrics import PSNR, SSIM
    g = load(name)
    n = g["pix"].shape[0] * g["pix"].shape[1]
    x, y = g["pix"].reshape(n, *g["pix"].shape[2:]).to(DEV), g["gt"].reshape(n, *g["pix"].shape[2:]).to(DEV)
    ps = PSNR(x, y, mean_flag=False)
    ss = SSIM()(x, y, mean_flag=False)
    assert float((ps.cpu() - g["psnr"]).abs().max()) <= 1e-4          # dB
    assert float((ss.cpu() - g["ssim"]).abs().max()) <= 1e-5
    assert abs(PSNR(x, y) - float(g["psnr"].mean())) <= 1e-4
    assert abs(float(SSIM()(x, y)) - float(g["ssim"].mean())) <= 1e-5


def test_metrics_full_size_properties():
    """Full Cityscapes frame size: identical images -> SSIM 1, PSNR = 80 dB (the 1e-8 floor); symmetry; range scaling."""
    from npvp_b200.metrics import PSNR, SSIM
    from oracle import post_oracle as P
    g = torch.Generator().manual_seed(5)
    x, y = torch.rand(6, 3, 128, 128, generator=g), torch.rand(6, 3, 128, 128, generator=g)
    xd, yd = x.to(DEV), y.to(DEV)
    assert float((SSIM()(xd, xd, mean_flag=False) - 1).abs().max()) <= 1e-6
    assert float((PSNR(xd, xd, mean_flag=False) - 80.0).abs().max()) <= 1e-4
    a, b = SSIM()(xd, yd, mean_flag=False), SSIM()(yd, xd, mean_flag=False)
    assert float((a - b).abs().max()) <= 1e-6
    assert float((a.cpu() - P.ssim(x, y)).abs().max()) <= 1e-5
    assert float((PSNR(xd, yd, mean_flag=False).cpu() - P.psnr(x, y)).abs().max()) <= 1e-4
    assert float((PSNR(xd * 255, yd * 255, data_range=255, mean_flag=False) - PSNR(xd, yd, mean_flag=False)).abs().max()) <= 1e-3
    with pytest.raises(NotImplementedError):
        PSNR(x, y)


def test_rollout_uint8_host_output():
    """rollout(out_host=uint8) == to_pixels(uint8=True) of the fp32 rollout, and to_pixels matches the oracle bit for bit."""
    from npvp_b200 import build_from_config
    from npvp_b200.config import preset
    from oracle import post_oracle as P
    model = build_from_config(preset("KITTI_VFP_NPVP-S"), device=DEV, seed=0)
    x = torch.rand(2, 4, 3, 128, 128, generator=torch.Generator().manual_seed(1)) * 2 - 1
    eps = [torch.randn(2, 512, 8, 8, generator=torch.Generator().manual_seed(10 + i)).to(DEV) for i in range(2)]
    host = torch.empty(2, 7, 3, 128, 128, dtype=torch.uint8).pin_memory()
    frames = model.rollout(x.to(DEV), 7, eps_list=eps, out_host=host)
    torch.cuda.synchronize()
    ref_u8 = model.to_pixels(frames, uint8=True)
    assert torch.equal(host, ref_u8.cpu())
    mean, std = model._renorm_constants()
    assert torch.equal(model.to_pixels(frames).cpu(), P.renormalize_clamp(frames.cpu(), mean, std))
    back = model.from_pixels(ref_u8)
    assert back.shape == frames.shape and back.dtype == torch.float32


@pytest.mark.parametrize("metric", ["psnr", "ssim"])
def test_best_of_k_selection_vs_oracle(metric):
    """npvp_b200.metrics.best_of_k (npvp_sample_scores + npvp_best_of_k): per-frame scores of K samples per clip against ONE
    ground truth, mean over time, argmax, winner's frames - against the oracle built from the reference's PSNR / SSIM."""
    from oracle import post_oracle as P
    from npvp_b200.metrics import best_of_k
    g = torch.Generator().manual_seed(5)
    N, K, T, C, H, W = 3, 5, 4, 3, 64, 48
    gt = torch.rand((N, T, C, H, W), generator=g)
    noise = torch.rand((N, K, 1, 1, 1, 1), generator=g) * 0.3 + 0.02           # every sample a different distance from the truth
    samples = (gt.unsqueeze(1) + noise * torch.randn((N, K, T, C, H, W), generator=g)).clamp(0, 1)
    best_ref, idx_ref, mean_ref = P.best_of_k(samples, gt, metric)
    best, idx, mean_scores, scores = best_of_k(samples.to(DEV), gt.to(DEV), metric)
    assert scores.shape == (N, K, T) and idx.dtype == torch.int32
    assert torch.equal(idx.cpu().long(), idx_ref)
    assert float((mean_scores.cpu() - mean_ref).abs().max()) <= (1e-3 if metric == "psnr" else 1e-5)
    assert torch.equal(best.cpu(), best_ref)
    # the per-sample scores equal the plain metric kernels on the replicated ground truth
    from npvp_b200.metrics import PSNR, SSIM
    flat = samples.reshape(N * K * T, C, H, W).to(DEV)
    rep = gt.unsqueeze(1).expand(N, K, T, C, H, W).reshape(N * K * T, C, H, W).contiguous().to(DEV)
    plain = PSNR(flat, rep, mean_flag=False) if metric == "psnr" else SSIM()(flat, rep, mean_flag=False)
    assert torch.equal(plain.reshape(N, K, T), scores)


def test_predict_best_of_k_pipeline():
    """NPVPInference.predict_best_of_k: 8 stochastic samples per clip (BASELINE config 3), best by PSNR against a ground truth that
    IS one of the samples - the selection must find it, and the returned frames are that sample's, bit for bit."""
    from npvp_b200.pipeline import build_from_config
    from util_init import seeded_rand, stress_init_
    model = build_from_config("BAIR_VFP_NPVP-S", device="cpu", seed=0)
    stress_init_(model.predictor, 3)
    model = model.cuda()
    x = (seeded_rand((2, 2, 3, 64, 64), 5) * 2 - 1).cuda()
    eps = torch.randn(16, 512, 8, 8, device="cuda", generator=torch.Generator("cuda").manual_seed(2))
    smp = model.predict_samples(x, 8, eps)
    gt = torch.stack([smp[0, 5], smp[1, 2]])                                   # clip 0: sample 5 is the truth, clip 1: sample 2
    best, idx, mean_scores = model.predict_best_of_k(x, gt, 8, "psnr", eps)
    assert idx.tolist() == [5, 2] and mean_scores.shape == (2, 8)
    assert torch.equal(best, gt)
    _, idx_s, _ = model.predict_best_of_k(x, gt, 8, "ssim", eps)
    assert idx_s.tolist() == [5, 2]


@pytest.mark.parametrize("preset_name", ["KITTI_VFP_NPVP-S", "SMMNIST_VFP_NPVP-D", "KTH_Unified_NPVP-S"])
def test_fused_pixel_ingest_and_epilogue(preset_name):
    """uint8 frames in (VidToTensor + VidNormalize inside the encoder's stem kernel) and uint8 frames out
    (VidReNormalize + clamp + ToPILImage's truncation inside the decoder's head kernel) are bit-identical to the stand-alone
    conversion kernels, which are bit-identical to the reference transforms (test_pixels_bit_exact_vs_reference)."""
    from npvp_b200.pipeline import build_from_config
    from util_init import stress_init_
    model = build_from_config(preset_name, device="cpu", seed=0)
    stress_init_(model.VPTR_Enc, 1)
    stress_init_(model.VPTR_Dec, 2)
    model = model.cuda()
    cfg = model.cfg
    To, ch, hw = cfg.Dataset.num_past_frames, cfg.Dataset.img_channels, cfg.Dataset.img_size
    if cfg.Predictor.rand_context:
        model.predictor.reset_pos_coor(model.to_list, model.tp_list)
    u8 = torch.randint(0, 256, (2, To, ch, hw, hw), dtype=torch.uint8, generator=torch.Generator().manual_seed(4)).cuda()
    eps = torch.randn(2, 512, 8, 8, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    x = model.from_pixels(u8)                                         # stand-alone kernel (reference operation order)
    # encoder: features of the uint8 frames == features of the normalised fp32 frames
    assert torch.equal(model.VPTR_Enc.forward_tokens(u8, norm=model._norm_constants()), model.VPTR_Enc.forward_tokens(x))
    # whole path, eager and captured: frames identical, fused uint8 == to_pixels(frames, uint8=True)
    ref = model.predict(x, eps)
    for graphs in (False, True):
        model.use_cuda_graphs(graphs)
        frames, frames_u8 = model.predict(u8, eps, pixels_u8=True)
        assert frames_u8.dtype == torch.uint8 and torch.equal(frames, ref)
        assert torch.equal(frames_u8, model.to_pixels(ref, uint8=True))
    model.use_cuda_graphs(False)
    only_u8 = model.VPTR_Dec.forward_tokens(model.predictor.forward_tokens(model.VPTR_Enc.forward_tokens(x), out16=model.VPTR_Dec._engine().dt),
                                            renorm=model._renorm_constants(), want_f32=False)
    assert only_u8[0] is None


def test_rollout_uint8_host_buffers_use_fused_epilogue():
    """rollout with a uint8 host input and a uint8 host output: pixels in, pixels out, equal to the fp32 rollout converted afterwards."""
    from npvp_b200.pipeline import build_from_config
    model = build_from_config("BAIR_VFP_NPVP-S", device="cuda", seed=0)
    u8 = torch.randint(0, 256, (2, 2, 3, 64, 64), dtype=torch.uint8, generator=torch.Generator().manual_seed(4))
    eps = [torch.randn(2, 512, 8, 8, device="cuda", generator=torch.Generator("cuda").manual_seed(i)) for i in range(3)]
    ref = model.rollout(model.from_pixels(u8.cuda()), 28, eps, last_block="query")
    ref_u8 = model.to_pixels(ref, uint8=True).cpu()
    host_in, host_out = u8.pin_memory(), torch.empty((2, 28, 3, 64, 64), dtype=torch.uint8).pin_memory()
    for graphs in (False, True):
        model.use_cuda_graphs(graphs)
        host_out.zero_()
        out = model.rollout(host_in, 28, eps, out_host=host_out, last_block="query")
        torch.cuda.synchronize()
        assert torch.equal(out, ref) and torch.equal(host_out, ref_u8)
