"""CPU restatement of the reference's pixel-space post-processing and evaluation metrics.

TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_golden_post.py and nothing in the product path.
Pinned by tests/golden/post_*.npz, which make_golden_post.py generates from the UNMODIFIED reference
(utils/metrics.py, utils/dataset.py) and torchvision 0.26 (the reference's ToTensor / Normalize / ToPILImage provider).

Plain torch fp32, one function per reference function, file:line cited.
"""
from __future__ import annotations

from math import exp

import torch
import torch.nn.functional as F


def renormalize_clamp(frames: torch.Tensor, mean, std) -> torch.Tensor:
    """Model space -> [0,1] pixel space.  frames (..., C, H, W) fp32.

    VidReNormalize (utils/dataset.py:860-886) composes Normalize(0, 1/std) and Normalize(-mean, 1), i.e. in fp32
    ``(x - 0) / fl32(1/std) - fl32(-mean)`` (torchvision Normalize = sub_(mean).div_(std)); callers then clamp to [0,1]
    (utils/train_summary.py:244-245)."""
    C = frames.shape[-3]
    inv_std = torch.tensor([1.0 / float(s) for s in std], dtype=torch.float32).view(C, 1, 1)
    inv_mean = torch.tensor([-float(m) for m in mean], dtype=torch.float32).view(C, 1, 1)
    x = frames.to(torch.float32)
    x = (x - 0.0) / inv_std
    x = (x - inv_mean) / 1.0
    return x.clamp(0.0, 1.0)


def to_uint8(pixels: torch.Tensor) -> torch.Tensor:
    """[0,1] fp32 -> uint8 the way the reference writes images: ToPILImage (utils/train_summary.py:246-248) =
    torchvision ``pic.mul(255).byte()``, i.e. truncation, not rounding."""
    return pixels.to(torch.float32).mul(255).to(torch.uint8)


def normalize_u8(frames_u8: torch.Tensor, mean, std) -> torch.Tensor:
    """uint8 pixels (..., C, H, W) -> model space: VidToTensor (x/255, utils/dataset.py:835-844) + VidNormalize
    ((x - mean) / std, utils/dataset.py:846-858)."""
    C = frames_u8.shape[-3]
    m = torch.tensor([float(v) for v in mean], dtype=torch.float32).view(C, 1, 1)
    s = torch.tensor([float(v) for v in std], dtype=torch.float32).view(C, 1, 1)
    x = frames_u8.to(torch.float32).div(255)
    return (x - m) / s


def psnr(x: torch.Tensor, y: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """Per-image PSNR, utils/metrics.py:12-30 with mean_flag=False.  x, y (N, C, H, W)."""
    x = x / float(data_range)
    y = y / float(data_range)
    mse = torch.mean((x - y) ** 2, dim=(1, 2, 3))
    return -10 * torch.log10(mse + 1e-8)


def gaussian_window(window_size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    """utils/metrics.py:79-88: normalised 1-D Gaussian, outer product -> (ws, ws) fp32."""
    g = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    return g.mm(g.t()).float()


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11) -> torch.Tensor:
    """Per-image SSIM, utils/metrics.py:90-109 with mean_flag=False: depthwise 11x11 Gaussian, zero padding 5."""
    C = img1.shape[1]
    w = gaussian_window(window_size).expand(C, 1, window_size, window_size).contiguous()
    p = window_size // 2
    mu1 = F.conv2d(img1, w, padding=p, groups=C)
    mu2 = F.conv2d(img2, w, padding=p, groups=C)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, w, padding=p, groups=C) - mu1_sq
    s2 = F.conv2d(img2 * img2, w, padding=p, groups=C) - mu2_sq
    s12 = F.conv2d(img1 * img2, w, padding=p, groups=C) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return torch.mean(m, dim=(1, 2, 3))


def best_of_k(samples: torch.Tensor, gt: torch.Tensor, metric: str = "psnr"):
    """Best-of-K selection: samples (N, K, T, C, H, W), gt (N, T, C, H, W), pixel space.  Per-frame metric with the functions
    above (the reference applies its metrics frame by frame, utils/metrics.py:127-135), mean over time per sample, argmax over
    the K samples of a clip.  Returns (best (N, T, C, H, W), best_idx (N,), mean_scores (N, K))."""
    N, K, T = samples.shape[:3]
    fn = psnr if metric == "psnr" else ssim
    flat = samples.reshape(N * K * T, *samples.shape[3:])
    ref = gt.unsqueeze(1).expand(N, K, *gt.shape[1:]).reshape(N * K * T, *gt.shape[2:])
    scores = fn(flat, ref).reshape(N, K, T)
    mean_scores = scores.double().mean(-1).float()
    idx = mean_scores.argmax(dim=1)
    return samples[torch.arange(N), idx], idx, mean_scores
