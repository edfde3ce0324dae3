"""On-device evaluation metrics with the reference's names and signatures (utils/metrics.py:12-109).

``PSNR(x, y, data_range=1.0, mean_flag=True)`` and ``SSIM(window_size=11)(img1, img2, mean_flag=True)`` take fp32 CUDA
tensors (N, C, H, W) and run one CUDA kernel each (npvp_psnr / npvp_ssim); there is no CPU path.
"""
from __future__ import annotations

from math import exp

import torch

from . import _lib


def _check(x, y):
    if not (isinstance(x, torch.Tensor) and isinstance(y, torch.Tensor) and x.is_cuda and y.is_cuda):
        raise NotImplementedError("npvp_b200.metrics: inputs must be CUDA tensors (there is no CPU fallback)")
    if x.dim() != 4 or x.shape != y.shape:
        raise ValueError(f"npvp_b200.metrics: expected two (N, C, H, W) tensors of the same shape, got {tuple(x.shape)} and {tuple(y.shape)}")
    return x.detach().to(torch.float32).contiguous(), y.detach().to(torch.float32).contiguous()


def PSNR(x, y, data_range=1.0, mean_flag: bool = True):
    """utils/metrics.py:12-30: average (mean_flag) or per-image PSNR of two batches of images."""
    x, y = _check(x, y)
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    _lib.ops().psnr(x, y, out, data_range)
    return torch.mean(out).item() if mean_flag else out


def gaussian_window_1d(window_size: int = 11, sigma: float = 1.5):
    """utils/metrics.py:79-81: the normalised 1-D Gaussian whose outer product is the reference's window (fp32 arithmetic)."""
    g = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return (g / g.sum()).tolist()


class SSIM(torch.nn.Module):
    """utils/metrics.py:47-109.  Only the reference's default window (11, sigma 1.5) is implemented."""

    def __init__(self, window_size: int = 11):
        super().__init__()
        if window_size != 11:
            raise NotImplementedError("npvp_b200.metrics.SSIM: the CUDA kernel implements the reference's 11x11 window only")
        self.window_size = window_size
        self.window = gaussian_window_1d(window_size, 1.5)
        self.__name__ = "SSIM"

    def forward(self, img1, img2, mean_flag: bool = True):
        x, y = _check(img1, img2)
        out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        _lib.ops().ssim(x, y, self.window, out)
        # the reference's mean_flag=True is the mean over every element of the SSIM map = mean of the per-image means
        return out.mean() if mean_flag else out


def best_of_k(samples, gt, metric: str = "psnr", data_range=1.0):
    """Best-of-K selection for stochastic prediction (NPVP-S, BASELINE config 3): ``samples`` (N, K, T, C, H, W) and ``gt``
    (N, T, C, H, W) are pixel-space fp32 CUDA tensors.  Every sample's frames are scored against the clip's ground truth with
    the reference's PSNR or SSIM (utils/metrics.py:12-109, per frame like ``pred_ave_metrics`` :111-140), averaged over time,
    and the best sample of every clip is returned:  (best (N, T, C, H, W), best_idx int32 (N,), mean_scores (N, K),
    scores (N, K, T)).  Two kernels (score, select); the ground truth is never replicated K times."""
    if not (isinstance(samples, torch.Tensor) and isinstance(gt, torch.Tensor) and samples.is_cuda and gt.is_cuda):
        raise NotImplementedError("npvp_b200.metrics.best_of_k: inputs must be CUDA tensors (there is no CPU fallback)")
    if samples.dim() != 6 or gt.dim() != 5 or tuple(samples.shape[:1] + samples.shape[2:]) != tuple(gt.shape):
        raise ValueError(f"best_of_k: expected samples (N, K, T, C, H, W) and gt (N, T, C, H, W), got {tuple(samples.shape)} and {tuple(gt.shape)}")
    if metric not in ("psnr", "ssim"):
        raise ValueError("best_of_k: metric must be 'psnr' or 'ssim'")
    x, y = samples.detach().to(torch.float32).contiguous(), gt.detach().to(torch.float32).contiguous()
    n, K, T = x.shape[:3]
    scores = torch.empty((n, K, T), dtype=torch.float32, device=x.device)
    _lib.ops().sample_scores(x, y, scores, gaussian_window_1d(11, 1.5) if metric == "ssim" else None, data_range)
    best = torch.empty_like(y)
    idx = torch.empty(n, dtype=torch.int32, device=x.device)
    mean_scores = torch.empty((n, K), dtype=torch.float32, device=x.device)
    _lib.ops().best_of_k(scores, x, idx, mean_scores, best)
    return best, idx, mean_scores, scores
