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
