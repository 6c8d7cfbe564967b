"""validation metrics on the GPU (uint8 HWC image batches, device resident).

psnr_u8  metrics/CalcPSNR.py:47-52,85-92 (RGB, 4-pixel border crop, data_range 255).
ssim_u8  metrics/CalcSSIM.py:47-62 = skimage.metrics.structural_similarity(multichannel=True, data_range=255) on the
         cropped pair: 7x7 uniform window, sample covariance.
Both return one float64 value per image pair (a CPU tensor: reading a metric is a host synchronisation by nature).
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import kernels as K


def _check_pair(a, b):
    for t in (a, b):
        if not (t.is_cuda and t.dtype == torch.uint8 and t.dim() == 4 and t.is_contiguous()):
            raise L.UeganError("expected contiguous CUDA uint8 tensors of shape (N, H, W, C)")
    if a.shape != b.shape:
        raise ValueError("Input images must have the same dimensions.")


def sse_u8(a: torch.Tensor, b: torch.Tensor, crop: int = 4) -> torch.Tensor:
    _check_pair(a, b)
    n, h, w, c = a.shape
    out = torch.empty(n, dtype=torch.int64, device=a.device)
    L.check(L.load().uegan_sse_u8(a.data_ptr(), b.data_ptr(), n, h, w, c, crop, out.data_ptr(), K._stream()), "sse_u8")
    K._count(1, "sse_u8")
    return out


def psnr_u8(a: torch.Tensor, b: torch.Tensor, crop: int = 4, data_range: float = 255.0) -> torch.Tensor:
    n, h, w, c = a.shape
    sse = sse_u8(a, b, crop).cpu().double()
    mse = sse / float((h - 2 * crop) * (w - 2 * crop) * c)
    return torch.where(mse == 0, torch.full_like(mse, float("inf")), 10.0 * torch.log10((data_range ** 2) / mse))


def ssim_u8(a: torch.Tensor, b: torch.Tensor, crop: int = 4) -> torch.Tensor:
    _check_pair(a, b)
    n, h, w, c = a.shape
    out = torch.empty(n, dtype=torch.float64, device=a.device)
    L.check(L.load().uegan_ssim_u8(a.data_ptr(), b.data_ptr(), n, h, w, c, crop, out.data_ptr(), K._stream()), "ssim_u8")
    K._count(1, "ssim_u8")
    return out.cpu() / float((h - 2 * crop - 6) * (w - 2 * crop - 6) * c)
