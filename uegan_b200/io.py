"""GPU-side input / output conversion around the hot path.

`pack_u8`    uint8 HWC batch -> (NHWC operand with halo, fp32 NCHW planes): transforms.ToTensor() + Normalize(0.5, 0.5)
             (data_loader.py:79-81,100-103) or the ImageNet normalisation of losses.py:26-27, bit-exact, in one pass.
`unpack_u8`  fp32 NCHW generator output -> uint8 HWC exactly as tester.py:70-75 saves it (denorm + save_image rounding).
`enhance_u8` models.Generator inference from uint8 to uint8: 3 bytes per pixel cross PCIe in each direction instead of
             the 12 + 12 of the fp32 NCHW staging.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from . import kernels as K

MEAN_05 = (0.5, 0.5, 0.5)
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _check_u8(img):
    if not (img.is_cuda and img.dtype == torch.uint8 and img.dim() == 4 and img.shape[3] == 3 and img.is_contiguous()):
        raise L.UeganError("expected a contiguous CUDA uint8 tensor of shape (N, H, W, 3)")


def pack_u8(img_u8: torch.Tensor, dst: Optional[K.NHWC] = None, planes: bool = True, mean=MEAN_05, std=MEAN_05,
            pad_mode: int = L.PAD_REFLECT, planes_out: Optional[torch.Tensor] = None):
    """Returns the fp32 NCHW tensor ToTensor + Normalize(mean, std) would give (or None) and fills `dst` (if given)."""
    _check_u8(img_u8)
    n, h, w, _ = img_u8.shape
    out = planes_out
    if out is None and planes:
        out = torch.empty(n, 3, h, w, dtype=torch.float32, device=img_u8.device)
    L.check(L.load().uegan_pack_input_u8(img_u8.data_ptr(), n, h, w, dst.ref() if dst is not None else None,
                                         out.data_ptr() if out is not None else None, pad_mode, L.float3(mean),
                                         L.float3(std), K._stream()), "pack_input_u8")
    K._count(1, "pack_input_u8", dst)
    return out


def unpack_u8(x_nchw: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if not (x_nchw.is_cuda and x_nchw.dtype == torch.float32 and x_nchw.dim() == 4 and x_nchw.is_contiguous()):
        raise L.UeganError("expected a contiguous CUDA fp32 tensor of shape (N, C, H, W)")
    n, c, h, w = x_nchw.shape
    if out is None:
        out = torch.empty(n, h, w, c, dtype=torch.uint8, device=x_nchw.device)
    L.check(L.load().uegan_unpack_output_u8(x_nchw.data_ptr(), out.data_ptr(), n, c, h, w, K._stream()),
            "unpack_output_u8")
    K._count(1, "unpack_output_u8")
    return out


@torch.no_grad()
def enhance_u8(G, img_u8: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """uint8 HWC in -> Generator (eval) -> uint8 HWC out, all on the GPU."""
    _check_u8(img_u8)
    n, h, w, _ = img_u8.shape
    P = G._plan(n, h, w, img_u8.device)
    x = P.get("x_planes")
    if x is None:
        x = P["x_planes"] = torch.empty(n, 3, h, w, dtype=torch.float32, device=img_u8.device)
    pack_u8(img_u8, P["x0"], planes_out=x)
    y = G.forward_native(x, packed=True)
    return unpack_u8(y, out)
