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


class U8Pipeline:
    """Double-buffered uint8 -> Generator -> uint8 inference from / to PINNED host memory: the H2D copy of batch i+1 and
    the D2H copy of batch i-1 run on their own streams while batch i computes, so the end-to-end rate approaches the
    resident rate (3 bytes per pixel cross PCIe each way instead of the 12 + 12 of fp32 NCHW staging).

        pipe = U8Pipeline(G)                       # G: uegan_b200.models.Generator in eval mode
        for src, dst in batches:                   # pinned uint8 (N, H, W, 3) tensors
            pipe.submit(src, dst)
        pipe.drain()                               # dst of every submitted batch is complete
    """

    def __init__(self, G, depth: int = 2):
        self.G, self.depth = G, depth
        self.slots, self.i = [], 0
        self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def _slot(self, shape, device):
        if len(self.slots) < self.depth:
            self.slots.append(dict(inp=torch.empty(shape, dtype=torch.uint8, device=device),
                                   out=torch.empty(shape, dtype=torch.uint8, device=device),
                                   e_in=torch.cuda.Event(), e_comp=torch.cuda.Event(), e_out=torch.cuda.Event(), used=False))
        return self.slots[self.i % self.depth]

    @torch.no_grad()
    def submit(self, src_host: torch.Tensor, dst_host: torch.Tensor):
        assert src_host.is_pinned() and dst_host.is_pinned() and src_host.dtype == torch.uint8
        main = torch.cuda.current_stream()
        sl = self._slot(tuple(src_host.shape), main.device)
        self.i += 1
        with torch.cuda.stream(self.s_in):
            if sl["used"]:
                self.s_in.wait_event(sl["e_comp"])   # the slot's previous batch has been consumed by the pack kernel
            sl["inp"].copy_(src_host, non_blocking=True)
            sl["e_in"].record(self.s_in)
        main.wait_event(sl["e_in"])
        if sl["used"]:
            main.wait_event(sl["e_out"])             # the slot's previous output has left for the host
        enhance_u8(self.G, sl["inp"], out=sl["out"])
        sl["e_comp"].record(main)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(sl["e_comp"])
            dst_host.copy_(sl["out"], non_blocking=True)
            sl["e_out"].record(self.s_out)
        sl["used"] = True

    def drain(self):
        self.s_out.synchronize()
        torch.cuda.current_stream().synchronize()
