"""ctypes binding of libuegan_sm100.so (the C ABI declared in include/uegan_sm100.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing or a call fails, this
module raises.  Build the library with `python -m uegan_b200.build` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libuegan_sm100.so")

F32, BF16, F16 = 0, 1, 2
ACT_NONE, ACT_LRELU, ACT_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3, 4
PAD_ZERO, PAD_REFLECT = 0, 1
ABI_VERSION = 4


class UeganError(RuntimeError):
    pass


class Tensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("halo", C.c_int32), ("dtype", C.c_int32), ("scale", C.c_void_p)]


class ScaleEntry(C.Structure):
    _fields_ = [("data", C.c_void_p), ("numel", C.c_int64), ("dtype", C.c_int32), ("reserved", C.c_int32),
                ("scale", C.c_void_p)]


class ConvDesc(C.Structure):
    _fields_ = [("x", Tensor), ("y", Tensor), ("y_c_off", C.c_int32), ("cout", C.c_int32), ("k", C.c_int32),
                ("stride", C.c_int32), ("pad", C.c_int32), ("act", C.c_int32), ("w_packed", C.c_void_p),
                ("w_scale", C.c_void_p), ("bias", C.c_void_p), ("alpha", C.c_void_p), ("mul", C.POINTER(Tensor)), ("out_nchw", C.c_void_p),
                ("residual_nchw", C.c_void_p), ("aux_nchw", C.c_void_p), ("y_mul", C.c_int32), ("y_off_h", C.c_int32), ("y_off_w", C.c_int32),
                ("mask", C.POINTER(Tensor)), ("mask_act", C.c_int32), ("in_stats", C.c_void_p), ("y_premul", C.POINTER(Tensor)), ("y_reflect_halo", C.c_int32), ("y_cls_c", C.c_int32)]


# name -> (restype, argtypes); every symbol include/uegan_sm100.h declares
SYMBOLS = {
    "uegan_abi_version": (C.c_int, []),
    "uegan_last_error": (C.c_char_p, []),
    "uegan_device_error": (C.c_int, []),
    "uegan_packed_weight_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "uegan_pack_conv_weight": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "uegan_pack_conv_weight_scaled": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 7 + [C.c_void_p, C.c_void_p]),
    "uegan_pack_conv_weight_dgrad_scaled": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 10 + [C.c_void_p, C.c_void_p]),
    "uegan_pack_conv_weight_dgrad4": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 7 + [C.c_void_p, C.c_void_p]),
    "uegan_scale_update": (C.c_int, [C.c_void_p, C.c_int32, C.c_float, C.c_int32, C.c_void_p]),
    "uegan_conv2d_fprop": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "uegan_conv2d_rowsum_supported": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "uegan_packed_weight_rowsum_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "uegan_pack_conv_weight_rowsum": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p]),
    "uegan_pack_conv_weight_rowsum_scaled": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 7 + [C.c_void_p, C.c_void_p]),
    "uegan_conv2d_fprop_rowsum": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p]),
    "uegan_pack_input": (C.c_int, [C.c_void_p, C.POINTER(Tensor), C.c_int32, C.POINTER(C.c_float),
                                   C.POINTER(C.c_float), C.c_void_p]),
    "uegan_halo_fill": (C.c_int, [C.POINTER(Tensor), C.c_int32, C.c_void_p]),
    "uegan_instance_norm": (C.c_int, [C.POINTER(Tensor), C.POINTER(Tensor), C.c_int32, C.c_float, C.c_void_p,
                                      C.c_void_p]),
    "uegan_instance_norm_apply": (C.c_int, [C.POINTER(Tensor), C.POINTER(Tensor), C.c_int32, C.c_float, C.c_void_p,
                                            C.c_void_p]),
    "uegan_spectral_sigma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "uegan_spectral_sigma_batch": (C.c_int, [C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                             C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p),
                                             C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p]),
    "uegan_instance_norm_stats": (C.c_int, [C.POINTER(Tensor), C.c_float, C.c_void_p, C.c_int32,
                                            C.POINTER(C.c_void_p), C.c_void_p]),
    "uegan_gan_loss_fwd": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_int64), C.c_void_p, C.c_void_p, C.c_void_p]),
    "uegan_gan_loss_phase": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "uegan_gan_loss_bwd": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_int64), C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.c_void_p, C.c_float, C.c_void_p]),
    "uegan_in_mse_fwd": (C.c_int, [C.POINTER(Tensor), C.POINTER(Tensor), C.c_void_p, C.c_void_p, C.c_float,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "uegan_msrec_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "uegan_spectral_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "uegan_pack_conv_weight_dgrad": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 10 + [C.c_void_p]),
    "uegan_conv2d_wgrad": (C.c_int, [C.POINTER(Tensor), C.POINTER(Tensor)] + [C.c_int32] * 7 +
                           [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_size_t, C.c_void_p]),
    "uegan_conv2d_wgrad_hstack": (C.c_int, [C.POINTER(Tensor), C.POINTER(Tensor)] + [C.c_int32] * 6 +
                                  [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_size_t, C.c_void_p]),
    "uegan_wgrad_last_launches": (C.c_int, []),
    "uegan_conv2d_wgrad_zwin_supported": (C.c_int, [C.c_int32] * 7),
    "uegan_conv2d_wgrad_zwin": (C.c_int, [C.POINTER(Tensor), C.POINTER(Tensor)] + [C.c_int32] * 6 +
                                [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_size_t, C.c_void_p]),
    "uegan_fold_inplace": (C.c_int, [C.POINTER(Tensor), C.c_void_p]),
    "uegan_dz_hstack": (C.c_int, [C.POINTER(Tensor), C.c_int32, C.c_int32, C.POINTER(Tensor), C.c_void_p]),
    "uegan_head_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(Tensor),
                                 C.c_void_p]),
    "uegan_grad_combine": (C.c_int, [C.POINTER(Tensor), C.c_int32, C.c_int32, C.POINTER(Tensor), C.c_int32, C.c_int32,
                                     C.c_int32, C.POINTER(Tensor), C.c_int32, C.POINTER(Tensor), C.c_int32,
                                     C.POINTER(Tensor), C.c_int32, C.c_int32, C.POINTER(Tensor), C.c_int32,
                                     C.c_void_p]),
    "uegan_grad_combine2": (C.c_int, [C.POINTER(Tensor), C.c_int32, C.c_int32, C.POINTER(Tensor), C.c_int32, C.c_int32,
                                      C.c_int32, C.POINTER(Tensor), C.c_int32, C.POINTER(Tensor), C.c_int32,
                                      C.POINTER(Tensor), C.c_int32, C.c_int32, C.POINTER(Tensor), C.c_int32,
                                      C.POINTER(Tensor), C.c_int32, C.POINTER(Tensor), C.c_int32, C.c_void_p]),
    "uegan_channel_sum": (C.c_int, [C.POINTER(Tensor), C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "uegan_instance_norm_bwd": (C.c_int, [C.POINTER(Tensor), C.c_int32, C.POINTER(Tensor), C.c_void_p,
                                          C.POINTER(Tensor), C.c_void_p, C.c_void_p]),
    "uegan_upsample2x_bwd": (C.c_int, [C.POINTER(Tensor), C.c_int32, C.POINTER(Tensor), C.c_void_p]),
    "uegan_maxpool2x2_bwd": (C.c_int, [C.POINTER(Tensor), C.POINTER(Tensor), C.POINTER(Tensor), C.c_void_p]),
    "uegan_in_mse_bwd": (C.c_int, [C.POINTER(Tensor), C.POINTER(Tensor), C.c_void_p, C.c_void_p, C.c_float,
                                   C.c_void_p, C.POINTER(Tensor), C.POINTER(Tensor), C.c_void_p, C.c_void_p]),
    "uegan_in_mse_joint": (C.c_int, [C.POINTER(Tensor), C.POINTER(Tensor), C.c_float, C.c_float, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.c_void_p]),
    "uegan_in_mse_bwd_apply": (C.c_int, [C.POINTER(Tensor), C.POINTER(Tensor), C.c_void_p, C.c_void_p, C.c_float,
                                         C.c_void_p, C.POINTER(Tensor), C.POINTER(Tensor), C.c_void_p, C.c_void_p]),
    "uegan_unpack_input_grad": (C.c_int, [C.POINTER(Tensor), C.POINTER(C.c_float), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "uegan_upsample2x": (C.c_int, [C.POINTER(Tensor), C.POINTER(Tensor), C.c_int32, C.c_void_p]),
    "uegan_cat_build": (C.c_int, [C.POINTER(Tensor), C.POINTER(Tensor), C.c_void_p, C.POINTER(Tensor), C.c_void_p]),
    "uegan_maxpool2x2": (C.c_int, [C.POINTER(Tensor), C.POINTER(Tensor), C.c_void_p]),
    "uegan_unpack_nchw": (C.c_int, [C.POINTER(Tensor), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "uegan_adam_step": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_void_p, C.c_void_p] + [C.c_float] * 4 + [C.c_void_p]),
    "uegan_adam_step_peers": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_void_p, C.c_void_p] + [C.c_float] * 4 + [C.c_void_p]),
    "uegan_peer_sum_f64": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_void_p]),
    "uegan_capture_status": (C.c_int, [C.c_void_p]),
    "uegan_memset_zero": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "uegan_pack_input_u8": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Tensor), C.c_void_p, C.c_int32,
                                      C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p]),
    "uegan_unpack_output_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "uegan_sse_u8": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p, C.c_void_p]),
    "uegan_ssim_u8": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 5 + [C.c_void_p, C.c_void_p]),
}

_lib = None


def load():
    """Loads the shared library (once) and declares every prototype.  Raises UeganError if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UeganError(f"{LIB_PATH} not found: the CUDA extension is not built "
                         "(run `python -m uegan_b200.build`); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.uegan_abi_version() != ABI_VERSION:
        raise UeganError(f"ABI mismatch: library {lib.uegan_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


_DEBUG_CAPTURE = os.environ.get("UEGAN_DEBUG_CAPTURE") == "1"
_last_ok = [""]


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().uegan_last_error()
        raise UeganError(f"{what}: {msg.decode() if msg else rc}")
    if _DEBUG_CAPTURE:  # pinpoints the call after which a CUDA-graph capture turned invalid
        import torch
        st = load().uegan_capture_status(torch.cuda.current_stream().cuda_stream)
        if st == 2 or st < 0:
            raise UeganError(f"capture invalid (status {st}) after `{what}` (last good call: `{_last_ok[0]}`)")
        _last_ok[0] = what


def float3(v):
    return (C.c_float * 3)(*[float(a) for a in v]) if v is not None else None
