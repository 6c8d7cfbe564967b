"""Drop-in `losses` module: PerceptualLoss (frozen VGG-19 tower), GANLoss, MultiscaleRecLoss with the reference's
constructors and call contracts (/root/reference/losses.py:12-36, 202-231, 255-411), computed by the sm_100a
kernels behind include/uegan_sm100.h.  Loss values come back as 0-d CUDA tensors, so `trainer.py:92-119`'s
arithmetic (`lambda * loss`, `+=`, `.item()`, `.backward()`) works unchanged on top."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import kernels as K

# torchvision vgg19.features[0:29]: conv index -> (cin, cout); 'M' = MaxPool2d(2,2).  Taps relu{1..5}_1 follow the
# convs at features index 0, 5, 10, 19, 28 (losses.py:62-110).
_VGG_LAYERS = [(0, 3, 64), (2, 64, 64), "M", (5, 64, 128), (7, 128, 128), "M", (10, 128, 256), (12, 256, 256),
               (14, 256, 256), (16, 256, 256), "M", (19, 256, 512), (21, 512, 512), (23, 512, 512), (25, 512, 512),
               "M", (28, 512, 512)]
_TAP_IDX = (0, 5, 10, 19, 28)
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class VGG19_relu(nn.Module):
    """Frozen VGG-19 feature tower up to relu5_1 (losses.py:39-164), fp16 storage / kind::f16 MMAs with fp32
    accumulation (10-bit operands like tf32, at the full 16-bit tensor rate; bf16 misses the 1e-3 loss tolerance).  The reference also evaluates relu5_2..5_4,
    whose outputs nothing reads (losses.py:30-34); they are not computed here."""

    def __init__(self, state_dict=None):
        super().__init__()
        import torchvision
        if state_dict is None:
            # same source as the reference (losses.py:43): the torchvision hub checkpoint
            cnn = torchvision.models.vgg19(pretrained=True)
        else:
            cnn = torchvision.models.vgg19(weights=None)
            cnn.load_state_dict(state_dict, strict=False)
        self.features = nn.Sequential(*list(cnn.features.children())[:29])  # parameter container only
        for p in self.parameters():
            p.requires_grad = False
        self._plans = {}
        self._wcache = {}

    def _packed(self, idx, cin_stored):
        key = (idx, cin_stored)
        conv = self.features[idx]
        tag = (conv.weight.data_ptr(), conv.weight._version)
        hit = self._wcache.get(key)
        if hit is None or hit[0] != tag:
            hit = (tag, K.packed_weight(conv.weight, cin_stored, L.F16))
            self._wcache[key] = hit
        return hit[1]

    def _plan(self, b, h, w, device, slot):
        key = (b, h, w, str(device), slot)
        pl = self._plans.get(key)
        if pl is None:
            # zero-initialised once: convs and pools never write the halo, so it IS the zero padding of every conv
            acts = [K.NHWC(b, h, w, 8, 1, L.F16, device, zero=True)]
            hh, ww = h, w
            for spec in _VGG_LAYERS:
                if spec == "M":
                    hh, ww = hh // 2, ww // 2
                    acts.append(K.NHWC(b, hh, ww, acts[-1].c, 1, L.F16, device, zero=True))
                else:
                    acts.append(K.NHWC(b, hh, ww, spec[2], 1, L.F16, device, zero=True))
            stats = [torch.empty(3 * b * c, dtype=torch.float64, device=device) for c in (64, 128, 256, 512, 512)]
            pl = dict(acts=acts, stats=stats)
            self._plans[key] = pl
        return pl

    @torch.no_grad()
    def run(self, x01, slot="x", eps=1e-5):
        """x01: (B,3,H,W) fp32 in [0,1].  Returns (taps, plan): taps = list of (NHWC activation, mean/rstd address)."""
        b, _, h, w = x01.shape
        if h % 16 or w % 16:
            raise ValueError("PerceptualLoss needs H, W multiples of 16")
        P = self._plan(b, h, w, x01.device, slot)
        acts = P["acts"]
        scale = [1.0 / s for s in IMAGENET_STD]
        shift = [-m / s for m, s in zip(IMAGENET_MEAN, IMAGENET_STD)]
        K.pack_input(x01.contiguous().float(), acts[0], L.PAD_ZERO, scale, shift)  # (x - mean) / std, losses.py:26-27
        taps, ti = [], 0
        for li, spec in enumerate(_VGG_LAYERS):
            src, dst = acts[li], acts[li + 1]
            if spec == "M":
                K.maxpool2x2(src, dst)
                continue
            idx, cin, cout = spec
            conv = self.features[idx]
            is_tap = idx in _TAP_IDX
            fused = is_tap and K.fused_stats_ok(dst.h, dst.w, cout)
            K.conv_fprop(src, self._packed(idx, src.c), cout, 3, 1, 1, dst, 0, conv.bias, None, L.ACT_RELU,
                         in_stats=P["stats"][ti] if fused else None)
            if is_tap:
                mr = K.instance_norm_stats(dst, P["stats"][ti], sums_ready=fused, eps=eps)
                taps.append((dst, mr))
                ti += 1
        return taps, P


class PerceptualLoss(nn.Module):
    """losses.py:12-36: sum_k w_k * MSE(IN(vgg_k(x)), IN(vgg_k(y))), w = 1/64, 1/64, 1/32, 1/32, 1; x, y in [0,1]."""

    def __init__(self, vgg_state_dict=None):
        super().__init__()
        self.add_module("vgg", VGG19_relu(vgg_state_dict))
        self.weights = [1.0 / 64, 1.0 / 64, 1.0 / 32, 1.0 / 32, 1.0 / 1]
        self.eps = 1e-5  # nn.InstanceNorm2d default (losses.py:18); exposed for conditioning tests
        self._accum = None

    def __call__(self, x, y):
        if not (x.is_cuda and y.is_cuda):
            raise L.UeganError("uegan_b200.losses.PerceptualLoss runs on CUDA (sm_100a) only; no CPU fallback")
        if x.shape[1] != 3:
            x, y = x.repeat(1, 3, 1, 1), y.repeat(1, 3, 1, 1)
        if self.vgg.features[0].weight.device != x.device:
            self.vgg.to(x.device)
        if torch.is_grad_enabled() and (x.requires_grad or y.requires_grad):
            from .autograd import perceptual_apply
            return perceptual_apply(self, x, y)
        return self.forward_native(x, y)

    @torch.no_grad()
    def forward_native(self, x, y):
        taps_x, _ = self.vgg.run(x, "x", self.eps)
        taps_y, _ = self.vgg.run(y, "y", self.eps)
        loss = torch.zeros(1, dtype=torch.float32, device=x.device)
        if self._accum is None or self._accum.device != x.device:
            self._accum = torch.zeros(1, dtype=torch.float64, device=x.device)
        for wgt, (tx, mx), (ty, my) in zip(self.weights, taps_x, taps_y):
            K.in_mse_fwd(tx, ty, mx, my, wgt, self._accum, loss)
        return loss[0]


class GANLoss(nn.Module):
    """losses.py:255-411.  Only the relativistic modes are reachable from trainer.py:92-104 (target_is_real=None);
    the other modes raise exactly where the reference does."""

    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0, tensor=torch.FloatTensor, opt=None):
        super().__init__()
        if gan_mode not in ("ls", "original", "w", "hinge", "rahinge", "rals"):
            raise ValueError("Unexpected gan_mode {}".format(gan_mode))
        self.gan_mode = gan_mode
        self.real_label, self.fake_label, self.Tensor, self.opt = target_real_label, target_fake_label, tensor, opt
        self.process_group = None  # set for data-parallel training: means / normalisation over the global batch

    def __call__(self, real_preds, fake_preds, target_is_real, for_real=None, for_fake=None, for_discriminator=True):
        if self.gan_mode not in K.GAN_MODES:
            raise NotImplementedError("nither for real_preds nor for fake_preds")  # losses.py:311/320/347/391
        if not isinstance(real_preds, (list, tuple)):
            real_preds, fake_preds = [real_preds], [fake_preds]
        real = [p[-1] if isinstance(p, (list, tuple)) else p for p in real_preds]
        fake = [p[-1] if isinstance(p, (list, tuple)) else p for p in fake_preds]
        if not all(t.is_cuda for t in real + fake):
            raise L.UeganError("uegan_b200.losses.GANLoss runs on CUDA (sm_100a) only; no CPU fallback")
        mode = K.GAN_MODES[self.gan_mode]
        if torch.is_grad_enabled() and any(t.requires_grad for t in real + fake):
            from .autograd import gan_loss_apply
            return gan_loss_apply(mode, bool(for_discriminator), real, fake, self.process_group)
        real = [t.detach().contiguous().float() for t in real]
        fake = [t.detach().contiguous().float() for t in fake]
        ws = torch.empty(48, dtype=torch.float64, device=real[0].device)
        loss = torch.empty(1, dtype=torch.float32, device=real[0].device)
        K.gan_loss_fwd(mode, bool(for_discriminator), real, fake, ws, loss, self.process_group)
        return loss[0]


class MultiscaleRecLoss(nn.Module):
    """losses.py:202-231: sum_i w_i * criterion(pool^i(pred), pool^i(gt)), w = 1, 1/2, 1/4, AvgPool2d(2,2)."""

    def __init__(self, scale=3, rec_loss_type="l1", multiscale=True):
        super().__init__()
        if rec_loss_type not in K.REC_TYPES:
            raise NotImplementedError("Loss [{}] is not implemented".format(rec_loss_type))
        self.rec_type = K.REC_TYPES[rec_loss_type]
        self.multiscale = multiscale
        self.scales = min(int(scale), 3) if multiscale else 1
        if multiscale:
            self.weights = [1.0, 1.0 / 2, 1.0 / 4][:scale]

    def forward(self, input, target):
        if not (input.is_cuda and target.is_cuda):
            raise L.UeganError("uegan_b200.losses.MultiscaleRecLoss runs on CUDA (sm_100a) only; no CPU fallback")
        if torch.is_grad_enabled() and input.requires_grad:
            from .autograd import msrec_apply
            return msrec_apply(self, input, target)
        pred, gt = input.detach().contiguous().float(), target.detach().contiguous().float()
        accum = torch.empty(3, dtype=torch.float64, device=pred.device)
        loss = torch.empty(1, dtype=torch.float32, device=pred.device)
        K.msrec_loss(pred, gt, self.rec_type, self.scales, accum, loss)
        return loss[0]
