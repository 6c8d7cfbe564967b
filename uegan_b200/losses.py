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
    accumulation (10-bit operands like tf32, at the full 16-bit tensor rate; bf16 misses the 1e-3 loss tolerance).  The
    reference also evaluates relu5_2..5_4, whose outputs nothing reads (losses.py:30-34); they are not computed here.

    Range: fp16 has 11 significant bits but only 2^-14 .. 2^16 of normal range, and the magnitudes inside a VGG depend
    on the checkpoint.  The tower is frozen and ReLU / max-pool are positively homogeneous, so every conv carries a
    POWER-OF-TWO factor f_l folded into its packed weight and bias: stored activation = s_l * true activation with
    s_l = f_1 ... f_l, chosen once (`calibrate`, first batch) so that each layer's largest stored value sits near 2^5.
    Powers of two commute with fp16 / fp32 rounding, so stored values carry exactly the bits the unscaled fp32-range
    computation would; InstanceNorm + MSE at the taps is scale-invariant except for eps, which becomes eps * s_l^2.
    `check_range` is the guard: it raises when a tap left the safe window instead of letting fp16 saturate silently."""

    TARGET_AMAX = 32.0

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
        self._calib = None  # {features idx: (tag, factor f_l, cumulative scale s_l, f_l * weight, s_l * bias)}

    # ---------------------------------------------------------------- power-of-two range folding
    def _tag(self):
        return tuple((self.features[spec[0]].weight.data_ptr(), self.features[spec[0]].weight._version)
                     for spec in _VGG_LAYERS if spec != "M")

    def layer(self, idx):
        """(effective fp32 weight, effective bias, cumulative scale) of the conv at features[idx]."""
        c = self._calib[idx]
        return c[2], c[3], c[1]

    def tap_scale(self, ti):
        return self._calib[_TAP_IDX[ti]][1]

    @torch.no_grad()
    def calibrate(self, x01):
        """Chooses the per-layer power-of-two factors on the batch x01 (layer by layer: run, read the largest stored
        value back, adjust, re-run).  One-off host-synchronising work; the tower is frozen afterwards."""
        import math
        b, _, h, w = x01.shape
        P = self._plan(b, h, w, x01.device, "x")
        acts = P["acts"]
        scale = [1.0 / s for s in IMAGENET_STD]
        shift = [-m / s for m, s in zip(IMAGENET_MEAN, IMAGENET_STD)]
        K.pack_input(x01.contiguous().float(), acts[0], L.PAD_ZERO, scale, shift)
        calib, s_prev = {}, 1.0
        for li, spec in enumerate(_VGG_LAYERS):
            src, dst = acts[li], acts[li + 1]
            if spec == "M":
                K.maxpool2x2(src, dst)
                continue
            idx, cin, cout = spec
            conv = self.features[idx]
            f = 1.0
            for attempt in range(24):
                w_eff = (conv.weight.detach() * f).contiguous()
                b_eff = (conv.bias.detach() * (s_prev * f)).contiguous()
                wp = K.packed_weight(w_eff, src.c, L.F16)
                K.conv_fprop(src, wp, cout, 3, 1, 1, dst, 0, b_eff, None, L.ACT_RELU)
                amax = float(dst.padded_view().float().abs().max())
                if not math.isfinite(amax) or amax > 2.0 ** 15:
                    f *= 2.0 ** -8
                elif amax == 0.0:
                    if attempt >= 6:  # a dead layer: nothing to scale
                        break
                    f *= 2.0 ** 8
                else:
                    g = 2.0 ** round(math.log2(self.TARGET_AMAX / amax))
                    if g == 1.0:
                        break
                    f *= g
            else:
                raise L.UeganError(f"VGG19_relu.calibrate: features[{idx}] does not settle in fp16 range")
            calib[idx] = (None, s_prev * f, w_eff, b_eff, f)
            s_prev *= f
        self._calib, self._calib_tag = calib, self._tag()
        self._wcache.clear()
        for pl in self._plans.values():
            pl.pop("grads", None)

    @torch.no_grad()
    def check_range(self, taps):
        """Raises if a tap's largest stored value left [2^-6, 2^15): the amax guard against silent fp16 saturation /
        underflow (host-synchronising; call it when a loss looks suspicious or periodically, not every step)."""
        import math
        for ti, (t, _) in enumerate(taps):
            amax = float(t.padded_view().float().abs().max())
            if not math.isfinite(amax) or amax >= 2.0 ** 15 or (0.0 < amax < 2.0 ** -6):
                raise L.UeganError(f"VGG19_relu: tap relu{ti + 1}_1 has |max| = {amax:g} in fp16 storage "
                                   f"(scale 2^{math.log2(self.tap_scale(ti)):.0f}); call calibrate() on a representative batch")

    def _packed(self, idx, cin_stored):
        key = (idx, cin_stored)
        hit = self._wcache.get(key)
        if hit is None:
            hit = K.packed_weight(self.layer(idx)[0], cin_stored, L.F16)
            self._wcache[key] = hit
        return hit

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
    def run(self, x01, slot="x", eps=1e-5, stats=True):
        """x01: (B,3,H,W) fp32 in [0,1].  Returns (taps, plan): taps = list of (NHWC activation, mean/rstd address).
        stats=False: no InstanceNorm statistics pass (address None): the caller measures both towers' taps jointly
        (kernels.in_mse_joint)."""
        b, _, h, w = x01.shape
        if h % 16 or w % 16:
            raise ValueError("PerceptualLoss needs H, W multiples of 16")
        if self._calib is None or self._calib_tag != self._tag():
            self.calibrate(x01)
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
            _, b_eff, s_l = self.layer(idx)
            is_tap = idx in _TAP_IDX
            fused = is_tap and stats and K.fused_stats_ok(dst.h, dst.w, cout)
            K.conv_fprop(src, self._packed(idx, src.c), cout, 3, 1, 1, dst, 0, b_eff, None, L.ACT_RELU,
                         in_stats=P["stats"][ti] if fused else None)
            if is_tap:
                # InstanceNorm of the TRUE activation: eps scales with the square of the stored scale
                mr = K.instance_norm_stats(dst, P["stats"][ti], sums_ready=fused, eps=eps * s_l * s_l) if stats else None
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
        if torch.is_grad_enabled() and y.requires_grad:
            # the trainer compares fake_exp with the DETACHED input (trainer.py:108); a gradient w.r.t. the second
            # argument is not implemented natively and must not be dropped silently
            raise NotImplementedError("uegan_b200.losses.PerceptualLoss: the second argument must not require grad")
        if torch.is_grad_enabled() and x.requires_grad:
            from .autograd import perceptual_apply
            return perceptual_apply(self, x, y)
        return self.forward_native(x, y)

    @staticmethod
    def joint_taps():
        """One joint pass per tap over both towers' feature maps (uegan_in_mse_joint) instead of separate statistics,
        MSE and backward-statistics passes; UEGAN_NO_JOINT_TAP=1 selects the separate passes."""
        import os
        return os.environ.get("UEGAN_NO_JOINT_TAP") != "1"

    def tap_terms(self, taps_x, taps_y, Px, loss):
        """loss += sum_k w_k MSE(IN(x_k), IN(y_k)) from the joint pass; returns per tap (mean/rstd x, mean/rstd y, sums)."""
        dev = loss.device
        if self._accum is None or self._accum.device != dev:
            self._accum = torch.zeros(1, dtype=torch.float64, device=dev)
        if "joint_ws" not in Px:
            Px["joint_ws"] = [torch.empty(9 * t.n * t.c, dtype=torch.float64, device=dev) for t, _ in taps_x]
        out = []
        for ti, (wgt, (tx, _), (ty, _)) in enumerate(zip(self.weights, taps_x, taps_y)):
            s_l = self.vgg.tap_scale(ti)
            out.append(K.in_mse_joint(tx, ty, self.eps * s_l * s_l, wgt, Px["joint_ws"][ti], self._accum, loss))
        return out

    @torch.no_grad()
    def forward_native(self, x, y):
        joint = self.joint_taps()
        taps_x, Px = self.vgg.run(x, "x", self.eps, stats=not joint)
        taps_y, _ = self.vgg.run(y, "y", self.eps, stats=not joint)
        loss = torch.zeros(1, dtype=torch.float32, device=x.device)
        if joint:
            self.tap_terms(taps_x, taps_y, Px, loss)
            return loss[0]
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
        if torch.is_grad_enabled() and target.requires_grad:
            raise NotImplementedError("uegan_b200.losses.MultiscaleRecLoss: the target must not require grad")
        if torch.is_grad_enabled() and input.requires_grad:
            from .autograd import msrec_apply
            return msrec_apply(self, input, target)
        pred, gt = input.detach().contiguous().float(), target.detach().contiguous().float()
        accum = torch.empty(3, dtype=torch.float64, device=pred.device)
        loss = torch.empty(1, dtype=torch.float32, device=pred.device)
        K.msrec_loss(pred, gt, self.rec_type, self.scales, accum, loss)
        return loss[0]


class TVLoss(nn.Module):
    """losses.py:167-184.  Imported by tester.py:9 but called nowhere in the reference (not on the hot path);
    kept so that `from losses import PerceptualLoss, TVLoss` resolves.  Plain tensor arithmetic on the caller's device."""

    def __init__(self, tv_loss_weight=1):
        super().__init__()
        self.tv_loss_weight = tv_loss_weight

    def forward(self, x):
        b, c, h, w = x.shape
        h_tv = (x[:, :, 1:, :] - x[:, :, :h - 1, :]).pow(2).sum()
        w_tv = (x[:, :, :, 1:] - x[:, :, :, :w - 1]).pow(2).sum()
        return self.tv_loss_weight * 2 * (h_tv / (c * (h - 1) * w) + w_tv / (c * h * (w - 1))) / b
