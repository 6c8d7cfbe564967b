"""CPU oracle for the UEGAN hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This file is a from-scratch fp32 CPU restatement (torch.nn.functional on CPU
tensors, explicit weight dictionaries keyed by the reference's state_dict
names) of the reference algorithm on the hot path:

    Generator forward            /root/reference/models.py:44-74
    GAM                          /root/reference/models.py:215-237
    Discriminator forward        /root/reference/models.py:139-155
    spectral norm (train mode)   torch/nn/utils/spectral_norm.py (hook semantics, pinned by
                                 models.py:185-188; 1 power iteration, dim=0, eps=1e-12)
    VGG19 taps + PerceptualLoss  /root/reference/losses.py:12-36, 39-164
    GANLoss (rahinge / rals)     /root/reference/losses.py:348-377, 393-409
    MultiscaleRecLoss            /root/reference/losses.py:202-231
    Trainer step                 /root/reference/trainer.py:75-119, Adam at trainer.py:337-338

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import this module.  The product path
(`uegan_b200/`) never does; it fails loudly when the CUDA library is missing.

Parity pin: the reference ships no tests / golden vectors, so
the oracle is pinned against outputs of the reference itself: `tests/golden/
make_golden.py` imports /root/reference/{models,losses}.py in the build
container, runs them on seeded inputs/weights, and commits the outputs as
fixtures; `tests/test_oracle.py` checks this restatement against those fixtures
(and, when /root/reference is present, against the live reference modules).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

# ----------------------------------------------------------------------------------------------
# deterministic synthetic weights (numpy PCG64 -> identical in every container, no torch RNG)
# ----------------------------------------------------------------------------------------------

G_CONVS = [  # (state_dict prefix, cin, cout, k)   models.py:16-42
    ("enc1.main.1", 3, 1, 7), ("enc2.main.1", 1, 2, 3), ("enc3.main.1", 2, 4, 3),
    ("enc4.main.1", 4, 8, 3), ("enc5.main.1", 8, 16, 3),
    ("upsample1.1.main.1", 16, 8, 1), ("upsample2.1.main.1", 8, 4, 1),
    ("upsample3.1.main.1", 4, 2, 1), ("upsample4.1.main.1", 2, 1, 1),
    ("dec1.main.1", 16, 8, 3), ("dec2.main.1", 8, 4, 3), ("dec3.main.1", 4, 2, 3),
    ("dec4.main.1", 2, 1, 3), ("dec5.0.main.1", 1, 1, 3), ("dec5.1.main.1", 1, -3, 7),
]
D_CONVS = [  # (k, cin_mult, cout_mult, kernel)   models.py:109-126
    (1, 3, 1, 7), (2, 1, 2, 7), (3, 2, 4, 7), (4, 4, 8, 5), (5, 8, 16, 5),
]
VGG_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M",
           512, 512, 512, 512]  # torchvision vgg19 features[0:36] (losses.py:43-116)
VGG_TAPS = {"relu1_1": 1, "relu2_1": 6, "relu3_1": 11, "relu4_1": 20, "relu5_1": 29}


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def _normal(rng, shape, std) -> Tensor:
    return torch.from_numpy((rng.standard_normal(shape) * std).astype(np.float32))


def make_generator_params(conv_dim: int = 32, seed: int = 0, regime: str = "o1") -> Params:
    """Synthetic Generator weights with the reference's state_dict keys.

    regime "o1": fan-in scaled normal weights (activations stay O(1));
    regime "tiny": N(0, 0.02/sqrt(fan_in))-like weights mimicking init_weights('orthogonal',0.02)
    (trainer.py:330-332) where activations decay to ~1e-10.
    """
    rng = _rng(seed)
    p: Params = {}
    gain = 1.0 if regime == "o1" else 0.02

    def conv(name, cin, cout, k, bias=True, g=1.0):
        fan_in = cin * k * k
        p[name + ".weight"] = _normal(rng, (cout, cin, k, k), g * gain * math.sqrt(2.0 / fan_in))
        if bias:
            p[name + ".bias"] = _normal(rng, (cout,), 0.05 if regime == "o1" else 0.0)

    for name, ci, co, k in G_CONVS:
        cin = 3 if (name == "enc1.main.1") else ci * conv_dim
        cout = 3 if co < 0 else co * conv_dim
        g = 0.3 if name.startswith("dec5.1") else 1.0  # keep tanh away from saturation
        conv(name, cin, cout, k, g=g)
    for i, mult in zip(range(1, 6), (1, 2, 4, 8, 16)):
        c = mult * conv_dim
        conv(f"ga{i}.conv.0", 2 * c, c // 8, 1, bias=False)
        conv(f"ga{i}.conv.2", c // 8, c, 1, bias=False)
        conv(f"ga{i}.fuse.0", 2 * c, c, 1, bias=True)
    return p


def make_discriminator_params(conv_dim: int = 32, seed: int = 1, regime: str = "o1") -> Params:
    """Synthetic Discriminator weights + spectral-norm buffers, reference keys."""
    rng = _rng(seed)
    p: Params = {}
    gain = 1.0 if regime == "o1" else 0.02
    for k, ci, co, ks in D_CONVS:
        cin = 3 if k == 1 else ci * conv_dim
        cout = co * conv_dim
        fan_in = cin * ks * ks
        p[f"d{k}.0.1.weight_orig"] = _normal(rng, (cout, cin, ks, ks), gain * math.sqrt(2.0 / fan_in))
        p[f"d{k}.0.1.bias"] = _normal(rng, (cout,), 0.05 if regime == "o1" else 0.0)
        u = _normal(rng, (cout,), 1.0)
        v = _normal(rng, (fan_in,), 1.0)
        p[f"d{k}.0.1.weight_u"] = u / u.norm().clamp_min(1e-12)
        p[f"d{k}.0.1.weight_v"] = v / v.norm().clamp_min(1e-12)
        p[f"d{k}_pred.0.1.weight"] = _normal(rng, (1, cout, ks, ks), gain * math.sqrt(1.0 / (cout * ks * ks)))
    return p


def make_vgg_params(seed: int = 2) -> Params:
    """Deterministic synthetic VGG-19 `features` weights (He-normal so activations stay O(1)).

    The ImageNet checkpoint losses.py:43 downloads is not available offline; both
    the oracle and the CUDA path load this same synthetic tower, keys `features.{idx}.{weight,bias}`.
    """
    rng = _rng(seed)
    p: Params = {}
    cin, idx = 3, 0
    for v in VGG_CFG:
        if v == "M":
            idx += 1
            continue
        p[f"features.{idx}.weight"] = _normal(rng, (v, cin, 3, 3), math.sqrt(2.0 / (cin * 9)))
        p[f"features.{idx}.bias"] = _normal(rng, (v,), 0.02)
        cin = v
        idx += 2
    return p


def make_images(shape: Sequence[int], seed: int) -> Tensor:
    """Synthetic images in [-1,1]: smooth low-frequency field + noise (so D/VGG see structure)."""
    rng = _rng(seed)
    b, c, h, w = shape
    coarse = torch.from_numpy(rng.uniform(-1, 1, (b, c, max(h // 16, 2), max(w // 16, 2))).astype(np.float32))
    img = F.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=False)
    img = 0.8 * img + 0.2 * torch.from_numpy(rng.uniform(-1, 1, (b, c, h, w)).astype(np.float32))
    return img.clamp(-1, 1).contiguous()


# ----------------------------------------------------------------------------------------------
# Generator (models.py:10-74)
# ----------------------------------------------------------------------------------------------

def _rconv(x: Tensor, w: Tensor, b: Optional[Tensor], stride: int = 1) -> Tensor:
    """ReflectionPad2d((k-1)//2) then Conv2d(padding=0)   (models.py:77-101)."""
    pad = (w.shape[-1] - 1) // 2
    if pad:
        x = F.pad(x, (pad, pad, pad, pad), mode="reflect")
    return F.conv2d(x, w, b, stride=stride)


def _act(x: Tensor, act_fun: str) -> Tensor:
    """get_act_fun (models.py:249-264)."""
    if act_fun == "LeakyReLU":
        return F.leaky_relu(x, 0.2)
    if act_fun == "ReLU":
        return F.relu(x)
    if act_fun == "Swish":
        return x * torch.sigmoid(x)
    if act_fun == "SELU":
        return F.selu(x)
    if act_fun == "none":
        return x
    raise NotImplementedError("activation function [%s] is not found" % act_fun)


def calc_mean_std(feat: Tensor, eps: float = 1e-5) -> Tuple[Tensor, Tensor]:
    """models.py:204-212 -- unbiased variance + eps."""
    n, c = feat.shape[:2]
    var = feat.reshape(n, c, -1).var(dim=2) + eps
    return feat.reshape(n, c, -1).mean(dim=2).view(n, c, 1, 1), var.sqrt().view(n, c, 1, 1)


def instance_norm(x: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.InstanceNorm2d(affine=False, track_running_stats=False): biased variance (models.py:227)."""
    mu = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps)


def gam_forward(p: Params, pre: str, x: Tensor) -> Tensor:
    """GAM.forward, full form (models.py:230-237), norm=True as built at models.py:38-42."""
    m, s = calc_mean_std(x)
    a = F.conv2d(torch.cat([m, s], dim=1), p[pre + ".conv.0.weight"])
    a = F.conv2d(F.relu(a), p[pre + ".conv.2.weight"])
    out = F.conv2d(torch.cat([x, a.expand_as(x)], dim=1), p[pre + ".fuse.0.weight"], p[pre + ".fuse.0.bias"])
    return instance_norm(out)


def gam_forward_simplified(p: Params, pre: str, x: Tensor) -> Tensor:
    """Algebraically identical form used by the CUDA path (SURVEY.md 8a "verified rewrites" (1)):
    the attention branch and the fuse bias add a per-(n,c) constant which the InstanceNorm
    that follows subtracts, so GAM(x) == IN(conv1x1(x, fuse.weight[:, :C]))."""
    c = x.shape[1]
    return instance_norm(F.conv2d(x, p[pre + ".fuse.0.weight"][:, :c]))


def generator_forward(p: Params, x: Tensor, act_fun: str = "LeakyReLU", simplified: bool = False,
                      return_all: bool = False):
    """Generator.forward (models.py:44-74) with norm_fun='none', use_sn=False (config.py:23-27)."""
    gam = gam_forward_simplified if simplified else gam_forward

    def block(name, t, stride=1):
        return _act(_rconv(t, p[name + ".weight"], p[name + ".bias"], stride), act_fun)

    def up(name, t):
        if simplified:  # conv1x1 and bilinear x2 commute (SURVEY.md 8a rewrite (2))
            t = F.conv2d(t, p[name + ".weight"], p[name + ".bias"])
            return F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)
        t = F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)
        return F.conv2d(t, p[name + ".weight"], p[name + ".bias"])

    x1 = block("enc1.main.1", x)
    x2 = block("enc2.main.1", x1, 2)
    x3 = block("enc3.main.1", x2, 2)
    x4 = block("enc4.main.1", x3, 2)
    x5 = block("enc5.main.1", x4, 2)
    x5 = gam(p, "ga5", x5)
    y1 = block("dec1.main.1", torch.cat([up("upsample1.1.main.1", x5), gam(p, "ga4", x4)], 1))
    y2 = block("dec2.main.1", torch.cat([up("upsample2.1.main.1", y1), gam(p, "ga3", x3)], 1))
    y3 = block("dec3.main.1", torch.cat([up("upsample3.1.main.1", y2), gam(p, "ga2", x2)], 1))
    y4 = block("dec4.main.1", torch.cat([up("upsample4.1.main.1", y3), gam(p, "ga1", x1)], 1))
    t = _rconv(y4 * x1, p["dec5.0.main.1.weight"], p["dec5.0.main.1.bias"])
    res = torch.tanh(_rconv(t, p["dec5.1.main.1.weight"], p["dec5.1.main.1.bias"]))
    out = torch.clamp(res + x, -1.0, 1.0)
    if return_all:
        return out, dict(x1=x1, x2=x2, x3=x3, x4=x4, x5=x5, y1=y1, y2=y2, y3=y3, y4=y4, t=t, res=res)
    return out


# ----------------------------------------------------------------------------------------------
# Discriminator (models.py:104-182) + spectral norm
# ----------------------------------------------------------------------------------------------

def _normalize(v: Tensor, eps: float = 1e-12) -> Tensor:
    return v / v.norm().clamp_min(eps)


def spectral_norm_weight(w_orig: Tensor, u: Tensor, v: Tensor, training: bool = True):
    """torch.nn.utils.spectral_norm hook, n_power_iterations=1, dim=0, eps=1e-12.
    Returns (W_sn, u_new, v_new, sigma).  In training mode u, v advance by one power iteration under
    no_grad and sigma = u.(W v) carries autograd through W only (u, v are treated as constants)."""
    wm = w_orig.reshape(w_orig.shape[0], -1)
    if training:
        with torch.no_grad():
            v = _normalize(torch.mv(wm.t(), u))
            u = _normalize(torch.mv(wm, v))
    sigma = torch.dot(u, torch.mv(wm, v))
    return w_orig / sigma, u, v, sigma


def discriminator_forward(p: Params, x: Tensor, training: bool = True, act_fun: str = "LeakyReLU",
                          adv_loss_type: str = "rahinge", update_buffers: bool = True) -> List[Tensor]:
    """Discriminator.forward (models.py:139-155); updates weight_u / weight_v in `p` in place when
    training (as the reference's forward-pre-hook does on every train-mode forward)."""
    if adv_loss_type in ("ls", "rals"):
        head = torch.sigmoid
    elif adv_loss_type in ("hinge", "rahinge"):
        head = torch.tanh
    else:
        raise NotImplementedError("Adversarial loss [{}] is not found".format(adv_loss_type))
    preds = []
    h = x
    for k in range(1, 6):
        w, u, v, _ = spectral_norm_weight(p[f"d{k}.0.1.weight_orig"], p[f"d{k}.0.1.weight_u"],
                                          p[f"d{k}.0.1.weight_v"], training)
        if training and update_buffers:
            p[f"d{k}.0.1.weight_u"], p[f"d{k}.0.1.weight_v"] = u.detach(), v.detach()
        h = _act(_rconv(h, w, p[f"d{k}.0.1.bias"], 2), act_fun)
        preds.append(head(_rconv(h, p[f"d{k}_pred.0.1.weight"], None)))
    return preds


# ----------------------------------------------------------------------------------------------
# VGG-19 tower + PerceptualLoss (losses.py:12-36, 39-164)
# ----------------------------------------------------------------------------------------------

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def vgg19_taps(vp: Params, x: Tensor, last: str = "relu5_1") -> Dict[str, Tensor]:
    """relu{1..5}_1 of torchvision vgg19.features (losses.py:120-140).  The reference also runs
    relu5_2..relu5_4, whose outputs nothing reads (losses.py:30-34) -- not restated."""
    taps: Dict[str, Tensor] = {}
    inv = {v: k for k, v in VGG_TAPS.items()}
    idx, h = 0, x
    for v in VGG_CFG:
        if v == "M":
            h = F.max_pool2d(h, 2, 2)
            idx += 1
            continue
        h = F.relu(F.conv2d(h, vp[f"features.{idx}.weight"], vp[f"features.{idx}.bias"], padding=1))
        if idx + 1 in inv:
            taps[inv[idx + 1]] = h
            if inv[idx + 1] == last:
                break
        idx += 2
    return taps


def perceptual_loss(vp: Params, x: Tensor, y: Tensor, eps: float = 1e-5) -> Tensor:
    """PerceptualLoss.__call__ (losses.py:22-36); x, y in [0,1]."""
    if x.shape[1] != 3:
        x, y = x.repeat(1, 3, 1, 1), y.repeat(1, 3, 1, 1)
    mean = torch.tensor(IMAGENET_MEAN).view(1, -1, 1, 1).to(x)
    std = torch.tensor(IMAGENET_STD).view(1, -1, 1, 1).to(x)
    xv, yv = vgg19_taps(vp, (x - mean) / std), vgg19_taps(vp, (y - mean) / std)
    weights = [1.0 / 64, 1.0 / 64, 1.0 / 32, 1.0 / 32, 1.0]
    loss = 0
    for wgt, k in zip(weights, ("relu1_1", "relu2_1", "relu3_1", "relu4_1", "relu5_1")):
        loss = loss + wgt * F.mse_loss(instance_norm(xv[k], eps), instance_norm(yv[k], eps))
    return loss


# ----------------------------------------------------------------------------------------------
# GANLoss / MultiscaleRecLoss (losses.py:202-231, 348-377, 393-409)
# ----------------------------------------------------------------------------------------------

def gan_loss(gan_mode: str, real_preds: List[Tensor], fake_preds: List[Tensor], for_discriminator: bool) -> Tensor:
    """Sum over scales of the relativistic average loss.  Means are over the WHOLE (B,1,h,w) tensor."""
    if gan_mode not in ("rahinge", "rals"):
        raise NotImplementedError("only rahinge / rals are reachable from trainer.py:92-104")
    loss = 0
    for r, f in zip(real_preds, fake_preds):
        rf = r - f.mean()
        fr = f - r.mean()
        if gan_mode == "rahinge":
            if for_discriminator:
                l = F.relu(1 - rf).mean() + F.relu(1 + fr).mean()
            else:
                l = F.relu(1 + rf).mean() + F.relu(1 - fr).mean()
        else:
            if for_discriminator:
                l = ((rf - 1) ** 2).mean() + ((fr + 1) ** 2).mean()
            else:
                l = ((rf + 1) ** 2).mean() + ((fr - 1) ** 2).mean()
        loss = loss + l / 2
    return loss


def multiscale_rec_loss(pred: Tensor, gt: Tensor, scale: int = 3, rec_loss_type: str = "l1") -> Tensor:
    """MultiscaleRecLoss.forward (losses.py:219-231), weights 1, 1/2, 1/4, AvgPool2d(2) between scales."""
    if rec_loss_type == "l1":
        crit = F.l1_loss
    elif rec_loss_type == "smoothl1":
        crit = F.smooth_l1_loss
    elif rec_loss_type == "l2":
        crit = F.mse_loss
    else:
        raise NotImplementedError("Loss [{}] is not implemented".format(rec_loss_type))
    weights = [1.0, 0.5, 0.25][:scale]
    loss = 0
    for i, w in enumerate(weights):
        loss = loss + w * crit(pred, gt)
        if i != len(weights) - 1:
            pred, gt = F.avg_pool2d(pred, 2, 2), F.avg_pool2d(gt, 2, 2)
    return loss


# ----------------------------------------------------------------------------------------------
# Adam (torch.optim.Adam with coupled weight decay, trainer.py:337-338) and the step (trainer.py:75-119)
# ----------------------------------------------------------------------------------------------

class AdamState:
    def __init__(self, names: Sequence[str]):
        self.t = 0
        self.m = {k: None for k in names}
        self.v = {k: None for k in names}


def adam_step(p: Params, grads: Params, st: AdamState, lr: float, beta1: float = 0.5, beta2: float = 0.999,
              eps: float = 1e-8, weight_decay: float = 1e-4) -> None:
    st.t += 1
    bc1, bc2 = 1 - beta1 ** st.t, 1 - beta2 ** st.t
    for k, g in grads.items():
        if g is None:
            continue
        g = g + weight_decay * p[k]
        st.m[k] = (1 - beta1) * g if st.m[k] is None else beta1 * st.m[k] + (1 - beta1) * g
        st.v[k] = (1 - beta2) * g * g if st.v[k] is None else beta2 * st.v[k] + (1 - beta2) * g * g
        denom = st.v[k].sqrt() / math.sqrt(bc2) + eps
        p[k] = p[k] - (lr / bc1) * st.m[k] / denom


G_TRAINABLE_SUFFIX = (".weight", ".bias")


def _trainable(p: Params) -> List[str]:
    return [k for k in p if not (k.endswith("weight_u") or k.endswith("weight_v"))]


def train_step(gp: Params, dp: Params, vp: Params, g_opt: AdamState, d_opt: AdamState, real_raw: Tensor,
               real_exp: Tensor, g_lr: float = 1e-4, d_lr: float = 4e-4, lambda_adv: float = 0.10,
               lambda_percep: float = 1.0, lambda_idt: float = 0.10, adv_input: bool = True,
               gan_mode: str = "rahinge", idt_loss_type: str = "l1") -> Dict[str, float]:
    """One iteration of the loop body trainer.py:75-119 with pool_size=0 (ImagePool.query is then the
    identity, utils.py:30-32).  Mutates gp / dp (weights and SN buffers) and the Adam states in place."""
    gk, dk = _trainable(gp), _trainable(dp)
    for k in gk:
        gp[k] = gp[k].detach().requires_grad_(True)
    for k in dk:
        dp[k] = dp[k].detach().requires_grad_(True)
    fake_exp = generator_forward(gp, real_raw)                                # trainer.py:85
    # ---- update D (trainer.py:89-98)
    real_preds = discriminator_forward(dp, real_exp)
    fake_preds = discriminator_forward(dp, fake_exp.detach())
    d_loss = gan_loss(gan_mode, real_preds, fake_preds, True)
    if adv_input:
        input_preds = discriminator_forward(dp, real_raw)
        d_loss = d_loss + gan_loss(gan_mode, real_preds, input_preds, True)
    dgrads = dict(zip(dk, torch.autograd.grad(d_loss, [dp[k] for k in dk], allow_unused=True)))
    with torch.no_grad():
        tmp = {k: dp[k].detach() for k in dk}
        adam_step(tmp, dgrads, d_opt, d_lr)
    for k in dk:
        dp[k] = tmp[k].detach().requires_grad_(True)
    # ---- update G (trainer.py:101-119)
    real_preds = discriminator_forward(dp, real_exp)
    fake_preds = discriminator_forward(dp, fake_exp)
    g_adv = lambda_adv * gan_loss(gan_mode, real_preds, fake_preds, False)
    g_percep = lambda_percep * perceptual_loss(vp, (fake_exp + 1.) / 2., (real_raw + 1.) / 2.)
    real_exp_idt = generator_forward(gp, real_exp)
    g_idt = lambda_idt * multiscale_rec_loss(real_exp_idt, real_exp, 3, idt_loss_type)
    g_loss = g_adv + g_percep + g_idt
    ggrads = dict(zip(gk, torch.autograd.grad(g_loss, [gp[k] for k in gk], allow_unused=True)))
    with torch.no_grad():
        tmp = {k: gp[k].detach() for k in gk}
        adam_step(tmp, ggrads, g_opt, g_lr)
    for k in gk:
        gp[k] = tmp[k].detach()
    for k in dk:
        dp[k] = dp[k].detach()
    return dict(d_loss=float(d_loss.detach()), g_adv_loss=float(g_adv.detach()), g_percep_loss=float(g_percep.detach()),
                g_idt_loss=float(g_idt.detach()), g_loss=float(g_loss.detach()))


def rel_err(a: Tensor, b: Tensor, floor: float = 1e-6) -> float:
    """max |a-b| / max(|b|_max, floor): the "relative" of north_star's 1e-3 (scale of the tensor)."""
    return float((a.double() - b.double()).abs().max() / max(float(b.abs().max()), floor))


# ----------------------------------------------------------------------------------------------
# Algebra of the tiny-Cout kernels (csrc/conv_rowsum.cu, conv_wgrad.cu vertical mode), restated on the CPU so that the
# transformation itself is pinned against F.conv2d / autograd independently of the GPU (tests/test_oracle.py).
# ----------------------------------------------------------------------------------------------
def conv_rowsum_restatement(xpad: Tensor, w: Tensor) -> Tensor:
    """Valid conv of xpad (N,C,Hp,Wp) with w (O,C,k,k) computed the way conv_rowsum_kernel does: horizontal taps as GEMM
    columns, D[n,(s,o),i,q] = sum_{r,c} xpad[n,c,i+r,q] * w[o,c,r,s]  (a k x 1 conv with k*O outputs), then the shifted sum
    y[n,o,i,x] = sum_s D[n,(s,o),i,x+s]  (the epilogue's __shfl_down)."""
    o, c, k, _ = w.shape
    wv = w.permute(3, 0, 1, 2).reshape(k * o, c, k, 1)            # rows (s, o); only the vertical taps remain
    d = F.conv2d(xpad, wv)                                        # (N, k*O, Hp-k+1, Wp)
    wo = xpad.shape[3] - k + 1
    return sum(d[:, s * o:(s + 1) * o, :, s:s + wo] for s in range(k))


def dz_hstack_restatement(dz: Tensor, k: int) -> Tensor:
    """uegan_dz_hstack: E[n,(s,o),y,q] = dz[n,o,y,q-s], q in [0, W+k-1) (zero outside the image)."""
    n, o, h, w = dz.shape
    e = dz.new_zeros(n, k * o, h, w + k - 1)
    for s in range(k):
        e[:, s * o:(s + 1) * o, :, s:s + w] = dz
    return e


def wgrad_hstack_restatement(xpad: Tensor, dz: Tensor, k: int) -> Tensor:
    """Weight gradient of y = conv(xpad, w) (stride 1) from the stacked gradient:
    dW[o,c,r,s] = sum_{n,y,q} xpad[n,c,y+r,q] * E[n,(s,o),y,q]  -- the wgrad of a k x 1 convolution with k*O outputs."""
    n, c, hp, wp = xpad.shape
    o = dz.shape[1]
    e = dz_hstack_restatement(dz, k)                              # (N, k*O, H, Wp)
    h = e.shape[2]
    dw = xpad.new_zeros(o, c, k, k)
    for r in range(k):
        t = torch.einsum("ncyq,njyq->jc", xpad[:, :, r:r + h, :], e)   # (k*O, C)
        dw[:, :, r, :] = t.reshape(k, o, c).permute(1, 2, 0)
    return dw


def fold_reflect_restatement(dxp: Tensor, pad: int) -> Tensor:
    """uegan_fold_inplace: adjoint of nn.ReflectionPad2d(pad) applied to the gradient of the padded tensor."""
    n, c, hp, wp = dxp.shape
    h, w = hp - 2 * pad, wp - 2 * pad
    g = dxp
    # separable: fold rows, then columns
    rows = g[:, :, pad:pad + h, :].clone()
    for y in range(1, pad + 1):
        rows[:, :, y, :] += g[:, :, pad - y, :]
        rows[:, :, h - 1 - y, :] += g[:, :, pad + h - 1 + y, :]
    out = rows[:, :, :, pad:pad + w].clone()
    for x in range(1, pad + 1):
        out[:, :, :, x] += rows[:, :, :, pad - x]
        out[:, :, :, w - 1 - x] += rows[:, :, :, pad + w - 1 + x]
    return out


# ----------------------------------------------------------------------------------------------
# SURVEY.md 8(f) "next" rows N3 / N4: the steps either side of the hot path (oracle only, for the next round's kernels)
# ----------------------------------------------------------------------------------------------
def to_tensor_normalize(img_u8_hwc: np.ndarray) -> Tensor:
    """transforms.ToTensor() + Normalize(0.5, 0.5) on a uint8 HWC RGB image -> fp32 CHW in [-1, 1]
    (/root/reference/data_loader.py:79-81, 100-103)."""
    x = torch.from_numpy(np.ascontiguousarray(img_u8_hwc)).permute(2, 0, 1).to(torch.float32).div(255.0)
    return (x - 0.5) / 0.5


def denorm(x: Tensor) -> Tensor:
    """/root/reference/utils.py:128-130."""
    return ((x + 1) / 2.0).clamp(0, 1)


def denorm_to_u8(x_chw: Tensor) -> np.ndarray:
    """denorm followed by torchvision.utils.save_image's quantisation (mul(255).add_(0.5).clamp_(0,255) -> uint8 HWC),
    the PNG the Tester writes (/root/reference/tester.py via utils.save_image)."""
    y = denorm(x_chw).mul(255).add(0.5).clamp(0, 255)
    return y.permute(1, 2, 0).to(torch.uint8).numpy()


def calculate_psnr(img1: np.ndarray, img2: np.ndarray, data_range: float = 255.0) -> float:
    """/root/reference/metrics/CalcPSNR.py:85-92 (inputs in [0, 255])."""
    a, b = img1.astype(np.float64), img2.astype(np.float64)
    mse = np.mean((a - b) ** 2, dtype=np.float64)
    if mse == 0:
        return float("inf")
    return float(10 * np.log10((data_range ** 2) / mse))


def structural_similarity(im1: np.ndarray, im2: np.ndarray, data_range: float = 255.0) -> float:
    """skimage.metrics.structural_similarity(im1, im2, multichannel=True, data_range=255) as metrics/CalcSSIM.py:62 calls
    it: the algorithm lives in the third-party dependency scikit-image (absent from /root/reference and from this image;
    the reference README pins no version -- the `multichannel=` keyword fixes it to 0.16 <= v < 0.19), restated here from
    its published source (skimage/metrics/_structural_similarity.py): win_size 7, uniform window through
    scipy.ndimage.uniform_filter, use_sample_covariance=True, K1 = 0.01, K2 = 0.03, mean of the SSIM map cropped by
    (win_size - 1) // 2, averaged over channels.  PARITY UNPINNED for this function (no skimage to run, the reference holds
    no SSIM fixture); it is anchored by its closed-form properties in tests/test_oracle.py."""
    from scipy.ndimage import uniform_filter
    assert im1.shape == im2.shape and im1.ndim == 3
    win, k1, k2 = 7, 0.01, 0.03
    npix = win * win
    cov_norm = npix / (npix - 1.0)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    pad = (win - 1) // 2
    vals = []
    for ch in range(im1.shape[2]):
        x, y = im1[..., ch].astype(np.float64), im2[..., ch].astype(np.float64)
        ux, uy = uniform_filter(x, size=win), uniform_filter(y, size=win)
        uxx, uyy, uxy = uniform_filter(x * x, size=win), uniform_filter(y * y, size=win), uniform_filter(x * y, size=win)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        a1, a2, b1, b2 = 2 * ux * uy + c1, 2 * vxy + c2, ux ** 2 + uy ** 2 + c1, vx + vy + c2
        s = (a1 * a2) / (b1 * b2)
        vals.append(s[pad:-pad, pad:-pad].mean(dtype=np.float64))
    return float(np.mean(vals))


def psnr_ssim_pair(gen_u8_hwc: np.ndarray, gt_u8_hwc: np.ndarray, crop_border: int = 4) -> Tuple[float, float]:
    """What calc_psnr / calc_ssim compute for one image pair (metrics/CalcPSNR.py:35-58, CalcSSIM.py:35-62): images / 255,
    border crop, * 255, then the metric."""
    a, b = gt_u8_hwc / 255., gen_u8_hwc / 255.
    a, b = a[crop_border:-crop_border, crop_border:-crop_border, :], b[crop_border:-crop_border, crop_border:-crop_border, :]
    return calculate_psnr(a * 255, b * 255), structural_similarity(a * 255, b * 255)


def to_tensor_normalize_imagenet(img_u8_hwc: np.ndarray) -> Tensor:
    """ToTensor() followed by the ImageNet normalisation PerceptualLoss applies to a [0, 1] image (losses.py:19-20,26-27)."""
    x = torch.from_numpy(np.ascontiguousarray(img_u8_hwc)).permute(2, 0, 1).to(torch.float32).div(255.0)
    mean = torch.tensor(IMAGENET_MEAN).view(-1, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(-1, 1, 1)
    return (x - mean) / std
