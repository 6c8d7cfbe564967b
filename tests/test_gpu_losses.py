"""GPU parity of the Discriminator, the VGG-19 perceptual tower and the loss reductions (forward) against the
golden vectors produced by the reference and against the CPU oracle.  Loss scalars: 1e-3 relative (north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import uegan_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(regime="o1"):
    return np.load(os.path.join(GOLD, f"golden_{regime}.npz"))


def _t(v):
    if isinstance(v, torch.Tensor):
        return v.detach().double().cpu()
    return torch.as_tensor(np.asarray(v)).double()


def rel(a, b):
    a, b = _t(a), _t(b)
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")


@pytest.mark.parametrize("regime", ["o1", "tiny"])
def test_discriminator_vs_golden(regime):
    need_gpu()
    from uegan_b200 import kernels as K
    from uegan_b200.models import Discriminator
    g = gold(regime)
    D = Discriminator(32, "none", "LeakyReLU", True, "rahinge")
    D.load_state_dict(O.make_discriminator_params(32, 1, regime))
    D = D.cuda().train()
    x = O.make_images((2, 3, 128, 128), 10).cuda()
    with torch.no_grad():
        preds = D(x)
    assert K.device_error() == 0
    for i, p in enumerate(preds):
        assert tuple(p.shape) == tuple(g[f"d128_pred{i+1}"].shape)
        e = rel(p, g[f"d128_pred{i+1}"])
        print(f"[{regime}] train pred{i+1} rel err {e:.3e}")
        assert e < 5e-3
    sd = D.state_dict()
    for k in range(1, 6):  # power iteration is fp32 CUDA-core math
        assert rel(sd[f"d{k}.0.1.weight_u"], g[f"d128_u{k}"]) < 1e-4
        assert rel(sd[f"d{k}.0.1.weight_v"], g[f"d128_v{k}"]) < 1e-4
    D.eval()
    with torch.no_grad():
        preds = D(x)
    for i, p in enumerate(preds):
        assert rel(p, g[f"d128_eval_pred{i+1}"]) < 5e-3
    for k in range(1, 6):  # eval mode must not advance u / v
        assert rel(D.state_dict()[f"d{k}.0.1.weight_u"], g[f"d128_u{k}"]) < 1e-4


def test_discriminator_rejects():
    from uegan_b200.models import Discriminator
    with pytest.raises(NotImplementedError):
        Discriminator(32, "none", "LeakyReLU", True, "wgan")
    if torch.cuda.is_available():
        D = Discriminator(8, "none", "LeakyReLU", True, "rahinge").cuda()
        with pytest.raises(ValueError), torch.no_grad():
            D(torch.zeros(1, 3, 64, 64, device="cuda"))


def test_gan_and_rec_losses_vs_golden():
    need_gpu()
    from uegan_b200.losses import GANLoss, MultiscaleRecLoss
    g = gold()
    rp = [torch.tanh(O.make_images((2, 1, s, s), 20 + i)).cuda() for i, s in enumerate((64, 32, 16, 8, 4))]
    fp = [torch.tanh(O.make_images((2, 1, s, s), 30 + i)).cuda() for i, s in enumerate((64, 32, 16, 8, 4))]
    with torch.no_grad():
        gl = GANLoss("rahinge")
        assert rel(gl(rp, fp, None, None, for_discriminator=True), g["rahinge_d"]) < 1e-5
        assert rel(gl(rp, fp, None, None, for_discriminator=False), g["rahinge_g"]) < 1e-5
        gl2 = GANLoss("rals")
        assert rel(gl2(rp, fp, None, None, for_discriminator=True), g["rals_d"]) < 1e-5
        assert rel(gl2(rp, fp, None, None, for_discriminator=False), g["rals_g"]) < 1e-5
        a = O.make_images((2, 3, 64, 64), 12).cuda()
        b = O.make_images((2, 3, 64, 64), 13).cuda()
        assert rel(MultiscaleRecLoss(3, "l1", True)(a, b), g["msl1"]) < 1e-5
        for t in ("l2", "smoothl1"):
            ref = O.multiscale_rec_loss(a.cpu(), b.cpu(), 3, t)
            assert rel(MultiscaleRecLoss(3, t, True)(a, b), ref) < 1e-5
    with pytest.raises(ValueError):
        GANLoss("bogus")
    with pytest.raises(NotImplementedError):
        MultiscaleRecLoss(3, "huber")
    with pytest.raises(NotImplementedError):
        GANLoss("hinge")(rp, fp, None)


def test_perceptual_vs_golden_and_oracle():
    need_gpu()
    from uegan_b200 import kernels as K
    from uegan_b200.losses import PerceptualLoss
    g = gold()
    vp = O.make_vgg_params()
    P = PerceptualLoss(vgg_state_dict=vp).cuda()
    a = O.make_images((2, 3, 64, 64), 12)
    b = O.make_images((2, 3, 64, 64), 13)
    with torch.no_grad():
        loss = P((a.cuda() + 1) / 2, (b.cuda() + 1) / 2)
        taps, _ = P.vgg.run((a.cuda() + 1) / 2, "x")
    assert K.device_error() == 0
    e = rel(loss, g["percep64"])
    print(f"perceptual 64x64: {float(loss):.6f} vs reference {float(g['percep64']):.6f} rel err {e:.3e}")
    assert e < 1e-3
    e5 = rel(taps[4][0].interior_nchw() / P.vgg.tap_scale(4), g["vgg64_relu5_1"])  # stored = 2^k * true activation
    print(f"relu5_1 (fp16 tower, stored scale {P.vgg.tap_scale(4):g}) rel err {e5:.3e}")
    assert e5 < 5e-2
    # larger size: exercises the statistics fused into the conv epilogue
    a = O.make_images((2, 3, 128, 160), 14)
    b = O.make_images((2, 3, 128, 160), 15)
    with torch.no_grad():
        loss = P((a.cuda() + 1) / 2, (b.cuda() + 1) / 2)
        ref = O.perceptual_loss(vp, (a + 1) / 2, (b + 1) / 2)
    e = rel(loss, ref)
    print(f"perceptual 128x160: {float(loss):.6f} vs oracle {float(ref):.6f} rel err {e:.3e}")
    assert e < 1e-3


@pytest.mark.parametrize("factor", [16.0, 1.0 / 16.0, "alternating"])
def test_perceptual_fp16_range_rescaled_tower(factor):
    """the fp16 tower must not depend on the synthetic He-normal checkpoint's dynamic range.  Every conv
    weight is multiplied by 16 (activations would reach 16^13 = 4.5e15 by relu5_1: far outside fp16), by 1/16 (2.2e-16:
    flushed to zero), or alternately by 64 and 1/64; the power-of-two calibration (losses.VGG19_relu.calibrate) must keep
    the loss, its image gradient and every tap within the same tolerances as the unscaled tower -- against the fp32 oracle
    evaluated on the SAME rescaled weights (biases are rescaled with the cumulative factor so that the rescaled fp32
    network is the original one up to per-layer scale; eps changes the loss only where s^2 * var ~ eps)."""
    need_gpu()
    from uegan_b200 import kernels as K
    from uegan_b200.losses import PerceptualLoss
    vp = {k: v.clone() for k, v in O.make_vgg_params().items()}
    keys = sorted({int(k.split(".")[1]) for k in vp})
    cum = 1.0
    for j, idx in enumerate(keys):
        f = factor if factor != "alternating" else (64.0 if j % 2 == 0 else 1.0 / 64.0)
        cum *= f
        vp[f"features.{idx}.weight"] *= f
        vp[f"features.{idx}.bias"] *= cum
    P = PerceptualLoss(vgg_state_dict=vp).cuda()
    a = ((O.make_images((2, 3, 64, 64), 12) + 1) / 2)
    b = ((O.make_images((2, 3, 64, 64), 13) + 1) / 2)
    x = a.cuda().requires_grad_(True)
    loss = P(x, b.cuda())
    loss.backward()
    assert K.device_error() == 0
    taps, _ = P.vgg.run(a.cuda(), "x")
    P.vgg.check_range(taps)  # the amax guard: must not raise on a calibrated tower
    ad = a.double().clone().requires_grad_(True)
    ref = O.perceptual_loss({k: v.double() for k, v in vp.items()}, ad, b.double())
    ref.backward()
    e = rel(loss, ref)
    g_err = float((x.grad.double().cpu() - ad.grad).norm() / ad.grad.norm())
    print(f"rescaled tower ({factor}): loss {float(loss):.6f} vs fp64 oracle {float(ref):.6f} rel {e:.2e}; "
          f"dL/dx rel-L2 {g_err:.3f}; stored tap scales {[P.vgg.tap_scale(i) for i in range(5)]}")
    assert torch.isfinite(loss) and torch.isfinite(x.grad).all()
    assert e < 1e-3
    assert g_err < 0.2  # same (mask-flip dominated) gate as the unscaled tower, tests/test_gpu_train.py


def test_perceptual_amax_guard_raises():
    """A tower whose activations left fp16's range after calibration (here: weights swapped under a stale calibration)
    is reported by check_range instead of saturating silently."""
    need_gpu()
    from uegan_b200 import _lib
    from uegan_b200.losses import PerceptualLoss
    P = PerceptualLoss(vgg_state_dict=O.make_vgg_params()).cuda()
    a = ((O.make_images((1, 3, 64, 64), 3) + 1) / 2).cuda()
    taps, _ = P.vgg.run(a, "x")
    P.vgg.check_range(taps)
    c = P.vgg._calib[0]
    P.vgg._calib[0] = (c[0], c[1], c[2] * 4096.0, c[3] * 4096.0, c[4])  # conv1_1 output x4096 -> overflow downstream
    P.vgg._wcache.clear()
    taps, _ = P.vgg.run(a, "x")
    with pytest.raises(_lib.UeganError):
        P.vgg.check_range(taps)
