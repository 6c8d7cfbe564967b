"""Full-resolution checks (3x512x512, the resolution of BASELINE.json configs[1..3]); the batch is kept small because every
image is processed independently (InstanceNorm, GAM and the convolutions are per-sample; only the relativistic GAN means
couple a batch).  Direct parity with the CPU oracle where it finishes in seconds, plus a size-independent exactness property.
(Last file of the suite on purpose: these cases were added after round 1's GPU budget was spent.)"""
import numpy as np
import pytest
import torch

from oracle import uegan_oracle as O

pytestmark = pytest.mark.gpu


def _rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def _generator(regime="o1"):
    from uegan_b200.models import Generator
    G = Generator(32, "none", "LeakyReLU", False)
    G.load_state_dict(O.make_generator_params(32, 0, regime))
    return G.cuda().eval()


def test_generator_512_vs_oracle():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from uegan_b200 import kernels as K
    G = _generator()
    x = O.make_images((1, 3, 512, 512), 21)
    with torch.no_grad():
        out = G(x.cuda()).cpu()
        ref = O.generator_forward(O.make_generator_params(32, 0, "o1"), x)
    assert K.device_error() == 0
    l2 = _rel_l2(out, ref)
    worst = float((out - ref).abs().max() / ref.abs().max())
    print(f"1x3x512x512: pixel rel-L2 {l2:.3e}, worst pixel {worst:.3e}")
    assert l2 < 1e-3 and worst < 5e-3  # north_star: 1e-3 relative on generator pixels


def test_generator_512_identity_when_last_conv_is_zero():
    """Size-independent exactness: with dec5.1 zeroed, res = tanh(0) = 0 and out = clamp(0 + x, -1, 1) = x bit for bit --
    every output pixel of every tile of the full-resolution launch grid must be written exactly once."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from uegan_b200 import kernels as K
    G = _generator()
    with torch.no_grad():
        G.dec5[1].main[1].weight.zero_()
        G.dec5[1].main[1].bias.zero_()
        x = O.make_images((2, 3, 512, 512), 22).cuda()
        out = torch.full_like(x, 7.0)
        out.copy_(G(x))
    assert K.device_error() == 0
    assert torch.equal(out, x)


def test_train_step_512_losses_vs_oracle():
    """One full training step (trainer.py:75-119) at 1x3x512x512 against the oracle's step: the five loss scalars."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from uegan_b200 import kernels as K
    from tests.test_gpu_train import build, train_step
    G, D, P, gl, ms = build()
    g_opt = torch.optim.Adam(G.parameters(), lr=1e-4, betas=[0.5, 0.999], weight_decay=0.0001)
    d_opt = torch.optim.Adam(D.parameters(), lr=4e-4, betas=[0.5, 0.999], weight_decay=0.0001)
    raw = O.make_images((1, 3, 512, 512), 50)
    exp = O.make_images((1, 3, 512, 512), 51)
    losses = train_step(G, D, P, gl, ms, g_opt, d_opt, raw.cuda(), exp.cuda())
    assert K.device_error() == 0
    gp, dp, vp = O.make_generator_params(32, 0, "o1"), O.make_discriminator_params(32, 1, "o1"), O.make_vgg_params()
    ref = O.train_step(gp, dp, vp, O.AdamState(O._trainable(gp)), O.AdamState(O._trainable(dp)), raw, exp)
    ref = [ref[k] for k in ("d_loss", "g_adv_loss", "g_percep_loss", "g_idt_loss", "g_loss")]
    errs = [abs(a - b) / abs(b) for a, b in zip(losses, ref)]
    print(f"512x512 step: losses {['%.6f' % v for v in losses]} ref {['%.6f' % v for v in ref]} rel {['%.2e' % e for e in errs]}")
    assert all(np.isfinite(losses))
    assert max(errs) < 1e-3  # north_star: 1e-3 relative on loss scalars


def test_config1_generator_inference_b32_512():
    """BASELINE.json configs[1] at its FULL size: Generator inference on 32 x 3x512x512.  Images are independent through the
    Generator, so the oracle checks a sample of the batch (first, middle, last) directly (1e-3 on pixels, north_star) and a
    size-independent property covers the rest: every image of the big batch equals the same image run in a batch of its own."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from uegan_b200 import kernels as K
    G = _generator()
    x = O.make_images((32, 3, 512, 512), 23)
    gp = O.make_generator_params(32, 0, "o1")
    with torch.no_grad():
        out = G(x.cuda()).clone()
        for i in (0, 13, 31):
            ref = O.generator_forward(gp, x[i:i + 1])
            l2 = _rel_l2(out[i:i + 1].cpu(), ref)
            print(f"configs[1] 32x3x512x512, image {i}: pixel rel-L2 vs oracle {l2:.3e}")
            assert l2 < 1e-3
        for i in (5, 20):
            solo = G(x[i:i + 1].cuda())
            assert float((solo - out[i:i + 1]).abs().max()) <= 1e-6, i
    assert K.device_error() == 0


def test_config4_generator_1024x1536_vs_oracle():
    """BASELINE.json configs[4] (mixed-resolution sweep): its largest shape, 1024 x 1536 (non-square, 3x the pixels of a tile
    grid the 512^2 cases exercise), against the oracle."""
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from uegan_b200 import kernels as K
    G = _generator()
    x = O.make_images((1, 3, 1024, 1536), 24)
    with torch.no_grad():
        out = G(x.cuda()).cpu()
        ref = O.generator_forward(O.make_generator_params(32, 0, "o1"), x)
    assert K.device_error() == 0
    l2 = _rel_l2(out, ref)
    print(f"configs[4] 1x3x1024x1536: pixel rel-L2 {l2:.3e}")
    assert l2 < 1e-3
    x2 = O.make_images((3, 3, 256, 384), 25)
    with torch.no_grad():
        l2 = _rel_l2(G(x2.cuda()).cpu(), O.generator_forward(O.make_generator_params(32, 0, "o1"), x2))
    print(f"configs[4] 3x3x256x384: pixel rel-L2 {l2:.3e}")
    assert l2 < 1e-3
