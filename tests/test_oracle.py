"""The oracle (oracle/uegan_oracle.py) against golden vectors produced by the reference itself
(tests/golden/make_golden.py, run in the build container where /root/reference exists)."""
import os

import numpy as np
import pytest
import torch

from oracle import uegan_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(regime):
    return np.load(os.path.join(GOLD, f"golden_{regime}.npz"))


def close(a, b, tol):
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    err = float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))
    assert err <= tol, f"rel err {err:.3e} > {tol:.1e}"


@pytest.mark.parametrize("regime", ["o1", "tiny"])
@pytest.mark.parametrize("simplified", [False, True])
def test_generator_matches_reference(regime, simplified):
    g = gold(regime)
    gp = O.make_generator_params(32, 0, regime)
    x = O.make_images((2, 3, 128, 128), 10)
    with torch.no_grad():
        out, inter = O.generator_forward(gp, x, simplified=simplified, return_all=True)
    # fp32 CPU vs fp32 CPU: only summation-order noise is allowed (1e-5 of the tensor scale); the
    # simplified (GAM-cancelled, hoisted-upsample) form is an algebraic identity -> 1e-4.
    tol = 1e-4 if simplified else 1e-5
    close(out, g["g128_out"], tol)
    close(inter["res"], g["g128_res"], 2e-3 if (simplified and regime == "tiny") else tol * 10)
    if regime == "o1":
        assert float(torch.as_tensor(g["g128_res"]).abs().mean()) > 0.05  # parity is not vacuous


@pytest.mark.parametrize("regime", ["o1", "tiny"])
def test_generator_nonsquare(regime):
    g = gold(regime)
    gp = O.make_generator_params(32, 0, regime)
    with torch.no_grad():
        out = O.generator_forward(gp, O.make_images((1, 3, 96, 160), 11))
    close(out, g["g96x160_out"], 1e-5)


@pytest.mark.parametrize("regime", ["o1", "tiny"])
def test_discriminator_matches_reference(regime):
    g = gold(regime)
    dp = O.make_discriminator_params(32, 1, regime)
    x = O.make_images((2, 3, 128, 128), 10)
    with torch.no_grad():
        preds = O.discriminator_forward(dp, x, training=True)
    for i, p in enumerate(preds):
        close(p, g[f"d128_pred{i+1}"], 2e-5)
    for k in range(1, 6):
        close(dp[f"d{k}.0.1.weight_u"], g[f"d128_u{k}"], 1e-5)
        close(dp[f"d{k}.0.1.weight_v"], g[f"d128_v{k}"], 1e-5)
    with torch.no_grad():
        preds = O.discriminator_forward(dp, x, training=False)
    for i, p in enumerate(preds):
        close(p, g[f"d128_eval_pred{i+1}"], 2e-5)


def test_discriminator_bad_loss_type():
    dp = O.make_discriminator_params(8, 1)
    with pytest.raises(NotImplementedError):
        O.discriminator_forward(dp, torch.zeros(1, 3, 96, 96), adv_loss_type="wgan")


def test_losses_match_reference():
    g = gold("o1")
    vp = O.make_vgg_params()
    a = O.make_images((2, 3, 64, 64), 12)
    b = O.make_images((2, 3, 64, 64), 13)
    with torch.no_grad():
        close(O.perceptual_loss(vp, (a + 1) / 2, (b + 1) / 2), g["percep64"], 1e-5)
        mean = torch.tensor(O.IMAGENET_MEAN).view(1, -1, 1, 1)
        std = torch.tensor(O.IMAGENET_STD).view(1, -1, 1, 1)
        taps = O.vgg19_taps(vp, ((a + 1) / 2 - mean) / std)
        close(taps["relu5_1"], g["vgg64_relu5_1"], 1e-5)
        close(taps["relu3_1"][0, :4], g["vgg64_relu3_1_n0c0"], 1e-5)
    rp = [torch.tanh(O.make_images((2, 1, s, s), 20 + i)) for i, s in enumerate((64, 32, 16, 8, 4))]
    fp = [torch.tanh(O.make_images((2, 1, s, s), 30 + i)) for i, s in enumerate((64, 32, 16, 8, 4))]
    close(O.gan_loss("rahinge", rp, fp, True), g["rahinge_d"], 1e-6)
    close(O.gan_loss("rahinge", rp, fp, False), g["rahinge_g"], 1e-6)
    close(O.gan_loss("rals", rp, fp, True), g["rals_d"], 1e-6)
    close(O.gan_loss("rals", rp, fp, False), g["rals_g"], 1e-6)
    close(O.multiscale_rec_loss(a, b), g["msl1"], 1e-6)
    with pytest.raises(NotImplementedError):
        O.multiscale_rec_loss(a, b, rec_loss_type="huber")


def test_train_step_matches_reference():
    """Two iterations of the trainer.py:75-119 loop body: loss scalars and post-step weights."""
    g = gold("o1")
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    gp = O.make_generator_params(32, 0, "o1")
    dp = O.make_discriminator_params(32, 1, "o1")
    vp = O.make_vgg_params()
    g_opt, d_opt = O.AdamState(O._trainable(gp)), O.AdamState(O._trainable(dp))
    raw = O.make_images((2, 3, 128, 128), 40)
    exp = O.make_images((2, 3, 128, 128), 41)
    for step in range(2):
        losses = O.train_step(gp, dp, vp, g_opt, d_opt, raw, exp)
        got = [losses[k] for k in ("d_loss", "g_adv_loss", "g_percep_loss", "g_idt_loss", "g_loss")]
        np.testing.assert_allclose(got, g[f"step{step}_losses"], rtol=2e-4)
    # Adam's first steps move every weight by ~lr regardless of gradient scale, so the post-step weight
    # DELTAS (in units of lr*steps) are a sharp check of gradient signs and optimizer arithmetic; elements
    # whose gradient is rounding noise may flip sign, hence a max bound < 1 and a tight mean bound.
    gp0, dp0 = O.make_generator_params(32, 0, "o1"), O.make_discriminator_params(32, 1, "o1")

    def delta_close(pre, mine, ref, unit, max_tol=0.5, mean_tol=0.01):
        d = ((mine - pre) - (torch.as_tensor(ref) - pre)).abs() / unit
        assert float(d.max()) <= max_tol and float(d.mean()) <= mean_tol, (float(d.max()), float(d.mean()))

    delta_close(gp0["enc1.main.1.weight"], gp["enc1.main.1.weight"], g["step_post_enc1_w"], 2e-4)
    delta_close(gp0["dec5.1.main.1.weight"], gp["dec5.1.main.1.weight"], g["step_post_dec5_1_w"], 2e-4)
    delta_close(gp0["dec5.1.main.1.bias"], gp["dec5.1.main.1.bias"], g["step_post_dec5_1_b"], 2e-4)
    delta_close(dp0["d1.0.1.weight_orig"], dp["d1.0.1.weight_orig"], g["step_post_d1_w"], 8e-4)
    delta_close(dp0["d5_pred.0.1.weight"], dp["d5_pred.0.1.weight"], g["step_post_d5_pred_w"], 8e-4)
    close(dp["d1.0.1.weight_u"], g["step_post_d1_u"], 1e-4)


# ---------------------------------------------------------------------------------------------------------------
# algebra of the tiny-Cout kernels and of the in-place fold, pinned on the CPU against F.conv2d / autograd
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cin,cout,k", [(8, 3, 7), (16, 1, 7), (8, 1, 5), (4, 1, 3)])
def test_rowsum_and_hstack_algebra(cin, cout, k):
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    pad = (k - 1) // 2
    x = torch.randn(2, cin, 13, 17, generator=g, dtype=torch.float64)
    w = torch.randn(cout, cin, k, k, generator=g, dtype=torch.float64, requires_grad=True)
    xpad = F.pad(x, (pad,) * 4, mode="reflect")
    ref = F.conv2d(xpad, w)
    close(O.conv_rowsum_restatement(xpad, w.detach()), ref.detach(), 1e-12)
    dz = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    ref.backward(dz)
    close(O.wgrad_hstack_restatement(xpad, dz, k), w.grad, 1e-12)
    e = O.dz_hstack_restatement(dz, k)
    assert e.shape == (2, k * cout, 13, 17 + k - 1)


@pytest.mark.parametrize("pad", [1, 2, 3])
def test_fold_reflect_algebra(pad):
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(6)
    xin = torch.zeros(2, 4, 9, 11, dtype=torch.float64, requires_grad=True)
    dxp = torch.randn(2, 4, 9 + 2 * pad, 11 + 2 * pad, generator=g, dtype=torch.float64)
    F.pad(xin, (pad,) * 4, mode="reflect").backward(dxp)
    close(O.fold_reflect_restatement(dxp, pad), xin.grad, 1e-12)


# ---------------------------------------------------------------------------------------------------------------
# SURVEY.md 8(f) N3 / N4 groundwork: input / output conversion and PSNR against the reference's own functions
# ---------------------------------------------------------------------------------------------------------------
def test_io_and_psnr_match_reference():
    import sys
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, size=(24, 40, 3), dtype=np.uint8)
    x = O.to_tensor_normalize(img)
    assert x.shape == (3, 24, 40) and float(x.min()) >= -1.0 and float(x.max()) <= 1.0
    assert np.array_equal(O.denorm_to_u8(x), img)  # ToTensor/Normalize -> denorm -> save_image quantisation round trip
    a = rng.integers(0, 256, size=(24, 40, 3)).astype(np.float64)
    b = np.clip(a + rng.normal(0, 4, size=a.shape), 0, 255)
    mine = O.calculate_psnr(a, b)
    assert abs(mine - 10 * np.log10(255.0 ** 2 / np.mean((a - b) ** 2))) < 1e-12
    assert O.calculate_psnr(a, a) == float("inf")
    if os.path.isdir("/root/reference"):  # live pin in the build container (cv2 / torchvision present there)
        sys.path.insert(0, "/root/reference")
        try:
            from metrics.CalcPSNR import calculate_psnr as ref_psnr
            import utils as ref_utils  # noqa: F401  (needs tensorflow / skimage shims in some images)
        except Exception:
            ref_psnr = None
        finally:
            sys.path.pop(0)
        if ref_psnr is not None:
            assert abs(mine - ref_psnr(a, b)) < 1e-12
        import torchvision.transforms as T
        from PIL import Image
        tf = T.Compose([T.ToTensor(), T.Normalize(mean=[0.5] * 3, std=[0.5] * 3)])  # data_loader.py:79-81
        close(x, tf(Image.fromarray(img)), 1e-7)


def test_ssim_restatement_properties():
    """structural_similarity (restated from scikit-image, parity unpinned: see its docstring) -- closed-form anchors."""
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (40, 48, 3), dtype=np.uint8)
    b = np.clip(a.astype(np.int32) + rng.integers(-20, 21, a.shape), 0, 255).astype(np.uint8)
    assert abs(O.structural_similarity(a.astype(np.float64), a.astype(np.float64)) - 1.0) < 1e-12
    s_ab, s_ba = O.structural_similarity(a.astype(np.float64), b.astype(np.float64)), O.structural_similarity(
        b.astype(np.float64), a.astype(np.float64))
    assert abs(s_ab - s_ba) < 1e-12 and 0.0 < s_ab < 1.0
    # constant images: S = (2 ux uy + C1) / (ux^2 + uy^2 + C1) everywhere (variances and covariance vanish)
    c1 = (0.01 * 255) ** 2
    x, y = np.full((16, 16, 3), 100.0), np.full((16, 16, 3), 140.0)
    assert abs(O.structural_similarity(x, y) - (2 * 100 * 140 + c1) / (100 ** 2 + 140 ** 2 + c1)) < 1e-12
    # brute-force evaluation of the definition at one interior pixel
    p, q = a[..., 0].astype(np.float64), b[..., 0].astype(np.float64)
    wy, wx = 10, 17
    P, Q = p[wy - 3:wy + 4, wx - 3:wx + 4], q[wy - 3:wy + 4, wx - 3:wx + 4]
    ux, uy = P.mean(), Q.mean()
    vx, vy, vxy = P.var(ddof=1), Q.var(ddof=1), ((P - ux) * (Q - uy)).sum() / 48.0
    c2 = (0.03 * 255) ** 2
    want = (2 * ux * uy + c1) * (2 * vxy + c2) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    from scipy.ndimage import uniform_filter
    cn = 49.0 / 48.0
    fx, fy = uniform_filter(p, 7), uniform_filter(q, 7)
    got = ((2 * fx * fy + c1) * (2 * cn * (uniform_filter(p * q, 7) - fx * fy) + c2) /
           ((fx ** 2 + fy ** 2 + c1) * (cn * (uniform_filter(p * p, 7) - fx * fx) + cn * (uniform_filter(q * q, 7) - fy * fy) + c2)))[wy, wx]
    assert abs(got - want) < 1e-9
