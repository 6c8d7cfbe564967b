"""SURVEY.md 8(f) N3 / N4 on the GPU, through the C ABI: uint8 <-> normalised tensor conversions (bit-exact against the
oracle's restatement of ToTensor / Normalize / denorm / save_image, which tests/test_oracle.py pins to the reference's own
functions) and PSNR / SSIM of uint8 pairs (metrics/CalcPSNR.py, metrics/CalcSSIM.py)."""
import numpy as np
import pytest
import torch

from oracle import uegan_oracle as O

pytestmark = pytest.mark.gpu


def _imgs(n, h, w, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    a[0, 0, 0] = (0, 255, 128)  # the extremes and the rounding midpoint are present
    return a


@pytest.mark.parametrize("shape", [(2, 32, 48), (1, 513, 259), (3, 512, 512)])
def test_pack_u8_bit_exact(shape):
    from uegan_b200 import _lib as L
    from uegan_b200 import io as IO
    from uegan_b200 import kernels as K
    n, h, w = shape
    a = _imgs(n, h, w, 1)
    dev = torch.from_numpy(a).cuda()
    dst = K.NHWC(n, h, w, 4, 3, L.F32, "cuda", zero=True)
    planes = IO.pack_u8(dev, dst)
    ref = torch.stack([O.to_tensor_normalize(a[i]) for i in range(n)])
    assert torch.equal(planes.cpu(), ref)  # bit for bit what the reference's loader yields
    # the NHWC operand: same values rounded to tf32, reflection halo of 3 (models.py:82 ReflectionPad2d(3))
    want = torch.nn.functional.pad(ref, (3, 3, 3, 3), mode="reflect").permute(0, 2, 3, 1)
    got = dst.padded_view()[..., :3].cpu()
    assert float((got - want).abs().max()) <= 2.0 ** -11  # tf32 rounding of values in [-1, 1]
    assert float(dst.padded_view()[..., 3].abs().max()) == 0.0
    # ImageNet variant (losses.py:26-27), fp16 operand with a zero halo (the VGG input)
    d16 = K.NHWC(n, h, w, 8, 1, L.F16, "cuda", zero=True)
    p2 = IO.pack_u8(dev, d16, mean=IO.IMAGENET_MEAN, std=IO.IMAGENET_STD, pad_mode=L.PAD_ZERO)
    ref2 = torch.stack([O.to_tensor_normalize_imagenet(a[i]) for i in range(n)])
    assert torch.equal(p2.cpu(), ref2)
    got16 = d16.padded_view()[:, 1:-1, 1:-1, :3].float().cpu()
    assert torch.equal(got16, ref2.permute(0, 2, 3, 1).half().float())
    assert float(d16.padded_view()[:, 0].abs().max()) == 0.0
    assert K.device_error() == 0


@pytest.mark.parametrize("shape", [(2, 3, 32, 48), (1, 3, 513, 259)])
def test_unpack_u8_bit_exact(shape):
    from uegan_b200 import io as IO
    x = O.make_images(shape, 5) * 1.2  # beyond [-1, 1]: the clamps are exercised
    x[0, 0, 0, :5] = torch.tensor([-1.0, 1.0, 0.0, 1.0 / 255 - 1.0, 0.003921568])
    got = IO.unpack_u8(x.cuda()).cpu().numpy()
    for i in range(shape[0]):
        assert np.array_equal(got[i], O.denorm_to_u8(x[i]))
    # round trip: u8 -> [-1, 1] -> u8 is the identity (ToTensor / Normalize then denorm / save_image)
    a = _imgs(2, 64, 64, 7)
    back = IO.unpack_u8(IO.pack_u8(torch.from_numpy(a).cuda())).cpu().numpy()
    assert np.array_equal(back, a)


def test_enhance_u8_matches_float_path():
    """uint8 in -> Generator -> uint8 out equals the reference-shaped path (loader tensor -> G -> denorm -> save_image)."""
    from uegan_b200 import io as IO
    from uegan_b200 import kernels as K
    from uegan_b200.models import Generator
    G = Generator(32, "none", "LeakyReLU", False)
    gp = O.make_generator_params(32, 0, "o1")
    G.load_state_dict(gp)
    G = G.cuda().eval()
    a = _imgs(2, 128, 160, 9)
    got = IO.enhance_u8(G, torch.from_numpy(a).cuda()).cpu().numpy()
    x = torch.stack([O.to_tensor_normalize(a[i]) for i in range(2)])
    with torch.no_grad():
        via_float = IO.unpack_u8(G(x.cuda())).cpu().numpy()
        ref = O.generator_forward(gp, x)
    assert np.array_equal(got, via_float)
    want = np.stack([O.denorm_to_u8(ref[i]) for i in range(2)])
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    print(f"enhance_u8 vs oracle PNG bytes: max |diff| {d.max()}, mean {d.mean():.4f}")
    # 1e-3 rel-L2 on pixels in [-1, 1] is ~0.1 grey level: never more than one level off, and a few per cent of the bytes
    # sit close enough to a rounding boundary to land on the other side (measured 3.5 %)
    assert d.max() <= 1 and d.mean() < 0.06
    assert K.device_error() == 0


@pytest.mark.parametrize("shape", [(2, 40, 56), (1, 512, 512), (3, 261, 135)])
def test_psnr_ssim_vs_oracle(shape):
    from uegan_b200 import metrics as MT
    n, h, w = shape
    a = _imgs(n, h, w, 11)
    rng = np.random.default_rng(12)
    b = np.clip(a.astype(np.int32) + rng.integers(-12, 13, a.shape), 0, 255).astype(np.uint8)
    b[-1] = a[-1]  # one identical pair: PSNR = inf, SSIM = 1
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    psnr, ssim = MT.psnr_u8(db, da), MT.ssim_u8(db, da)
    for i in range(n):
        p_ref, s_ref = O.psnr_ssim_pair(b[i], a[i])
        if np.isinf(p_ref):
            assert np.isinf(float(psnr[i])) and abs(float(ssim[i]) - 1.0) < 1e-12
            continue
        assert abs(float(psnr[i]) - p_ref) < 1e-9 * abs(p_ref), (float(psnr[i]), p_ref)
        assert abs(float(ssim[i]) - s_ref) < 1e-9, (float(ssim[i]), s_ref)
    # exact integer property: the sum of squared differences equals numpy's on the cropped region
    sse = MT.sse_u8(da, db).cpu().numpy()
    for i in range(n):
        d = a[i, 4:-4, 4:-4].astype(np.int64) - b[i, 4:-4, 4:-4].astype(np.int64)
        assert int(sse[i]) == int((d * d).sum())


def test_metrics_reject_bad_input():
    from uegan_b200 import _lib as L
    from uegan_b200 import metrics as MT
    a = torch.zeros(1, 10, 10, 3, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        MT.psnr_u8(a, torch.zeros(1, 10, 12, 3, dtype=torch.uint8, device="cuda"))
    with pytest.raises(L.UeganError):
        MT.ssim_u8(a, a)  # 10 - 8 < 7: smaller than the SSIM window after the border crop
    with pytest.raises(L.UeganError):
        MT.psnr_u8(a.cpu(), a.cpu())
