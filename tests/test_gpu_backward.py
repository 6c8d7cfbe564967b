"""GPU parity of the backward kernels (dgrad through the fprop kernel, tcgen05 MN-major wgrad, reflect-fold / masks,
InstanceNorm / upsample / head backward) against torch.autograd in fp32 (double where cheap).  Operands are
quantised to tf32 first, so what remains is fp32 accumulation order plus the tf32 rounding of stored results."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests.test_gpu_kernels import fill_nhwc, relerr, tf32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from uegan_b200 import kernels, _lib
    _lib.load()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return kernels


BWD_CASES = [
    # name, n, cin, h, w, cout, k, stride, pad_mode
    ("k3s1", 2, 64, 24, 32, 32, 3, 1, "reflect"),
    ("k3s2", 2, 32, 32, 32, 64, 3, 2, "reflect"),
    ("k7s2", 2, 32, 48, 48, 64, 7, 2, "reflect"),
    ("k5s2", 2, 128, 12, 16, 256, 5, 2, "reflect"),
    ("k1", 2, 128, 16, 24, 64, 1, 1, "reflect"),
    ("k7s1_cout3", 2, 32, 24, 40, 3, 7, 1, "reflect"),
    ("k7s1_cin3", 2, 3, 32, 48, 32, 7, 1, "reflect"),
    ("k7s2_cin3", 2, 3, 96, 96, 32, 7, 2, "reflect"),
    ("k3_zero", 2, 64, 16, 16, 128, 3, 1, "zero"),
    ("big_c", 1, 512, 8, 8, 256, 3, 1, "reflect"),
    # patch ("Toeplitz descriptor") wgrad: stride 1, Cout <= 64, C multiple of 32
    ("patch_32_32_k3", 2, 32, 40, 24, 32, 3, 1, "reflect"),
    ("patch_64_64_k3", 2, 64, 16, 32, 64, 3, 1, "reflect"),
    ("patch_128_64_k3", 1, 128, 24, 16, 64, 3, 1, "reflect"),
    ("patch_64_1_k7", 2, 64, 16, 24, 1, 7, 1, "reflect"),
    ("patch_256_1_k5", 2, 256, 12, 16, 1, 5, 1, "reflect"),
]


@pytest.mark.parametrize("case", BWD_CASES, ids=[c[0] for c in BWD_CASES])
def test_conv_dgrad_wgrad(K, case):
    from uegan_b200 import _lib as L
    name, n, cin, h, w, cout, k, stride, pad_mode = case
    g = torch.Generator(device="cuda").manual_seed(99)
    pad = (k - 1) // 2
    pm = L.PAD_REFLECT if pad_mode == "reflect" else L.PAD_ZERO
    x = tf32(torch.randn(n, cin, h, w, device="cuda", generator=g))
    wgt = tf32(torch.randn(cout, cin, k, k, device="cuda", generator=g) / math.sqrt(cin * k * k))
    xr = x.clone().requires_grad_(True)
    wr = wgt.clone().double().requires_grad_(True)
    xpad = F.pad(xr.double(), (pad,) * 4, mode="reflect" if pad_mode == "reflect" else "constant") if pad else xr.double()
    y = F.conv2d(xpad, wr, stride=stride)
    ho, wo = y.shape[2], y.shape[3]
    dzv = tf32(torch.randn(n, cout, ho, wo, device="cuda", generator=g))
    y.backward(dzv.double())
    # ---- native tensors
    c_x = 4 if cin == 3 else cin
    xt = fill_nhwc(K, x, c_x, pad, pm, L.F32)
    kq = (k + stride - 1) // stride
    c_dz = 4 if cout <= 4 else cout
    dz_d = fill_nhwc(K, dzv, c_dz, kq - 1, L.PAD_ZERO, L.F32)          # operand of dgrad (zero halo)
    # ---- dgrad -> gradient w.r.t. the padded input, then fold
    if cin != 3:
        dxp = K.NHWC(n, h + 2 * pad, w + 2 * pad, cin, 0, L.F32, "cuda", zero=True)
        K.conv_dgrad(dz_d, wgt, k, stride, dxp)
        dx = K.NHWC(n, h, w, cin, 1, L.F32, "cuda", zero=False)
        dx.buf.fill_(7.0)  # halo must come back as zeros
        K.grad_combine(dx, cin, src_a=dxp, pad_a=pad, pad_mode_a=pm)
        assert K.device_error() == 0
        e = relerr(dx.interior_nchw(), xr.grad)
        print(f"{name}: dgrad rel err {e:.3e}")
        assert e < 1e-3
        pv = dx.padded_view()
        assert float(pv[:, 0].abs().max()) == 0.0 and float(pv[:, :, 0].abs().max()) == 0.0
    # ---- wgrad
    c_dzw = (cout + 31) // 32 * 32
    dz_w = fill_nhwc(K, dzv, c_dzw, 0, L.PAD_ZERO, L.F32)
    dw = torch.zeros(cout, cin, k, k, device="cuda")
    K.conv_wgrad(xt, dz_w, dw, k, stride, pad)
    assert K.device_error() == 0
    e = relerr(dw, wr.grad)
    print(f"{name}: wgrad rel err {e:.3e}")
    assert e < 1e-4  # exact operands, fp32 accumulation in TMEM + fp32 atomics across the K split


def test_wgrad_slice_alpha_scale(K):
    """GAM fuse: only the first half of the input channels of a (C, 2C, 1, 1) weight gets a gradient."""
    from uegan_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(5)
    n, c, h, w = 2, 64, 16, 16
    x = tf32(torch.randn(n, c, h, w, device="cuda", generator=g))
    dz = tf32(torch.randn(n, c, h, w, device="cuda", generator=g))
    xt = fill_nhwc(K, x, c, 1, L.PAD_REFLECT, L.F32)
    dzt = fill_nhwc(K, dz, c, 0, L.PAD_ZERO, L.F32)
    dw = torch.zeros(c, 2 * c, 1, 1, device="cuda")
    alpha = torch.tensor([0.5], device="cuda")
    K.conv_wgrad(xt, dzt, dw, 1, 1, 0, cin_first=0, cin=c, alpha=alpha, scale=2.0)
    ref = torch.einsum("nohw,nchw->oc", dz.double(), x.double())
    assert relerr(dw[:, :c, 0, 0], ref) < 1e-4
    assert float(dw[:, c:].abs().max()) == 0.0


def test_elementwise_backward(K):
    from uegan_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(11)
    n, c, h, w = 2, 32, 12, 20
    # grad_combine with mask + mul + extra adds
    a = tf32(torch.randn(n, c, h + 2, w + 2, device="cuda", generator=g))
    b = tf32(torch.randn(n, 2 * c, h, w, device="cuda", generator=g))
    fwd = tf32(torch.randn(n, c, h, w, device="cuda", generator=g))
    mul = tf32(torch.randn(n, c, h, w, device="cuda", generator=g))
    ta = fill_nhwc(K, a, c, 0, L.PAD_ZERO, L.F32)
    tb = fill_nhwc(K, b, 2 * c, 0, L.PAD_ZERO, L.F32)
    tf_ = fill_nhwc(K, fwd, c, 1, L.PAD_REFLECT, L.F32)
    tm = fill_nhwc(K, mul, c, 1, L.PAD_REFLECT, L.F32)
    dst = K.NHWC(n, h, w, c, 2, L.F32, "cuda")
    K.grad_combine(dst, c, src_a=ta, pad_a=1, add_b=tb, b_c_off=c, mask=tf_, act=L.ACT_LRELU, mul=tm)
    xin = torch.zeros(n, c, h, w, device="cuda", dtype=torch.double, requires_grad=True)
    F.pad(xin, (1,) * 4, mode="reflect").backward(a.double())
    ref = (xin.grad + b[:, c:].double()) * mul.double() * torch.where(fwd > 0, 1.0, 0.2).double()
    assert relerr(dst.interior_nchw(), ref) < 1e-3
    assert float(dst.padded_view()[:, :2].abs().max()) == 0.0
    # channel sum
    out = torch.empty(c, device="cuda")
    K.channel_sum(tb, out, c_off=c)
    assert relerr(out, b[:, c:].double().sum(dim=(0, 2, 3))) < 1e-5
    # instance norm backward
    z = tf32(torch.randn(n, c, h, w, device="cuda", generator=g) * 2 + 0.5)
    dout = tf32(torch.randn(n, 2 * c, h, w, device="cuda", generator=g))
    tz = fill_nhwc(K, z, c, 0, L.PAD_ZERO, L.F32)
    td = fill_nhwc(K, dout, 2 * c, 0, L.PAD_ZERO, L.F32)
    ws = torch.empty(3 * n * c, dtype=torch.float64, device="cuda")
    mr = K.instance_norm_stats(tz, ws)
    zr = z.double().requires_grad_(True)
    F.instance_norm(zr, eps=1e-5).backward(dout[:, c:].double())
    for halo in (0, 2):
        dz = K.NHWC(n, h, w, c, halo, L.F32, "cuda")
        dz.buf.fill_(5.0)
        K.instance_norm_bwd(td, c, tz, mr, dz, torch.empty(2 * n * c, dtype=torch.float64, device="cuda"))
        assert relerr(dz.interior_nchw(), zr.grad) < 1e-3
        if halo:
            pv = dz.padded_view()
            assert float(pv[:, :halo].abs().max()) == 0.0 and float(pv[:, :, -halo:].abs().max()) == 0.0
    # upsample backward
    up = tf32(torch.randn(n, 2 * c, 2 * h, 2 * w, device="cuda", generator=g))
    tu = fill_nhwc(K, up, 2 * c, 0, L.PAD_ZERO, L.F32)
    ds = K.NHWC(n, h, w, c, 0, L.F32, "cuda")
    K.upsample2x_bwd(tu, c, ds)
    sr = torch.zeros(n, c, h, w, device="cuda", dtype=torch.double, requires_grad=True)
    F.interpolate(sr, scale_factor=2, mode="bilinear", align_corners=True).backward(up[:, c:].double())
    assert relerr(ds.interior_nchw(), sr.grad) < 1e-3
    # heads
    res = torch.tanh(torch.randn(n, 3, h, w, device="cuda", generator=g))
    xx = torch.rand(n, 3, h, w, device="cuda", generator=g) * 2 - 1
    do = torch.randn(n, 3, h, w, device="cuda", generator=g)
    dzt = K.NHWC(n, h, w, 4, 6, L.F32, "cuda")
    K.head_bwd(do, res, xx, 2, dzt)
    s = res + xx
    ref = do * ((s >= -1) & (s <= 1)).float() * (1 - res * res)
    assert relerr(dzt.interior_nchw()[:, :3], ref) < 1e-3
    assert float(dzt.padded_view()[:, :6].abs().max()) == 0.0 and float(dzt.interior_nchw()[:, 3].abs().max()) == 0.0
    assert K.device_error() == 0


HSTACK_CASES = [
    # n, cin, h, w, cout, k
    (2, 32, 24, 40, 3, 7),    # G's last conv
    (1, 32, 19, 29, 1, 7),    # ragged extents
    (2, 64, 16, 24, 1, 7),    # two channel chunks
    (1, 128, 16, 16, 1, 7),   # chunks split over two slices
    (2, 256, 12, 16, 1, 5),   # k5 head
    (1, 512, 8, 8, 1, 5),     # deepest head: 16 chunks, 6 slices
]


@pytest.mark.parametrize("case", HSTACK_CASES)
def test_wgrad_hstack(K, case):
    """Tiny-Cout weight gradient through the horizontally unrolled gradient (uegan_dz_hstack + vertical patch wgrad)."""
    from uegan_b200 import _lib as L
    n, cin, h, w, cout, k = case
    g = torch.Generator(device="cuda").manual_seed(123 + cin)
    pad = (k - 1) // 2
    x = tf32(torch.randn(n, cin, h, w, device="cuda", generator=g))
    wr = torch.zeros(cout, cin, k, k, device="cuda", dtype=torch.double, requires_grad=True)
    dzv = tf32(torch.randn(n, cout, h, w, device="cuda", generator=g))
    F.conv2d(F.pad(x.double(), (pad,) * 4, mode="reflect"), wr).backward(dzv.double())
    xt = fill_nhwc(K, x, cin, pad + 1, L.PAD_REFLECT, L.F32)  # halo > pad exercises the patch offset
    dz4 = fill_nhwc(K, dzv, 4, k - 1, L.PAD_ZERO, L.F32)
    e = K.NHWC(n, h, w + k - 1, 32, 0, L.F32, "cuda")
    e.buf.fill_(9.0)
    K.dz_hstack(dz4, cout, k, e)
    ev = e.interior_nchw()
    for s in range(k):
        for o in range(cout):
            assert torch.equal(ev[:, s * cout + o, :, s:s + w], dzv[:, o])
    assert float(ev[:, k * cout:].abs().max()) == 0.0
    dw = torch.zeros(cout, cin, k, k, device="cuda")
    alpha = torch.tensor([0.5], device="cuda")
    K.conv_wgrad_hstack(xt, e, dw, k, pad, alpha=alpha, scale=2.0)
    assert K.device_error() == 0
    err = relerr(dw, wr.grad)
    print(f"hstack wgrad {case}: rel err {err:.3e}")
    assert err < 1e-4


ZWIN_CASES = [
    # n, cin, h, w, cout, dz stored channels, k
    (2, 32, 24, 40, 3, 8, 7),      # G's last conv: dz is the 8-channel head gradient
    (1, 32, 19, 29, 1, 8, 7),      # ragged extents, D head
    (2, 256, 12, 16, 1, 8, 5),     # k5 head, four channel chunks
    (1, 512, 8, 8, 1, 8, 5),       # deepest head
    (2, 32, 40, 24, 32, 32, 3),    # dec5.0: N = 96
    (2, 64, 16, 32, 32, 32, 3),    # dec4
    (1, 128, 24, 16, 64, 64, 3),   # dec3: N = 192, two slices of channel chunks
]


@pytest.mark.parametrize("case", ZWIN_CASES)
def test_wgrad_zwin_f16(K, case):
    """Weight gradient with the horizontally stacked gradient read as a sliding TMA window over the zero-haloed dz
    (uegan_conv2d_wgrad_zwin) == torch's conv2d weight gradient on the same fp16-quantised operands; the alpha / scale
    factors and a per-tensor scale on both operands ride along."""
    from uegan_b200 import _lib as L
    n, cin, h, w, cout, cdz, k = case
    g = torch.Generator(device="cuda").manual_seed(321 + cin + cout)
    pad = (k - 1) // 2
    x = torch.randn(n, cin, h, w, device="cuda", generator=g).half().float()
    dzv = torch.randn(n, cout, h, w, device="cuda", generator=g).half().float()
    wr = torch.zeros(cout, cin, k, k, device="cuda", dtype=torch.double, requires_grad=True)
    F.conv2d(F.pad(x.double(), (pad,) * 4, mode="reflect"), wr).backward(dzv.double())
    sx, sz = torch.tensor([4.0], device="cuda"), torch.tensor([0.25], device="cuda")
    xt = K.NHWC(n, h, w, cin, pad + 1, L.F16, "cuda", zero=True, scale=sx)  # halo > pad exercises the patch offset
    xt.padded_view()[...] = (F.pad(x, (pad + 1,) * 4, mode="reflect") * 4.0).permute(0, 2, 3, 1).half()
    dzt = K.NHWC(n, h, w, cdz, k - 1, L.F16, "cuda", zero=True, scale=sz)
    dzt.padded_view()[:, k - 1:k - 1 + h, k - 1:k - 1 + w, :cout] = (dzv * 0.25).permute(0, 2, 3, 1).half()
    assert K.zwin_ok(cout, xt, dzt, k)
    dw = torch.full((cout, cin, k, k), 1.0, device="cuda")  # accumulates
    alpha = torch.tensor([0.5], device="cuda")
    K.conv_wgrad(xt, dzt, dw, k, 1, pad, alpha=alpha, scale=2.0, dz_zero_halo=True)
    assert K.device_error() == 0
    err = relerr(dw - 1.0, wr.grad)
    print(f"zwin wgrad {case}: rel err {err:.3e}")
    assert err < 1e-4
    # bit-reproducible, and equal to the generic path within fp32 summation order
    dw2 = torch.full((cout, cin, k, k), 1.0, device="cuda")
    K.conv_wgrad(xt, dzt, dw2, k, 1, pad, alpha=alpha, scale=2.0, dz_zero_halo=True)
    assert torch.equal(dw, dw2)


@pytest.mark.parametrize("shape", [(2, 32, 24, 40, 3), (2, 64, 12, 20, 1), (1, 32, 8, 7, 3), (2, 128, 9, 33, 2)])
def test_fold_inplace(K, shape):
    """In-place reflect-pad adjoint == grad_combine(src_a) == autograd of F.pad(reflect); halo zeroed."""
    from uegan_b200 import _lib as L
    n, c, h, w, pad = shape
    g = torch.Generator(device="cuda").manual_seed(31)
    a = tf32(torch.randn(n, c, h + 2 * pad, w + 2 * pad, device="cuda", generator=g))
    ta = fill_nhwc(K, a, c, 0, L.PAD_ZERO, L.F32)
    ref_t = K.NHWC(n, h, w, c, 0, L.F32, "cuda")
    K.grad_combine(ref_t, c, src_a=ta, pad_a=pad)
    v = K.fold_inplace(ta, pad)
    assert K.device_error() == 0
    assert (v.n, v.h, v.w, v.halo) == (n, h, w, pad)
    xin = torch.zeros(n, c, h, w, device="cuda", dtype=torch.double, requires_grad=True)
    F.pad(xin, (pad,) * 4, mode="reflect").backward(a.double())
    assert relerr(v.interior_nchw(), xin.grad) < 1e-3
    assert torch.equal(v.interior_nchw(), ref_t.interior_nchw())  # same summation order, same tf32 rounding
    pv = v.padded_view()
    assert float(pv[:, :pad].abs().max()) == 0.0 and float(pv[:, -pad:].abs().max()) == 0.0
    assert float(pv[:, :, :pad].abs().max()) == 0.0 and float(pv[:, :, -pad:].abs().max()) == 0.0
