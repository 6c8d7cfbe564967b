"""GPU parity tests of the individual sm_100a kernels (through the C ABI) against fp32 PyTorch ops.

Tolerances: kind::tf32 operands carry 10 mantissa bits (rounded to nearest at the producer) with fp32
accumulation -> ~1e-3 of the output scale per element for K up to a few thousand; bf16 (8 bits) -> ~1e-2.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from uegan_b200 import kernels
    from uegan_b200 import _lib
    _lib.load()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return kernels


def fill_nhwc(K, x_nchw, c_stored, halo, pad_mode, dtype):
    """test helper: NCHW torch tensor -> NHWC-with-halo buffer (halo written with torch ops)."""
    from uegan_b200 import _lib as L
    n, c, h, w = x_nchw.shape
    t = K.NHWC(n, h, w, c_stored, halo, dtype, x_nchw.device, zero=True)
    xp = x_nchw
    if halo:
        xp = F.pad(x_nchw, (halo,) * 4, mode="reflect" if pad_mode == L.PAD_REFLECT else "constant")
    t.padded_view()[..., :c] = xp.permute(0, 2, 3, 1).to(t.buf.dtype)
    return t


def tf32(x):
    """cvt.rna.tf32.f32 emulation: round to nearest (ties away) onto 10 mantissa bits."""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def quant(x, dtype):
    from uegan_b200 import _lib as L
    if dtype == L.F32:
        return tf32(x)
    return x.bfloat16().float() if dtype == L.BF16 else x.half().float()


def dt(name):
    from uegan_b200 import _lib as L
    return {"f32": L.F32, "bf16": L.BF16, "f16": L.F16}[name]


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.abs().max()), 1e-30))


CONV_CASES = [
    # (name, n, cin, h, w, cout, k, stride, pad_mode)
    ("enc1_like", 2, 3, 32, 48, 32, 7, 1, "reflect"),
    ("enc2_like", 2, 32, 32, 32, 64, 3, 2, "reflect"),
    ("enc5_like", 2, 256, 16, 16, 512, 3, 2, "reflect"),
    ("dec1_like", 1, 512, 16, 16, 256, 3, 1, "reflect"),
    ("dec4_like", 1, 64, 64, 64, 32, 3, 1, "reflect"),
    ("fuse_1x1", 2, 128, 16, 24, 64, 1, 1, "reflect"),
    ("d1_like", 2, 3, 96, 96, 32, 7, 2, "reflect"),
    ("d2_like", 2, 32, 48, 48, 64, 7, 2, "reflect"),
    ("d4_like", 2, 128, 12, 12, 256, 5, 2, "reflect"),
    ("d5_like_tiny", 2, 256, 6, 6, 512, 5, 2, "reflect"),
    ("vgg_like", 2, 64, 32, 32, 64, 3, 1, "zero"),
    ("vgg_first", 2, 3, 32, 32, 64, 3, 1, "zero"),
    ("ragged", 3, 32, 40, 24, 48, 3, 1, "reflect"),
    ("tiny_2x2", 3, 64, 4, 4, 32, 3, 2, "reflect"),
    # patch mode with STREAMED weights (weights do not fit next to >= 3 patch stages)
    ("stream_128_64", 2, 128, 32, 24, 64, 3, 1, "reflect"),
    ("stream_nsplit", 2, 256, 16, 16, 256, 3, 1, "zero"),      # few tiles: N split 256 -> 64, four N tiles per patch
    ("stream_vgg2_1", 1, 64, 32, 32, 128, 3, 1, "zero"),
    ("stream_k5", 1, 64, 20, 28, 48, 5, 1, "reflect"),
]


@pytest.mark.parametrize("dtype_name", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_fprop(K, case, dtype_name, monkeypatch):
    from uegan_b200 import _lib as L
    name, n, cin, h, w, cout, k, stride, pad_mode = case
    if name.startswith("stream_"):
        monkeypatch.setenv("UEGAN_STREAM_MAXN", "128")  # the streamed-weight patch mode is opt-in (read per launch)
    dtype = dt(dtype_name)
    pm = L.PAD_REFLECT if pad_mode == "reflect" else L.PAD_ZERO
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = quant(torch.randn(n, cin, h, w, device="cuda", generator=g), dtype)
    wgt = quant(torch.randn(cout, cin, k, k, device="cuda", generator=g) / math.sqrt(cin * k * k), dtype)
    bias = torch.randn(cout, device="cuda", generator=g) * 0.1
    pad = (k - 1) // 2
    vec = 4 if dtype == L.F32 else 8
    c_stored = (cin + vec - 1) // vec * vec
    xt = fill_nhwc(K, x, c_stored, pad, pm, dtype)
    ho = (h + 2 * pad - k) // stride + 1
    wo = (w + 2 * pad - k) // stride + 1
    y = K.NHWC(n, ho, wo, cout + 16, 1, dtype, "cuda", zero=True)  # written into a channel slice at offset 16
    wp = K.packed_weight(wgt, c_stored, dtype)
    K.conv_fprop(xt, wp, cout, k, stride, pad, y, 16, bias, None, L.ACT_LRELU)
    assert K.device_error() == 0
    xpad = F.pad(x, (pad,) * 4, mode="reflect" if pad_mode == "reflect" else "constant") if pad else x
    ref = F.leaky_relu(F.conv2d(xpad.double(), wgt.double(), bias.double(), stride=stride), 0.2).float()
    got = y.interior_nchw()[:, 16:]
    # operands are exactly representable, accumulation is fp32: what is left is the output rounding to the
    # storage format (tf32: 2^-11, bf16: 2^-8 of the value) and fp32 summation order
    tol = 5e-3 if dtype == L.BF16 else 6e-4
    assert relerr(got, ref) < tol, f"{name}: rel err {relerr(got, ref):.3e}"
    assert float(y.interior_nchw()[:, :16].abs().max()) == 0.0  # neighbouring slice untouched
    assert float(y.padded_view()[:, 0].abs().max()) == 0.0  # halo untouched by the conv


def test_conv_planar_head(K):
    """cout=1 prediction head with tanh and cout=3 residual+clamp head (planar fp32 NCHW epilogue)."""
    from uegan_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(7)
    n, cin, h, w = 2, 32, 24, 40
    x = tf32(torch.randn(n, cin, h, w, device="cuda", generator=g))
    for cout, k, resid in ((1, 7, False), (3, 7, True), (1, 5, False)):
        wgt = tf32(torch.randn(cout, cin, k, k, device="cuda", generator=g) / math.sqrt(cin * k * k))
        bias = torch.randn(cout, device="cuda", generator=g) * 0.1
        pad = (k - 1) // 2
        xt = fill_nhwc(K, x, cin, pad, L.PAD_REFLECT, L.F32)
        out = torch.zeros(n, cout, h, w, device="cuda")
        res = torch.rand(n, cout, h, w, device="cuda", generator=g) * 2 - 1 if resid else None
        K.conv_fprop(xt, K.packed_weight(wgt, cin, L.F32), cout, k, 1, pad, None, 0, bias, None, L.ACT_TANH, None,
                     out, res)
        assert K.device_error() == 0
        ref = torch.tanh(F.conv2d(F.pad(x, (pad,) * 4, mode="reflect").double(), wgt.double(), bias.double())).float()
        if resid:
            ref = torch.clamp(ref + res, -1, 1)
        assert relerr(out, ref) < 2e-5  # planar fp32 output is not rounded: fp32 accumulation error only


def test_conv_alpha_and_mul(K):
    from uegan_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(8)
    n, cin, h, w, cout = 2, 64, 16, 16, 32
    x = tf32(torch.randn(n, cin, h, w, device="cuda", generator=g))
    m = tf32(torch.randn(n, cout, h, w, device="cuda", generator=g))
    wgt = tf32(torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / math.sqrt(cin * 9))
    bias = torch.randn(cout, device="cuda", generator=g) * 0.1
    alpha = torch.tensor([0.37], device="cuda")
    xt = fill_nhwc(K, x, cin, 1, L.PAD_REFLECT, L.F32)
    mt = fill_nhwc(K, m, cout, 1, L.PAD_REFLECT, L.F32)
    y = K.NHWC(n, h, w, cout, 0, L.F32, "cuda", zero=True)
    K.conv_fprop(xt, K.packed_weight(wgt, cin, L.F32), cout, 3, 1, 1, y, 0, bias, alpha, L.ACT_LRELU, mt)
    assert K.device_error() == 0
    ref = F.leaky_relu(0.37 * F.conv2d(F.pad(x, (1,) * 4, mode="reflect"), wgt) + bias.view(1, -1, 1, 1), 0.2) * m
    assert relerr(y.interior_nchw(), ref) < 6e-4


@pytest.mark.parametrize("dtype_name", ["f32", "bf16", "f16"])
def test_elementwise(K, dtype_name):
    from uegan_b200 import _lib as L
    dtype = dt(dtype_name)
    tol = 1e-2 if dtype == L.BF16 else 1e-3
    g = torch.Generator(device="cuda").manual_seed(3)
    # pack_input (reflect halo, affine) ------------------------------------------------
    x = torch.rand(2, 3, 20, 28, device="cuda", generator=g) * 2 - 1
    t = K.NHWC(2, 20, 28, 4 if dtype == L.F32 else 8, 3, dtype, "cuda", zero=True)
    K.pack_input(x, t, L.PAD_REFLECT, scale=[0.5, 0.25, 2.0], shift=[0.1, -0.2, 0.3])
    ref = x * torch.tensor([0.5, 0.25, 2.0], device="cuda").view(1, 3, 1, 1) + \
        torch.tensor([0.1, -0.2, 0.3], device="cuda").view(1, 3, 1, 1)
    refp = F.pad(ref, (3,) * 4, mode="reflect").permute(0, 2, 3, 1)
    assert relerr(t.padded_view()[..., :3].float(), refp) < tol
    assert float(t.padded_view()[..., 3:].abs().max()) == 0.0
    t2 = K.NHWC(2, 20, 28, 4 if dtype == L.F32 else 8, 1, dtype, "cuda", zero=True)
    K.pack_input(x, t2, L.PAD_ZERO)
    assert relerr(t2.padded_view()[..., :3].float(), F.pad(x, (1,) * 4).permute(0, 2, 3, 1)) < tol
    # halo fill ------------------------------------------------------------------------
    a = quant(torch.randn(2, 32, 12, 20, device="cuda", generator=g), dtype)
    for halo in (1, 2, 3):
        ta = fill_nhwc(K, a, 32, 0, L.PAD_REFLECT, dtype)
        tb = K.NHWC(2, 12, 20, 32, halo, dtype, "cuda", zero=True)
        tb.padded_view()[:, halo:halo + 12, halo:halo + 20] = ta.padded_view()
        K.halo_fill(tb, L.PAD_REFLECT)
        refh = F.pad(ta.interior_nchw(), (halo,) * 4, mode="reflect").permute(0, 2, 3, 1)
        assert relerr(tb.padded_view().float(), refh) < 1e-6
        K.halo_fill(tb, L.PAD_ZERO)
        refz = F.pad(ta.interior_nchw(), (halo,) * 4).permute(0, 2, 3, 1)
        assert relerr(tb.padded_view().float(), refz) < 1e-6
    # instance norm into a channel slice ------------------------------------------------
    for c in (32, 128, 512):
        a = torch.randn(2, c, 10, 14, device="cuda", generator=g) * 3 + 1.5
        ta = fill_nhwc(K, a, c, 1, L.PAD_REFLECT, dtype)
        td = K.NHWC(2, 10, 14, 2 * c, 1, dtype, "cuda", zero=True)
        ws = torch.empty(3 * 2 * c, dtype=torch.float64, device="cuda")
        K.instance_norm(ta, td, c, ws)
        refn = F.instance_norm(ta.interior_nchw(), eps=1e-5)
        assert relerr(td.interior_nchw()[:, c:], refn) < tol
        assert float(td.interior_nchw()[:, :c].abs().max()) == 0.0
    # eps-dominated regime (variance << eps, models.py:227 with orthogonal(0.02) init)
    a = torch.randn(1, 32, 8, 8, device="cuda", generator=g) * 1e-8
    ta = fill_nhwc(K, a, 32, 0, L.PAD_REFLECT, L.F32)
    td = K.NHWC(1, 8, 8, 32, 0, L.F32, "cuda", zero=True)
    K.instance_norm(ta, td, 0, torch.empty(3 * 32, dtype=torch.float64, device="cuda"))
    assert relerr(td.interior_nchw(), F.instance_norm(ta.interior_nchw(), eps=1e-5)) < 2e-3
    # bilinear x2 align_corners=True ---------------------------------------------------------
    a = quant(torch.randn(2, 64, 6, 10, device="cuda", generator=g), dtype)
    ta = fill_nhwc(K, a, 64, 0, L.PAD_REFLECT, dtype)
    td = K.NHWC(2, 12, 20, 128, 1, dtype, "cuda", zero=True)
    K.upsample2x(ta, td, 64)
    refu = F.interpolate(ta.interior_nchw(), scale_factor=2, mode="bilinear", align_corners=True)
    assert relerr(td.interior_nchw()[:, 64:], refu) < tol
    # max pool --------------------------------------------------------------------------
    ta = fill_nhwc(K, a, 64, 1, L.PAD_ZERO, dtype)
    td = K.NHWC(2, 3, 5, 64, 1, dtype, "cuda", zero=True)
    K.maxpool2x2(ta, td)
    assert relerr(td.interior_nchw(), F.max_pool2d(ta.interior_nchw(), 2, 2)) < 1e-6
    assert relerr(K.unpack_nchw(ta, 8, 16), ta.interior_nchw()[:, 8:24]) < 1e-6
    assert K.device_error() == 0


@pytest.mark.parametrize("case", [
    # (n, cin, h, w, cout, k, residual)
    (2, 32, 24, 40, 3, 7, True),     # G's last conv: N = 21 -> 32, ragged 26-column tiles, 16-row tiles
    (1, 32, 37, 61, 1, 7, False),    # odd extents: partial tiles in both directions
    (2, 64, 16, 32, 1, 7, False),    # two channel chunks per patch
    (1, 128, 20, 28, 1, 7, False),   # four chunks
    (2, 256, 12, 36, 1, 5, False),   # k5 head, weights force the 8-row tile variant
    (1, 32, 8, 8, 1, 3, False),      # k3, image smaller than a tile
])
def test_conv_rowsum(K, case):
    """Row-sum kernel (csrc/conv_rowsum.cu) vs fp64 conv2d on tf32-representable operands, with alpha / bias / tanh /
    residual + clamp / aux exactly as the generic planar epilogue."""
    from uegan_b200 import _lib as L
    n, cin, h, w, cout, k, resid = case
    g = torch.Generator(device="cuda").manual_seed(100 + cin + k)
    assert K.rowsum_supported(cout, cin, k, L.F32)
    x = tf32(torch.randn(n, cin, h, w, device="cuda", generator=g))
    wgt = tf32(torch.randn(cout, cin, k, k, device="cuda", generator=g) / math.sqrt(cin * k * k))
    bias = torch.randn(cout, device="cuda", generator=g) * 0.1
    alpha = torch.tensor([0.73], device="cuda")
    pad = (k - 1) // 2
    xt = fill_nhwc(K, x, cin, pad + 1, L.PAD_REFLECT, L.F32)  # halo larger than pad: exercises patch_off
    out = torch.full((n, cout, h, w), 7.0, device="cuda")
    aux = torch.full((n, cout, h, w), 7.0, device="cuda")
    res = torch.rand(n, cout, h, w, device="cuda", generator=g) * 2 - 1 if resid else None

    class Cache:
        def get(self, key, param, fn):
            return fn()
    K.conv_planar(xt, wgt, Cache(), "t", k, pad, bias, alpha, L.ACT_TANH, out, res, aux)
    assert K.device_error() == 0
    ref = torch.tanh(0.73 * F.conv2d(F.pad(x, (pad,) * 4, mode="reflect").double(), wgt.double())
                     + bias.double().view(1, -1, 1, 1)).float()
    assert relerr(aux, ref) < 2e-5
    if resid:
        ref = torch.clamp(ref + res, -1, 1)
    assert relerr(out, ref) < 2e-5


def test_in_mse_bwd_direct(K):
    """Gradient of weight * mean((IN(x) - IN(y))^2) w.r.t. fp16 features x (+ deep gradient, ReLU mask), zero halo."""
    from uegan_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(21)
    n, c, h, w = 2, 64, 12, 20
    x = torch.relu(torch.randn(n, c, h, w, device="cuda", generator=g) + 0.3).half().float()
    y = torch.relu(torch.randn(n, c, h, w, device="cuda", generator=g) + 0.3).half().float()
    deep = (torch.randn(n, c, h, w, device="cuda", generator=g) * 0.1).half().float()
    tx = fill_nhwc(K, x, c, 1, L.PAD_ZERO, L.F16)
    ty = fill_nhwc(K, y, c, 1, L.PAD_ZERO, L.F16)
    tdeep = fill_nhwc(K, deep, c, 1, L.PAD_ZERO, L.F16)
    wsx = torch.empty(3 * n * c, dtype=torch.float64, device="cuda")
    wsy = torch.empty(3 * n * c, dtype=torch.float64, device="cuda")
    mx, my = K.instance_norm_stats(tx, wsx), K.instance_norm_stats(ty, wsy)
    weight = 50.0
    gs = torch.tensor([0.5], device="cuda")
    for dp in (None, tdeep):
        dx = K.NHWC(n, h, w, c, 1, L.F16, "cuda")
        dx.buf.fill_(3.0)
        K.in_mse_bwd(tx, ty, mx, my, weight, gs, dp, dx, torch.empty(2 * n * c, dtype=torch.float64, device="cuda"))
        assert K.device_error() == 0
        xr = x.double().requires_grad_(True)
        loss = weight * 0.5 * F.mse_loss(F.instance_norm(xr, eps=1e-5), F.instance_norm(y.double(), eps=1e-5))
        loss.backward()
        ref = xr.grad
        if dp is not None:
            ref = ref + deep.double()
        ref = torch.where(x > 0, ref, torch.zeros_like(ref))
        assert relerr(dx.interior_nchw(), ref) < 2e-3
        pv = dx.padded_view().float()
        assert float(pv[:, 0].abs().max()) == 0.0 and float(pv[:, :, 0].abs().max()) == 0.0
        assert float(pv[:, -1].abs().max()) == 0.0 and float(pv[:, :, -1].abs().max()) == 0.0


@pytest.mark.parametrize("case", [(2, 32, 24, 40, 32, 3, 1, 3), (1, 8, 32, 48, 32, 7, 1, 1), (2, 32, 32, 32, 64, 3, 2, 1),
                                  (1, 64, 16, 24, 32, 3, 1, 2), (1, 128, 40, 24, 256, 5, 2, 2)])
def test_conv_epilogue_reflect_halo(K, case):
    """uegan_conv_desc.y_reflect_halo: the conv epilogue writes the reflection-padding halo of its output == the same conv
    followed by uegan_halo_fill(REFLECT), bit for bit over the whole padded tensor (every tile mode: RGB / 64-byte /
    128-byte patch, plain stride 2, N tiles > 1)."""
    from uegan_b200 import _lib as L
    n, cin, h, w, cout, k, stride, yhalo = case
    g = torch.Generator(device="cuda").manual_seed(17 + cin + cout)
    pad = (k - 1) // 2
    real_cin = 3 if cin == 8 else cin
    x = torch.randn(n, real_cin, h, w, device="cuda", generator=g).half().float()
    wgt = (torch.randn(cout, real_cin, k, k, device="cuda", generator=g) / math.sqrt(real_cin * k * k)).half().float()
    bias = torch.randn(cout, device="cuda", generator=g)
    xt = fill_nhwc(K, x, cin, pad, L.PAD_REFLECT, L.F16)
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    wp = K.packed_weight(wgt, cin, L.F16)
    ya = K.NHWC(n, ho, wo, cout, yhalo, L.F16, "cuda", zero=True)
    yb = K.NHWC(n, ho, wo, cout, yhalo, L.F16, "cuda", zero=True)
    K.conv_fprop(xt, wp, cout, k, stride, pad, ya, bias=bias, act=L.ACT_LRELU, reflect_halo=True)
    K.conv_fprop(xt, wp, cout, k, stride, pad, yb, bias=bias, act=L.ACT_LRELU)
    K.halo_fill(yb, L.PAD_REFLECT)
    assert K.device_error() == 0
    assert torch.equal(ya.padded_view(), yb.padded_view())
    ref = F.leaky_relu(F.conv2d(F.pad(x, (pad,) * 4, mode="reflect").double(), wgt.double(), bias.double(), stride=stride), 0.2)
    assert relerr(ya.interior_nchw(), ref) < 1e-3


@pytest.mark.parametrize("shape", [(2, 32, 12, 20), (1, 128, 9, 7)])
def test_cat_build(K, shape):
    """uegan_cat_build (one-pass decoder concat, opt-in) == uegan_upsample2x into the first half + uegan_instance_norm into
    the second half, bit for bit, and both == torch's bilinear x2 (align_corners) / instance_norm."""
    from uegan_b200 import _lib as L
    n, c, h, w = shape
    g = torch.Generator(device="cuda").manual_seed(5 + c)
    u = torch.randn(n, c, h, w, device="cuda", generator=g).half().float()
    z = (torch.randn(n, c, 2 * h, 2 * w, device="cuda", generator=g) * 3 + 1).half().float()
    su, sz, sd = (torch.tensor([v], device="cuda") for v in (2.0, 0.5, 4.0))
    tu = K.NHWC(n, h, w, c, 0, L.F16, "cuda", zero=True, scale=su)
    tu.padded_view()[...] = (u * 2.0).permute(0, 2, 3, 1).half()
    tz = K.NHWC(n, 2 * h, 2 * w, c, 0, L.F16, "cuda", zero=True, scale=sz)
    tz.padded_view()[...] = (z * 0.5).permute(0, 2, 3, 1).half()
    a = K.NHWC(n, 2 * h, 2 * w, 2 * c, 1, L.F16, "cuda", zero=True, scale=sd)
    b = K.NHWC(n, 2 * h, 2 * w, 2 * c, 1, L.F16, "cuda", zero=True, scale=sd)
    st = torch.empty(3 * n * c, dtype=torch.float64, device="cuda")
    K.cat_build(tu, tz, st, a)
    K.upsample2x(tu, b, 0)
    K.instance_norm(tz, b, c, st)
    assert K.device_error() == 0
    assert torch.equal(a.padded_view(), b.padded_view())
    ref = torch.cat([F.interpolate(u.double(), scale_factor=2, mode="bilinear", align_corners=True),
                     F.instance_norm(z.double(), eps=1e-5)], 1)
    assert relerr(a.interior_nchw(), ref) < 2e-3


@pytest.mark.parametrize("shape", [(2, 64, 12, 20), (1, 128, 33, 17), (2, 512, 4, 4)])
def test_in_mse_joint(K, shape):
    """The joint tap pass (five raw moments per (n, c) in one read of x and y) == torch's
    weight * mse(instance_norm(x), instance_norm(y)) and, through uegan_in_mse_bwd_apply, its gradient; channels with a
    tiny variance next to a large mean (the conditioning risk of raw moments) and dead channels included."""
    from uegan_b200 import _lib as L
    n, c, h, w = shape
    g = torch.Generator(device="cuda").manual_seed(77 + c)
    x = torch.relu(torch.randn(n, c, h, w, device="cuda", generator=g) + 0.3)
    y = torch.relu(torch.randn(n, c, h, w, device="cuda", generator=g) + 0.3)
    x[:, 1] = 3.0 + 0.05 * x[:, 1]   # low variance on a large mean
    y[:, 1] = 2.0 + 0.05 * y[:, 1]
    x[:, 2] = 0.0                    # a dead channel in one map, in both
    x[:, 3] = 0.0
    y[:, 3] = 0.0
    x, y = x.half().float(), y.half().float()
    tx = fill_nhwc(K, x, c, 1, L.PAD_ZERO, L.F16)
    ty = fill_nhwc(K, y, c, 1, L.PAD_ZERO, L.F16)
    weight, eps = 0.75, 1e-5
    ws = torch.empty(9 * n * c, dtype=torch.float64, device="cuda")
    accum = torch.zeros(1, dtype=torch.float64, device="cuda")
    loss = torch.full((1,), 2.0, dtype=torch.float32, device="cuda")  # accumulates
    mx, my, sm = K.in_mse_joint(tx, ty, eps, weight, ws, accum, loss)
    assert K.device_error() == 0
    xr = x.double().requires_grad_(True)
    ref = weight * F.mse_loss(F.instance_norm(xr, eps=eps), F.instance_norm(y.double(), eps=eps))
    ref.backward()
    e = abs(float(loss[0]) - 2.0 - float(ref)) / float(ref)
    print(f"in_mse_joint {shape}: loss rel err {e:.2e}")
    assert e < 2e-5
    assert float(accum[0]) == 0.0  # scratch handed back zeroed
    # the separate-pass path computes the same numbers
    wsx = torch.empty(3 * n * c, dtype=torch.float64, device="cuda")
    wsy = torch.empty(3 * n * c, dtype=torch.float64, device="cuda")
    mx2, my2 = K.instance_norm_stats(tx, wsx, eps=eps), K.instance_norm_stats(ty, wsy, eps=eps)
    loss2 = torch.zeros(1, dtype=torch.float32, device="cuda")
    K.in_mse_fwd(tx, ty, mx2, my2, weight, accum, loss2)
    assert abs(float(loss2[0]) - (float(loss[0]) - 2.0)) / float(ref) < 2e-5
    # gradient: joint sums + apply  vs  autograd, and vs the statistics-pass variant
    scale = 4096.0
    gs = torch.tensor([0.5], device="cuda")
    dx = K.NHWC(n, h, w, c, 1, L.F16, "cuda", zero=True)
    K.in_mse_bwd_apply(tx, ty, mx, my, weight * scale, gs, None, dx, sm)
    dx2 = K.NHWC(n, h, w, c, 1, L.F16, "cuda", zero=True)
    K.in_mse_bwd(tx, ty, mx2, my2, weight * scale, gs, None, dx2, torch.empty(2 * n * c, dtype=torch.float64, device="cuda"))
    assert K.device_error() == 0
    refg = torch.where(x > 0, xr.grad * (scale * 0.5), torch.zeros_like(xr.grad))
    eg = relerr(dx.interior_nchw(), refg)
    eg2 = relerr(dx2.interior_nchw(), refg)
    print(f"in_mse_joint {shape}: grad rel err {eg:.2e} (separate passes {eg2:.2e})")
    assert eg < 2e-3 and eg <= 1.5 * eg2 + 1e-4


@pytest.mark.parametrize("case", [("f32", 64, 32, 1), ("f32", 32, 64, 3), ("f16", 64, 128, 3)])
def test_conv_fused_instance_norm_stats(K, case):
    """Per-(n, c) sum / sum of squares accumulated by the conv epilogue (in_stats), consumed by instance_norm_apply: the
    opt-in alternative (UEGAN_FUSED_STATS) to a separate statistics pass."""
    from uegan_b200 import _lib as L
    dtype_name, cin, cout, k = case
    dtype = dt(dtype_name)
    g = torch.Generator(device="cuda").manual_seed(17)
    n, h, w = 2, 16, 24
    pad = (k - 1) // 2
    x = quant(torch.randn(n, cin, h, w, device="cuda", generator=g), dtype)
    wgt = quant(torch.randn(cout, cin, k, k, device="cuda", generator=g) / math.sqrt(cin * k * k), dtype)
    xt = fill_nhwc(K, x, cin, pad, L.PAD_ZERO, dtype)
    z = K.NHWC(n, h, w, cout, 0, dtype, "cuda")
    dst = K.NHWC(n, h, w, cout, 0, dtype, "cuda")
    stats = torch.empty(3 * n * cout, dtype=torch.float64, device="cuda")
    K.conv_fprop(xt, K.packed_weight(wgt, cin, dtype), cout, k, 1, pad, z, in_stats=stats)
    K.instance_norm_apply(z, dst, 0, stats)
    assert K.device_error() == 0
    ref = F.instance_norm(z.interior_nchw(), eps=1e-5)  # statistics of the STORED (rounded) conv output
    assert relerr(dst.interior_nchw(), ref) < (2e-3 if dtype != L.F32 else 1e-3)


@pytest.mark.parametrize("case", [
    # (n, cin, h, w, cout, k)
    (2, 32, 24, 40, 3, 7),      # G's last conv: 32 stored channels = half of a 64-channel fp16 chunk (zero-filled)
    (2, 64, 16, 32, 1, 7),
    (2, 256, 12, 36, 1, 5),
    (2, 256, 8, 8, 1, 5),       # the p4 head at 128x128 input
    (2, 512, 16, 16, 1, 5),     # the deepest head (fits in fp16 only): 8 chunks, 8-row tiles
    (2, 512, 4, 4, 1, 5),       # ... at 128x128 input: image smaller than one tile
])
def test_conv_rowsum_f16(K, case):
    """fp16 row-sum kernel with scaled operands (x stored = 2^a x, packed w = 2^b w) vs fp64 conv2d on fp16-representable
    operands: the planar output is the TRUE value."""
    from uegan_b200 import _lib as L
    n, cin, h, w, cout, k = case
    g = torch.Generator(device="cuda").manual_seed(200 + cin + k + h)
    assert K.rowsum_supported(cout, cin, k, L.F16)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g).half().float()
    wgt = (torch.randn(cout, cin, k, k, device="cuda", generator=g) / math.sqrt(cin * k * k)).half().float()
    bias = torch.randn(cout, device="cuda", generator=g) * 0.1
    alpha = torch.tensor([0.73], device="cuda")
    pad = (k - 1) // 2
    sx = torch.tensor([4.0], device="cuda")
    sw = torch.tensor([16.0], device="cuda")
    xt = K.NHWC(n, h, w, cin, pad, L.F16, "cuda", zero=True, scale=sx)
    xt.padded_view()[..., :cin] = (F.pad(x, (pad,) * 4, mode="reflect") * 4.0).permute(0, 2, 3, 1).half()
    out = torch.full((n, cout, h, w), 7.0, device="cuda")

    class Cache:
        def get(self, key, param, fn):
            return fn()
    K.conv_planar(xt, wgt, Cache(), "t", k, pad, bias, alpha, L.ACT_TANH, out, None, None, w_scale=sw)
    assert K.device_error() == 0
    ref = torch.tanh(0.73 * F.conv2d(F.pad(x, (pad,) * 4, mode="reflect").double(), wgt.double())
                     + bias.double().view(1, -1, 1, 1)).float()
    e = relerr(out, ref)
    print(f"rowsum f16 {case}: rel err {e:.3e}")
    assert e < 2e-5
