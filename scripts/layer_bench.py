"""Per-layer timing + correctness of the weight-gradient and stride-2 data-gradient launches of one training step
(fp16 operands, B=16, 512x512 network): CUDA-event times (min / median of 7 after 2 warm-ups) next to the
algorithmic TFLOP/s, and the error against torch's fp32 convolution backward on the same fp16-quantised operands
at n=2.     python scripts/layer_bench.py [wgrad|dgrad|all] [name-filter]
Under `ncu --profile-from-start off` only the LAST timed launch of each case sits between cudaProfilerStart/Stop."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from uegan_b200 import _lib as L
from uegan_b200 import kernels as K

what = sys.argv[1] if len(sys.argv) > 1 else "all"
flt = sys.argv[2] if len(sys.argv) > 2 else ""
B = int(os.environ.get("LB_BATCH", "16"))
dev = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

# name, cin, h (= w) of x, cout, k, stride, launches per step
WGRAD = [
    ("G.enc1  3->32 k7s1", 3, 512, 32, 7, 1, 2), ("G.enc2 32->64 k3s2", 32, 512, 64, 3, 2, 2),
    ("G.enc3 64->128 k3s2", 64, 256, 128, 3, 2, 2), ("G.enc4 128->256 k3s2", 128, 128, 256, 3, 2, 2),
    ("G.enc5 256->512 k3s2", 256, 64, 512, 3, 2, 2), ("G.dec1 512->256 k3s1", 512, 64, 256, 3, 1, 2),
    ("G.dec2 256->128 k3s1", 256, 128, 128, 3, 1, 2), ("G.dec3 128->64 k3s1", 128, 256, 64, 3, 1, 2),
    ("G.dec4 64->32 k3s1", 64, 512, 32, 3, 1, 2), ("G.dec5.0 32->32 k3s1", 32, 512, 32, 3, 1, 2),
    ("G.up1 512->256 k1", 512, 32, 256, 1, 1, 2), ("G.ga 256->256 k1", 256, 64, 256, 1, 1, 2),
    ("G.ga1 32->32 k1", 32, 512, 32, 1, 1, 2), ("G.up4 64->32 k1", 64, 256, 32, 1, 1, 2),
    ("D.d1  3->32 k7s2", 3, 512, 32, 7, 2, 3), ("D.d2 32->64 k7s2", 32, 256, 64, 7, 2, 3),
    ("D.d3 64->128 k7s2", 64, 128, 128, 7, 2, 3), ("D.d4 128->256 k5s2", 128, 64, 256, 5, 2, 3),
    ("D.d5 256->512 k5s2", 256, 32, 512, 5, 2, 3),
    ("G.dec5.1 32->3 k7s1", 32, 512, 3, 7, 1, 2), ("D.p1 32->1 k7s1", 32, 256, 1, 7, 1, 3),
    ("D.p2 64->1 k7s1", 64, 128, 1, 7, 1, 3), ("D.p3 128->1 k7s1", 128, 64, 1, 7, 1, 3),
    ("D.p4 256->1 k5s1", 256, 32, 1, 5, 1, 3), ("D.p5 512->1 k5s1", 512, 16, 1, 5, 1, 3),
]
# stride-2 data gradients: name, cin (dx channels), h of x, cout (dz channels), k, launches per step
DGRAD = [
    ("G.enc2 dx32 k3s2", 32, 512, 64, 3, 2), ("G.enc3 dx64 k3s2", 64, 256, 128, 3, 2),
    ("G.enc4 dx128 k3s2", 128, 128, 256, 3, 2), ("G.enc5 dx256 k3s2", 256, 64, 512, 3, 2),
    ("D.d2 dx32 k7s2", 32, 256, 64, 7, 4), ("D.d3 dx64 k7s2", 64, 128, 128, 7, 4),
    ("D.d4 dx128 k5s2", 128, 64, 256, 5, 4), ("D.d5 dx256 k5s2", 256, 32, 512, 5, 4),
]


def fill(x_nchw, c_stored, halo, reflect):
    n, c, h, w = x_nchw.shape
    t = K.NHWC(n, h, w, c_stored, halo, L.F16, dev, zero=True)
    xp = F.pad(x_nchw, (halo,) * 4, mode="reflect" if reflect else "constant") if halo else x_nchw
    t.padded_view()[..., :c] = xp.permute(0, 2, 3, 1).half()
    return t


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.abs().max()), 1e-30))


def timeit(fn, reps=7):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for i in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if i == reps - 1:
            torch.cuda.profiler.start()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if i == reps - 1:
            torch.cuda.profiler.stop()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def run_wgrad():
    tot = 0.0
    for name, cin, h, cout, k, stride, per_step in WGRAD:
        if flt and flt not in name:
            continue
        pad = (k - 1) // 2
        g = torch.Generator(device=dev).manual_seed(7)
        cx = 8 if cin == 3 else cin
        zw = stride == 1 and k > 1          # dgrad operands: zero halo of k - 1 (the sliding-window stack path)
        cdz = 8 if cout <= 8 else (cout + 31) // 32 * 32
        zh = k - 1 if zw else 0
        halo = max(pad, 1)
        # correctness at n = 2 on a reduced extent (same channel structure, same code path)
        hs = min(h, 64)
        x = torch.randn(2, cin, hs, hs, device=dev, generator=g).half().float()
        ho = (hs + 2 * pad - k) // stride + 1
        dz = torch.randn(2, cout, ho, ho, device=dev, generator=g).half().float()
        xt, dzt = fill(x, cx, halo, True), fill(dz, cdz, zh, False)
        dw = torch.zeros(cout, cin, k, k, device=dev)
        K.conv_wgrad(xt, dzt, dw, k, stride, pad, dz_zero_halo=zw)
        xpad = F.pad(x, (pad,) * 4, mode="reflect") if pad else x
        ref = torch.nn.grad.conv2d_weight(xpad.double(), (cout, cin, k, k), dz.double(), stride=stride)
        err = relerr(dw, ref)
        assert K.device_error() == 0
        # timing at the step's shape
        xb = K.NHWC(B, h, h, cx, halo, L.F16, dev); xb.buf.normal_()
        hob = (h + 2 * pad - k) // stride + 1
        dzb = K.NHWC(B, hob, hob, cdz, zh, L.F16, dev, zero=True)
        dzb.padded_view()[:, zh:zh + hob, zh:zh + hob, :].normal_()
        dwb = torch.zeros(cout, cin, k, k, device=dev)
        tmin, tmed = timeit(lambda: K.conv_wgrad(xb, dzb, dwb, k, stride, pad, dz_zero_halo=zw))
        gf = 2.0 * B * hob * hob * cout * cin * k * k / 1e9
        tot += tmed * per_step
        print(f"wgrad {name:24s} err {err:.2e}  min {tmin:.3f} med {tmed:.3f} ms  {gf / tmed:7.1f} TF/s  x{per_step} = {tmed * per_step:.3f} ms", flush=True)
        del xb, dzb
    print(f"wgrad total per step (listed launches): {tot:.3f} ms")


def run_dgrad():
    tot = 0.0
    for name, cin, h, cout, k, per_step in DGRAD:
        if flt and flt not in name:
            continue
        pad = (k - 1) // 2
        kq = (k + 1) // 2
        g = torch.Generator(device=dev).manual_seed(9)
        hs = min(h, 64)
        ho = (hs + 2 * pad - k) // 2 + 1
        wgt = (torch.randn(cout, cin, k, k, device=dev, generator=g) / math.sqrt(cin * k * k)).half().float()
        dz = torch.randn(2, cout, ho, ho, device=dev, generator=g).half().float()
        dzt = fill(dz, cout, kq - 1, False)
        dxp = K.NHWC(2, hs + 2 * pad, hs + 2 * pad, cin, 0, L.F16, dev, zero=True)
        K.conv_dgrad(dzt, wgt, k, 2, dxp)
        ref = torch.nn.grad.conv2d_input((2, cin, hs + 2 * pad, hs + 2 * pad), wgt.double(), dz.double(), stride=2)
        got = dxp.padded_view().permute(0, 3, 1, 2).float()
        err = relerr(got, ref)
        assert K.device_error() == 0
        hob = (h + 2 * pad - k) // 2 + 1
        dzb = K.NHWC(B, hob, hob, cout, kq - 1, L.F16, dev, zero=True); dzb.buf.normal_()
        dxb = K.NHWC(B, h + 2 * pad, h + 2 * pad, cin, 0, L.F16, dev, zero=True)

        class Cache(dict):
            def get(self, key, param, fn):
                if key not in self:
                    self[key] = fn()
                return self[key]
        cache = Cache()
        tmin, tmed = timeit(lambda: K.conv_dgrad(dzb, wgt, k, 2, dxb, cache, "w"))
        gf = 2.0 * B * hob * hob * cout * cin * k * k / 1e9
        tot += tmed * per_step
        print(f"dgrad {name:24s} err {err:.2e}  min {tmin:.3f} med {tmed:.3f} ms  {gf / tmed:7.1f} TF/s  x{per_step} = {tmed * per_step:.3f} ms", flush=True)
        del dzb, dxb
    print(f"stride-2 dgrad total per step (listed launches): {tot:.3f} ms")


# stride-1 forward-kernel launches on low-channel full-resolution tensors: name, stored cin, h, cout, k, launches per step
FPROP = [
    ("G.enc1 8->32 k7", 8, 512, 32, 7, 2), ("G.dec5.0 32->32 k3", 32, 512, 32, 3, 2), ("G.dgrad dec4 32->64 k3", 32, 512, 64, 3, 2),
    ("G.dgrad dec5.1 8->32 k7", 8, 512, 32, 7, 2), ("VGG conv1_1 8->64 k3", 8, 512, 64, 3, 2), ("G.dec4 64->32 k3", 64, 512, 32, 3, 2),
    ("VGG conv1_2 64->64 k3", 64, 512, 64, 3, 2),
]


def run_fprop():
    tot = 0.0
    for name, cin, h, cout, k, per_step in FPROP:
        if flt and flt not in name:
            continue
        pad = (k - 1) // 2
        g = torch.Generator(device=dev).manual_seed(5)
        real_cin = 3 if cin == 8 else cin
        hs = 64
        x = torch.randn(2, real_cin, hs, hs, device=dev, generator=g).half().float()
        wgt = (torch.randn(cout, real_cin, k, k, device=dev, generator=g) / math.sqrt(real_cin * k * k)).half().float()
        xt = fill(x, cin, pad, True)
        y = K.NHWC(2, hs, hs, cout, 0, L.F16, dev, zero=True)
        wp = K.packed_weight(wgt, cin, L.F16)
        K.conv_fprop(xt, wp, cout, k, 1, pad, y, act=L.ACT_LRELU)
        ref = F.leaky_relu(F.conv2d(F.pad(x, (pad,) * 4, mode="reflect").double(), wgt.double()), 0.2)
        err = relerr(y.interior_nchw(), ref)
        assert K.device_error() == 0
        xb = K.NHWC(B, h, h, cin, pad, L.F16, dev); xb.buf.normal_()
        yb = K.NHWC(B, h, h, cout, 0, L.F16, dev)
        tmin, tmed = timeit(lambda: K.conv_fprop(xb, wp, cout, k, 1, pad, yb, act=L.ACT_LRELU))
        gf = 2.0 * B * h * h * cout * real_cin * k * k / 1e9
        gb = B * h * h * (cin + cout) * 2 / 1e9
        tot += tmed * per_step
        print(f"fprop {name:26s} err {err:.2e}  min {tmin:.3f} med {tmed:.3f} ms  {gf / tmed:7.1f} TF/s  {gb / tmed * 1e3:6.0f} GB/s  x{per_step} = {tmed * per_step:.3f} ms", flush=True)
        del xb, yb
    print(f"low-channel fprop-kernel total per step (listed launches): {tot:.3f} ms")


if what in ("fprop", "all"):
    run_fprop()
if what in ("wgrad", "all"):
    run_wgrad()
if what in ("dgrad", "all"):
    run_dgrad()
