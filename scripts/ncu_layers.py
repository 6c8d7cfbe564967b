"""Representative single launches of each GEMM kernel family at the bench shapes (B=16, 512x512 network), for
    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/layers python scripts/ncu_layers.py
Each layer runs twice untimed (warm-up: tensor maps, attributes), then once between cudaProfilerStart/Stop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uegan_b200 import _lib as L
from uegan_b200 import kernels as K

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = "cuda"


def T(h, w, c, halo=0, dtype=L.F32):
    t = K.NHWC(B, h, w, c, halo, dtype, dev)
    t.buf.normal_()
    return t


def W(cout, cin, k):
    return torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5


class NoCache:
    def get(self, key, param, fn):
        return fn()


layers = []

# 1. d3's data gradient, one parity class: plain mode, N = 64, k4, C = 128, dz 64x64   ("dgrad tf32 128->64 k4s1 @64")
dz3, w3 = T(64, 64, 128, 3), W(128, 64, 7)
dxp3 = T(128 + 6, 128 + 6, 64)
wp3 = K.packed_weight_dgrad(w3, 128, L.F32, 2, 0, 0)
layers.append(("dgrad_d3_plain_n64", lambda: K.conv_generic(dz3, wp3, 64, 4, 1, 3, dxp3, y_mul=2, real_taps=16)))
# 2. G dec4 forward: resident-weight patch mode, 64 -> 32 k3 @512
cat4, w4, y4 = T(512, 512, 64, 1), W(32, 64, 3), T(512, 512, 32)
wp4 = K.packed_weight(w4, 64, L.F32)
layers.append(("fprop_dec4_patch_n32", lambda: K.conv_fprop(cat4, wp4, 32, 3, 1, 1, y4, act=L.ACT_LRELU)))
# 3. G's last conv: row-sum kernel, 32 -> 3 k7 @512
t5, w5 = T(512, 512, 32, 3), W(3, 32, 7)
out5 = torch.empty(B, 3, 512, 512, device=dev)
layers.append(("fprop_last_rowsum", lambda: K.conv_planar(t5, w5, NoCache(), "x", 7, 3, None, None, L.ACT_TANH, out5)))
# 4. d2's weight gradient: generic kernel, 32 -> 64 k7 s2, x @256
x2, dz2 = T(256, 256, 32, 3), T(128, 128, 64)
dw2 = torch.zeros(64, 32, 7, 7, device=dev)
layers.append(("wgrad_d2_generic_k7s2", lambda: K.conv_wgrad(x2, dz2, dw2, 7, 2, 3)))
# 5. dec4's weight gradient: patch kernel, 64 -> 32 k3 @512
dzy4 = T(512, 512, 32)
dw4 = torch.zeros(32, 64, 3, 3, device=dev)
layers.append(("wgrad_dec4_patch", lambda: K.conv_wgrad(cat4, dzy4, dw4, 3, 1, 1)))
# 6. VGG conv2_1: fp16, 64 -> 128 k3 @256 (resident weights leave two patch stages)
v1, wv, v2 = T(256, 256, 64, 1, L.F16), W(128, 64, 3), T(256, 256, 128, 1, L.F16)
wpv = K.packed_weight(wv, 64, L.F16)
layers.append(("fprop_vgg2_1_f16_n128", lambda: K.conv_fprop(v1, wpv, 128, 3, 1, 1, v2, act=L.ACT_RELU)))
# 7. VGG conv3_2: fp16, 256 -> 256 k3 @128, plain mode N = 256 (the fast case, for contrast)
v3, wv3, v4 = T(128, 128, 256, 1, L.F16), W(256, 256, 3), T(128, 128, 256, 1, L.F16)
wpv3 = K.packed_weight(wv3, 256, L.F16)
layers.append(("fprop_vgg3_2_f16_n256", lambda: K.conv_fprop(v3, wpv3, 256, 3, 1, 1, v4, act=L.ACT_RELU)))
# 8. G enc2 forward: stride 2, 32 -> 64 k3, x @512 (plain mode)
x1, we2, x2o = T(512, 512, 32, 1), W(64, 32, 3), T(256, 256, 64, 1)
wpe2 = K.packed_weight(we2, 32, L.F32)
layers.append(("fprop_enc2_s2_plain", lambda: K.conv_fprop(x1, wpe2, 64, 3, 2, 1, x2o, act=L.ACT_LRELU)))

for name, fn in layers:
    fn(); fn()
torch.cuda.synchronize()
assert K.device_error() == 0
ev = []
torch.cuda.profiler.start()
for name, fn in layers:
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record()
    ev.append((name, a, b))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
for name, a, b in ev:
    print(f"{name}: {a.elapsed_time(b):.3f} ms")
