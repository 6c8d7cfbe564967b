"""Debug: per-layer comparison of the native VGG backward chain with torch.autograd (fp32, GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import uegan_oracle as O
from uegan_b200.losses import PerceptualLoss, _VGG_LAYERS
from uegan_b200 import autograd as A

torch.backends.cudnn.allow_tf32 = False
vp = O.make_vgg_params()
P = PerceptualLoss(vgg_state_dict=vp).cuda()
a = ((O.make_images((2, 3, 128, 128), 40) + 1) / 2).cuda()
b = ((O.make_images((2, 3, 128, 128), 41) + 1) / 2).cuda()
# torch reference with retained grads
mean = torch.tensor(O.IMAGENET_MEAN, device="cuda").view(1, -1, 1, 1); std = torch.tensor(O.IMAGENET_STD, device="cuda").view(1, -1, 1, 1)
q16 = lambda t: t + (t.half().float() - t).detach()
def tower(x, keep):
    h = q16((x - mean) / std)
    acts = [h]
    for spec in _VGG_LAYERS:
        if spec == "M":
            h = F.max_pool2d(h, 2, 2)
        else:
            idx = spec[0]
            h = q16(F.relu(F.conv2d(h, vp[f"features.{idx}.weight"].cuda().half().float(), vp[f"features.{idx}.bias"].cuda(), padding=1)))
        if keep: h.retain_grad()
        acts.append(h)
    return acts
ar = a.clone().requires_grad_(True)
ax = tower(ar, True); ay = tower(b, False)
taps = [1, 4, 7, 12, 17]
w = [1/64, 1/64, 1/32, 1/32, 1.0]
loss = sum(wi * F.mse_loss(O.instance_norm(ax[t]), O.instance_norm(ay[t])) for wi, t in zip(w, taps))
loss.backward()
an = a.clone().requires_grad_(True)
ln = P(an, b)
ln.backward()
print("loss", float(loss), float(ln))
# find the plan
plan = [v for k, v in P.vgg._plans.items() if k[-1] == "x"][0]
G = plan["grads"]
t5 = plan["acts"][17]
scale = float(t5.n * t5.c * t5.h * t5.w) / 2.0
def rl2(x, y): return float((x.double() - y.double()).norm() / y.double().norm())
for j in range(len(G)):
    ref = ax[j].grad if j > 0 else None
    if ref is None: continue
    mine = G[j].interior_nchw()[:, :ref.shape[1]] / scale
    # G[j] of conv outputs is masked by relu' except taps (unmasked 'deep'); torch's .grad of a relu output is unmasked
    act = ax[j]
    masked_ref = ref * (act > 0)
    print(j, _VGG_LAYERS[j-1], "vs unmasked %.3e  vs masked %.3e" % (rl2(mine, ref), rl2(mine, masked_ref)), "|ref| %.3e" % float(ref.norm()))
print("dx rel-L2", rl2(an.grad, ar.grad))

for j in (1, 4, 7, 12, 17):
    print("fwd act", j, rl2(plan["acts"][j].interior_nchw(), ax[j].detach()), "exact-equal frac", float((plan["acts"][j].interior_nchw() == ax[j].detach()).float().mean()))
