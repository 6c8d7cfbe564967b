"""CPU study for DESIGN.md 7a item 1: which operand formats keep the Generator within 1e-3 of the fp32 reference?
Every convolution of the oracle's simplified Generator forward gets its INPUT and WEIGHT rounded to the candidate format
(fp32 accumulation, like the tensor cores), in both weight regimes of the golden fixtures:
    tf32            10 explicit mantissa bits, fp32 range          (what the sm_100a path does today)
    bf16             7 bits, fp32 range
    fp16            10 bits, 5-bit exponent: under/overflows in the orthogonal(0.02) regime
    fp16 + scale    per-tensor power-of-two scale so that amax -> 2^14 before rounding (activations AND weights)
Run: python scripts/precision_study.py   (about a minute on 8 cores; no GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import uegan_oracle as O


def q_tf32(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def q_bf16(t):
    return t.bfloat16().float()


def q_fp16(t):
    return t.half().float()


def q_fp16_scaled(t):
    amax = float(t.abs().max())
    if amax == 0.0:
        return t
    s = 2.0 ** (14 - torch.ceil(torch.log2(torch.tensor(amax))).item())
    return (t * s).half().float() / s


def run(q, regime, shape, seed):
    gp = O.make_generator_params(32, 0, regime)
    x = O.make_images(shape, seed)
    conv2d, rconv = F.conv2d, O._rconv

    def conv_q(inp, w, b=None, stride=1, *a, **k):
        return conv2d(q(inp), q(w), b, stride, *a, **k)
    with torch.no_grad():
        ref, ri = O.generator_forward(gp, x, simplified=True, return_all=True)
        F.conv2d = conv_q
        try:
            out, oi = O.generator_forward(gp, x, simplified=True, return_all=True)
        finally:
            F.conv2d = conv2d
    l2 = float((out - ref).norm() / ref.norm())
    res = float((oi["res"] - ri["res"]).norm() / ri["res"].norm().clamp_min(1e-30))
    return l2, float((out - ref).abs().max()), res


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    print(f"{'format':14s} {'regime':6s} {'pixels rel-L2':>14s} {'max abs':>10s} {'residual rel-L2':>16s}")
    for regime in ("o1", "tiny"):
        for name, q in (("tf32", q_tf32), ("bf16", q_bf16), ("fp16", q_fp16), ("fp16+scale", q_fp16_scaled)):
            l2, mx, res = run(q, regime, (2, 3, 128, 128), 10)
            print(f"{name:14s} {regime:6s} {l2:14.3e} {mx:10.3e} {res:16.3e}")


# ------------------------------------------------------------------------------------------------------------------
# Training side: the same formats on the BACKWARD GEMMs as well (dgrad: q(dz) * q(w); wgrad: q(x) * q(dz)), every tensor
# with its own power-of-two scale in the fp16+scale case.  Loss = sum(G(x) * r) / mean over the five D maps.
# ------------------------------------------------------------------------------------------------------------------
class QConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, q):
        xq, wq = q(x), q(w)
        ctx.save_for_backward(xq, wq)
        ctx.stride, ctx.q, ctx.has_b = stride, q, b is not None
        return torch.nn.functional._orig_conv2d(xq, wq, b, stride)

    @staticmethod
    def backward(ctx, dz):
        xq, wq = ctx.saved_tensors
        dq = ctx.q(dz)
        dx = torch.nn.grad.conv2d_input(xq.shape, wq, dq, stride=ctx.stride)
        dw = torch.nn.grad.conv2d_weight(xq, wq.shape, dq, stride=ctx.stride)
        db = dz.sum(dim=(0, 2, 3)) if ctx.has_b else None
        return dx, dw, db, None, None


def grads(q, regime, which):
    torch.manual_seed(0)
    x = O.make_images((2, 3, 128, 128), 40)
    if which == "G":
        p = {k: v.clone().requires_grad_(True) for k, v in O.make_generator_params(32, 0, regime).items()}
        r = O.make_images((2, 3, 128, 128), 77)
        loss_fn = lambda: (O.generator_forward(p, x, simplified=True) * r).sum()
    else:
        p = {k: v.clone() for k, v in O.make_discriminator_params(32, 1, regime).items()}
        for k in O._trainable(p):
            p[k].requires_grad_(True)
        loss_fn = lambda: sum(m.mean() for m in O.discriminator_forward(p, x, training=True, update_buffers=False))
    F._orig_conv2d = F.conv2d
    if q is not None:
        F.conv2d = lambda inp, w, b=None, stride=1, *a, **k: QConv.apply(inp, w, b, stride if isinstance(stride, int) else stride[0], q)
    try:
        loss = loss_fn()
        names = [k for k in p if p[k].requires_grad]
        g = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    finally:
        F.conv2d = F._orig_conv2d
    return {k: v for k, v in zip(names, g) if v is not None}


def grad_study():
    print(f"\n{'format':14s} {'net':3s} {'regime':6s} {'worst rel-L2 of a weight gradient':>34s} {'median':>10s}")
    for which in ("G", "D"):
        for regime in ("o1", "tiny"):
            ref = grads(None, regime, which)
            for name, q in (("tf32", q_tf32), ("bf16", q_bf16), ("fp16", q_fp16), ("fp16+scale", q_fp16_scaled)):
                g = grads(q, regime, which)
                errs = sorted(float((g[k] - ref[k]).norm() / ref[k].norm().clamp_min(1e-30)) for k in ref
                              if k.endswith("weight") or k.endswith("weight_orig"))
                errs = [e for e in errs if e == e]
                print(f"{name:14s} {which:3s} {regime:6s} {errs[-1]:34.3e} {errs[len(errs)//2]:10.3e}")


if __name__ == "__main__":
    grad_study()
