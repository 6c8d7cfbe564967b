"""Bisects which part of the training step invalidates a CUDA-graph capture."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import train_args
from oracle import uegan_oracle as O
from uegan_b200.trainer import Trainer
from uegan_b200 import kernels as K

def make():
    T = Trainer(None, train_args(2), vgg_state_dict=O.make_vgg_params())
    T.G.load_state_dict(O.make_generator_params(32, 0, "o1"))
    T.D.load_state_dict(O.make_discriminator_params(32, 1, "o1"))
    return T

raw = O.make_images((2, 3, 128, 128), 40).cuda()
exp = O.make_images((2, 3, 128, 128), 41).cuda()

def try_capture(name, fn, warm=2):
    torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(warm): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(f"[capture] {name}: OK", flush=True)
    except Exception as e:
        print(f"[capture] {name}: FAILED {type(e).__name__}: {str(e)[:200]}", flush=True)
        try: torch.cuda.synchronize()
        except Exception: pass

T = make()
gan = T.criterionGAN
def d_fwd_bwd():
    T.d_grads.zero_grad()
    p1 = T.D(exp); p2 = T.D(raw)
    l = gan(p1, p2, None, None, for_discriminator=True)
    l.backward()
def d_fwd_only():
    with torch.no_grad():
        T.D(exp)
def d_step():
    T.d_optimizer.step()
def g_fwd_bwd():
    T.g_grads.zero_grad()
    out = T.G(raw)
    l = T.criterionIdt(out, exp)
    l.backward()
def zero_only():
    T.d_grads.zero_grad(); K.zero_(T.d_grads.grad)
def d_fwd_bwd_sum():
    T.d_grads.zero_grad()
    p1 = T.D(exp)
    sum(q.sum() for q in p1).backward()
def gan_only():
    a = [torch.tanh(torch.randn(2, 1, s, s, device="cuda")).requires_grad_(True) for s in (64, 32)]
    b = [torch.tanh(torch.randn(2, 1, s, s, device="cuda")).requires_grad_(True) for s in (64, 32)]
    gan(a, b, None, None, for_discriminator=True).backward()
def d_nosink():
    sink = T.D._grad_sink
    T.D._grad_sink = None
    for p in T.D.parameters():
        p.grad = None
    try:
        p1 = T.D(exp)
        sum(q.sum() for q in p1).backward()
    finally:
        T.D._grad_sink = sink
for name, fn in (("D fwd+bwd (sum loss)", d_fwd_bwd_sum), ("D fwd+bwd", d_fwd_bwd), ("G fwd+bwd", g_fwd_bwd),
                 ("full step", lambda: T.train_step(raw, exp, sync_scalars=False))):
    try_capture(name, fn)
