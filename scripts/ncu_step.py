"""One eager training step (or one Generator forward) between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python scripts/ncu_step.py [batch] [train|inference]
(the launch list behind profiles/*_launches.md) and for `ncu --set full -k regex:...` captures of single kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import train_args
from oracle import uegan_oracle as O  # deterministic synthetic weights only

b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
mode = sys.argv[2] if len(sys.argv) > 2 else "train"
x = torch.rand(b, 3, 512, 512, device="cuda") * 2 - 1
if mode == "train":
    from uegan_b200.trainer import Trainer
    T = Trainer(None, train_args(b), vgg_state_dict=O.make_vgg_params())
    T.G.load_state_dict(O.make_generator_params(32, 0, "o1"))
    T.D.load_state_dict(O.make_discriminator_params(32, 1, "o1"))
    y = torch.rand(b, 3, 512, 512, device="cuda") * 2 - 1
    step = lambda: T.train_step(x, y, sync_scalars=False)
else:
    from uegan_b200.models import Generator
    G = Generator(32, "none", "LeakyReLU", False)
    G.load_state_dict(O.make_generator_params(32, 0, "o1"))
    G = G.cuda().eval()
    def step():
        with torch.no_grad():
            G(x)
for _ in range(3):
    step()
torch.cuda.synchronize()
from uegan_b200 import kernels as K
K._Counters.trace = []
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
import json
out = os.environ.get("UEGAN_TRACE_OUT")
if out:
    with open(out, "w") as f:
        json.dump(K._Counters.trace, f)
