"""Kernel-time breakdown of one eager training step with torch.profiler (CUPTI): cheap alternative to an ncu launch list.
    python scripts/profile_step.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from bench import train_args
from oracle import uegan_oracle as O
from uegan_b200.trainer import Trainer

b = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = Trainer(None, train_args(b), vgg_state_dict=O.make_vgg_params())
T.G.load_state_dict(O.make_generator_params(32, 0, "o1"))
T.D.load_state_dict(O.make_discriminator_params(32, 1, "o1"))
x = torch.rand(b, 3, 512, 512, device="cuda") * 2 - 1
y = torch.rand(b, 3, 512, 512, device="cuda") * 2 - 1
for _ in range(3):
    T.train_step(x, y, sync_scalars=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    T.train_step(x, y, sync_scalars=False)
    torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type is not None and "cuda" in str(e.device_type).lower():
        n = e.name.split("(")[0][:60]
        r = rows.setdefault(n, [0, 0.0]); r[0] += 1; r[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in rows.values())
print(f"total kernel time {tot/1e3:.2f} ms in {sum(v[0] for v in rows.values())} launches")
for k, v in sorted(rows.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k:62s} n={v[0]:4d} {v[1]/1e3:8.3f} ms {100*v[1]/tot:5.1f}%")
