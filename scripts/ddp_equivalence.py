"""Data-parallel equivalence (SURVEY.md 8e / BASELINE.json configs[3] in miniature): N ranks x B images must take the
same training step as 1 rank x (N*B) images -- the relativistic means are global, gradients are summed over ranks.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/ddp_equivalence.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import train_args  # noqa: E402
from oracle import uegan_oracle as O  # noqa: E402
from uegan_b200.trainer import Trainer  # noqa: E402


def make_trainer(batch, group, peer=True):
    a = train_args(batch)
    a.peer_reduce = peer
    T = Trainer(None, a, process_group=group, vgg_state_dict=O.make_vgg_params())
    T.G.load_state_dict(O.make_generator_params(32, 0, "o1"))
    T.D.load_state_dict(O.make_discriminator_params(32, 1, "o1"))
    return T


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    b, res = 2, 128
    raw = O.make_images((world * b, 3, res, res), 40).cuda()
    exp = O.make_images((world * b, 3, res, res), 41).cuda()
    # sharded step: reductions over peer memory inside our kernels (default) or NCCL (UEGAN_DDP_NCCL=1)
    peer = os.environ.get("UEGAN_DDP_NCCL") != "1"
    T = make_trainer(b, dist.group.WORLD, peer)
    if rank == 0:
        print(f"world {world}: gradient / loss reductions via", "peer memory (uegan_adam_step_peers, uegan_peer_sum_f64)"
              if T.comm is not None else "NCCL all-reduce")
    v_dp = T.train_step(raw[rank * b:(rank + 1) * b].contiguous(), exp[rank * b:(rank + 1) * b].contiguous())
    # every rank must hold bit-identical weights after the step
    import hashlib
    flat = torch.cat([p.detach().flatten() for p in list(T.G.parameters()) + list(T.D.parameters())]).cpu().numpy()
    hashes = [None] * world
    dist.all_gather_object(hashes, hashlib.sha256(flat.tobytes()).hexdigest()[:16])
    # mean-type losses are per-rank means: average them over ranks for the comparison
    t = torch.tensor([v_dp["g_percep_loss"], v_dp["g_idt_loss"]], device="cuda", dtype=torch.float64)
    dist.all_reduce(t)
    t /= world
    ok = True
    if rank == 0:
        # the same step on the whole batch in one process (no group)
        T1 = make_trainer(world * b, None)
        v1 = T1.train_step(raw, exp)
        got = dict(d_loss=v_dp["d_loss"], g_adv_loss=v_dp["g_adv_loss"], g_percep_loss=float(t[0]), g_idt_loss=float(t[1]))
        for k, v in got.items():
            e = abs(v - v1[k]) / abs(v1[k])
            print(f"{k}: dp {v:.6f} single {v1[k]:.6f} rel {e:.2e}")
            ok &= e < 1e-4
        # post-step weights: gradients of both paths agree up to fp32 summation order / atomics
        for (n1, p1), (_, p2) in zip(T1.G.named_parameters(), T.G.named_parameters()):
            d = (p1 - p2).abs().max().item()
            if d > 2.5e-4:  # > lr*2: a real disagreement, not an Adam sign flip on a ~zero gradient
                print("G param mismatch", n1, d); ok = False
        dG = torch.cat([(p1 - p2).flatten() for p1, p2 in zip(T1.G.parameters(), T.G.parameters())]).abs()
        dD = torch.cat([(p1 - p2).flatten() for p1, p2 in zip(T1.D.parameters(), T.D.parameters())]).abs()
        print(f"post-step |dW|: G mean {dG.mean().item():.2e} (lr 1e-4)  D mean {dD.mean().item():.2e} (lr 4e-4)")
        ok &= dG.mean().item() < 2e-5 and dD.mean().item() < 4e-5
        # spectral-norm buffers: the same power iterations in both runs.  The two D forwards of the G step see the weights
        # AFTER D's Adam update, where a near-zero gradient may flip a weight's update sign (+-lr = 4e-4 on a few of d5's
        # 3.3 M weights): u = normalize(W v) then moves by ~1e-5 (measured 1.6e-5) -- a different iteration COUNT would show
        # as O(0.1)
        for k in range(1, 6):
            du = float((T1.D.state_dict()[f"d{k}.0.1.weight_u"] - T.D.state_dict()[f"d{k}.0.1.weight_u"]).abs().max())
            print(f"weight_u d{k}: max |du| {du:.2e}")
            ok &= du < 1e-4
        print("replica weight checksums:", hashes)
        ok &= len(set(hashes)) == 1
        print("DDP_EQUIVALENCE", "PASS" if ok else "FAIL")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
