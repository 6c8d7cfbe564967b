"""N > 1 host logic on CPU with the gloo backend, world_size 2 (no GPU):
* the three-phase relativistic GAN loss protocol (local sums -> all-reduce -> terms -> all-reduce -> finalise) equals the
  loss on the gathered global batch (what the reference's nn.DataParallel computes), including gradients;
* the flat gradient bucket of uegan_b200.trainer all-reduces every parameter gradient in ONE collective."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import uegan_oracle as O
    torch.manual_seed(0)
    sizes = (16, 8, 4)
    real_all = [torch.tanh(torch.randn(world * 2, 1, s, s)) for s in sizes]
    fake_all = [torch.tanh(torch.randn(world * 2, 1, s, s)) for s in sizes]
    ok = True
    for mode in ("rahinge", "rals"):
        for for_d in (True, False):
            # ---- reference: global batch in one process
            rg = [t.clone().requires_grad_(True) for t in real_all]
            fg = [t.clone().requires_grad_(True) for t in fake_all]
            ref = O.gan_loss(mode, rg, fg, for_d)
            ref.backward()
            # ---- protocol on this rank's shard (same arithmetic as csrc/losses.cu gan_*_kernel)
            r = [t[rank * 2:(rank + 1) * 2].clone() for t in real_all]
            f = [t[rank * 2:(rank + 1) * 2].clone() for t in fake_all]
            sums = torch.tensor([[t.sum() for t in r], [t.sum() for t in f]], dtype=torch.float64)
            dist.all_reduce(sums)                                           # all-reduce #1: 2*nscales doubles
            sr, sf = (-1.0, 1.0) if for_d else (1.0, -1.0)
            terms = torch.zeros(4, len(sizes), dtype=torch.float64)
            loc = []
            for i in range(len(sizes)):
                n = r[i].numel() * world
                mr, mf = sums[0, i] / n, sums[1, i] / n
                zr, zf = 1 + sr * (r[i] - mf), 1 + sf * (f[i] - mr)
                if mode == "rahinge":
                    tr, tf_, gr, gf = zr.clamp_min(0), zf.clamp_min(0), (zr > 0) * sr, (zf > 0) * sf
                else:
                    tr, tf_, gr, gf = zr * zr, zf * zf, 2 * zr * sr, 2 * zf * sf
                terms[:, i] = torch.tensor([tr.sum(), tf_.sum(), gr.sum(), gf.sum()], dtype=torch.float64)
                loc.append((n, gr, gf))
            dist.all_reduce(terms)                                          # all-reduce #2: 4*nscales doubles
            loss = sum(0.5 * (terms[0, i] + terms[1, i]) / loc[i][0] for i in range(len(sizes)))
            ok &= abs(float(loss) - float(ref)) < 1e-6
            for i, (n, gr, gf) in enumerate(loc):
                d_r = 0.5 / n * (gr - terms[3, i] / n)
                d_f = 0.5 / n * (gf - terms[2, i] / n)
                ok &= float((d_r - rg[i].grad[rank * 2:(rank + 1) * 2]).abs().max()) < 1e-7
                ok &= float((d_f - fg[i].grad[rank * 2:(rank + 1) * 2]).abs().max()) < 1e-7
    # ---- flat gradient bucket
    from uegan_b200.trainer import _FlatGrads
    from uegan_b200.models import Discriminator
    torch.manual_seed(1)
    D = Discriminator(32, "none", "LeakyReLU", True, "rahinge")
    fg_ = _FlatGrads(D, dist.group.WORLD)
    assert fg_.flat.numel() == 4633632  # every D parameter lives in the one bucket
    for i, p in enumerate(fg_.params):
        p.grad.fill_(float(rank + 1) * (i + 1))
    fg_.all_reduce()
    for i, p in enumerate(fg_.params):
        ok &= bool(torch.all(p.grad == float(sum(range(1, world + 1))) * (i + 1)))
        ok &= p.grad.data_ptr() >= fg_.flat.data_ptr()  # still a view of the bucket
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


@pytest.mark.gpu
def test_ddp_equivalence_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import subprocess
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "scripts", "ddp_equivalence.py")],
                         capture_output=True, text=True, timeout=900)
    print(out.stdout[-3000:], out.stderr[-2000:])
    assert "DDP_EQUIVALENCE PASS" in out.stdout
