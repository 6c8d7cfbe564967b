"""Host-side logic of the flat buckets / fused Adam (uegan_b200.optim) on CPU tensors: no kernel is launched here -- the
update itself is checked on the GPU (tests/test_gpu_optim.py) -- but the layout, the views and the checkpoint format are
plain torch bookkeeping and must hold everywhere: parameters stay views of ONE buffer through state_dict round trips, and
a `torch.optim.Adam` state_dict (what the reference writes into its .pth, trainer.py:186-210) loads into FlatAdam and
comes back out in the same layout."""
import torch


def _nets():
    from uegan_b200.models import Discriminator
    torch.manual_seed(3)
    return Discriminator(8, "none", "LeakyReLU", True, "rahinge")


def test_flat_bucket_views_and_alignment():
    from uegan_b200.optim import FlatBucket
    D = _nets()
    before = {k: v.clone() for k, v in D.state_dict().items()}
    b = FlatBucket(D)
    assert b.numel % 4 == 0 and all(o % 4 == 0 for o in b.offsets)
    for (name, p), off in zip(D.named_parameters(), b.offsets):
        assert p.data_ptr() == b.flat.data_ptr() + 4 * off and p.grad.data_ptr() == b.grad.data_ptr() + 4 * off
        assert D._grad_sink[name] is p.grad
    assert all(torch.equal(before[k], v) for k, v in D.state_dict().items())
    # load_state_dict writes THROUGH the views
    sd = {k: torch.full_like(v, 0.25) if v.dtype.is_floating_point else v for k, v in before.items()}
    D.load_state_dict(sd)
    used = sum(p.numel() for p in D.parameters())
    assert float(b.flat.sum()) == 0.25 * used
    b.grad.fill_(1.0)
    b.zero_grad()
    assert all(float(p.grad.abs().sum()) == 0.0 for p in D.parameters())


def test_flat_adam_speaks_torch_adam_checkpoints():
    from uegan_b200.optim import FlatAdam, FlatBucket
    D, D2 = _nets(), _nets()
    ref = torch.optim.Adam(D2.parameters(), lr=4e-4, betas=[0.5, 0.999], weight_decay=1e-4)
    for p in D2.parameters():
        p.grad = torch.randn_like(p)
    ref.step(); ref.step()
    ck = ref.state_dict()
    opt = FlatAdam(FlatBucket(D), lr=4e-4, betas=[0.5, 0.999], weight_decay=1e-4)
    opt.load_state_dict(ck)
    assert float(opt.dev_state[0]) == 2.0
    for p, q in zip(D.parameters(), D2.parameters()):
        st = opt.state[p]
        assert torch.equal(st["exp_avg"], ref.state[q]["exp_avg"]) and torch.equal(st["exp_avg_sq"], ref.state[q]["exp_avg_sq"])
        assert st["exp_avg"].data_ptr() >= opt.exp_avg.data_ptr()  # a view of the flat moment buffer again
    out = opt.state_dict()
    assert out["param_groups"][0]["betas"] == (0.5, 0.999) and out["param_groups"][0]["weight_decay"] == 1e-4
    assert sorted(out["state"]) == sorted(ck["state"])
    for i in ck["state"]:
        assert float(out["state"][i]["step"]) == 2.0
        assert torch.equal(out["state"][i]["exp_avg"], ck["state"][i]["exp_avg"])
    # and torch's own Adam accepts what FlatAdam wrote
    torch.optim.Adam(_nets().parameters(), lr=4e-4, betas=[0.5, 0.999], weight_decay=1e-4).load_state_dict(out)
    # a scheduler drives the device-side learning rate through param_groups
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda e: 0.5)
    opt.sync_lr()
    assert abs(float(opt.lr_dev) - 2e-4) < 1e-10  # a float32 device scalar
    del sched
