"""The reference's own modules driven through the loop body of trainer.py:75-119  --  TEST / BASELINE INFRASTRUCTURE.

`make_step(workload, batch, device)` imports `models.py` and `losses.py` of the vendored, unmodified reference
(oracle/_ref, built by oracle/make_ref.py; /root/reference in the build container), builds Generator / Discriminator /
PerceptualLoss / GANLoss / MultiscaleRecLoss / torch.optim.Adam exactly as trainer.py:313-354 does, loads the same
deterministic synthetic weights the native arm uses, and returns a closure that runs ONE iteration.  Every tensor
operation inside that closure is the reference's code; only the ~20 lines of sequencing are restated here
(cited line by line) because the reference has no step() method.

device "cpu"  : the CPU baseline (`bench.py --impl reference`, `cpu_baseline.kind == "reference"`).
device "cuda" : the same code on the GPU -- eager PyTorch + cuDNN with `cudnn.benchmark = True` (main.py:16) and torch's
                default TF32 convolutions: the library baseline "to beat on the same box" (last bullet).
Used by bench.py and tests/ only; uegan_b200/ never imports it.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_mods = {}


def reference_modules():
    """(models, losses) of the unmodified reference, imported under private names (no sys.path games)."""
    if _mods:
        return _mods["models"], _mods["losses"]
    from oracle.make_ref import ref_dir
    from oracle.run_reference import seed_vgg_checkpoint
    d = ref_dir()
    if d is None:
        raise RuntimeError("reference sources not found (oracle/_ref is built by `python oracle/make_ref.py`)")
    os.environ.setdefault("TORCH_HOME", os.path.join(os.environ.get("TMPDIR", "/tmp"), "uegan_torch_home"))
    seed_vgg_checkpoint(os.environ["TORCH_HOME"])
    for name in ("models", "losses"):
        spec = importlib.util.spec_from_file_location("uegan_reference_" + name, os.path.join(d, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _mods[name] = m
    return _mods["models"], _mods["losses"]


def make_step(workload, batch, device="cpu", res=512, seed=0):
    import warnings
    import torch
    from oracle import uegan_oracle as O
    warnings.filterwarnings("ignore")
    M, Ls = reference_modules()
    dev = torch.device(device)
    if dev.type == "cuda":
        torch.backends.cudnn.benchmark = True  # main.py:16
    G = M.Generator(32, "none", "LeakyReLU", False)
    G.load_state_dict(O.make_generator_params(32, 0, "o1"))
    G = G.to(dev)
    x = O.make_images((batch, 3, res, res), seed).to(dev)
    if workload == "inference":
        G.eval()

        def step():
            with torch.no_grad():
                return G(x)
        return step
    D = M.Discriminator(32, "none", "LeakyReLU", True, "rahinge")
    D.load_state_dict(O.make_discriminator_params(32, 1, "o1"))
    D = D.to(dev)
    g_opt = torch.optim.Adam(params=G.parameters(), lr=1e-4, betas=[0.5, 0.999], weight_decay=0.0001)  # trainer.py:337
    d_opt = torch.optim.Adam(params=D.parameters(), lr=4e-4, betas=[0.5, 0.999], weight_decay=0.0001)  # trainer.py:338
    percep = Ls.PerceptualLoss().to(dev)                              # trainer.py:54
    idt = Ls.MultiscaleRecLoss(scale=3, rec_loss_type="l1", multiscale=True)   # trainer.py:55
    gan = Ls.GANLoss("rahinge", tensor=torch.FloatTensor)            # trainer.py:56
    y = O.make_images((batch, 3, res, res), seed + 1).to(dev)
    lam_adv, lam_percep, lam_idt = 0.10, 1.0, 0.10                   # config.py:46-48

    def step():
        G.train(); D.train()                                          # trainer.py:77-78
        real_raw, real_exp = x, y
        fake_exp = G(real_raw)                                        # :85
        fake_store = fake_exp                                         # :86, pool_size = 0
        d_opt.zero_grad()                                             # :89
        real_preds = D(real_exp)                                      # :90
        fake_preds = D(fake_store.detach())                           # :91
        d_loss = gan(real_preds, fake_preds, None, None, for_discriminator=True)           # :92
        d_loss = d_loss + gan(real_preds, D(real_raw), None, None, for_discriminator=True)  # :93-95 (adv_input)
        d_loss.backward()                                             # :96
        d_opt.step()                                                  # :97
        g_opt.zero_grad()                                             # :101
        real_preds = D(real_exp)                                      # :102
        fake_preds = D(fake_exp)                                      # :103
        g_adv = lam_adv * gan(real_preds, fake_preds, None, None, for_discriminator=False)  # :104
        g_percep = lam_percep * percep((fake_exp + 1.) / 2., (real_raw + 1.) / 2.)          # :108
        g_idt = lam_idt * idt(G(real_exp), real_exp)                  # :112-113
        g_loss = g_adv + g_percep + g_idt
        g_loss.backward()                                             # :117
        g_opt.step()                                                  # :118
        return dict(d_loss=d_loss.item(), g_adv_loss=g_adv.item(), g_percep_loss=g_percep.item(),
                    g_idt_loss=g_idt.item(), g_loss=g_loss.item())    # :98-119 (.item() syncs as in the reference)
    return step
