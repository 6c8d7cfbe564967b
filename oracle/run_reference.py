#!/usr/bin/env python
"""Runs the UNMODIFIED reference `main.py` (oracle/_ref or /root/reference)  --  TEST INFRASTRUCTURE.

    python oracle/run_reference.py [--dropin DIR[:DIR...]] [--cpu] -- <main.py arguments>

`--dropin dropin` puts the repo's `dropin/` directory in front of the reference on sys.path, so that the reference's
`from models import Generator, Discriminator` / `from losses import ...` / `from trainer import Trainer`
(main.py:5, trainer.py:9-11, tester.py:9-11) bind to uegan_b200 instead -- the drop-in boundary of SURVEY.md 8(b).
`--dropin dropin/kernels_only` replaces only `models` and `losses`: the reference's own trainer.py / tester.py then
drive the native kernels.

The reference imports four packages this image does not have; they are stubbed at import time exactly as
SURVEY.md 8(c) describes (none is touched on the code path that runs): `munch.Munch` (attribute dict),
`tensorflow` (only `utils.Logger`, --use_tensorboard False), `scipy.misc` (unused import), `skimage.metrics`
(only metrics/CalcSSIM.py, --is_test_psnr_ssim False).  torchvision's `vgg19(pretrained=True)` (losses.py:43) finds a
deterministic synthetic checkpoint (oracle.make_vgg_params) seeded at $TORCH_HOME/hub/checkpoints: there is no
network, and both sides of every comparison load that same file.
"""
import importlib.machinery
import os
import runpy
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def install_shims():
    import torch.utils.tensorboard  # noqa: F401  (must be imported before the fake tensorflow is visible)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class Munch(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

    if "munch" not in sys.modules:
        try:
            import munch  # noqa: F401
        except ImportError:
            mod("munch", Munch=Munch)
    try:
        import tensorflow  # noqa: F401
    except ImportError:
        mod("tensorflow")
    import scipy
    try:
        import scipy.misc  # noqa: F401
    except ImportError:
        scipy.misc = mod("scipy.misc")
    try:
        import skimage.metrics  # noqa: F401
    except ImportError:
        def structural_similarity(*a, **k):
            raise RuntimeError("skimage is not installed (stub)")
        sk = mod("skimage")
        sk.metrics = mod("skimage.metrics", structural_similarity=structural_similarity)


def seed_vgg_checkpoint(torch_home):
    """torchvision's hub cache entry for vgg19 (the hash in the file name is only verified on download)."""
    import torch
    path = os.path.join(torch_home, "hub", "checkpoints", "vgg19-dcbb9e9d.pth")
    if not os.path.exists(path):
        sys.path.insert(0, ROOT)
        from oracle import uegan_oracle as O
        import torchvision
        os.makedirs(os.path.dirname(path), exist_ok=True)
        net = torchvision.models.vgg19(weights=None)
        net.load_state_dict(O.make_vgg_params(), strict=False)  # features.*; the classifier keeps its seeded init
        torch.save(net.state_dict(), path + ".tmp")
        os.replace(path + ".tmp", path)
    return path


def reference_dir():
    sys.path.insert(0, ROOT)
    from oracle.make_ref import ref_dir
    d = ref_dir()
    if d is None:
        raise SystemExit("reference sources not found: run `python oracle/make_ref.py` where /root/reference exists")
    return d


def main():
    argv = sys.argv[1:]
    dropins, cpu = [], False
    while argv and argv[0] != "--":
        a = argv.pop(0)
        if a == "--dropin":
            dropins = [os.path.join(ROOT, d) if not os.path.isabs(d) else d for d in argv.pop(0).split(":") if d]
        elif a == "--cpu":
            cpu = True
        else:
            raise SystemExit(f"unknown option {a}")
    if argv and argv[0] == "--":
        argv.pop(0)
    if cpu:
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
    os.environ.setdefault("TORCH_HOME", os.path.join(os.environ.get("TMPDIR", "/tmp"), "uegan_torch_home"))
    ref = reference_dir()
    import torch
    torch.manual_seed(0)
    seed_vgg_checkpoint(os.environ["TORCH_HOME"])
    install_shims()
    import warnings
    warnings.filterwarnings("ignore")
    sys.path[:0] = dropins + [ref]
    if ROOT not in sys.path:
        sys.path.append(ROOT)  # `import uegan_b200` from dropin/*
    os.chdir(ref)  # config.py's default directories are relative (./data/fivek/...)
    sys.argv = [os.path.join(ref, "main.py")] + argv
    runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
