"""The drop-in boundary, proven end to end: the UNMODIFIED reference `main.py` is run
through oracle/run_reference.py
  (a) as is, on the CPU                                   -- the parity anchor (fp32 torch-CPU, the reference's own code)
  (b) with `dropin/` in front of it on sys.path            -- models + losses + trainer bound to uegan_b200 (sm_100a)
  (c) with `dropin/kernels_only/` in front of it           -- the reference's OWN trainer.py / tester.py on the native kernels
for `--mode train` (2 iterations on the bundled data/fivek PNGs, seed 1990, torch default init so that the CPU and GPU
runs start from bit-identical weights) and `--mode test`, and the printed losses / written PNGs are compared.

Checkpoint compatibility rides on the same runs: the reference-written `.pth` (state_dict + Adam state +
spectral-norm u/v) is loaded by the native Tester path and by the native Trainer's resume, and a natively written `.pth`
is loaded by the reference's own tester.py on the CPU.
"""
import os
import re
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUN = os.path.join(ROOT, "oracle", "run_reference.py")
LOSS_RE = re.compile(r"D_loss:([\d.]+), G_loss:([\d.]+), G_percep_loss:([\d.]+), G_adv_loss:([\d.]+), G_idt_loss:([\d.]+)")

pytestmark = [pytest.mark.gpu]

TRAIN = ["--mode", "train", "--train_batch_size", "2", "--resize_size", "128", "--num_workers", "0", "--is_test_nima",
         "False", "--info_step", "1", "--init_type", "", "--pool_size", "0", "--version", "t", "--sample_step", "1000",
         "--is_print_network", "False"]
TEST = ["--mode", "test", "--test_img_size", "128", "--num_workers", "0", "--is_test_nima", "False", "--version", "t",
        "--init_type", "", "--is_print_network", "False"]


def run_main(tmp, name, opts, args, timeout=900):
    root = os.path.join(tmp, name)
    cmd = [sys.executable, RUN] + opts + ["--"] + args + ["--save_root_dir", root]
    env = dict(os.environ, TORCH_HOME=os.path.join(tmp, "torch_home"), PYTHONWARNINGS="ignore")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    out = res.stdout + res.stderr
    assert res.returncode == 0, f"{name}: main.py exited {res.returncode}\n{out[-4000:]}"
    return root, [tuple(float(v) for v in m) for m in LOSS_RE.findall(out)], out


def have_reference():
    sys.path.insert(0, ROOT)
    from oracle.make_ref import ref_dir
    return ref_dir() is not None


def read_pngs(folder):
    import cv2
    out = {}
    for f in sorted(os.listdir(folder)):
        out[f.split("_")[0]] = cv2.imread(os.path.join(folder, f)).astype(np.int32)
    return out


@pytest.fixture(scope="module")
def runs(tmp_path_factory):
    if not have_reference():
        pytest.skip("no reference sources (oracle/_ref is built by `python oracle/make_ref.py`)")
    tmp = str(tmp_path_factory.mktemp("dropin_main"))
    r = {}
    r["ref_dir"], r["ref"], _ = run_main(tmp, "ref", ["--cpu"], TRAIN + ["--total_epochs", "2"])
    r["full_dir"], r["full"], r["full_out"] = run_main(tmp, "full", ["--dropin", "dropin"], TRAIN + ["--total_epochs", "2"])
    r["ko_dir"], r["ko"], _ = run_main(tmp, "ko", ["--dropin", "dropin/kernels_only"], TRAIN + ["--total_epochs", "2"])
    r["tmp"] = tmp
    return r


def _close(a, b, rel):
    return abs(a - b) <= rel * max(abs(b), 1e-3) + 1.01e-4  # the reference prints four decimals


@pytest.mark.parametrize("which", ["full", "ko"])
def test_unmodified_main_trains_on_native_path(runs, which):
    ref, got = runs["ref"], runs[which]
    assert len(ref) == 2 and len(got) == 2, (ref, got)
    names = ("D_loss", "G_loss", "G_percep_loss", "G_adv_loss", "G_idt_loss")
    for n, a, b in zip(names, got[0], ref[0]):  # step 1: identical weights and data on both sides
        assert _close(a, b, 1e-3), f"step 1 {n}: native {a} vs reference-CPU {b}"
    for n, a, b in zip(names, got[1], ref[1]):  # step 2: after one Adam update of both networks (sign-sensitive)
        assert _close(a, b, 3e-2), f"step 2 {n}: native {a} vs reference-CPU {b}"


def test_checkpoints_are_written_under_the_reference_names(runs):
    for d in (runs["ref_dir"], runs["full_dir"], runs["ko_dir"]):
        names = sorted(os.listdir(os.path.join(d, "t", "models")))
        assert names == ["t_rahinge_1.0.pth", "t_rahinge_2.0.pth"], (d, names)


def _copy_ckpt(src_root, dst_root, epoch):
    os.makedirs(os.path.join(dst_root, "t", "models"), exist_ok=True)
    name = f"t_rahinge_{epoch}.pth"
    shutil.copyfile(os.path.join(src_root, "t", "models", name), os.path.join(dst_root, "t", "models", name))


def test_mode_test_reference_checkpoint_on_native_generator(runs):
    """tester.py:133-146 loads the reference-written checkpoint into the native Generator; PNGs (denorm + save_image,
    utils.py:128-130) must equal the reference-CPU ones up to one grey level."""
    tmp = runs["tmp"]
    ref_root, _, _ = run_main(tmp, "ref", ["--cpu"], TEST + ["--pretrained_model", "2.0"])
    _copy_ckpt(runs["ref_dir"], os.path.join(tmp, "test_native"), "2.0")
    nat_root, _, _ = run_main(tmp, "test_native", ["--dropin", "dropin"], TEST + ["--pretrained_model", "2.0"])
    a = read_pngs(os.path.join(ref_root, "t", "test", "test_results"))
    b = read_pngs(os.path.join(nat_root, "t", "test", "test_results"))
    assert sorted(a) == sorted(b) and len(a) == 3
    for k in a:
        d = np.abs(a[k] - b[k])
        assert d.max() <= 1 and d.mean() < 0.02, (k, int(d.max()), float(d.mean()))


def test_native_checkpoint_loads_in_reference_tester(runs):
    """A `.pth` written by uegan_b200.trainer.Trainer is read by the reference's own tester.py on the CPU."""
    tmp = runs["tmp"]
    _copy_ckpt(runs["full_dir"], os.path.join(tmp, "test_ref_on_native_ckpt"), "1.0")
    root, _, out = run_main(tmp, "test_ref_on_native_ckpt", ["--cpu"], TEST + ["--pretrained_model", "1.0"])
    assert "loaded trained models" in out
    assert len(os.listdir(os.path.join(root, "t", "test", "test_results"))) == 3


def test_native_trainer_resumes_from_reference_checkpoint(runs):
    """trainer.py:402-423: G, D (incl. weight_u/v), both Adam states and the schedulers come from the reference's
    epoch-1 checkpoint; the resumed native step 2 must reproduce the reference's step 2 -- same state, same batch order."""
    tmp = runs["tmp"]
    _copy_ckpt(runs["ref_dir"], os.path.join(tmp, "resume"), "1.0")
    _, got, out = run_main(tmp, "resume", ["--dropin", "dropin"], TRAIN + ["--total_epochs", "2", "--pretrained_model", "1.0"])
    assert len(got) == 1, out[-2000:]
    # the resumed run draws its first batch from a fresh loader, i.e. the reference's STEP-1 batch with the step-1 weights:
    # compare against a reference-CPU resume of the same checkpoint
    _copy_ckpt(runs["ref_dir"], os.path.join(tmp, "resume_ref"), "1.0")
    _, ref, _ = run_main(tmp, "resume_ref", ["--cpu"], TRAIN + ["--total_epochs", "2", "--pretrained_model", "1.0"])
    for n, a, b in zip(("D_loss", "G_loss", "G_percep_loss", "G_adv_loss", "G_idt_loss"), got[0], ref[0]):
        assert _close(a, b, 2e-3), f"resumed step {n}: native {a} vs reference-CPU {b}"
