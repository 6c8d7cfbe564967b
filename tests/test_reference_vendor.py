"""oracle/_ref is an UNMODIFIED copy of the reference (oracle/make_ref.py), and driving its modules through the loop
body of trainer.py:75-119 (oracle/ref_step.py, used by bench.py's reference / library-baseline legs) gives the same
five losses as the oracle restatement that the golden fixtures pin."""
import hashlib
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sha(p):
    return hashlib.sha256(open(p, "rb").read()).hexdigest()


def _ref():
    from oracle.make_ref import ref_dir
    d = ref_dir()
    if d is None:
        pytest.skip("no reference sources")
    return d


@pytest.mark.reference
def test_vendored_copy_is_byte_identical_to_the_reference():
    from oracle import make_ref
    dst = make_ref.vendor()
    man = json.load(open(os.path.join(dst, "MANIFEST.json")))["files"]
    assert {"main.py", "trainer.py", "tester.py", "models.py", "losses.py", "utils.py", "config.py",
            "data_loader.py"} <= set(man)
    for rel, sha in man.items():
        assert _sha(os.path.join(dst, rel)) == sha == _sha(os.path.join("/root/reference", rel)), rel


def test_manifest_matches_vendored_files():
    d = _ref()
    if not os.path.exists(os.path.join(d, "MANIFEST.json")):
        pytest.skip("running against /root/reference directly")
    for rel, sha in json.load(open(os.path.join(d, "MANIFEST.json")))["files"].items():
        assert _sha(os.path.join(d, rel)) == sha, rel


def test_reference_step_equals_oracle_step():
    _ref()
    from oracle import ref_step
    from oracle import uegan_oracle as O
    got = ref_step.make_step("train", 1, "cpu", res=96)()
    gp, dp, vp = O.make_generator_params(32, 0, "o1"), O.make_discriminator_params(32, 1, "o1"), O.make_vgg_params()
    x, y = O.make_images((1, 3, 96, 96), 0), O.make_images((1, 3, 96, 96), 1)
    want = O.train_step(gp, dp, vp, O.AdamState(O._trainable(gp)), O.AdamState(O._trainable(dp)), x, y)
    for k in ("d_loss", "g_adv_loss", "g_percep_loss", "g_idt_loss", "g_loss"):
        assert abs(got[k] - float(want[k])) <= 1e-5 * max(1.0, abs(float(want[k]))), (k, got[k], want[k])
