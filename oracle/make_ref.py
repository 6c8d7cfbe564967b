#!/usr/bin/env python
"""Vendors the UNMODIFIED reference (eezkni/UEGAN, pure Python) into oracle/_ref/  --  TEST INFRASTRUCTURE.

/root/reference exists only in the build container; the GPU box gets `oracle/_ref/` with the repo snapshot
(git-ignored, not gpurun-ignored).  This recipe copies, byte for byte, the files the hot path and its callers need:

    *.py                      main, config, trainer, tester, models, losses, utils, data_loader
    metrics/                  CalcPSNR, CalcSSIM, NIMA/{CalcNIMA, mobile_net_v2}.py  (imported by trainer.py:12-14;
                              the 8.8 MB NIMA checkpoint is NOT copied: every run here passes --is_test_nima False)
    data/fivek/               the 3+3+3 bundled PNG pairs (train / val / test)

Nothing is edited: `oracle/_ref/MANIFEST.json` records the sha256 of every copied file next to the source's, and
tests/test_reference_vendor.py re-checks them when /root/reference is present.  Users: tests/ (drop-in boundary and
checkpoint tests), bench.py's `--impl reference` / cpu_baseline / gpu_library_baseline legs.  Never imported by uegan_b200/.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("UEGAN_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")

SKIP_DIRS = {"figures", ".git", "__pycache__"}
SKIP_FILES = {"pretrain-model.pth"}


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def wanted():
    out = []
    for dirpath, dirs, files in os.walk(SRC):
        dirs[:] = sorted(d for d in dirs if d not in SKIP_DIRS)
        rel = os.path.relpath(dirpath, SRC)
        for f in sorted(files):
            if f in SKIP_FILES or f.endswith((".pyc", ".md", ".m")):
                continue
            out.append(os.path.normpath(os.path.join(rel, f)))
    return out


def vendor(force=False):
    """Copies the reference into oracle/_ref (idempotent).  Returns the destination, or None when neither the source nor
    a previous copy exists."""
    manifest = os.path.join(DST, "MANIFEST.json")
    if not os.path.isdir(SRC):
        return DST if os.path.exists(manifest) else None
    files = wanted()
    if not force and os.path.exists(manifest):
        try:
            have = json.load(open(manifest))["files"]
            if sorted(have) == sorted(files) and all(os.path.exists(os.path.join(DST, f)) for f in files):
                return DST
        except Exception:
            pass
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    rec = {}
    for rel in files:
        s, d = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        rec[rel] = _sha(d)
        assert rec[rel] == _sha(s), rel
    with open(manifest, "w") as f:
        json.dump({"source": SRC, "files": rec}, f, indent=1, sort_keys=True)
    return DST


def ref_dir():
    """Directory holding the reference sources: the vendored copy when it exists, else /root/reference, else None."""
    if os.path.exists(os.path.join(DST, "MANIFEST.json")):
        return DST
    if os.path.isdir(SRC):
        return SRC
    return None


if __name__ == "__main__":
    d = vendor(force="--force" in sys.argv)
    print(d if d else "reference source not found and no vendored copy present")
    sys.exit(0 if d else 1)
