"""CPU-side checks of the boundary: the shared library loads, exports every symbol include/uegan_sm100.h declares,
and the native modules keep the reference's parameter names / counts.  No compute calls."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "uegan_sm100.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(uegan_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from uegan_b200 import _lib
    lib = _lib.load()
    declared = header_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes prototype in uegan_b200/_lib.py"
    assert set(_lib.SYMBOLS) == set(declared)
    assert lib.uegan_abi_version() == _lib.ABI_VERSION


def test_sass_is_blackwell_native():
    """tcgen05.mma -> UTC*MMA, TMA -> UTMALDG, tcgen05.ld -> LDTM must be present in the shipped SASS."""
    import shutil
    import subprocess
    from uegan_b200 import _lib
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic


def test_generator_state_dict_matches_reference_keys():
    from oracle import uegan_oracle as O
    from uegan_b200.models import Generator
    G = Generator(32, "none", "LeakyReLU", False)
    ref = O.make_generator_params(32, 0)
    sd = G.state_dict()
    assert set(sd) == set(ref) and len(sd) == 50
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    assert sum(p.numel() for p in G.parameters()) == 4158435  # trainer.py:393-399 printout of the reference
    # init_weights (trainer.py:357-390) keys on class names containing 'Conv'
    convs = [m for m in G.modules() if m.__class__.__name__.find("Conv") != -1 and hasattr(m, "weight")]
    assert len(convs) == 30


def test_product_path_has_no_cpu_fallback():
    from uegan_b200 import _lib
    from uegan_b200.models import Generator
    G = Generator(8, "none", "LeakyReLU", False)
    with pytest.raises(_lib.UeganError):
        G(torch.zeros(1, 3, 32, 32))
    for bad in (dict(norm_fun="BatchNorm"), dict(act_fun="Swish"), dict(use_sn=True), dict(norm_fun="nope")):
        kw = dict(conv_dim=8, norm_fun="none", act_fun="LeakyReLU", use_sn=False)
        kw.update(bad)
        with pytest.raises(NotImplementedError):
            Generator(**kw)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "uegan_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("no oracle", ""), f


def test_shape_policy_entry_points_without_gpu(monkeypatch):
    """Host-only policy functions of the C ABI (no CUDA calls): which convolutions take the row-sum kernels.  The
    deepest head fits in fp16 only."""
    import types
    from uegan_b200 import _lib as L
    from uegan_b200 import kernels as K
    lib = L.load()
    monkeypatch.delenv("UEGAN_NO_ROWSUM", raising=False)
    # tiny-Cout row-sum (validated on B200): G's last conv and D's heads qualify, the deepest head's weights do not fit
    assert lib.uegan_conv2d_rowsum_supported(3, 32, 7, L.F32) == 1
    assert lib.uegan_conv2d_rowsum_supported(1, 256, 5, L.F32) == 1
    assert lib.uegan_conv2d_rowsum_supported(1, 512, 5, L.F32) == 0
    assert lib.uegan_conv2d_rowsum_supported(3, 32, 7, L.F16) == 1  # fp16 operands: 64-channel chunks
    assert lib.uegan_conv2d_rowsum_supported(1, 512, 5, L.F16) == 1  # half the bytes: the deepest head fits too
    assert lib.uegan_conv2d_rowsum_supported(32, 32, 3, L.F32) == 0
    monkeypatch.setenv("UEGAN_NO_ROWSUM", "1")
    assert lib.uegan_conv2d_rowsum_supported(3, 32, 7, L.F32) == 0
    assert lib.uegan_packed_weight_rowsum_bytes(3, 32, 7) == 32 * 7 * 64 * 4  # sized for fp32 rows or fp16 rows padded to 64 channels
    # statistics policy (kernels.fused_stats_ok): separate pass by default
    monkeypatch.delenv("UEGAN_FUSED_STATS", raising=False)
    assert K.fused_stats_ok(512, 512, 64) is False
    monkeypatch.setenv("UEGAN_FUSED_STATS", "all")
    assert K.fused_stats_ok(512, 512, 64) is True and K.fused_stats_ok(4, 4, 64) is False
    monkeypatch.setenv("UEGAN_FUSED_STATS", "32")
    assert K.fused_stats_ok(512, 512, 64) is False and K.fused_stats_ok(512, 512, 32) is True
    # sliding-window weight gradient (uegan_conv2d_wgrad_zwin): fp16, stride 1, odd k, a zero halo >= k - 1 on dz, stacked
    # columns k * dz_c <= 256 -- G's dec5.1 (3 of 8 stored channels, k 7), dec5.0 / dec4 (32 ch, k 3), dec3 (64 ch, k 3),
    # the heads (1 of 8, k 3 / 5 / 7); not the 128-channel layers, not stride 2, not tf32
    z = lib.uegan_conv2d_wgrad_zwin_supported
    assert z(3, 8, 6, 32, 7, 1, L.F16) == 1 and z(32, 32, 2, 64, 3, 1, L.F16) == 1 and z(64, 64, 2, 128, 3, 1, L.F16) == 1
    assert z(1, 8, 4, 256, 5, 1, L.F16) == 1
    assert z(128, 128, 2, 256, 3, 1, L.F16) == 0  # 3 * 128 stacked columns > 256
    assert z(32, 32, 1, 64, 3, 1, L.F16) == 0     # halo too small for the horizontal window
    assert z(32, 32, 2, 64, 3, 2, L.F16) == 0 and z(32, 32, 2, 64, 3, 1, L.F32) == 0 and z(32, 32, 3, 64, 4, 1, L.F16) == 0
    assert z(32, 32, 2, 8, 3, 1, L.F16) == 0      # x rows must be whole 64-byte (32-channel) units


def test_struct_layouts_match_the_header(tmp_path):
    """include/uegan_sm100.h compiled as plain C (gcc, no CUDA headers): sizeof / offsetof of every struct that crosses the
    C ABI must equal the ctypes mirror in uegan_b200/_lib.py field by field -- a field added on one side only would shift
    every later field silently."""
    import shutil
    import subprocess
    import ctypes as C
    from uegan_b200 import _lib as L
    if shutil.which("gcc") is None:
        pytest.skip("gcc not on PATH")
    structs = {"uegan_tensor": L.Tensor, "uegan_conv_desc": L.ConvDesc, "uegan_scale_entry": L.ScaleEntry}
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "uegan_sm100.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = {}
    for ln in out.splitlines():
        s, f, v = ln.split()
        got[(s, f)] = int(v)
    for cname, cls in structs.items():
        assert got[(cname, "size")] == C.sizeof(cls), (cname, got[(cname, "size")], C.sizeof(cls))
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
    # and the header declares no field the mirror lacks (sizes equal + last field's end == size up to padding)
    hdr = open(os.path.join(ROOT, "include", "uegan_sm100.h")).read()
    body = hdr[hdr.index("typedef struct uegan_conv_desc {"):hdr.index("} uegan_conv_desc;")]
    import re
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    declared = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        stmt = stmt.split("{")[-1].strip()
        if not stmt:
            continue
        names = stmt.replace("*", " ").split(",")
        declared.append(names[0].split()[-1])
        declared += [n.strip() for n in names[1:]]
    assert declared == [f for f, _ in L.ConvDesc._fields_], declared
