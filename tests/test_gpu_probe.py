"""Hardware probe: tcgen05.mma reading a K-major SWIZZLE_128B operand through shifted descriptor windows.
The aligned case (row_shift multiple of 8, SBO 1024) must be exact; the other variants are recorded in
gpurun_out/probe_umma.json for DESIGN.md (they decide the shared-memory window-reuse design)."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def load_probe_lib():
    """tests/native/libuegan_probe.so: test-only hardware probes (built by __graft_entry__.build())."""
    import ctypes as C
    from tests.native.build_probe import build
    lib = C.CDLL(build())
    lib.uegan_probe_umma_window.restype = C.c_int
    lib.uegan_probe_umma_window.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p]
    lib.uegan_probe_device_error.restype = C.c_int
    return lib


def run_probe(lib, a, b, n, row_shift, base_offset, sbo):
    out = torch.zeros(128, n, device="cuda")
    rc = lib.uegan_probe_umma_window(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.shape[0], n, 1, row_shift,
                                     base_offset, sbo, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    assert lib.uegan_probe_device_error() == 0
    return out


def test_probe_umma_window():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    lib = load_probe_lib()
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(5)
    n = 32
    a_rows = 400
    # values exactly representable in tf32 so that any mismatch is an addressing effect, not rounding
    a = torch.randint(-8, 9, (a_rows, 32), device="cuda", generator=g).float()
    b = torch.randint(-8, 9, (n, 32), device="cuda", generator=g).float()
    results = {}

    def expected(row_shift, sbo):
        rows = torch.tensor([row_shift + (m // 8) * (sbo // 128) + m % 8 for m in range(128)], device="cuda")
        return a[rows] @ b.t()

    for row_shift, base_offset, sbo in [(0, 0, 1024), (8, 0, 1024), (16, 0, 1024),
                                        (1, 0, 1024), (1, 1, 1024), (3, 0, 1024), (3, 3, 1024), (5, 5, 1024),
                                        (0, 0, 1280), (0, 0, 2048), (0, 0, 2304), (2, 2, 2304), (2, 0, 2304)]:
        out = run_probe(lib, a, b, n, row_shift, base_offset, sbo)
        ok = bool(torch.equal(out, expected(row_shift, sbo)))
        results[f"shift{row_shift}_bo{base_offset}_sbo{sbo}"] = ok
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe_umma.json", "w") as f:
        json.dump(results, f, indent=1)
    print("PROBE", json.dumps(results))
    assert results["shift0_bo0_sbo1024"], "canonical aligned SW128 K-major operand mismatch"
    assert results["shift8_bo0_sbo1024"] and results["shift16_bo0_sbo1024"]
