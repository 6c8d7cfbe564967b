"""Builds tests/native/libuegan_probe.so (hardware probes for the test-suite) in-tree with nvcc for sm_100a."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libuegan_probe.so")
ROOT = os.path.dirname(os.path.dirname(HERE))


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("probe_lib.cu", "probe.cu")] + [
        os.path.join(ROOT, "uegan_b200", "csrc", f) for f in ("common.cuh", "host_util.cu", "host_util.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler",
           "-fPIC", "-o", LIB, os.path.join(HERE, "probe_lib.cu"), "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libuegan_probe.so")
    return LIB


if __name__ == "__main__":
    print(build(force=True))
