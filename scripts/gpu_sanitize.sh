#!/bin/bash
# compute-sanitizer memcheck over the kernel-level parity tests that exercise the round-2 kernels (bounded: 8 minutes)
mkdir -p gpurun_out
timeout 480 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_backward.py -m gpu -q -x -p no:cacheprovider -k "zwin or dgrad_wgrad" > gpurun_out/sanitize_backward.log 2>&1
echo "memcheck backward exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_backward.log | head -8
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python scripts/layer_bench.py fprop "G.enc1" > gpurun_out/sanitize_fprop.log 2>&1
echo "memcheck fprop exit $?"; grep -E "ERROR SUMMARY|Invalid|out of bounds|^fprop" gpurun_out/sanitize_fprop.log | head -6
