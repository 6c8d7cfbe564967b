#!/bin/bash
# Runs the GPU test files in separate processes (a trapped kernel poisons only its own process) and
# collects the logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
for t in "$@"; do
  name=$(basename "$t" .py)
  echo "=== $t"
  timeout 600 python -m pytest "$t" -m gpu -q -s --durations=4 --no-header -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  echo "exit $?" >> "gpurun_out/$name.log"
  tail -n 25 "gpurun_out/$name.log"
done
