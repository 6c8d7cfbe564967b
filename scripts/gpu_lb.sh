#!/bin/bash
# layer bench pass: usage gpu_lb.sh TAG [what] [filter]
mkdir -p gpurun_out
python scripts/layer_bench.py ${2:-all} "${3:-}" > gpurun_out/${1}_layer_bench.txt 2>&1
cat gpurun_out/${1}_layer_bench.txt
