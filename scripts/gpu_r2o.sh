#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py -m gpu -q -x -s -p no:cacheprovider -k "in_mse or zwin or dgrad_wgrad" 2>&1 | grep -E "in_mse_joint|passed|failed|Error|error|assert" | head -40
python scripts/layer_bench.py dgrad > gpurun_out/r2o_layer_bench.txt 2>&1; cat gpurun_out/r2o_layer_bench.txt | tail -12
