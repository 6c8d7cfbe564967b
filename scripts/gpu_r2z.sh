#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_backward.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -1
python scripts/layer_bench.py wgrad 2>&1 | tee gpurun_out/r2z_wgrad_bench.txt | grep -vE "D.p|dec3|dec4|dec5"
echo "== occ 1"; UEGAN_WGRAD_OCC=1 python scripts/layer_bench.py wgrad 2>&1 | grep -vE "D.p|dec3|dec4|dec5"
