#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_backward.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
python scripts/layer_bench.py wgrad 2>&1 | tee gpurun_out/r2y_wgrad_bench.txt | grep -E "dec3|dec4|dec5|D.p|total"
echo "== single issuer"; UEGAN_WGRAD_ISSUERS=1 python scripts/layer_bench.py wgrad 2>&1 | grep -E "dec3|dec4|dec5|D.p|total"
