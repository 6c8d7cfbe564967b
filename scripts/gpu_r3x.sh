#!/bin/bash
python -m pytest tests/test_gpu_backward.py tests/test_gpu_pinned_chain.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -1
python scripts/layer_bench.py wgrad 2>&1 | grep -E "enc2|D.d2|total"
echo "== no wgrad64"; UEGAN_NO_WGRAD64=1 python scripts/layer_bench.py wgrad 2>&1 | grep -E "enc2|D.d2|total"
