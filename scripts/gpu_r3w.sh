#!/bin/bash
python -m pytest tests/test_gpu_backward.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -1
python scripts/layer_bench.py wgrad 2>&1 | grep -E "dec5|D.p1|total"
echo "== no zwin64"; UEGAN_NO_ZWIN64=1 python scripts/layer_bench.py wgrad 2>&1 | grep -E "dec5|D.p1|total"
