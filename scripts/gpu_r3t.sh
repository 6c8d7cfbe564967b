#!/bin/bash
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "cat_build or instance_norm or upsample" 2>&1 | tail -1
for m in "" "UEGAN_CAT_BUILD=1" "" "UEGAN_CAT_BUILD=1"; do
echo "== $m"
env $m python bench.py --workload inference --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('infer', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
