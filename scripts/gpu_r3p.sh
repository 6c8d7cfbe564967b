#!/bin/bash
python -m pytest tests/test_gpu_generator_f16.py tests/test_gpu_generator.py tests/test_gpu_train.py tests/test_gpu_pinned_chain.py tests/test_gpu_zz_fullsize.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
for m in "UEGAN_NO_CAT_BUILD=1" ""; do
echo "== $m"
env $m python bench.py --workload inference --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('infer', d['value'], d['ms_per_step'], d['e2e']['value'])"
env $m python bench.py --steps 10 --warmup 3 --lib-baseline 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | head -c 600
