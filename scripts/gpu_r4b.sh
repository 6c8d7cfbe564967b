#!/bin/bash
python scripts/layer_bench.py dgrad 2>&1 | tail -9
python -m pytest tests/test_gpu_backward.py tests/test_gpu_train.py tests/test_gpu_pinned_chain.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 --lib-baseline 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
