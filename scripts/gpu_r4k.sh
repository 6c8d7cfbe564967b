#!/bin/bash
# r4k: item-indexed fold_inplace + per-tap upsample2x_bwd restored: parity tests, then same-box A/B old vs new .so
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py tests/test_gpu_pinned_chain.py tests/test_gpu_train.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
cp uegan_b200/libuegan_sm100.so /tmp/new.so
for rep in 1 2; do
for v in old new; do
if [ $v = old ]; then cp scratch_ab/old.so uegan_b200/libuegan_sm100.so; else cp /tmp/new.so uegan_b200/libuegan_sm100.so; fi
timeout 300 python bench.py --steps 20 --warmup 3 --lib-baseline 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train $v', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']['sm_mhz'])"
done
done
cp /tmp/new.so uegan_b200/libuegan_sm100.so
