#!/bin/bash
# r4g: one batched G pass for G(real_raw) and G(real_exp): trainer tests, then same-box A/B (UEGAN_BATCH_G=0|1)
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_dropin_main.py tests/test_gpu_zz_fullsize.py tests/test_gpu_optim.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
for rep in 1 2; do
for v in 0 1; do
UEGAN_BATCH_G=$v timeout 300 python bench.py --steps 20 --warmup 3 --lib-baseline 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train batch_g=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']['sm_mhz'], d['roofline']['frac'])"
done
done
