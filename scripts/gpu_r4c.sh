#!/bin/bash
python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "reflect_halo" 2>&1 | tail -3
python -m pytest tests/test_gpu_generator_f16.py tests/test_gpu_generator.py tests/test_gpu_losses.py tests/test_gpu_train.py tests/test_gpu_pinned_chain.py tests/test_gpu_zz_fullsize.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -1
python bench.py --workload inference --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('infer', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
python bench.py --steps 20 --warmup 3 --lib-baseline 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
