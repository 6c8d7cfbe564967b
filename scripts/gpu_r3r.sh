#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py tests/test_gpu_generator_f16.py tests/test_gpu_generator.py tests/test_gpu_losses.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -1
python bench.py --workload inference --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('infer', d['value'], d['ms_per_step'], d['e2e']['value'])"
python bench.py --steps 10 --warmup 3 --lib-baseline 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 400 ncu --profile-from-start off --set full --clock-control none -k "regex:grad_combine|affine_apply|strip_reduce|upsample2x_bwd|maxpool|fold_inplace|head_bwd|halo_fill" -c 40 -o /tmp/r3r -f python scripts/ncu_step.py 16 train > gpurun_out/r3r.log 2>&1
ncu -i /tmp/r3r.ncu-rep --page raw --csv > gpurun_out/r3r_raw.csv 2>/dev/null
tail -1 gpurun_out/r3r.log
