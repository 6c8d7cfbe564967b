#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py tests/test_gpu_generator_f16.py tests/test_gpu_generator.py tests/test_gpu_losses.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
python scripts/layer_bench.py fprop 2>&1 | tee gpurun_out/r3b_fprop_bench.txt
python scripts/layer_bench.py dgrad 2>&1 | tee gpurun_out/r3b_dgrad_bench.txt | tail -9
for m in "UEGAN_CONV_OCC=1" ""; do
echo "== $m"
env $m python bench.py --workload inference --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('infer', d['value'], d['ms_per_step'], d['e2e']['value'])"
env $m python bench.py --steps 10 --warmup 3 --lib-baseline 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train', d['value'], d['ms_per_step'], d['roofline']['by_kind_ms_tflops'])"
done
