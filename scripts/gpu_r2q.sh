#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py tests/test_gpu_generator_f16.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15
for m in "" "UEGAN_NO_PATCH16=1" "UEGAN_NO_PATCH16=1 UEGAN_NO_PATCH64=1"; do
  echo "== $m"; env $m python bench.py --workload inference --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
done
