#!/bin/bash
TAG=r3e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 3 --no-header -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|error" gpurun_out/${TAG}_pytest_gpu.log | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/${TAG}_pytest_gpu.log | head -20
for m in "UEGAN_SIDE_STREAM=0" "UEGAN_SIDE_STREAM=0 UEGAN_NO_PREMUL=1" ""; do
echo "== $m"
env $m timeout 300 python bench.py --steps 10 --warmup 3 --lib-baseline 0 2>gpurun_out/${TAG}_bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('train', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_kind_ms_tflops'])" || tail -20 gpurun_out/${TAG}_bench.err
done
