#!/bin/bash
# Quick GPU visit: GPU tests (3 processes), then the two bench lines.  usage: gpu_quick.sh TAG [extra pytest args]
mkdir -p gpurun_out
TAG=${1:-quick}
timeout 600 python -m pytest tests -m gpu -q -n 3 --no-header -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|error" gpurun_out/${TAG}_pytest_gpu.log | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/${TAG}_pytest_gpu.log | head -20
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_train.json 2> gpurun_out/${TAG}_bench_train.err
tail -c 300 gpurun_out/${TAG}_bench_train.err; head -c 400 gpurun_out/${TAG}_bench_train.json; echo
timeout 200 python bench.py --workload inference --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_infer.json 2> gpurun_out/${TAG}_bench_infer.err
tail -c 300 gpurun_out/${TAG}_bench_infer.err; head -c 300 gpurun_out/${TAG}_bench_infer.json; echo
timeout 200 python scripts/profile_step.py 16 > gpurun_out/${TAG}_torchprof.txt 2>&1
head -n 24 gpurun_out/${TAG}_torchprof.txt
