#!/bin/bash
# round-2 GPU pass A: the whole GPU suite (new evidence tests included) + both bench workloads
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_nvsmi.txt 2>&1
python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -s > gpurun_out/r2a_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2a_pytest_gpu.log
tail -5 gpurun_out/r2a_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_train.json 2> gpurun_out/r2a_bench_train.err
tail -c 1500 gpurun_out/r2a_bench_train.json
python bench.py --workload inference --steps 20 --warmup 5 > gpurun_out/r2a_bench_infer.json 2> gpurun_out/r2a_bench_infer.err
tail -c 600 gpurun_out/r2a_bench_infer.json
