#!/bin/bash
# One GPU-box visit: bench lines, torch-profiler breakdown, ncu launch list, ncu full capture of the top kernels, GPU tests.
mkdir -p gpurun_out
TAG=${1:-r1d}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvsmi.txt 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_train.json 2> gpurun_out/${TAG}_bench_train.err
tail -c 600 gpurun_out/${TAG}_bench_train.err; head -c 700 gpurun_out/${TAG}_bench_train.json; echo
timeout 300 python bench.py --workload inference --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_infer.json 2> gpurun_out/${TAG}_bench_infer.err
head -c 500 gpurun_out/${TAG}_bench_infer.json; echo
timeout 300 python scripts/profile_step.py 16 > gpurun_out/${TAG}_torchprof.txt 2>&1
head -n 30 gpurun_out/${TAG}_torchprof.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/${TAG}_launches.csv python scripts/ncu_step.py 16 train > gpurun_out/${TAG}_ncu_list.log 2>&1
wc -l gpurun_out/${TAG}_launches.csv
if [ "${FULL:-1}" = "1" ]; then
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:conv_wgrad -c 12 -o gpurun_out/${TAG}_wgrad_full -f python scripts/ncu_step.py 4 train > gpurun_out/${TAG}_ncu_full1.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on \
   -k regex:conv_fprop -c 16 -o gpurun_out/${TAG}_fprop_full -f python scripts/ncu_step.py 4 train > gpurun_out/${TAG}_ncu_full2.log 2>&1
ls -la gpurun_out/*.ncu-rep
fi
timeout 900 python -m pytest tests -m gpu -q -x -n 3 --no-header -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -n 8 gpurun_out/${TAG}_pytest_gpu.log
