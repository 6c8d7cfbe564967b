#!/bin/bash
# full pass: GPU tests, train bench, ncu launch list of one eager step.  usage: gpu_full.sh TAG
TAG=${1:-full}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 3 --no-header -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|error" gpurun_out/${TAG}_pytest_gpu.log | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/${TAG}_pytest_gpu.log | head -20
timeout 300 python bench.py --steps 10 --warmup 3 --lib-baseline 0 > gpurun_out/${TAG}_bench_train.json 2> gpurun_out/${TAG}_bench_train.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_train.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["gemm_ms_per_step"], d["config"]["cuda_graph"], d["roofline"]["by_kind_ms_tflops"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${TAG}_bench_train.err").read()[-3000:])
PY
UEGAN_TRACE_OUT=gpurun_out/${TAG}_trace.json timeout 1200 ncu --profile-from-start off \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
  --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/ncu_step.py 16 train > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
python scripts/ncu_join.py gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_trace.json > gpurun_out/${TAG}_train_step_launches.md 2> gpurun_out/${TAG}_join.err
cat gpurun_out/${TAG}_join.err; head -5 gpurun_out/${TAG}_train_step_launches.md
