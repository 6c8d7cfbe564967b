#!/bin/bash
# final round-2 pass on one GPU: all GPU tests, training + inference + sweep bench lines, reference arm, ncu launch list
TAG=${1:-r3m}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/${TAG}_nvsmi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -n 3 --no-header -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|error" gpurun_out/${TAG}_pytest_gpu.log | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/${TAG}_pytest_gpu.log | head -20
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_train.json 2> gpurun_out/${TAG}_bench_train.err
python - <<PY
import json
for w in ("train",):
    try:
        d = json.load(open(f"gpurun_out/${TAG}_bench_{w}.json"))
        print(w, {k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"] if d.get("roofline") else None, d.get("gpu_library_baseline"), d["cpu_baseline"]["value"], d["clocks"])
    except Exception as e:
        print(w, "parse failed", e); print(open(f"gpurun_out/${TAG}_bench_{w}.err").read()[-1500:])
PY
UEGAN_TRACE_OUT=gpurun_out/${TAG}_trace.json timeout 1200 ncu --profile-from-start off \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
  --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/ncu_step.py 16 train > gpurun_out/${TAG}_ncu.log 2>&1
python scripts/ncu_join.py gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_trace.json > gpurun_out/${TAG}_train_step_launches.md 2> gpurun_out/${TAG}_join.err
cat gpurun_out/${TAG}_join.err; head -3 gpurun_out/${TAG}_train_step_launches.md
