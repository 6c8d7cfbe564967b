#!/bin/bash
TAG=${1:-r2n}
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
