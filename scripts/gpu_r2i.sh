#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider -s > gpurun_out/r2i_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2i_pytest_gpu.log
grep -E "FAILED|passed|failed" gpurun_out/r2i_pytest_gpu.log | tail -25
python bench.py --steps 10 --warmup 3 --lib-baseline 0 > gpurun_out/r2i_bench_train.json 2> gpurun_out/r2i_bench_train.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2i_bench_train.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["gemm_ms_per_step"], d["config"]["cuda_graph"], d["roofline"]["by_kind_ms_tflops"], d["roofline"]["traffic_over_algorithmic"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r2i_bench_train.err").read()[-3000:])
PY
