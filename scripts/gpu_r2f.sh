#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider -s > gpurun_out/r2f_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2f_pytest_gpu.log
grep -E "FAILED|passed|failed" gpurun_out/r2f_pytest_gpu.log | tail -25
python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_train.json 2> gpurun_out/r2f_bench_train.err
python bench.py --workload inference --steps 20 --warmup 5 > gpurun_out/r2f_bench_infer.json 2> gpurun_out/r2f_bench_infer.err
python - <<'PY'
import json
for f in ("r2f_bench_train", "r2f_bench_infer"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, {k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "dtype")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["gemm_ms_per_step"], d["config"]["cuda_graph"], d["roofline"]["by_kind_ms_tflops"], d.get("gpu_library_baseline", {}).get("value"))
    except Exception as e:
        print(f, "parse failed", e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
