#!/bin/bash
mkdir -p gpurun_out
python scripts/debug_capture.py 2>&1 | grep -E "capture\]" | cut -c1-200
python -m pytest tests/test_gpu_train.py tests/test_gpu_optim.py -q --timeout=900 -p no:cacheprovider -s 2>&1 | grep -E "^\[f16\]|step [01]:|two runs|passed|failed|Error|FAILED|assert " | cut -c1-330 | head -40
UEGAN_GD_DTYPE=f16 python bench.py --steps 10 --warmup 3 --lib-baseline 0 > gpurun_out/r2e_bench_train_f16.json 2> gpurun_out/r2e_bench_train_f16.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2e_bench_train_f16.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "dtype")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["gemm_ms_per_step"], d["config"]["cuda_graph"], d["roofline"]["by_kind_ms_tflops"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r2e_bench_train_f16.err").read()[-3000:])
PY
