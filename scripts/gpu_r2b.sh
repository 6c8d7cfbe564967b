#!/bin/bash
# round-2 GPU pass B: whole GPU suite with the fused Adam / gradient sink / deterministic split-K build + training bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -s > gpurun_out/r2b_pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2b_pytest_gpu.log
grep -E "FAILED|passed|failed|pinned chain|reproducible" gpurun_out/r2b_pytest_gpu.log | tail -25
python bench.py --steps 10 --warmup 3 --lib-baseline 0 > gpurun_out/r2b_bench_train.json 2> gpurun_out/r2b_bench_train.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2b_bench_train.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"], d["roofline"]["frac"], d["roofline"]["gemm_ms_per_step"], d["config"]["cuda_graph"], d["replicas"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/r2b_bench_train.err").read()[-3000:])
PY
UEGAN_DETERMINISTIC=0 python bench.py --steps 10 --warmup 3 --lib-baseline 0 > gpurun_out/r2b_bench_train_atomics.json 2> gpurun_out/r2b_bench_train_atomics.err
python -c "
import json; d=json.load(open('gpurun_out/r2b_bench_train_atomics.json')); print('atomics:', d['value'], d['ms_per_step'])"
