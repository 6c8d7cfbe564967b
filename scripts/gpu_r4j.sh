#!/bin/bash
# r4j: inference with the halo_fill launches restored: parity tests of the inference path, final inference + sweep bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_generator.py tests/test_gpu_generator_f16.py tests/test_gpu_io_metrics.py tests/test_gpu_zz_fullsize.py tests/test_gpu_dropin_main.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
timeout 300 python bench.py --workload inference --steps 30 --warmup 5 > gpurun_out/r4z_bench_infer.json 2> gpurun_out/r4z_bench_infer.err
timeout 300 python bench.py --workload sweep --steps 5 --warmup 3 > gpurun_out/r4z_bench_sweep.json 2> gpurun_out/r4z_bench_sweep.err
python - <<'PY'
import json
for w in ("infer", "sweep"):
    d = json.load(open(f"gpurun_out/r4z_bench_{w}.json"))
    print(w, d["value"], d["ms_per_step"], d.get("e2e"), d["gpu_launches"], d["clocks"], d.get("gpu_library_baseline", {}).get("value"))
PY
