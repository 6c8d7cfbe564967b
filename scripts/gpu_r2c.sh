#!/bin/bash
mkdir -p gpurun_out
UEGAN_DEBUG_CAPTURE=1 python scripts/debug_capture.py 2>&1 | grep -E "capture|Error" | cut -c1-600 | head -20
python -m pytest tests/test_gpu_generator_f16.py tests/test_gpu_generator.py -q --timeout=600 -p no:cacheprovider -s 2>&1 | tail -40 | cut -c1-400
UEGAN_GD_DTYPE=f16 python bench.py --workload inference --steps 20 --warmup 5 --lib-baseline 0 > gpurun_out/r2c_bench_infer_f16.json 2> gpurun_out/r2c_bench_infer_f16.err
python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_infer_f16.json')); print('f16 infer:', d['value'], d['ms_per_step'], d['e2e'], d['roofline']['by_kind_ms_tflops'])" || tail -5 gpurun_out/r2c_bench_infer_f16.err
