#!/bin/bash
mkdir -p gpurun_out
UEGAN_TRACE_OUT=gpurun_out/r3h_trace_infer.json timeout 600 ncu --profile-from-start off \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
  --clock-control none --csv --log-file gpurun_out/r3h_launches_infer.csv python scripts/ncu_step.py 32 inference > gpurun_out/r3h_ncu_infer.log 2>&1
python scripts/ncu_join.py gpurun_out/r3h_launches_infer.csv gpurun_out/r3h_trace_infer.json > gpurun_out/r3h_infer_launches.md 2> gpurun_out/r3h_join.err
cat gpurun_out/r3h_join.err; cat gpurun_out/r3h_infer_launches.md | cut -c1-170
