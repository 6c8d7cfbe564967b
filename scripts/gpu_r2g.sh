#!/bin/bash
# round-2 pass G: fp16 row-sum unit cases, D golden, then the ncu launch list (time + DRAM bytes + tensor pipe) of ONE fp16 training step
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_losses.py -q -k "rowsum_f16 or discriminator_vs_golden" -s -p no:cacheprovider 2>&1 | grep -E "rowsum f16|pred[0-9]|passed|failed|Error" | head -40
UEGAN_TRACE_OUT=gpurun_out/r2g_trace.json timeout 1200 ncu --profile-from-start off \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
  --clock-control none --csv --log-file gpurun_out/r2g_launches.csv python scripts/ncu_step.py 16 train > gpurun_out/r2g_ncu.log 2>&1
tail -3 gpurun_out/r2g_ncu.log
python scripts/ncu_join.py gpurun_out/r2g_launches.csv gpurun_out/r2g_trace.json > gpurun_out/r2g_train_step_launches.md 2> gpurun_out/r2g_join.err
cat gpurun_out/r2g_join.err; head -50 gpurun_out/r2g_train_step_launches.md
