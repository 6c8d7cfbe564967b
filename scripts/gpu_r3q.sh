#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --profile-from-start off --set full --clock-control none -k "regex:in_apply|upsample2x|maxpool2x2_kernel|grad_combine" -c 8 -o /tmp/r3q -f python scripts/ncu_step.py 16 train > gpurun_out/r3q.log 2>&1
ncu -i /tmp/r3q.ncu-rep --page raw --csv > gpurun_out/r3q_raw.csv 2>/dev/null
tail -2 gpurun_out/r3q.log
