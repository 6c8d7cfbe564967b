#!/bin/bash
# exactly the driver's command: one process, -x
mkdir -p gpurun_out
time python -m pytest tests/ -x -q -m gpu > gpurun_out/r3_pytest_gpu_serial.log 2>&1
echo "exit $?"; grep -E "passed|failed|error|real" gpurun_out/r3_pytest_gpu_serial.log | tail -5
nvidia-smi --query-gpu=memory.used --format=csv
