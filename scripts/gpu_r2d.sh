#!/bin/bash
mkdir -p gpurun_out
UEGAN_DEBUG_CAPTURE=1 python scripts/debug_capture.py 2>&1 | grep -E "capture\]|Error" | cut -c1-300 | head -20
python -m pytest tests/test_gpu_pinned_chain.py tests/test_gpu_backward.py tests/test_gpu_kernels.py -q --timeout=600 -p no:cacheprovider -s -x 2>&1 | grep -E "pinned chain|passed|failed|Error|error|assert" | cut -c1-300 | head -40
