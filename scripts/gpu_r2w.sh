#!/bin/bash
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_backward.py tests/test_gpu_generator_f16.py tests/test_gpu_generator.py -m gpu -q -x -p no:cacheprovider 2>&1 | grep -E "^(FAILED|ERROR)|Error|error|assert|rel err|failed|passed" | head -30
