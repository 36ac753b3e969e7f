#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_matching.py -q -x -p no:cacheprovider 2>&1 | tail -15
echo skip

