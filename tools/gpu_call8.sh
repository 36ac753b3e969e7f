#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tcg.py tests/test_gpu_regularization.py tests/test_gpu_network.py tests/test_gpu_embedding.py -q -x -p no:cacheprovider 2>&1 | tail -8
PDS_B200_PROFILE_DETAIL=1 timeout 300 python tools/bench_detail.py > gpurun_out/c8_layers.txt 2>&1
grep -E "splitk|6x18x30|3x9x15|12x36x60|total" gpurun_out/c8_layers.txt
