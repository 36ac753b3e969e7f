#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matching_tc.py tests/test_gpu_matching.py -q -s -x -p no:cacheprovider > gpurun_out/c7_pytest_matching.log 2>&1
echo "pytest exit $?" >> gpurun_out/c7_pytest_matching.log
tail -c 2000 gpurun_out/c7_pytest_matching.log
PDS_B200_PROFILE_DETAIL=1 timeout 300 python tools/bench_detail.py 2>&1 | head -8
