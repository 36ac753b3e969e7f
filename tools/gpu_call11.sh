#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -q -x -p no:cacheprovider 2>&1 | tail -8
timeout 600 python tools/train_step_bench.py > gpurun_out/r02_train_step.json 2> gpurun_out/r02_train_step.err; echo "exit $?"
cat gpurun_out/r02_train_step.json; tail -3 gpurun_out/r02_train_step.err
timeout 600 python tools/train_step_profile.py 2>&1 | cut -c1-230 > gpurun_out/train_profile.txt
