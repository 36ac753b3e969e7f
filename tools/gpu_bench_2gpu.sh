#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
echo "exit $?"; wc -c gpurun_out/r02_bench_2gpu.json; tail -5 gpurun_out/r02_bench_2gpu.err | cut -c1-300
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_2gpu_ref.json 2>/dev/null
echo "exit $?"; cut -c1-300 gpurun_out/r02_bench_2gpu_ref.json
