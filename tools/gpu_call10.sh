#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/e2e_probe.py 2> gpurun_out/c10_probe.err | tee gpurun_out/c10_probe_$N.json
tail -3 gpurun_out/c10_probe.err | cut -c1-300
