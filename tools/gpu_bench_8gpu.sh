#!/bin/bash
# 8 ranks on one box: value vs e2e (NUMA pinning, uint8 images) -- the driver's SCALE run at N = 8
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 30 --warmup 3 --extra-configs '' > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
echo "exit $?"; tail -3 gpurun_out/r02_bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
d = json.loads(open('gpurun_out/r02_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ratio', round(d['e2e']['value'] / d['value'], 4), 'latency', d.get('latency_ms'), d['clocks'])
print({k: v for k, v in d.items() if k in ('numa', 'host', 'impl_config')})
PY
nproc; numactl --hardware 2>/dev/null | head -5; nvidia-smi topo -m 2>/dev/null | head -14
