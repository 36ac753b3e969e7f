#!/bin/bash
# split-K on/off: bench line (value, latency_ms) without the extra configs
mkdir -p gpurun_out
for sk in 2 4 8 0; do
  PDS_B200_TCG_SPLITK=$sk python bench.py --steps 30 --warmup 5 --extra-configs '' > gpurun_out/c9_bench_sk$sk.json 2> gpurun_out/c9_bench_sk$sk.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/c9_bench_sk$sk.json').read().strip().splitlines()[-1])
print('splitk=$sk', 'value', d['value'], 'e2e', d['e2e']['value'], 'latency_ms', d.get('latency_ms'), '1stream', d.get('value_1stream'), 'launches', d.get('gpu_launches'))
PY
done
