#!/bin/bash
mkdir -p gpurun_out
SECONDS=0; timeout 1200 python bench.py > gpurun_out/r02_bench_C2.json 2> gpurun_out/r02_bench_C2.err
tail -4 gpurun_out/r02_bench_C2.err | cut -c1-200
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_C2.json'))
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'e2e other', round(d['e2e_other_images']['value'],1), '1stream', round(d['value_1stream'],1), 'latency', round(d['latency_ms'],3), 'launches', d['gpu_launches'], d['clocks'])
print('roofline', {k: d['roofline'][k] for k in ('kernel','bound','achieved','peak','frac','traffic')})
print('cpu', d.get('cpu_baseline'))
for o in d['other_configs']: print(o)
PY
echo "bench wall seconds: $SECONDS"
nvidia-smi --query-gpu=memory.used --format=csv
