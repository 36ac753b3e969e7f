#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/final_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/final_pytest.log
tail -c 600 gpurun_out/final_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2>/dev/null; echo "ref exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/final_bench.json'))
print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), '1stream', round(d['value_1stream'],1), 'latency', round(d['latency_ms'],3), 'launches', d['gpu_launches'], d['clocks'])
print([(o['config'], round(o.get('value',0),1), o.get('error')) for o in d['other_configs']])
r=json.load(open('gpurun_out/final_bench_ref.json'))
print('reference arm', round(r['value'],2), r['cpu_baseline']['cores'], r['config']==d['config'], sorted(r.keys()))
PY
