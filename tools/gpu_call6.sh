#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/c6_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c6_pytest.log
tail -c 1500 gpurun_out/c6_pytest.log
timeout 600 python bench.py --no-cpu-baseline --extra-configs '' > gpurun_out/c6_bench_graphs.json 2> gpurun_out/c6_bench_graphs.err
timeout 600 python bench.py --no-cpu-baseline --extra-configs '' --graphs 0 > gpurun_out/c6_bench_eager.json 2> gpurun_out/c6_bench_eager.err
tail -3 gpurun_out/c6_bench_graphs.err
python - <<'PY'
import json
for f in ('c6_bench_graphs.json','c6_bench_eager.json'):
    try:
        d=json.load(open('gpurun_out/'+f))
        print(f, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'e2e f32', round(d['e2e_other_images']['value'],1), '1stream', round(d['value_1stream'],1), 'latency', round(d['latency_ms'],3), 'launches', d['gpu_launches'], d['clocks'])
    except Exception as e: print(f, 'failed', e)
PY
