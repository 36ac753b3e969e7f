#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/c5_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c5_pytest.log
tail -c 1200 gpurun_out/c5_pytest.log
PDS_B200_PROFILE_DETAIL=1 timeout 300 python tools/bench_detail.py 2>&1 | head -12
timeout 600 python bench.py --no-cpu-baseline --extra-configs '' > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
PDS_B200_DYNAMIC_CONV=0 timeout 600 python bench.py --no-cpu-baseline --extra-configs '' > gpurun_out/c5_bench_static.json 2> gpurun_out/c5_bench_static.err
python - <<'PY'
import json
for f in ('c5_bench.json','c5_bench_static.json'):
    try:
        d=json.load(open('gpurun_out/'+f))
        print(f, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), '1stream', round(d['value_1stream'],1), 'latency', round(d['latency_ms'],3), d['clocks'])
    except Exception as e: print(f, 'failed', e)
PY
