#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matching_tc.py tests/test_gpu_matching.py -q -s -x -p no:cacheprovider > gpurun_out/c2_pytest_matching.log 2>&1
echo "pytest exit $?" >> gpurun_out/c2_pytest_matching.log
tail -c 2500 gpurun_out/c2_pytest_matching.log
if grep -q "pytest exit 0" gpurun_out/c2_pytest_matching.log; then
  timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/c2_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/c2_pytest.log
  tail -c 800 gpurun_out/c2_pytest.log
  timeout 600 python bench.py --no-cpu-baseline --extra-configs '' > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
  PDS_B200_FUSE_NORM=0 timeout 600 python bench.py --no-cpu-baseline --extra-configs '' > gpurun_out/c2_bench_nofuse.json 2> gpurun_out/c2_bench_nofuse.err
  PDS_B200_PROFILE_DETAIL=1 timeout 300 python tools/bench_detail.py > gpurun_out/c2_layers.txt 2>&1
  python - <<'PY'
import json
for f in ('c2_bench.json','c2_bench_nofuse.json'):
    try:
        d=json.load(open('gpurun_out/'+f))
        print(f, 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), '1stream', round(d['value_1stream'],1), 'latency', round(d['latency_ms'],3), d['clocks'])
    except Exception as e: print(f, 'failed', e)
PY
  head -30 gpurun_out/c2_layers.txt
fi
