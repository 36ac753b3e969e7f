#!/bin/bash
# GPU call 1 of round 2: state after the parity / default-precision changes + measurements the judge asked for.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/c1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c1_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/c1_smoke.log 2>&1
for w in C2 C3 C4; do
  timeout 600 python tools/precision_study.py --workload $w --precisions fp16x2,bf16x3,fp32 --json gpurun_out/c1_precision_$w.json > gpurun_out/c1_precision_$w.txt 2>&1
done
timeout 600 python tools/aten_gpu_bench.py --workload C2 --json gpurun_out/c1_aten_gpu_C2.json > gpurun_out/c1_aten_gpu_C2.txt 2>&1
timeout 300 python tools/matching_volume_bench.py --workload C2 --json gpurun_out/c1_matching_volume_C2.json > gpurun_out/c1_matching_volume_C2.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"matching_(concat|stack)" -c 6 --csv --log-file gpurun_out/c1_ncu_matching_volume.csv python tools/matching_volume_bench.py --workload C2 --reps 1 > /dev/null 2>&1
timeout 900 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
tail -c 1500 gpurun_out/c1_pytest.log
cat gpurun_out/c1_smoke.log | tail -3
