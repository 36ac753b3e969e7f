#!/bin/bash
# Round-2 profile set (runs on the GPU box): bench line, per-layer times, ncu launch list of one step,
# ncu --set full of one forward (NVTX-scoped) summarised on the box.
TAG=${1:-r02}
mkdir -p gpurun_out /tmp/ncu
timeout 900 python bench.py > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench_C2.err
PDS_B200_PROFILE_DETAIL=1 timeout 300 python tools/bench_detail.py > gpurun_out/${TAG}_layers_fp16x2_C2.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --graphs 0 --extra-configs '' > gpurun_out/${TAG}_launches_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv 3 > gpurun_out/${TAG}_launches_network_fp16x2.txt 2>&1
rm -f gpurun_out/${TAG}_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "capture/" \
    -o /tmp/ncu/${TAG}_network -f \
    python tools/profile_stages.py --precision fp16x2 --reps 1 --stages network --nvtx > gpurun_out/${TAG}_ncu_network.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${TAG}_network.ncu-rep > gpurun_out/${TAG}_ncu_network_fp16x2.txt 2>&1
tail -3 gpurun_out/${TAG}_bench_C2.err
head -c 600 gpurun_out/${TAG}_bench_C2.json; echo
tail -5 gpurun_out/${TAG}_launches_network_fp16x2.txt
head -12 gpurun_out/${TAG}_ncu_network_fp16x2.txt
