#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of one bench run + ncu --set full capture of one
# network forward, summarised ON THE BOX (the .ncu-rep files exceed what gpurun copies back).
#   tools/profile_gpu.sh [precision] [tag]
P=${1:-fp16x2}
TAG=${2:-r01}
mkdir -p gpurun_out /tmp/ncu
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${TAG}_launches_${P}.csv \
    python bench.py --precision $P --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches_${P}.csv 3 > gpurun_out/${TAG}_launches_${P}.txt 2>&1
rm -f gpurun_out/${TAG}_launches_${P}.csv
# one network forward, marked by an NVTX range (the earlier forwards prepare weights and warm up)
ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "capture/" \
    -o /tmp/ncu/${TAG}_network_${P} -f \
    python tools/profile_stages.py --precision $P --reps 1 --stages network --nvtx > gpurun_out/${TAG}_ncu_network.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${TAG}_network_${P}.ncu-rep > gpurun_out/${TAG}_ncu_network_${P}.txt 2>&1
ls -la gpurun_out | tail -8
