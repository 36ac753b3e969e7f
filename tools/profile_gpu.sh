#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of one bench run + ncu --set full captures of
# the dominant kernels.  Outputs land in gpurun_out/; summaries are copied to profiles/ by hand.
#   tools/profile_gpu.sh [precision] [tag]
P=${1:-fp16x2}
TAG=${2:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/${TAG}_launches_${P}.csv \
    python bench.py --precision $P --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'conv3x3_tc|tc_norm_split' -s 8 -c 8 -o gpurun_out/${TAG}_matching_${P} -f \
    python tools/profile_stages.py --precision $P --reps 1 --stages matching > gpurun_out/${TAG}_ncu_matching.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'subpixel_map|hourglass_tail|conv_igemm_f32|instance_norm' -s 24 -c 24 -o gpurun_out/${TAG}_regest_${P} -f \
    python tools/profile_stages.py --precision $P --reps 1 --stages estimator,regularization > gpurun_out/${TAG}_ncu_regest.log 2>&1
ls -la gpurun_out
