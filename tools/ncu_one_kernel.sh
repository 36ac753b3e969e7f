#!/bin/bash
# ncu --set full of ONE kernel class of a network forward, key metrics printed on the box.
#   tools/ncu_one_kernel.sh <kernel regex> [precision] [skip] [count]   (count > 1: the LONGEST captured launch is printed)
K=${1:-hourglass_tail}
P=${2:-fp16x2}
SKIP=${3:-2}
COUNT=${4:-1}
mkdir -p /tmp/ncu gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c $COUNT -o /tmp/ncu/one -f \
    python tools/profile_stages.py --precision $P --reps 1 --stages network > gpurun_out/ncu_one.log 2>&1
ncu -i /tmp/ncu/one.ncu-rep --page raw --csv > /tmp/ncu/one_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/ncu/one_raw.csv')))
hdr, units, vals = rows[0], rows[1], rows[2:]
want = ('gpu__time_duration.sum', 'sm__throughput', 'issue_active', 'inst_executed.sum', 'warps_active', 'registers',
        'issue_stalled', 'dram__bytes', 'lts__t_bytes.sum', 'l1tex__data_bank_conflicts', 'shared', 'pipe_fma', 'pipe_lsu',
        'pipe_alu', 'inst_executed_pipe', 'sm__inst_executed_pipe', 'achieved_occupancy', 'eligible', 'lsu_mem_shared',
        'pipe_tensor', 'uniform', 'l1tex__data_pipe', 'smsp__warp_issue_stalled', 'tmem', 'cycles_active')
di = hdr.index('gpu__time_duration.sum')
vals = [max(vals, key=lambda v: float(v[di].replace(',', '')))]
for v in vals:
    print('==', v[hdr.index('Kernel Name')][:70], 'grid', v[hdr.index('Grid Size')] if 'Grid Size' in hdr else '')
    for h, u, x in zip(hdr, units, v):
        if any(w in h for w in want) and x not in ('', '0'):
            print('  %-90s %16s %s' % (h, x, u))
PY
