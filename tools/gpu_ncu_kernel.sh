#!/bin/bash
# ncu --set full of one kernel (regex $1, skip $2, count 1) of a network forward; key metrics + top stall sites
K=${1:-conv_last_kernel}; SKIP=${2:-1}
mkdir -p /tmp/ncu gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $SKIP -c ${3:-1} -o /tmp/ncu/one -f \
    python tools/profile_stages.py --precision fp16x2 --reps 1 --stages network > gpurun_out/ncu_one.log 2>&1
ncu -i /tmp/ncu/one.ncu-rep --page raw --csv > /tmp/ncu/one_raw.csv 2>/dev/null
ncu -i /tmp/ncu/one.ncu-rep --page source --csv --print-source sass > /tmp/ncu/one_src.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/ncu/one_raw.csv')))
hdr, units, vals = rows[0], rows[1], rows[2:]
want = ('gpu__time_duration.sum', 'sm__throughput', 'issue_active', 'inst_executed.sum', 'registers',
        'dram__bytes', 'lts__t_sector_hit_rate.pct', 'pipe_tensor', 'l1tex__data_pipe', 'smsp__warp_issue_stalled', 'cycles_active',
        'lts__throughput', 'gpu__dram_throughput', 'l1tex__data_bank_conflicts', 'smsp__average_warp')
for v in vals:
    print('==', v[hdr.index('Kernel Name')][:90])
    for h, u, x in zip(hdr, units, v):
        if any(w in h for w in want) and x not in ('', '0'):
            print('  %-100s %18s %s' % (h, x, u))
rows = list(csv.reader(open('/tmp/ncu/one_src.csv')))
H = rows[1]; data = [r for r in rows[2:] if len(r) == len(H)]
si = H.index('# Samples'); src = H.index('Source'); ie = H.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(float(r[si] or 0) for r in data)
print('total samples', tot)
for r in sorted(data, key=lambda r: -float(r[si] or 0))[:45]:
    st = sorted(((float(r[i] or 0), H[i]) for i in stall_cols), reverse=True)[:2]
    print('%7.0f %5.1f%% exec %9s  %-70s %s' % (float(r[si]), 100 * float(r[si]) / tot, r[ie], r[src][:70], ' '.join('%s=%.0f' % (n[6:], v) for v, n in st if v)))
PY
