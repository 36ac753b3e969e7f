#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_kernel -s 16 -c 3 -o /tmp/ncu/fused -f python tools/profile_stages.py --precision fp16x2 --reps 1 --stages network > gpurun_out/c3_ncu.log 2>&1
ncu -i /tmp/ncu/fused.ncu-rep --page raw --csv > /tmp/ncu/fused_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/ncu/fused_raw.csv')))
hdr, units, vals = rows[0], rows[1], rows[2:]
want = ('gpu__time_duration.sum', 'sm__throughput', 'issue_active', 'inst_executed.sum', 'warps_active', 'registers',
        'dram__bytes', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate', 'lts__t_sectors_srcunit_tex_op_read', 'pipe_tensor', 'l1tex__data_pipe', 'smsp__warp_issue_stalled', 'cycles_active',
        'lts__t_sectors_op_atom', 'lts__t_sectors_op_red', 'l1tex__t_bytes', 'lts__throughput', 'gpu__dram_throughput', 'l1tex__m_xbar2l1tex_read_bytes', 'lts__t_requests')
with open('gpurun_out/c3_ncu_fused.txt', 'w') as f:
    for v in vals:
        f.write('== %s grid %s\n' % (v[hdr.index('Kernel Name')][:90], v[hdr.index('Grid Size')] if 'Grid Size' in hdr else ''))
        for h, u, x in zip(hdr, units, v):
            if any(w in h for w in want) and x not in ('', '0'):
                f.write('  %-100s %18s %s\n' % (h, x, u))
PY
grep -E "^==|gpu__time_duration.sum|dram__bytes_(read|write).sum |pipe_tensor_cycles_active.avg.pct|lts__t_sector_hit_rate.pct|lts__throughput.avg.pct" gpurun_out/c3_ncu_fused.txt | head -40
tail -5 gpurun_out/c3_ncu.log
ncu -i /tmp/ncu/fused.ncu-rep --page source --csv --print-source sass --kernel-id :::2 > /tmp/ncu/fused_src.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/ncu/fused_src.csv')))
hdr = rows[0]
print(hdr[:40])
PY
cp /tmp/ncu/fused_src.csv gpurun_out/c3_fused_src.csv
ls -la gpurun_out/c3_fused_src.csv
