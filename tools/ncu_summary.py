"""Summarises an .ncu-rep (read on the CPU box with `ncu -i`) into a per-launch table of the
metrics the roofline discussion needs.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ('gpu__time_duration.sum', 'dur_us', 1e-3),
    ('dram__bytes_read.sum', 'dram_rd_MB', 1e-6),
    ('dram__bytes_write.sum', 'dram_wr_MB', 1e-6),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_%', 1),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_%', 1),
    ('sm__inst_executed_pipe_tensor.sum', 'tensor_inst', 1),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_%', 1),
    ('lts__t_sector_hit_rate.pct', 'l2_hit_%', 1),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ_%', 1),
    ('launch__registers_per_thread', 'regs', 1),
    ('launch__grid_size', 'grid', 1),
    ('smsp__cycles_active.avg', 'act_cyc', 1),
    ('sm__cycles_elapsed.max', 'ela_cyc', 1),
]
UNIT_SCALE = {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6,
              'Gbyte': 1e9}

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
print(f'# {rep}: {len(data)} launches (ncu --set full --clock-control none; cold cache, serialised)')
print('%-58s' % 'kernel' + ''.join('%12s' % m[1] for m in METRICS))
for r in data:
    name = r[col['Kernel Name']].replace('pds::', '')[:56]
    out = '%-58s' % name
    for key, _, scale in METRICS:
        if key not in col or r[col[key]] in ('', 'n/a'):
            out += '%12s' % '-'
            continue
        v = float(r[col[key]].replace(',', '')) * UNIT_SCALE.get(units[col[key]], 1.0) * scale
        out += '%12.1f' % v
    print(out)
