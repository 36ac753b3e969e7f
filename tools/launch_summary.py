"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list into
per-kernel totals for ONE step (the launches between two subpixel_map calls).

    python tools/launch_summary.py gpurun_out/launches.csv [step_index]
"""
import collections
import csv
import sys

path = sys.argv[1]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
H = rows[hdr]
data = [r for r in rows[hdr + 1:] if r[0].isdigit()]
ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
names = [r[ki] for r in data]
scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}
vals = [float(r[vi].replace(',', '')) * scale[r[ui]] for r in data]
idx = [i for i, n in enumerate(names) if 'subpixel' in n]
print(f'{len(data)} launches, {len(idx)} steps in the file; showing step {step}')
a, b = idx[step - 1] + 1, idx[step] + 1
agg = collections.OrderedDict()
for n, v in zip(names[a:b], vals[a:b]):
    n = n.replace('void ', '').replace('<unnamed>::', '').replace('unnamed>::', '')[:78]
    agg.setdefault(n, [0, 0.0])
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v for _, v in agg.values())
print(f'{"kernel":80s} {"n":>4s} {"us":>10s} {"share":>6s}')
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{n:80s} {c:4d} {v:10.1f} {100 * v / tot:5.1f}%')
print(f'{"total":80s} {b - a:4d} {tot:10.1f}')
