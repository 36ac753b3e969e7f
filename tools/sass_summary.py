"""Per-kernel count of the SASS instructions that prove a Blackwell-native path (runs without a
GPU: cuobjdump on the built library): UTC*MMA (tcgen05.mma), UTMALDG / UBLKCP (TMA), LDTM / STTM
(tcgen05.ld / st), UTCBAR (tcgen05.commit), SYNCS (mbarrier), HMMA (legacy mma.sync: must be 0).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'practicaldeepstereo_nips2018_b200', 'libpds_b200.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'UTMALDG', 'UBLKCP', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'HMMA', 'FFMA', 'RED', 'ATOM']

sass = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
names = subprocess.run(['cu++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)),
                       stdout=subprocess.PIPE, text=True).stdout.splitlines()
counts, order, current, total = {}, [], None, collections.Counter()
it = iter(names)
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        current = next(it)
        current = current.replace('(anonymous namespace)::', '').replace('<unnamed>::', '').replace('void ', '')
        current = re.sub(r'\((int|bool|unsigned int)\)', '', current)
        current = re.sub(r'\(.*\)$', '', current)
        while current in counts:
            current += "'"

        counts[current] = collections.Counter()
        order.append(current)
        continue
    if current is None:
        continue
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1).split('.')[0]
        counts[current]['_all'] += 1
        for k in KEYS:
            if op == k or (k in ('HMMA',) and op.startswith(k)):
                counts[current][k] += 1
                total[k] += 1
print(f'cuobjdump -sass {os.path.relpath(LIB, ROOT)}: {len(order)} kernels (sm_100a)')
print('library totals: ' + ', '.join(f'{k} {total[k]}' for k in KEYS))
print()
print(f'{"kernel":110s} {"instr":>7s} ' + ' '.join(f'{k:>8s}' for k in KEYS))
for name in sorted(order, key=lambda n: -(counts[n]['UTCHMMA'] + counts[n]['UTCQMMA'])):
    c = counts[name]
    print(f'{name[:110]:110s} {c["_all"]:7d} ' + ' '.join(f'{c[k]:8d}' for k in KEYS))
