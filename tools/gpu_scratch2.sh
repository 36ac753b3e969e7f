python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/ncu/one_src.csv')))
H = rows[1]; data = [r for r in rows[2:] if len(r) == len(H)]
si = H.index('# Samples'); src = H.index('Source'); ie = H.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h]
idx = [i for i, r in enumerate(data) if 'UTCHMMA' in r[src]]
lo, hi = max(0, idx[0] - 70), min(len(data), idx[-1] + 40)
for r in data[lo:hi]:
    st = sorted(((float(r[i] or 0), H[i]) for i in stall_cols), reverse=True)[:2]
    print('%6.0f exec %8s  %-90s %s' % (float(r[si] or 0), r[ie], r[src][:90], ' '.join('%s=%.0f' % (n[6:], v) for v, n in st if v)))
PY
