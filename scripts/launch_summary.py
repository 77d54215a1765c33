"""Per-kernel totals of an ncu launch list (gpu__time_duration.sum CSV): python scripts/launch_summary.py file.csv [n_steps]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
d = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    v = float(r[vi].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ui], 1.0)
    name = r[ki].split('(')[0][:70]
    d[name][0] += 1
    d[name][1] += v
    tot += v
print(f'{"kernel":70s} {"launches/step":>13s} {"us/step":>10s} {"share":>7s}')
for k, (n, t) in sorted(d.items(), key=lambda x: -x[1][1]):
    print(f'{k:70s} {n / steps:13.1f} {t / steps:10.1f} {100 * t / tot:6.1f}%')
print(f'{"total":70s} {"":13s} {tot / steps:10.1f}')
