"""Per-launch list of ONE tile batch of the second Predictor pass out of an ncu launch-list CSV (gpurun_out/launches_pred.csv):
python scripts/pred_layers.py file.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
ki, vi, ui, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('Grid Size')
seq = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    v = float(r[vi].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ui], 1.0)
    seq.append((r[ki].split('(')[0][-34:], v, r[gi]))
packs = [i for i, s in enumerate(seq) if s[0].endswith('pack_kernel')]
start, end = packs[-2], packs[-1]            # the last complete tile batch
tot = sum(s[1] for s in seq[start:end])
for s in seq[start:end]:
    print('%-36s %9.1f us %5.1f%%  grid %s' % (s[0], s[1], 100 * s[1] / tot, s[2]))
print('one tile batch: %.1f us' % tot)
