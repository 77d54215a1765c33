"""Per-instruction stall samples of one kernel of an ncu report (--set full): top instructions and the cumulative sample
count at every barrier / mbarrier wait, to see which phase of a kernel the time goes to.
usage: python scripts/ncu_hotspots.py report.ncu-rep [launch-skip]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
skip = sys.argv[2] if len(sys.argv) > 2 else '0'
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--launch-skip', skip,
                      '--launch-count', '1'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
print(rows[0][1][:90])
hdr = rows[1]
i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
data = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == 'Address' or r[0] == 'Kernel Name':
        break
    data.append(r)
tot = sum(int(r[i_s]) for r in data)
print('samples', tot, 'instructions', len(data), 'warp instructions executed', sum(int(r[i_ex]) for r in data))
agg = {}
for r in data:
    for k in stalls:
        v = r[hdr.index(k)]
        if v and v != '0':
            agg[k] = agg.get(k, 0) + int(v)
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
top = sorted(range(len(data)), key=lambda i: -int(data[i][i_s]))[:24]
for i in sorted(top):
    r = data[i]
    st = {k: int(r[hdr.index(k)]) for k in stalls if r[hdr.index(k)] not in ('', '0')}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:2])
    print(i, r[i_src].strip()[:50], r[i_s], r[i_ex], st)
cum, out = 0, []
for i, r in enumerate(data):
    cum += int(r[i_s])
    if any(t in r[i_src] for t in ('BAR.SYNC', 'EXIT', 'SYNCS.PHASECHK')):
        out.append((i, r[i_src].strip()[:14], cum))
print(out)
