"""Condense an `ncu --set full` report into the handful of per-launch numbers DESIGN.md / bench.py quote.

  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.csv
"""
import csv
import subprocess
import sys

COLS = [
    ('Kernel Name', 'kernel'),
    ('Grid Size', 'grid'),
    ('Block Size', 'block'),
    ('gpu__time_duration.sum', 'time_us'),
    ('dram__bytes_read.sum', 'dram_read_MB'),
    ('dram__bytes_write.sum', 'dram_write_MB'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram_pct'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor_pipe_pct_active'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor_pipe_pct_elapsed'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex_smem_pct'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l2_pct'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_pct'),
    ('launch__registers_per_thread', 'regs'),
    ('launch__shared_mem_per_block_dynamic', 'dyn_smem'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct'),
    ('sm__cycles_elapsed.avg.per_second', 'sm_ghz'),
]


def to_unit(val, unit, want):
    v = float(val.replace(',', ''))
    scale = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'second': 1e6,
             'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}
    return v * scale.get(unit, 1.0)


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    w = csv.writer(sys.stdout)
    w.writerow([name for _, name in COLS] + ['dram_total_MB', 'dram_GBps'])
    for r in rows[2:]:
        out = []
        vals = {}
        for key, name in COLS:
            i = idx.get(key)
            if i is None:
                out.append('')
                continue
            v = r[i]
            if name in ('time_us', 'dram_read_MB', 'dram_write_MB'):
                v = to_unit(v, units[i], name)
                vals[name] = v
                v = f'{v:.3f}'
            elif name == 'kernel':
                v = v.split('(')[0]
            out.append(v)
        tot = vals.get('dram_read_MB', 0) + vals.get('dram_write_MB', 0)
        out += [f'{tot:.3f}', f'{tot / vals["time_us"] * 1e3:.1f}' if vals.get('time_us') else '']
        w.writerow(out)


if __name__ == '__main__':
    main()
