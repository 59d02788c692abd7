"""Per-launch table from an `ncu --metrics ... --csv` log (long format, one row per launch and metric) of ONE eager step:
    python profiles/metrics_table.py gpurun_out/metrics_r2n_b8_192.csv > profiles/r02_metrics_b8.md
Columns: duration (cold, serialised: compare shares), DRAM read + write bytes, tensor-pipe activity on the active SMs, SM and DRAM
throughput over the whole GPU, warps active, grid, registers."""
import csv
import re
import sys
from collections import OrderedDict

SHORT = OrderedDict([
    ('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'rd MB'), ('dram__bytes_write.sum', 'wr MB'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm %'), ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps %'), ('launch__grid_size', 'grid'),
    ('launch__registers_per_thread', 'regs'), ('smsp__inst_executed.sum', 'Minst')])
UNIT = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3, 'ms': 1e3}


def short_name(k):
    k = re.sub(r'^void ', '', k)
    k = k.replace('clslam::', '')
    m = re.match(r'([\w:]+)(<[^(]*>)?\(', k)
    if not m:
        return k[:60]
    t = (m.group(2) or '')
    t = re.sub(r'\(bool\)', '', t).replace('(int)', '')
    return (m.group(1) + t)[:70]


def main():
    rows = OrderedDict()
    with open(sys.argv[1]) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if 'clslam' not in r['Kernel Name']:
            continue
        e = rows.setdefault(r['ID'], {'name': short_name(r['Kernel Name'])})
        m = r['Metric Name']
        if m in SHORT:
            v = float(r['Metric Value'].replace(',', ''))
            v *= UNIT.get(r['Metric Unit'], 1.0)
            if m == 'smsp__inst_executed.sum':
                v *= 1e-6
            e[m] = v
    total = sum(e.get('gpu__time_duration.sum', 0.0) for e in rows.values())
    print(f'# {sys.argv[1]}: {len(rows)} launches of one eager step, {total:.0f} us summed (cold caches, serialised)\n')
    print('| # | kernel | ' + ' | '.join(SHORT.values()) + ' |')
    print('|---:|---|' + '---:|' * len(SHORT))
    fams = OrderedDict()
    for i, e in enumerate(rows.values()):
        cells = []
        for m in SHORT:
            v = e.get(m)
            cells.append('' if v is None else (f'{v:.0f}' if m in ('launch__grid_size', 'launch__registers_per_thread') else f'{v:.1f}'))
        print(f'| {i} | `{e["name"]}` | ' + ' | '.join(cells) + ' |')
        fam = e['name'].split('<')[0]
        f = fams.setdefault(fam, [0, 0.0, 0.0])
        f[0] += 1; f[1] += e.get('gpu__time_duration.sum', 0.0); f[2] += e.get('dram__bytes_read.sum', 0.0) + e.get('dram__bytes_write.sum', 0.0)
    print('\n| kernel family | launches | us | share | DRAM MB / launch |\n|---|---:|---:|---:|---:|')
    for fam, (n, us, mb) in sorted(fams.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{fam}` | {n} | {us:.0f} | {us / total:.3f} | {mb / n:.2f} |')


if __name__ == '__main__':
    main()
