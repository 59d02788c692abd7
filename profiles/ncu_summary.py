"""Dump the metrics the roofline block cites from an .ncu-rep (run here, no GPU needed):
    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<kernel>_ncu.md"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum']


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [hdr.index(w) for w in WANT if w in hdr]
    k = hdr.index('Kernel Name')
    print(f'# {rep}\n')
    print('| kernel | ' + ' | '.join(f'{hdr[c]} [{units[c]}]' for c in cols) + ' |')
    print('|---|' + '---:|' * len(cols))
    for r in rows[2:]:
        print(f'| `{r[k][:70]}` | ' + ' | '.join(r[c] for c in cols) + ' |')


if __name__ == '__main__':
    main()
