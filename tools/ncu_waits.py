"""Summarise an ncu report of a warp-specialised kernel: headline metrics + spin counts of every mbarrier wait."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'lts__t_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for d in data:
    print(d[hdr.index('Kernel Name')][:70])
    for w in want:
        if w in hdr:
            i = hdr.index(w); print(f'  {w:70s} {d[i]} {units[i]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
i_src, i_samp, i_exec = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
body = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        break
    if len(r) > i_samp:
        body.append(r)
print('total samples', sum(int(r[i_samp] or 0) for r in body))
agg = {}
for r in body:
    s = r[i_src]
    if 'TRYWAIT' in s:
        off = s.split('+')[-1].split(']')[0]
        agg[off] = agg.get(off, 0) + int(r[i_exec] or 0)
for k in sorted(agg):
    print('  wait', k, agg[k])
top = sorted(body, key=lambda r: -int(r[i_samp] or 0))[:12]
for r in top:
    print('  ', r[i_samp], r[i_exec], r[i_src].strip()[:90])
