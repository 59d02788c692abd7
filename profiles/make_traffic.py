"""profiles/r01_traffic.json: average DRAM bytes (read + write) per launch of each kernel family, from ncu --set full
reports:  python profiles/make_traffic.py gpurun_out/prof_step_b8.ncu-rep:8 gpurun_out/prof_step_b1.ncu-rep:1"""
import csv, io, json, subprocess, sys
from pathlib import Path

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
out = {}
p = Path(__file__).with_name('r01_traffic.json')
if p.exists():
    out = json.loads(p.read_text())
for arg in sys.argv[1:]:
    rep, batch = arg.rsplit(':', 1)
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    k, ir, iw = hdr.index('Kernel Name'), hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    it = hdr.index('gpu__time_duration.sum')
    acc = {}
    for r in data:
        name = r[k]
        fam = next((f for f in ('conv3x3_tma_kernel', 'wgrad3x3_tma_kernel', 'photo_fwd_kernel', 'photo_bwd_kernel', 'gemm_tc_kernel')
                    if f in name), None)
        if fam is None:
            continue
        if fam == 'gemm_tc_kernel':
            fam = 'gemm_tc_kernel (gather)'
        b = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
        a = acc.setdefault(fam, [0.0, 0, 0.0])
        a[0] += b; a[1] += 1; a[2] += float(r[it])
    for fam, (b, n, t) in acc.items():
        out[f'{fam}@b{batch}'] = round(b / n)
        print(f'{fam}@b{batch}: {n} launches, {b / n / 1e6:.2f} MB / launch, {t / n:.1f} us / launch (under ncu)')
p.write_text(json.dumps(out, indent=1, sort_keys=True))
