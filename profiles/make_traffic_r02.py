"""profiles/r02_traffic.json: average DRAM bytes (read + write) per launch of each kernel family, from the per-launch ncu metric
passes of one eager step (tools/r2_evidence.sh: dram__bytes_read.sum + dram__bytes_write.sum per launch):
    python profiles/make_traffic_r02.py gpurun_out/metrics_r2n_b1_192.csv:b1@192x640 gpurun_out/metrics_r2n_b8_192.csv:b8@192x640 ...
Keys are '<kernel family as bench.py names it>@b<batch>@<H>x<W>' (read by bench.py for `roofline.traffic`)."""
import csv
import json
import sys
from pathlib import Path

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
FAMILIES = [('conv3x3_tma_kernel', 'conv3x3_tma_kernel'), ('wgrad3x3_tma_kernel', 'wgrad3x3_tma_kernel'), ('photo_fwd', 'photo_fwd_kernel'),
            ('photo_bwd', 'photo_bwd_kernel'), ('photo_identity', 'photo_identity_kernel'), ('gemm_tc_kernel', 'gemm_tc_kernel (gather)'),
            ('dispconv_fwd_kernel', 'dispconv_fwd_kernel'), ('dispconv_wgrad_kernel', 'dispconv_wgrad_kernel'),
            ('dispconv_dgrad_act_kernel', 'dispconv_dgrad_act_kernel'), ('disp_grad_kernel', 'disp_grad_kernel'),
            ('dgrad_fold_kernel', 'dgrad_fold_kernel'), ('adam_kernel', 'adam_kernel'), ('wgrad_reduce_kernel', 'wgrad_reduce_kernel')]
out = {}
p = Path(__file__).with_name('r02_traffic.json')
for arg in sys.argv[1:]:
    path, tag = arg.rsplit(':', 1)
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    per = {}
    for r in csv.DictReader(lines):
        if r['Metric Name'] not in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            continue
        fam = next((name for key, name in FAMILIES if key in r['Kernel Name']), None)
        if fam is None:
            continue
        e = per.setdefault(fam, {})
        e[r['ID']] = e.get(r['ID'], 0.0) + float(r['Metric Value'].replace(',', '')) * UNIT[r['Metric Unit']]
    for fam, ids in per.items():
        out[f'{fam}@{tag}'] = round(sum(ids.values()) / len(ids))
        print(f'{fam}@{tag}: {len(ids)} launches, {sum(ids.values()) / len(ids) / 1e6:.2f} MB / launch')
p.write_text(json.dumps(out, indent=1, sort_keys=True))
