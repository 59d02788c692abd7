"""Attribute a kernel's executed instructions and warp-stall samples to its phases (the code between two BAR.SYNC),
from the source page of an ncu report taken with --set full --import-source on:

    ncu -i REPORT.ncu-rep --page source --csv --kernel-name regex:photo_bwd > bwd.csv
    python profiles/phase_breakdown.py bwd.csv [phase names ...]

Prints a markdown table: share of executed warp instructions, share of stall samples, dominant stall reasons and the
most frequent opcodes per phase."""
import csv
import sys


def main():
    path, names = sys.argv[1], sys.argv[2:]
    rows = list(csv.reader(open(path)))
    kernel = rows[0][1] if rows and len(rows[0]) > 1 else '?'
    hdr, data = rows[1], rows[2:]
    i_src, i_samp, i_exec = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    stalls = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
    i_stall = {n: hdr.index(n) for n in stalls}
    segs, cur = [], None
    for k, r in enumerate(data):
        if cur is None:
            cur = {'n': 0, 'inst': 0, 'samp': 0, 'ops': {}, 'stall': {n: 0 for n in stalls}}
        src = r[i_src].strip()
        tok = src.split()
        op = (tok[1] if tok[0].startswith('@') else tok[0]).split('.')[0] if tok else '?'
        ex = int(r[i_exec])
        cur['n'] += 1; cur['inst'] += ex; cur['samp'] += int(r[i_samp])
        cur['ops'][op] = cur['ops'].get(op, 0) + ex
        for n in stalls:
            v = r[i_stall[n]]
            cur['stall'][n] += int(v) if v.isdigit() else 0
        if 'BAR.SYNC' in src or k == len(data) - 1:
            segs.append(cur); cur = None
    tot_i = sum(s['inst'] for s in segs) or 1
    tot_s = sum(s['samp'] for s in segs) or 1
    print(f'`{kernel}`: {tot_i} warp instructions, {tot_s} stall samples\n')
    print('| phase | SASS instr | share of executed instr | share of stall samples | top stall reasons | top opcodes |')
    print('|---|---:|---:|---:|---|---|')
    for k, s in enumerate(segs):
        if s['inst'] < 0.005 * tot_i and s['samp'] < 0.005 * tot_s:
            continue
        name = names[k] if k < len(names) else f'#{k}'
        st = sum(s['stall'].values()) or 1
        top_st = ', '.join(f"{n[6:]} {100 * v // st}%" for n, v in sorted(s['stall'].items(), key=lambda x: -x[1])[:3])
        top_op = ', '.join(f"{n} {100 * v // max(s['inst'], 1)}%" for n, v in sorted(s['ops'].items(), key=lambda x: -x[1])[:5])
        print(f"| {name} | {s['n']} | {100 * s['inst'] / tot_i:.1f} % | {100 * s['samp'] / tot_s:.1f} % | {top_st} | {top_op} |")


if __name__ == '__main__':
    main()
