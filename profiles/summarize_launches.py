"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
    python profiles/summarize_launches.py gpurun_out/launches.csv [steps] > profiles/rNN_launches.md
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    hdr, data = None, []
    for r in csv.reader(open(path, errors='replace')):
        if 'Kernel Name' in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        name = d['Kernel Name']
        short = re.sub(r'^void ', '', name)
        m = re.match(r'(clslam::\w+)<(.*)', short)
        if m:
            ops = re.findall(r'clslam::(\w+)', m.group(2))
            nums = re.findall(r'^\(int\)(\d+), \(int\)(\d+), \(int\)(\d+)', m.group(2))
            tile = 'x'.join(nums[0][:2]) if nums else ''
            short = f'{m.group(1)}<{tile} {",".join(ops)}>'
        else:
            short = re.sub(r'\(.*', '', short)[:90]
        v = float(d['Metric Value'].replace(',', ''))
        unit = d['Metric Unit']
        us = v / 1000 if unit in ('ns', 'nsecond') else (v * 1000 if unit in ('ms', 'msecond') else v)
        a = agg.setdefault(short, [0, 0.0, d['Grid Size'], d['Block Size']])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f'# {path}: {len(data)} launches over {steps} step(s), {total / steps:.1f} us of kernel time per step\n')
    print('| kernel | launches/step | us/step | share |')
    print('|---|---:|---:|---:|')
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{k}` | {a[0] / steps:.1f} | {a[1] / steps:.1f} | {a[1] / total:.3f} |')


if __name__ == '__main__':
    main()
