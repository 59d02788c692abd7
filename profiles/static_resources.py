"""cuobjdump -res-usage / -sass of the in-tree library -> markdown table (registers, spills, static shared memory and the
tcgen05 / TMA mnemonics of every kernel).  python profiles/static_resources.py > profiles/r01_static_resources.md"""
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else 'clslam_b200/libclslam_b200.so'
MNEMONICS = ('UTCHMMA', 'UTMALDG', 'UTCBAR', 'LDTM')      # tcgen05.mma, cp.async.bulk.tensor, tcgen05.commit, tcgen05.ld


def main():
    res = subprocess.run(['cuobjdump', '-res-usage', LIB], capture_output=True, text=True, check=True).stdout
    usage, cur = {}, None
    for line in res.splitlines():
        m = re.search(r'Function (\S+):', line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r'REG:(\d+) STACK:(\d+) SHARED:(\d+)', line)
        if m and cur:
            usage[cur] = tuple(map(int, m.groups()))
            cur = None
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = dict.fromkeys(MNEMONICS, 0)
            continue
        if cur:
            for k in MNEMONICS:
                if re.search(r'\b' + k + r'\b', line):
                    counts[cur][k] += 1
    dem = subprocess.run(['c++filt'], input='\n'.join(usage), capture_output=True, text=True, check=True).stdout.splitlines()
    names = dict(zip(usage, dem))

    def short(n):
        n = n.replace('clslam::', '').replace('void ', '')
        return re.sub(r'\(.*$', '', n)[:90]

    print('# Static resource usage and SASS evidence (cuobjdump on the in-tree build, sm_100a)\n')
    print(f'`cuobjdump -res-usage {LIB}` (registers / spill stack bytes / static shared memory; the dynamic shared memory of the')
    print('tcgen05 kernels is chosen at launch, up to 227 KB) and how often the tcgen05 / TMA instructions appear in each kernel\'s')
    print('SASS: UTCHMMA = `tcgen05.mma.kind::f16`, UTMALDG = `cp.async.bulk.tensor` (TMA load), UTCBAR = `tcgen05.commit`,')
    print('LDTM = `tcgen05.ld` (B200_PROFILING.md, "SASS mnemonics").  `conv3x3_tma_kernel<KC, BN, T, true>` are the split-K')
    print('instantiations, `photo_*_occ3` / `photo_bwd_reuse*` the opt-in variants (DESIGN.md section 9).\n')
    print('| kernel | regs | stack B | static smem B | ' + ' | '.join(MNEMONICS) + ' |')
    print('|---|---:|---:|---:|' + '---:|' * len(MNEMONICS))
    for k in sorted(usage, key=lambda k: names[k]):
        r, st, sh = usage[k]
        c = counts.get(k, {})
        print(f"| `{short(names[k])}` | {r} | {st} | {sh} | " + ' | '.join(str(c.get(m, 0)) for m in MNEMONICS) + ' |')


if __name__ == '__main__':
    main()
