"""SASS opcode evidence for profiles/: per kernel of libdiffndm_b200.so the counts of the Blackwell-native mnemonics
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA loads/stores, UBLKCP/UBLKPF = bulk copies /
prefetch, SYNCS = mbarrier, MUFU, FFMA2, HFMA2) plus registers / spills from ptxas -v.

    python scripts/sass_summary.py > profiles/r2_sass_summary.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'diffndm_b200', 'lib', 'libdiffndm_b200.so')
OPS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UBLKPF', 'UTCATOMSWS', 'SYNCS', 'MUFU', 'FFMA2', 'HFMA2', 'HADD2',
       'HMMA', 'ACQBULK', 'NANOSLEEP']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
            kernels[cur] = {}
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and cur:
            op = m.group(1)
            kernels[cur][op] = kernels[cur].get(op, 0) + 1
            kernels[cur]['_total'] = kernels[cur].get('_total', 0) + 1
    print(f'# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a); counts of static instructions per kernel')
    print('kernel | total | ' + ' | '.join(OPS))
    for k, c in sorted(kernels.items()):
        print(f'{k} | {c.get("_total", 0)} | ' + ' | '.join(str(c.get(o, 0)) for o in OPS))
    tot = {o: sum(c.get(o, 0) for c in kernels.values()) for o in OPS}
    print('ALL | - | ' + ' | '.join(str(tot[o]) for o in OPS))


if __name__ == '__main__':
    main()
