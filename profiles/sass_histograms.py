#!/usr/bin/env python
"""Per-kernel SASS opcode histograms of the shipped objects (cuobjdump -sass on build/obj/*.o), written as markdown.
Evidence for "which instructions does each kernel really run": TMA bulk copies (UBLKCP), cp.async (LDGSTS), mbarrier
(SYNCS), packed FP32 (FFMA2/FMUL2/FADD2), vector reductions (REDG ... F32x4), multimem (LDGMC), MUFU, etc.

    python profiles/sass_histograms.py > profiles/r2_sass_histograms.md
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OBJ = ROOT / 'faster-gaussian-splatting_b200' / 'build' / 'obj'
INTEREST = ['UBLKCP', 'LDGSTS', 'SYNCS', 'FFMA2', 'FMUL2', 'FADD2', 'REDG', 'RED', 'ATOMG', 'ATOMS', 'LDGMC', 'MUFU', 'LDG', 'STG', 'LDS', 'STS',
            'SHFL', 'VOTE', 'MATCH', 'BAR', 'FFMA', 'FMUL', 'FADD', 'IMAD', 'LDL', 'STL', 'UTMALDG', 'UTCHMMA', 'HMMA', 'NANOSLEEP']


def demangle(name):
    try:
        return subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


def main():
    print('# SASS opcode histograms per kernel (sm_100a objects as shipped in libfgs_b200.so)\n')
    print('`python profiles/sass_histograms.py`; counts are static instruction counts of the whole kernel body.\n')
    for obj in sorted(OBJ.glob('*.o')):
        out = subprocess.run(['cuobjdump', '-sass', str(obj)], capture_output=True, text=True).stdout
        kernels = re.split(r'\n\s*Function : ', out)[1:]
        if not kernels:
            continue
        print(f'## {obj.stem}.cu\n')
        for k in kernels:
            name = demangle(k.split('\n', 1)[0].strip())
            short = name.replace('(anonymous namespace)::', '').replace('void ', '')
            short = re.sub(r'\(.*', '', short)
            ops = collections.Counter()
            full = collections.Counter()
            for m in re.finditer(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_x]+)*)', k):
                ops[m.group(1)] += 1
                full[m.group(1) + m.group(2)] += 1
            total = sum(ops.values())
            if total == 0:
                continue
            hl = ', '.join(f'{o} {ops[o]}' for o in INTEREST if ops.get(o))
            vec = ', '.join(f'{o} {c}' for o, c in sorted(full.items()) if any(t in o for t in ('.128', 'F32x4', 'x2', '.64', 'BULK', 'EX2', 'RCP', 'RSQ', 'LG2', 'SQRT')) and c)
            print(f'* `{short}`: {total} instructions; {hl}')
            if vec:
                print(f'  * vector / special forms: {vec}')
        print()


if __name__ == '__main__':
    main()
