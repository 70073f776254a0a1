#!/usr/bin/env python
"""Pipe-class stream of the innermost hot loop of a kernel (static SASS).

usage: sass_stream.py <object-or-lib> <mangled-name substring> [min_fp64]
D = FP64 pipe, M = IMAD.WIDE, m = IMAD, a = ALU, X = XU (MUFU/I2F/FLO), L = LDS/LDG/LDC,
B = branch unit, u = uniform datapath.  The loop is the smallest backward-branch
range holding at least `min_fp64` FP64 instructions (cold out-of-line blocks that
ptxas placed inside the range show up as solid integer runs after a B).
"""
import re
import subprocess
import sys
from collections import Counter


def cls(t):
    t = re.sub(r'^@!?U?P\d+\s+', '', t)
    op = t.split()[0].split('.')[0]
    if op in ('DFMA', 'DMUL', 'DADD', 'DSETP'):
        return 'D'
    if op == 'IMAD':
        return 'M' if 'WIDE' in t else 'm'
    if op in ('LOP3', 'SHF', 'IADD3', 'VIADD', 'ISETP', 'FSEL', 'SEL', 'LEA', 'PRMT',
              'VIMNMX', 'MOV', 'IADD', 'PLOP3', 'BREV', 'POPC'):
        return 'a'
    if op in ('MUFU', 'I2F', 'FLO', 'F2I', 'FRND', 'F2F'):
        return 'X'
    if op in ('LDS', 'STS', 'LDG', 'STG', 'LDC', 'LDCU', 'LDGSTS', 'LDSM', 'ST', 'LD'):
        return 'L'
    if op in ('BRA', 'BSSY', 'BSYNC', 'EXIT', 'WARPSYNC', 'BAR', 'CALL', 'RET', 'BRX'):
        return 'B'
    if op.startswith('U') or op in ('R2UR', 'S2UR'):
        return 'u'
    return '?'


def main():
    obj, key = sys.argv[1], sys.argv[2]
    need = int(sys.argv[3]) if len(sys.argv) > 3 else 80
    txt = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    cur, ins = None, []
    for line in txt.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            continue
        if cur and key in cur:
            m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for a, t in ins:
        if 'BRA' in t:
            m = re.search(r'0x([0-9a-f]+)', t)
            if m and int(m.group(1), 16) < a:
                loops.append((int(m.group(1), 16), a))
    best = None
    for lo, hi in loops:
        body = [t for a, t in ins if lo <= a <= hi]
        nd = sum(cls(t) == 'D' for t in body)
        if nd >= need and (best is None or len(body) < len(best[2])):
            best = (lo, hi, body)
    lo, hi, body = best
    s = ''.join(cls(t) for t in body)
    print('loop %#x..%#x: %d instructions' % (lo, hi, len(body)), dict(Counter(s)))
    print(s)
    if '-v' in sys.argv:
        print('\n'.join(body))


if __name__ == '__main__':
    main()
