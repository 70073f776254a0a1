#!/usr/bin/env python
"""Opcode histogram of one kernel (or of its hottest loop) from `cuobjdump -sass`.

usage: sass_stats.py <lib.so> <substring of mangled name> [--loop]
With --loop, restricts to the innermost backward-branch region that contains
the most FP64 instructions (the per-step loop of the integration kernel).
"""
import collections
import re
import subprocess
import sys

FP64 = ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX')


def functions(lib):
    txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    cur, out = None, {}
    for line in txt.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1); out[cur] = []
        elif cur is not None:
            m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
            if m:
                out[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opcode(ins):
    ins = re.sub(r'^@!?U?P\d+\s+', '', ins)
    return ins.split()[0].split('.')[0]


def main():
    lib, key = sys.argv[1], sys.argv[2]
    loop = '--loop' in sys.argv
    for name, ins in functions(lib).items():
        if key not in name:
            continue
        region = ins
        if loop:
            best = None
            for addr, text in ins:
                m = re.search(r'BRA\S*\s+.*?0x([0-9a-f]+)', text)
                if m and int(m.group(1), 16) < addr:
                    lo, hi = int(m.group(1), 16), addr
                    body = [(a, t) for a, t in ins if lo <= a <= hi]
                    n64 = sum(opcode(t) in FP64 for _, t in body)
                    if best is None or n64 > best[0] or (n64 == best[0] and len(body) < len(best[1])):
                        best = (n64, body)
            region = best[1] if best else ins
        hist = collections.Counter(opcode(t) for _, t in region)
        n64 = sum(hist[k] for k in FP64)
        print('%s\n  instructions=%d  fp64-pipe=%d' % (name, len(region), n64))
        print('  ' + '  '.join('%s:%d' % kv for kv in hist.most_common(40)))


if __name__ == '__main__':
    main()
