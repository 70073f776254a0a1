
"""Instruction mix of the per-step Philox hot path of an integrate kernel.

usage: hotloop.py <lib.so> <mangled-name substring>
Anchors: from the uniform first Philox round of a step (UIMAD.WIDE.U32 with the
Philox multiplier) to the R2UR that reads the step's store row.
"""
import collections
import re
import subprocess
import sys

FP64 = ('DFMA', 'DMUL', 'DADD', 'DSETP')


def main():
    lib, key = sys.argv[1], sys.argv[2]
    txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    if 'Function' not in txt:
        txt = subprocess.run(['nvdisasm', '-c', lib], capture_output=True, text=True).stdout
    cur, ins = None, []
    for line in txt.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            continue
        if cur and key in cur:
            m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
            if m:
                ins.append(m.group(2).strip())
    start = next(i for i, t in enumerate(ins) if t.startswith('UIMAD.WIDE.U32') and '0x326172a9' in t)
    end = next(i for i in range(start, len(ins)) if ins[i].startswith('R2UR'))
    body = ins[start:end + 1]
    # drop the optional dW dump block (guarded stores)
    hist = collections.Counter(re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0] for t in body)
    n64 = sum(hist[k] for k in FP64)
    print('hot path: %d instr, %d FP64-pipe, %d other' % (len(body), n64, len(body) - n64))
    print(' '.join('%s:%d' % kv for kv in hist.most_common()))
    if '-v' in sys.argv:
        print('\n'.join(body))


if __name__ == '__main__':
    main()
