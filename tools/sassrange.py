"""Opcode histogram of a SASS address range: sassrange.py <cubin|so> <name-substr> <lo-hex> <hi-hex> [-v]"""
import collections, re, subprocess, sys
lib, key, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
cur, ins = None, []
for line in txt.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1); continue
    if cur and key in cur:
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if m and lo <= int(m.group(1), 16) < hi:
            ins.append((m.group(1), m.group(2).strip()))
hist = collections.Counter(re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0] for _, t in ins)
n64 = sum(hist[k] for k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))
print('%d instr, %d FP64-pipe, %d other' % (len(ins), n64, len(ins) - n64))
print(' '.join('%s:%d' % kv for kv in hist.most_common()))
if '-v' in sys.argv:
    for a, t in ins: print(a, t)
