"""List the loops (backward branches) of a kernel with their instruction mix.
usage: loops.py <cubin|so> <name substring> [min_fp64]"""
import collections, re, subprocess, sys
lib, key = sys.argv[1], sys.argv[2]
min64 = int(sys.argv[3]) if len(sys.argv) > 3 else 20
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
cur, ins = None, []
for line in txt.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1); continue
    if cur and key in cur:
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
def op(t):
    return re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0]
FP64 = ('DFMA', 'DMUL', 'DADD', 'DSETP')
for addr, t in ins:
    m = re.search(r'BRA\S*\s+(?:!?U?P\d+,\s+)?`?\(?0x([0-9a-f]+)', t)
    if m and int(m.group(1), 16) <= addr:
        lo = int(m.group(1), 16)
        body = [x for a, x in ins if lo <= a <= addr]
        h = collections.Counter(op(x) for x in body)
        n64 = sum(h[k] for k in FP64)
        if n64 >= min64:
            print('loop %05x..%05x: %d instr, %d fp64, %d other | %s' % (
                lo, addr, len(body), n64, len(body) - n64,
                ' '.join('%s:%d' % kv for kv in h.most_common(14))))
