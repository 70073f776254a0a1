"""Warp instructions executed per opcode from an `ncu --page source --csv` dump.

    python tools/ncu_opcodes.py <source.csv> <paths> <steps> > profiles/rNN_ncu_pipe_instructions.csv
"""
import collections
import csv
import re
import sys


def main():
    path, paths, steps = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
    rows = list(csv.reader(open(path)))
    name = rows[0][1]
    hdr = rows[1]
    si, ci = hdr.index('Source'), hdr.index('Instructions Executed')
    stall = hdr.index('Warp Stall Sampling (All Samples)')
    cnt, samples = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        if len(r) <= ci:
            continue
        op = re.sub(r'^@!?U?P\d+\s+', '', r[si].strip()).split()[0].split('.')[0]
        cnt[op] += int(r[ci])
        samples[op] += int(r[stall] or 0)
    ws = paths*steps/32
    print('# %s, %g paths x %g steps: warp instructions executed per opcode' % (name, paths, steps))
    print('# (ncu --set full --import-source on, source page, summed over SASS lines); '
          'per_warp_step = count / %.4g; stall_samples = warp-stall samples attributed to the opcode' % ws)
    print('opcode,instructions_executed,per_warp_step,stall_samples')
    for op, c in cnt.most_common():
        if c:
            print('%s,%d,%.2f,%d' % (op, c, c/ws, samples[op]))
    print('TOTAL,%d,%.2f,%d' % (sum(cnt.values()), sum(cnt.values())/ws, sum(samples.values())))


if __name__ == '__main__':
    main()
