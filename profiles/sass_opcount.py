#!/usr/bin/env python
"""Executed-instruction histogram by SASS opcode from `ncu --page source --csv` of one kernel:
   python profiles/sass_opcount.py report.ncu-rep <kernel regex>"""
import collections
import csv
import subprocess
import sys


def main(rep, regex):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{regex}'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
    hdr = rows[hi]
    ie, src, smp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
    ops = collections.Counter()
    samples = collections.Counter()
    tot = 0
    for r in rows[hi + 1:]:
        if len(r) <= ie or not r[ie].isdigit():
            continue
        toks = r[src].split()
        op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
        op = op.split('.')[0]
        n = int(r[ie])
        ops[op] += n
        samples[op] += int(r[smp]) if r[smp].isdigit() else 0
        tot += n
    print(f"total warp instructions {tot}")
    st = sum(samples.values())
    for op, n in ops.most_common(28):
        print(f"{op:12s} {n:12d} {100 * n / tot:5.1f}%   stall-samples {100 * samples[op] / max(st, 1):5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
