#!/usr/bin/env python
"""Executed warp-instructions per SOURCE LINE of one kernel.

Joins the per-SASS-instruction counts of an ncu report (`--page source --csv`, SASS view) with the
line table of the cubin (`nvdisasm -g`): both list the kernel's instructions in the same order.

    python profiles/line_profile.py report.ncu-rep <kernel regex> <cubin> <mangled-name substring> [topN]
"""
import collections
import csv
import re
import subprocess
import sys


def ncu_counts(rep, regex):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{regex}'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
    hdr = rows[hi]
    ie, src, smp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
    res = []
    for r in rows[hi + 1:]:
        if len(r) > ie and r[ie].isdigit():
            res.append((r[src].split()[0] if not r[src].split()[0].startswith('@') else r[src].split()[1],
                        int(r[ie]), int(r[smp]) if r[smp].isdigit() else 0))
    return res


def line_table(cubin, name):
    out = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
    res, cur, active = [], None, False
    for l in out:
        if l.startswith('.text.'):
            active = name in l
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
            res.append(cur)
    return res


def main(rep, regex, cubin, name, top=40):
    counts = ncu_counts(rep, regex)
    lines = line_table(cubin, name)
    n = min(len(counts), len(lines))
    print(f"# {len(counts)} SASS instructions in the report, {len(lines)} in the cubin")
    agg, smp = collections.Counter(), collections.Counter()
    for (op, c, s), ln in zip(counts[:n], lines[:n]):
        agg[ln] += c
        smp[ln] += s
    tot, st = sum(agg.values()), max(sum(smp.values()), 1)
    srcs = {}
    for ln, c in agg.most_common(int(top)):
        if ln is None:
            continue
        f, no = ln
        if f not in srcs:
            import glob
            cand = glob.glob(f'/root/repo/**/{f}', recursive=True)
            srcs[f] = open(cand[0]).read().splitlines() if cand else []
        text = srcs[f][no - 1].strip()[:100] if 0 < no <= len(srcs[f]) else ''
        print(f"{100 * c / tot:5.1f}% inst {100 * smp[ln] / st:5.1f}% stall  {f}:{no:<4d} {text}")


if __name__ == "__main__":
    main(*sys.argv[1:])
