#!/usr/bin/env python
"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: python profiles/launch_summary.py in.csv"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith('==')))
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        d = agg.setdefault(r[ki].split('(')[0], [0, 0.0])
        d[0] += 1
        d[1] += float(r[vi].replace(',', ''))
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:90]}` | {n} | {t / 1e3:.1f} | {t / tot:.3f} |")


if __name__ == "__main__":
    main(sys.argv[1])
