#!/usr/bin/env python
"""Lane utilisation and shared-memory wavefront excess per SOURCE LINE of one kernel.

Companion of line_profile.py: joins the per-SASS-instruction columns of an ncu report (`--page source --csv`) --
'Instructions Executed', 'Thread Instructions Executed', 'L1 Wavefronts Shared' / '... Ideal' -- with the cubin's
line table.  A line whose warps run with few active lanes is work that a different thread mapping would halve
(how the per-frame window lists of photo_bwd_kernel were found: its adjoint ran at 16 of 32 lanes).

    python profiles/lane_profile.py report.ncu-rep <kernel regex> <cubin> <mangled-name substring> [topN]
"""
import collections
import csv
import subprocess
import sys

from line_profile import line_table


def main(rep, regex, cubin, name, top=30):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{regex}'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
    h = rows[hi]
    ie, te = h.index('Instructions Executed'), h.index('Thread Instructions Executed')
    ws, wi = h.index('L1 Wavefronts Shared'), h.index('L1 Wavefronts Shared Ideal')
    recs = [(int(r[ie]), int(r[te]), int(r[ws] or 0), int(r[wi] or 0)) for r in rows[hi + 1:] if len(r) > te and r[ie].isdigit()]
    lines = line_table(cubin, name)
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    for rec, ln in zip(recs, lines):
        for k in range(4):
            agg[ln][k] += rec[k]
    tot = sum(v[0] for v in agg.values())
    tth = sum(v[1] for v in agg.values())
    print(f"# {tot} warp-instructions, {tth / max(tot, 1):.1f} active lanes on average; "
          f"shared wavefronts {sum(v[2] for v in agg.values())} (ideal {sum(v[3] for v in agg.values())})")
    srcs = {}
    for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(top)]:
        text = ''
        if ln:
            try:
                if ln[0] not in srcs:
                    import glob
                    cand = glob.glob(f'**/{ln[0]}', recursive=True)
                    srcs[ln[0]] = open(cand[0]).read().split('\n') if cand else []
                text = srcs[ln[0]][ln[1] - 1].strip()[:90] if srcs[ln[0]] else ''
            except Exception:
                pass
        print(f"{100 * v[0] / tot:5.1f}% inst  {v[1] / max(v[0], 1):5.1f} lanes  smem x{v[2] / max(v[3], 1):4.2f}  "
              f"{ln[0] if ln else '?'}:{ln[1] if ln else 0}  {text}")


if __name__ == '__main__':
    sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
    main(*sys.argv[1:])
