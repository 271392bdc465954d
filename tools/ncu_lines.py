#!/usr/bin/env python3
"""Per-source-line instruction counts / stall samples / shared wavefronts from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass [-k ...]`.
usage: ncu_lines.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, agg, line = None, None, {}, None
for r in rows:
    if len(r) >= 2 and r[0] in ('File Name', 'File Path'):
        if r[0] == 'File Name':
            cur = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or not r:
        continue
    if r[0] != '':
        try:
            line = (cur, int(r[0]), r[1].strip()[:100])
        except ValueError:
            continue
        agg.setdefault(line, [0, 0, 0, 0])
    elif line:
        try:
            ie = int(r[hdr.index('Instructions Executed')])
            smp = int(r[4])
            wf = int(r[hdr.index('L1 Wavefronts Shared')]) if 'L1 Wavefronts Shared' in hdr else 0
            wfi = (int(r[hdr.index('L1 Wavefronts Shared Ideal')])
                   if 'L1 Wavefronts Shared Ideal' in hdr else 0)
        except (ValueError, IndexError):
            continue
        a = agg[line]
        a[0] += ie; a[1] += smp; a[2] += wf; a[3] += wfi
tot = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print('total warp instructions', tot, 'stall samples', ts)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%5.1f%% inst %5.1f%% smp  shwf %9d/%9d  %s:%d  %s' % (100 * a[0] / tot, 100 * a[1] / ts, a[2], a[3], k[0], k[1], k[2]))
