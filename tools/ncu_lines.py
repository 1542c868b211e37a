#!/usr/bin/env python
"""Warp-stall samples of `ncu --page source --csv --print-source cuda,sass` aggregated by CUDA source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python tools/ncu_lines.py src.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur = None
hdr = None
agg = collections.Counter()
reasons = collections.defaultdict(collections.Counter)
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split('/')[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None or r[0] == "":
        continue
    try:
        n = int(r[4] or 0)
    except ValueError:
        continue
    key = (cur, int(r[0]), r[1].strip()[:100])
    agg[key] += n
    for j, h in enumerate(hdr[:len(r)]):
        if h.startswith('stall_') and 'Not Issued' not in h:
            try:
                reasons[key][h[6:]] += int(r[j] or 0)
            except ValueError:
                pass
tot = sum(agg.values())
print('total samples', tot)
for k, n in agg.most_common(top):
    print("%6d %5.1f%%  %s:%d  %s   %s" % (n, 100.0 * n / max(tot, 1), k[0], k[1], k[2], dict(reasons[k].most_common(3))))
