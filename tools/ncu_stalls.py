#!/usr/bin/env python
"""Aggregate the warp-stall samples of one kernel section of `ncu --page source --csv` output by opcode.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_stalls.py src.csv [section]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
heads = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
hi = heads[sec]
end = heads[sec + 1] - 1 if sec + 1 < len(heads) else len(rows)
hdr = rows[hi]
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
print(rows[hi - 1][:2])
tot = sum(int(r[2] or 0) for r in data)
print("total samples", tot)
agg = collections.Counter()
st = collections.defaultdict(collections.Counter)
allst = collections.Counter()
for r in data:
    toks = r[1].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LDS', 'LDG', 'STG', 'SYNCS', 'DMMA')) and '.' in op else '')
    n = int(r[2] or 0)
    agg[op] += n
    for j in range(30, 47):
        st[op][hdr[j]] += int(r[j] or 0)
        allst[hdr[j]] += int(r[j] or 0)
print("by reason:", ", ".join("%s %.1f%%" % (k.replace('stall_', ''), 100 * v / tot) for k, v in allst.most_common(8)))
for op, n in agg.most_common(16):
    print("%-14s %6d %5.1f%%  " % (op, n, 100 * n / tot),
          ", ".join("%s %d" % (k.replace('stall_', ''), v) for k, v in st[op].most_common(4)))
for r in sorted(data, key=lambda r: -int(r[2] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 16]:
    print(r[2], r[1][:80], {hdr[j].replace('stall_', ''): int(r[j]) for j in range(30, 47) if int(r[j] or 0) > 0.15 * int(r[2])})
