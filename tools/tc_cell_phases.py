#!/usr/bin/env python
"""Per-chunk phase durations of the cell-mode layer-5 epilogue from a B200BEV_TC_TRACE dump (events 0x250.., 0x270..0x274)."""
import collections
import sys

ev = []
for l in open(sys.argv[1]):
    r, i, c = l.split()
    if r == '1':
        ev.append((int(i, 16), int(c)))
d = collections.defaultdict(list)
prev = None
names = {0x250: "accumulator seen", 0x270: "in registers", 0x271: "transposed tile stored", 0x272: "barrier 1 passed",
         0x273: "runs walked", 0x274: "barrier 2 passed"}
for i, c in ev:
    k = 0x250 if 0x250 <= i <= 0x257 else i
    if prev and prev[0] in names and k in names:
        d[(prev[0], k)].append(c - prev[1])
    prev = (k, c)
for (a, b), v in sorted(d.items()):
    print(f"{names[a]:24s} -> {names[b]:24s} n {len(v):4d}  avg {sum(v) / len(v):7.0f} clk")
