#!/usr/bin/env python
"""Averages the B200BEV_TC_TRACE clock stamps of CTA 0 into a per-tile timeline (tests/trace_tc.py)."""
import collections
import sys


def load(p):
    ev = {0: [], 1: []}
    for l in open(p):
        r, i, c = l.split()
        ev[int(r)].append((int(i, 16), int(c)))
    return ev


def split(events, marker):
    out, cur = [], []
    for i, c in events:
        if i == marker and cur:
            out.append(cur)
            cur = []
        cur.append((i, c))
    out.append(cur)
    return out


for name in sys.argv[1:]:
    ev = load(name)
    print("=====", name)
    for role, marker, label in ((1, 0x200, "epi"), (0, 0x100, "mma")):
        tiles = split(ev[role], marker)
        acc = collections.OrderedDict()
        durs = []
        for t in range(5, min(len(tiles) - 1, 30)):
            t0 = tiles[t][0][1]
            durs.append(tiles[t + 1][0][1] - t0)
            for i, c in tiles[t]:
                acc.setdefault(i, []).append(c - t0)
        print(f"{label}: tile {sum(durs) / len(durs):.0f} clk (min {min(durs)}, max {max(durs)})")
        prev = 0
        for i, v in acc.items():
            m = sum(v) / len(v)
            print(f"   {label} {i:03x}  +{m:7.0f}  (d {m - prev:6.0f})")
            prev = m
