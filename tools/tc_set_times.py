#!/usr/bin/env python
"""Cell-mode trace (B200BEV_TC_TRACE): per own layer-5 chunk of the traced epilogue warp, time from 'accumulator seen' to 'runs walked'
and the idle time before the next own chunk.   python tools/tc_set_times.py gpurun_out/trace_tc_cell.txt"""
import statistics
import sys

for path in sys.argv[1:]:
    ev = []
    for l in open(path):
        r, i, c = l.split()
        if int(r) == 1:
            ev.append((int(i, 16), int(c)))
    tiles, cur = [], []
    for i, c in ev:
        if i == 0x200 and cur:
            tiles.append(cur)
            cur = []
        cur.append((i, c))
    tiles.append(cur)
    T, gaps, chunks = [], [], None
    for t in tiles[2:-1]:
        seen = [(i, c) for i, c in t if 0x250 <= i < 0x258]
        done = [c for i, c in t if i == 0x273]
        if len(seen) == 4 and len(done) == 4:
            chunks = [i - 0x250 for i, _ in seen]
            T.append([d - s[1] for s, d in zip(seen, done)])
            gaps.append([seen[k + 1][1] - done[k] for k in range(3)])
    print(path, "tiles", len(T), "own chunks", chunks)
    for k in range(4):
        xs = sorted(x[k] for x in T)
        print(f"  own chunk {chunks[k]}: seen -> walked  median {statistics.median(xs):6.0f}  p90 {xs[int(0.9 * len(xs))]:6d}  max {xs[-1]:6d}")
    for k in range(3):
        xs = sorted(x[k] for x in gaps)
        print(f"  idle before own chunk {chunks[k + 1]}: median {statistics.median(xs):6.0f}  p10 {xs[int(0.1 * len(xs))]:6d}  min {xs[0]:6d}")
