#!/usr/bin/env python
"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_shares.py profiles/r01_launches_bench_c.csv [--ours]
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    ours_only = "--ours" in sys.argv
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    head = rows[0]
    k_name, k_val, k_unit = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Unit")
    tot = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[k_name]).replace("b200bev::<unnamed>::", "").replace("void ", "")
        us = float(r[k_val].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[k_unit], 1.0)
        if ours_only and "b200bev" not in r[k_name]:
            continue
        c = tot.setdefault(name, [0, 0.0])
        c[0] += 1
        c[1] += us
    total = sum(v[1] for v in tot.values())
    print(f"{'kernel':60s} {'launches':>8s} {'total us':>10s} {'avg us':>9s} {'share':>7s}")
    for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:60]:60s} {n:8d} {us:10.1f} {us / n:9.1f} {us / total:7.1%}")


if __name__ == "__main__":
    main()
