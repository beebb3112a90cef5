#!/usr/bin/env python
"""Summarises an .ncu-rep (ncu --set full) into a small CSV kept under profiles/: one row per captured
launch with the counters the roofline uses.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.csv ["header comment"]
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head = rows[0]
    idx = [head.index(k) for k in KEEP if k in head]
    with open(out, "w", newline="") as f:
        if note:
            f.write(f'"# {note}"\n')
        w = csv.writer(f)
        for r in rows:
            if len(r) >= len(head):
                w.writerow([r[i].replace("b200bev::<", "").replace("b200bev::", "") for i in idx])
    print(f"{out}: {len(rows) - 2} launches")


if __name__ == "__main__":
    main()
