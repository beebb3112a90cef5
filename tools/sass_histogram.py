#!/usr/bin/env python
"""SASS opcode histogram of libb200bev.so per kernel file: the Blackwell-native instructions (B200_PROFILING.md:
`tcgen05.mma` -> UTC*MMA, `tcgen05.ld/st` -> LDTM/STTM, bulk copies -> UBLKCP, tcgen05.commit -> UTCBAR, ...) counted in
the object code that ships.  Runs without a GPU (cuobjdump on the in-tree objects).

    python tools/sass_histogram.py [> profiles/rNN_sass_histogram.txt]
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
NATIVE = ROOT / "bevfusion_multimodal_3d_object_detection_b200" / "_native"
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "ELECT",
         "HMMA", "HGMMA", "FADD2", "FFMA2", "F2FP", "LDGSTS", "RED", "ATOMG", "ATOM", "MATCH", "SHFL", "FFMA", "HFMA2", "FMNMX", "BAR", "UCGABAR_ARV", "UCGABAR_WAIT"]


def histogram(obj: Path):
    out = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
    per_kernel, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per_kernel.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    return per_kernel


def demangle(name: str) -> str:
    r = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    r = re.sub(r"b200bev::\(anonymous namespace\)::", "", r)
    return re.sub(r"\(.*", "", r)[:80]


def main():
    archs = subprocess.run(["cuobjdump", "--list-elf", str(NATIVE / "libb200bev.so")], capture_output=True, text=True).stdout
    print("libb200bev.so ELF images:", sorted(set(re.findall(r"sm_\d+a?", archs))))
    total = collections.Counter()
    for obj in sorted(NATIVE.glob("*.o")):
        per_kernel = histogram(obj)
        if not per_kernel:
            continue
        print(f"\n== {obj.stem}.cu")
        for k, c in per_kernel.items():
            watched = {w: c[w] for w in WATCH if c[w]}
            total.update(watched)
            print(f"  {demangle(k):80s} {sum(c.values()):6d} instr  " + "  ".join(f"{w}={n}" for w, n in watched.items()))
    print("\n== whole library (watched opcodes)")
    print("  " + "  ".join(f"{w}={n}" for w, n in total.items() if n))
    legacy = total["HMMA"] + total["HGMMA"]
    print(f"  tcgen05 MMA instructions: {total['UTCHMMA'] + total['UTCQMMA'] + total['UTCIMMA']}, legacy mma.sync/wgmma: {legacy}")


if __name__ == "__main__":
    sys.exit(main())
