"""Builds libb200bev.so (the C-ABI library of sm_100a kernels) in-tree with nvcc.

    python -m bevfusion_multimodal_3d_object_detection_b200.build [--force] [--verbose]

The library lands in ``bevfusion_multimodal_3d_object_detection_b200/_native/`` — git-ignored,
but it travels with a gpurun snapshot, so the GPU box never needs to compile.  nvcc cross-compiles
without a GPU.  No torch headers are involved: the ABI is plain C (include/b200bev.h).
"""
from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OUT_DIR = PKG / "_native"
LIB = OUT_DIR / "libb200bev.so"
STAMP = OUT_DIR / "libb200bev.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-I", str(ROOT / "include"), "-I", str(CSRC),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libb200bev.so cannot be built (set NVCC=/path/to/nvcc)")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _fingerprint(extra: str = "") -> str:
    h = hashlib.sha256()
    h.update(extra.encode())
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "b200bev.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    # flags without the checkout-dependent include paths: the .so built here must count as current
    # on the GPU box, where the repo lives under another directory
    h.update(" ".join(f for f in NVCC_FLAGS if not f.startswith(str(ROOT))).encode())
    return h.hexdigest()


def is_current(debug_env: bool = False) -> bool:
    return LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == _fingerprint("debug-env" if debug_env else "")


def build(force: bool = False, verbose: bool = False, debug_env: bool = False) -> Path:
    """Compile every csrc/*.cu for sm_100a and link libb200bev.so. Returns the library path.
    debug_env: -DB200BEV_DEBUG_ENV, the build in which the experiment / trace environment switches exist
    (csrc/common.cuh: debug_env); tools/tc_timeline.py and tests/trace_tc.py need it.  The release build has none."""
    if not force and is_current(debug_env):
        return LIB
    OUT_DIR.mkdir(exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in sources():
        obj = OUT_DIR / (src.stem + ".o")
        # B200BEV_NVCC_EXTRA: extra flags for experiment builds on the GPU box (-D switches of a kernel under study)
        cmd = [nvcc, *NVCC_FLAGS, *(["-DB200BEV_DEBUG_ENV"] if debug_env else []), *os.environ.get("B200BEV_NVCC_EXTRA", "").split(),
               "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src.name} ---\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"--- {src.name} ---\n{out}\n")
    if failed:
        raise RuntimeError("nvcc failed; see messages above")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs),
            "-cudart", "static", "-lcuda"]
    if verbose:
        print(" ".join(link), flush=True)
    subprocess.run(link, check=True)
    STAMP.write_text(_fingerprint("debug-env" if debug_env else "") + "\n")
    return LIB


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--debug-env", action="store_true", help="build with the experiment / trace environment switches")
    args = ap.parse_args()
    print(build(force=args.force, verbose=args.verbose, debug_env=args.debug_env))


if __name__ == "__main__":
    main()
