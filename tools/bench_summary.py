#!/usr/bin/env python
"""Prints the essentials of a bench.py JSON line (file argument or stdin)."""
import json
import sys

text = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = json.loads([l for l in text.strip().splitlines() if l.startswith("{")][-1])
print(f"value {d['value']:.1f} {d['unit']}  ms/step {d['ms_per_step']:.3f}  dtype {d['dtype']}  n_gpus {d['n_gpus']}  launches/step {d.get('launches_per_step')}")
print("roofline", {k: v for k, v in (d.get("roofline") or {}).items() if k not in ("note", "peak_source", "peak_kind")})
print("eager_stage_ms", d.get("eager_stage_ms"), "sum", d.get("eager_ms_per_step"))
for key in ("e2e", "e2e_features_on_device"):
    e = d.get(key)
    if e:
        print(key, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in e.items() if k != "api"})
for key in ("fp32_path", "fp32_cudnn_path", "bf16_path"):
    o = d.get(key)
    if o:
        print(key, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in o.items() if k not in ("note", "pointnet_encode", "pointnet_encode_global_only")})
        for k in ("pointnet_encode", "pointnet_encode_global_only"):
            if o.get(k):
                print("   ", k, {a: b for a, b in o[k].items() if a not in ("note", "peak_kind")})
print("cpu_baseline", d.get("cpu_baseline"))
print("clocks", d.get("clocks"), "fp32_fma_peak", d.get("fp32_fma_peak"))
for k, v in (d.get("kernels") or {}).items():
    if isinstance(v, dict):
        print("  ", k, {a: b for a, b in v.items() if a not in ("note", "peak_source", "peak_kind")})
for c in d.get("configs") or []:
    if "error" in c:
        print(c["name"], "ERROR", c["error"])
        continue
    e, ed = c.get("e2e"), c.get("e2e_features_on_device")
    print(f"{c['name']:12s} F/gpu {c['frames_per_gpu']:3d} value {c['value']:9.1f} ms {c['ms_per_step']:7.3f} "
          f"e2e {e['value'] if e else 0:8.1f} ({e['ms_per_step'] if e else 0:.2f} ms, h2d ceiling {e.get('h2d_ceiling_ms') or 0:.2f} ms) "
          f"e2e_dev {ed['value'] if ed else 0:8.1f}  stages {c.get('eager_stage_ms')}")
