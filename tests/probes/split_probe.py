"""Debug / timing probe of the fp32-accuracy tensor-core MLP (csrc/pointnet_mlp_split.cu): runs the C-ABI entry with its own
workspace, compares every layer's activations left in the workspace with a float64 numpy chain, then times the full sizes."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from bevfusion_multimodal_3d_object_detection_b200 import _lib, ops  # noqa: E402
from bevfusion_multimodal_3d_object_detection_b200 import synthetic as syn  # noqa: E402
from oracle import bev_oracle as orc  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.lib()
layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
ws64, bs64 = orc.fold_layers(layers)
blob, dims = ops.pack_mlp_params([torch.from_numpy(w) for w in ws64], [torch.from_numpy(b) for b in bs64], dev)
img = ops.pack_mlp_params_split(blob, dims)
torch.cuda.synchronize()
print("image bytes", img.numel(), flush=True)


def chain64(pts):
    x = pts.astype(np.float64)
    acts = []
    for w, b in zip(ws64, bs64):
        x = np.maximum(x @ w.T + b, 0.0)
        acts.append(x)
    return acts


def run(B, N, cell=False, W=50):
    pts = syn.lidar_batch(900 + N, B, n_valid=max(N - N // 50 - 1, 1), n_total=N)
    d = torch.from_numpy(pts).to(dev)
    Npad = (N + 255) // 256 * 256
    nbytes = lib.b200bev_pointnet_split_workspace_bytes(B, N)
    wsb = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    out = torch.full((B, 1024), -1.0, device=dev)
    perm = off = canvas = None
    if cell:
        _, perm, off = ops.bin_sort(d, W, W)
        canvas = torch.empty((B, W * W, 1024), device=dev)
    dd = (C.c_int32 * 6)(*dims)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
    rc = lib.b200bev_pointnet_encode_split(p(d), B, N, 4, dd, 5, p(perm), p(off), W * W if cell else 0, p(img), p(out), p(canvas),
                                           p(wsb), nbytes, st)
    torch.cuda.synchronize()
    print(f"B={B} N={N} cell={cell} rc={rc}", flush=True)
    acts = chain64(pts.reshape(-1, 4)) if B * N <= 20000 else None
    stats = wsb[:16].view(torch.float32).cpu().numpy()
    print("  layer maxima (device):", stats, flush=True)
    if acts is not None and not cell:
        print("  layer maxima (float64):", [float(a.max()) for a in acts[:4]])
        rows = B * Npad
        cid_off = 256
        a_off = (cid_off + rows * 4 + 1023) // 1024 * 1024
        b_off = a_off + rows * 256 * 4
        bufA = wsb[a_off:a_off + rows * 256 * 4].view(torch.float32).view(rows, 256).cpu().numpy()
        bufB = wsb[b_off:b_off + rows * 512 * 4].view(torch.float32).view(rows, 512).cpu().numpy()
        for name, buf, act in (("act3", bufA, acts[2]), ("act4", bufB, acts[3])):
            got = np.concatenate([buf[b * Npad:b * Npad + N] for b in range(B)])
            err = np.abs(got - act).max() / act.max()
            print(f"  {name}: max_rel {err:.3e}  (nan: {np.isnan(got).any()})", flush=True)
    ref = orc.pointnet_global(pts, layers) if B * N <= 80000 else None
    if ref is not None:
        g = out.cpu().numpy()
        print(f"  global: max_rel {np.abs(g - ref).max() / np.abs(ref).max():.3e}", flush=True)
        if cell:
            rc_ = orc.cell_index(pts, syn.PC_RANGE, W, W)
            cref = orc.pointnet_cell_max(pts[0], layers, rc_[0], W * W)
            print(f"  canvas[0]: max_rel {np.abs(canvas[0].cpu().numpy() - cref).max() / np.abs(cref).max():.3e}", flush=True)
    return d, perm, off


for (B, N) in [(1, 256), (1, 200), (2, 1000)]:
    run(B, N)
run(2, 2011, cell=True)
run(1, 35000)
run(1, 35000, cell=True)

# timing at the bench shapes
pts = torch.from_numpy(syn.lidar_batch(42, 32, n_valid=34720, n_total=35000)).to(dev)
_, perm, off = ops.bin_sort(pts, 50, 50)
for label, kw in (("global", {}), ("cell+global", dict(perm=perm, offsets=off, n_cells=2500))):
    for _ in range(2):
        ops.pointnet_encode(pts, blob, dims, precision=_lib.F32, tc_params=img, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.pointnet_encode(pts, blob, dims, precision=_lib.F32, tc_params=img, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"32 x 35000 {label}: {ms:.3f} ms  ({32 * 35000 * 2 * 696576 / ms / 1e9:.1f} algorithmic TFLOP/s, x3 on the tensor pipe)", flush=True)
g_split = ops.pointnet_encode(pts, blob, dims, precision=_lib.F32, tc_params=img)
g_ffma = ops.pointnet_encode(pts, blob, dims)
print("32 x 35000 split vs FFMA: max_rel", float((g_split - g_ffma).abs().max() / g_ffma.abs().max()))
print("SPLIT-PROBE-DONE")
