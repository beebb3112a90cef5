"""One eager pass of the bench step (bf16 and f32) for an `ncu --metrics gpu__time_duration.sum` launch list."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

dev = torch.device("cuda:0")
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 32
# the fp32-accuracy MLP alone, global-only and cell + global (same points)
from bevfusion_multimodal_3d_object_detection_b200 import _lib, ops  # noqa: E402
hp = bench.HotPath("lidar_only", frames, "f32", dev, 42)
prec, blob, dims, tc = hp.lidar_params()
lidar = hp.inputs["lidar"]
_, perm, off = ops.bin_sort(lidar, 50, 50)
for _ in range(2):
    ops.pointnet_encode(lidar, blob, dims, precision=prec, tc_params=tc)
    ops.pointnet_encode(lidar, blob, dims, perm=perm, offsets=off, n_cells=2500, precision=prec, tc_params=tc)
torch.cuda.synchronize()
print("done mlp-only", flush=True)
del hp
for precision in ("bf16", "f32"):
    hp = bench.HotPath("step", frames, precision, dev, 42)
    for _ in range(2):
        hp.step(hp.inputs)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push(f"step-{precision}")
    hp.step(hp.inputs)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    print("done", precision, flush=True)
    del hp
    torch.cuda.empty_cache()
