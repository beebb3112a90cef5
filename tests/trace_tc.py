"""Debug aid: dumps the clock timeline of CTA 0 of the tcgen05 MLP kernel (B200BEV_TC_TRACE)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bevfusion_multimodal_3d_object_detection_b200 import _lib, ops  # noqa: E402
from bevfusion_multimodal_3d_object_detection_b200 import synthetic as syn  # noqa: E402
from oracle import bev_oracle as orc  # noqa: E402

dev = torch.device("cuda:0")
lw, lb = orc.fold_layers(syn.mlp_weights(101, syn.LIDAR_DIMS))
blob, dims = ops.pack_mlp_params([torch.from_numpy(w) for w in lw], [torch.from_numpy(b) for b in lb], dev)
tc = ops.pack_mlp_params_bf16(blob, dims)
pts = torch.from_numpy(syn.lidar_batch(42, 32)).to(dev)
for _ in range(3):
    ops.pointnet_encode(pts, blob, dims, precision=_lib.BF16_TENSOR, tc_params=tc)
torch.cuda.synchronize()
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace.txt"
os.environ["B200BEV_TC_TRACE"] = out
ops.pointnet_encode(pts, blob, dims, precision=_lib.BF16_TENSOR, tc_params=tc)
torch.cuda.synchronize()
_, perm, off = ops.bin_sort(pts, 50, 50)
os.environ["B200BEV_TC_TRACE"] = out.replace(".txt", "_cell.txt")
ops.pointnet_encode(pts, blob, dims, perm=perm, offsets=off, n_cells=2500, precision=_lib.BF16_TENSOR, tc_params=tc)
torch.cuda.synchronize()
print("trace written")
