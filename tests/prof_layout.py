"""ncu target: the layout kernels of the bf16 path at the bench shapes (python tests/prof_layout.py)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bevfusion_multimodal_3d_object_detection_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(4)
feats = torch.relu(torch.randn((32, 6, 512, 57, 100), device=dev, generator=g))
x = torch.randn((32, 768, 50, 50), device=dev, generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(2):
    flush.zero_()
    a = ops.camera_mean_nhwc_bf16(feats)
    flush.zero_()
    b = ops.nchw_to_nhwc_bf16([x])
torch.cuda.synchronize()
for name, fn in (("camera_mean_nhwc_bf16", lambda: ops.camera_mean_nhwc_bf16(feats)), ("camera_mean", lambda: ops.camera_mean(feats)),
                 ("nchw_to_nhwc_bf16 768x50x50", lambda: ops.nchw_to_nhwc_bf16([x]))):
    ts = []
    for _ in range(5):
        flush.zero_()
        flush.view(torch.int64).sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{name:32s} median {sorted(ts)[2] * 1e3:8.1f} us")
print(float(a.float().abs().max()), float(b.float().abs().max()))
