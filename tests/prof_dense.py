"""ncu target: the 164 MB dense layer alone, FFMA streaming kernel and tensor-core kernel (python tests/prof_dense.py [batch])."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bevfusion_multimodal_3d_object_detection_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(2)
w2 = torch.randn((80000, 512), device=dev, generator=g) * 0.04
b2 = torch.zeros(80000, device=dev)
img = ops.dense_pack_split(w2, b2)
for B in [int(a) for a in sys.argv[1:]] or [32]:
    hid = torch.rand((B, 512), device=dev, generator=g)
    for _ in range(2):
        out = ops.dense_layer(hid, w2, b2)
        tc = ops.dense_layer_split(hid, img, 80000)
    torch.cuda.synchronize()
    print(B, float(out.abs().max()), float((tc - out).abs().max()))
