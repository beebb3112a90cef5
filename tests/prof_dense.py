"""ncu target: the weight-streaming dense layer alone (python tests/prof_dense.py [batch])."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bevfusion_multimodal_3d_object_detection_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(2)
w2 = torch.randn((80000, 512), device=dev, generator=g) * 0.04
b2 = torch.zeros(80000, device=dev)
for B in [int(a) for a in sys.argv[1:]] or [32]:
    hid = torch.rand((B, 512), device=dev, generator=g)
    for _ in range(2):
        out = ops.dense_layer(hid, w2, b2)
    torch.cuda.synchronize()
    print(B, float(out.abs().max()))
