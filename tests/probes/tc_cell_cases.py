"""Debug probe: the bf16 tcgen05 MLP in cell mode on a list of shapes, each in its own process (a launch failure poisons the context)."""
import subprocess
import sys

CASE = r'''
import sys, numpy as np, torch
sys.path.insert(0, ".")
from bevfusion_multimodal_3d_object_detection_b200 import _lib, ops
from bevfusion_multimodal_3d_object_detection_b200 import synthetic as syn
from oracle import bev_oracle as orc
B, N, G = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0")
layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
ws, bs = orc.fold_layers(layers)
blob, dims = ops.pack_mlp_params([torch.from_numpy(w) for w in ws], [torch.from_numpy(b) for b in bs], dev)
tc = ops.pack_mlp_params_bf16(blob, dims)
p = syn.lidar_batch(50 + N, B, n_valid=max(N - 7, 1), n_total=N)
d = torch.from_numpy(p).to(dev)
_, perm, off = ops.bin_sort(d, G, G)
g, c = ops.pointnet_encode(d, blob, dims, perm=perm, offsets=off, n_cells=G * G, precision=_lib.BF16_TENSOR, tc_params=tc)
torch.cuda.synchronize()
ref = orc.pointnet_global(p, layers)
print("ok", B, N, G, float(np.abs(g.cpu().numpy() - ref).max() / np.abs(ref).max()))
'''
for B, N, G in [(1, 128, 50), (1, 100, 50), (2, 333, 50), (3, 1000, 50), (1, 35000, 50), (4, 35000, 50), (32, 35000, 50), (2, 5000, 100)]:
    r = subprocess.run([sys.executable, "-c", CASE, str(B), str(N), str(G)], capture_output=True, text=True, timeout=300)
    print((r.stdout.strip() or "FAILED " + str((B, N, G)) + " :: " + r.stderr.strip().splitlines()[-1][:200]), flush=True)
