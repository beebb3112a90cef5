"""Runs every hot-path stage of the bench workload a few times — the target of the ncu captures kept
under profiles/ (one process, one GPU; see B200_PROFILING.md).

    ncu --set full --clock-control none --import-source on -k regex:<kernel> --launch-skip 2 -c 1 \
        -o gpurun_out/<name> python tests/prof_stages.py [--frames 32] [--reps 3] [--only stage,...]
"""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from bevfusion_multimodal_3d_object_detection_b200 import _lib, ops  # noqa: E402
from bevfusion_multimodal_3d_object_detection_b200 import synthetic as syn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--only", default="")
    ap.add_argument("--grid", type=int, default=50)
    ap.add_argument("--points", type=int, default=35000)
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    want = lambda n: not only or n in only
    dev = torch.device("cuda:0")
    F, G = args.frames, args.grid
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def run(fn):
        for _ in range(args.reps):
            flush.zero_()
            fn()
        torch.cuda.synchronize()

    if want("bin_sort") or want("mlp_tc_cell") or want("mlp_tc_global") or want("mlp_f32"):
        lw, lb = syn.fold_mlp(syn.mlp_weights(101, syn.LIDAR_DIMS))
        blob, dims = ops.pack_mlp_params([torch.from_numpy(w) for w in lw], [torch.from_numpy(b) for b in lb], dev)
        tc = ops.pack_mlp_params_bf16(blob, dims)
        pts = to(syn.lidar_batch(42, F, n_valid=args.points - args.points // 125, n_total=args.points))
        _, perm, off = ops.bin_sort(pts, G, G)
        if want("bin_sort"):
            run(lambda: ops.bin_sort(pts, G, G))
        if want("mlp_tc_global"):
            run(lambda: ops.pointnet_encode(pts, blob, dims, precision=_lib.BF16_TENSOR, tc_params=tc))
        if want("mlp_tc_cell"):
            run(lambda: ops.pointnet_encode(pts, blob, dims, perm=perm, offsets=off, n_cells=G * G,
                                            precision=_lib.BF16_TENSOR, tc_params=tc))
        if want("mlp_f32"):
            run(lambda: ops.pointnet_encode(pts[:2], blob, dims, perm=perm[:2], offsets=off[:2], n_cells=G * G))
    if want("prepare"):
        rows = 43000
        raw = to(np.concatenate([syn.raw_sweep(900 + i, rows) for i in range(F)], axis=0))
        off = torch.tensor([rows * i for i in range(F + 1)], dtype=torch.int64, device=dev)
        run(lambda: ops.lidar_prepare(raw, off, 35000, syn.PC_RANGE, max_frame_rows=rows))
    if want("radar"):
        rw, rb = syn.fold_mlp(syn.mlp_weights(111, syn.RADAR_DIMS))
        rblob, rdims = ops.pack_mlp_params([torch.from_numpy(w) for w in rw], [torch.from_numpy(b) for b in rb], dev)
        fcw, fcb = (to(a) for a in syn.linear_weights(112, 1280, 256))
        radars = [to(r) for r in syn.radar_batch(43, F)]
        run(lambda: ops.radar_encode(radars, rblob, rdims, "concat", fcw, fcb))
    if want("camera_mean") or want("bilinear_resize") or want("camera_project"):
        g = torch.Generator(device=dev).manual_seed(1)
        feats = torch.relu(torch.randn((F, 6, 512, 57, 100), device=dev, generator=g))
        K, E = syn.camera_rig()
        Kd, Ed = to(K), to(E)
        mean = ops.camera_mean(feats)
        x = mean.view(F * 2, 256, 57, 100)[:F]
        if want("camera_mean"):
            run(lambda: ops.camera_mean(feats))
        if want("bilinear_resize"):
            run(lambda: ops.bilinear_resize(x, (G, G)))
        if want("camera_project"):
            run(lambda: ops.camera_project(feats, Kd, Ed, (1600.0, 900.0), (G, G)))
    if want("decode"):
        maps = {k: to(v) for k, v in syn.head_maps(44, F, 10, G, G).items()}
        run(lambda: ops.centernet_decode(maps["heatmap"], maps["offset"], maps["size"], maps["rot"], maps["vel"], 100, 2.048))
    if want("dense"):
        g = torch.Generator(device=dev).manual_seed(2)
        w1 = torch.randn((512, 1024), device=dev, generator=g) * 0.03
        w2 = torch.randn((80000, 512), device=dev, generator=g) * 0.04
        b1, b2 = torch.zeros(512, device=dev), torch.zeros(80000, device=dev)
        x = torch.rand((F, 1024), device=dev, generator=g)
        run(lambda: ops.lidar_init(x, w1, b1, w2, b2))
    if want("conv"):
        g = torch.Generator(device=dev).manual_seed(3)
        for cin, cout, k, H, W in ((768, 512, 3, G, G), (256, 320, 3, G, G), (512, 256, 1, 57, 100)):
            x = torch.randn((F, cin, H, W), device=dev, generator=g)
            w = torch.randn((cout, cin, k, k), device=dev, generator=g) / (cin * k * k) ** 0.5
            b = torch.randn(cout, device=dev, generator=g)
            img = ops.conv_pack(w)
            run(lambda: ops.conv_bn_relu_bf16(ops.nchw_to_nhwc_bf16([x]), img, b, cout, k * k))
    print("prof_stages done", flush=True)


if __name__ == "__main__":
    main()
