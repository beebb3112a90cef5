"""Kernel micro-timings (CUDA events, L2 flushed between runs) — a development aid, not the bench.

    python tests/perf_kernels.py [mlp|binsort|camera|decode|all] [--frames 32]
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
from oracle import bev_oracle as orc  # noqa: E402

dev = torch.device("cuda:0")
flush = None


def timeit(fn, reps=10, warm=3):
    global flush
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        # write a buffer larger than L2, then read it back: the write alone leaves L2 full of DIRTY lines, and a short
        # kernel that streams > 100 MB then pays for their write-back (lidar_init: 55 us instead of 30)
        flush.zero_()
        flush.view(torch.int64).sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--grid", type=int, default=50)
    ap.add_argument("--points", type=int, default=35000, help="points per frame (300000 = the 10-sweep stress config)")
    args = ap.parse_args()
    NP = args.points
    F = args.frames
    G = args.grid
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    if args.what in ("mlp", "all"):
        lw, lb = orc.fold_layers(syn.mlp_weights(101, syn.LIDAR_DIMS))
        blob, dims = ops.pack_mlp_params([torch.from_numpy(w) for w in lw], [torch.from_numpy(b) for b in lb], dev)
        tc = ops.pack_mlp_params_bf16(blob, dims)
        pts = to(syn.lidar_batch(42, F, n_valid=NP - NP // 125, n_total=NP))
        _, perm, off = ops.bin_sort(pts, G, G)
        flops = F * NP * 2.0 * 696576
        for name, fn in (
            ("mlp f32 global", lambda: ops.pointnet_encode(pts, blob, dims)),
            ("mlp f32 canvas+global", lambda: ops.pointnet_encode(pts, blob, dims, perm=perm, offsets=off, n_cells=G * G)),
            ("mlp bf16 tcgen05 global", lambda: ops.pointnet_encode(pts, blob, dims, precision=_lib.BF16_TENSOR, tc_params=tc)),
            ("mlp bf16 tcgen05 canvas+global", lambda: ops.pointnet_encode(pts, blob, dims, perm=perm, offsets=off, n_cells=G * G,
                                                                          precision=_lib.BF16_TENSOR, tc_params=tc)),
        ):
            try:
                med, best = timeit(fn, reps=5 if "f32" in name else 20)
                print(f"{name:34s} median {med:8.3f} ms  best {best:8.3f} ms  {flops / med / 1e9:8.1f} TFLOP/s", flush=True)
            except _lib.B200BevError as e:
                print(f"{name:34s} {e}", flush=True)
    if args.what in ("binsort", "all"):
        pts = to(syn.lidar_batch(42, F, n_valid=NP - NP // 125, n_total=NP))
        med, best = timeit(lambda: ops.bin_sort(pts, G, G), reps=20)
        by = F * (24.0 * NP + 4 * (G * G + 1))
        print(f"bin_sort {F}x{NP} {G}x{G}            median {med * 1e3:8.1f} us  best {best * 1e3:8.1f} us  {by / med / 1e6:8.1f} GB/s", flush=True)
    if args.what in ("prepare", "all"):
        rows = 43000                                            # ~35k in range after the filter
        sweeps = [syn.raw_sweep(900 + i, rows) for i in range(F)]
        raw = to(np.concatenate(sweeps, axis=0))
        off = torch.tensor([rows * i for i in range(F + 1)], dtype=torch.int64, device=dev)
        med, best = timeit(lambda: ops.lidar_prepare(raw, off, 35000, syn.PC_RANGE, max_frame_rows=rows), reps=20)
        by = F * 16.0 * (rows + 35000)
        print(f"lidar_prepare {F}x{rows} -> 35000     median {med * 1e3:8.1f} us  best {best * 1e3:8.1f} us  {by / med / 1e6:8.1f} GB/s", flush=True)
    if args.what in ("radar", "all"):
        rw, rb = orc.fold_layers(syn.mlp_weights(111, syn.RADAR_DIMS))
        rblob, rdims = ops.pack_mlp_params([torch.from_numpy(w) for w in rw], [torch.from_numpy(b) for b in rb], dev)
        fcw, fcb = (to(a) for a in syn.linear_weights(112, 1280, 256))
        radars = [to(r) for r in syn.radar_batch(43, F)]
        med, best = timeit(lambda: ops.radar_encode(radars, rblob, rdims, "concat", fcw, fcb), reps=20)
        fl = F * 625 * 2.0 * (7 * 32 + 32 * 64 + 64 * 128 + 128 * 256) + F * 2.0 * 1280 * 256
        print(f"radar_encode {F}x5x125               median {med * 1e3:8.1f} us  best {best * 1e3:8.1f} us  {fl / med / 1e9:8.2f} TFLOP/s", flush=True)
    if args.what in ("camera", "all"):
        g = torch.Generator(device=dev).manual_seed(1)
        feats = torch.relu(torch.randn((F, 6, 512, 57, 100), device=dev, generator=g))
        K, E = syn.camera_rig()
        Kd, Ed = to(K), to(E)
        mean = ops.camera_mean(feats)
        x = mean.view(F * 2, 256, 57, 100)[:F]
        table = orc.project_cells(K, E, (1600.0, 900.0), (57, 100), (G, G), syn.PC_RANGE)
        hits = int(table[:, :, 2].sum())
        for name, fn, by in (
            ("camera_mean", lambda: ops.camera_mean(feats), F * 4.0 * 512 * 5700 * 7),
            ("bilinear_resize", lambda: ops.bilinear_resize(x, (G, G)), F * 4.0 * 256 * (5700 + G * G)),
            ("camera_project", lambda: ops.camera_project(feats, Kd, Ed, (1600.0, 900.0), (G, G)),
             F * 4.0 * 512 * (min(6 * 5700, 4 * hits) + G * G)),
        ):
            med, best = timeit(fn, reps=20)
            print(f"{name:34s} median {med * 1e3:8.1f} us  best {best * 1e3:8.1f} us  {by / med / 1e6:8.1f} GB/s", flush=True)
    if args.what in ("dense", "all"):
        g = torch.Generator(device=dev).manual_seed(2)
        w1 = torch.randn((512, 1024), device=dev, generator=g) * 0.03
        w2 = torch.randn((80000, 512), device=dev, generator=g) * 0.04
        b1, b2 = torch.zeros(512, device=dev), torch.zeros(80000, device=dev)
        for Bd in sorted({1, 8, 16, F}):
            x = torch.rand((Bd, 1024), device=dev, generator=g)
            hid = ops.dense_layer(x, w1, b1, relu=True)
            med, best = timeit(lambda: ops.lidar_init(x, w1, b1, w2, b2), reps=20)
            med2, best2 = timeit(lambda: ops.dense_layer(hid, w2, b2), reps=20)
            by = 4.0 * (80000 * 512 + 80000 + Bd * (512 + 80000))
            print(f"lidar_init batch {Bd:3d} (both layers) median {med * 1e3:8.1f} us  best {best * 1e3:8.1f} us | 512->80000 alone "
                  f"median {med2 * 1e3:8.1f} us  best {best2 * 1e3:8.1f} us  {by / med2 / 1e6:8.1f} GB/s  "
                  f"{2.0 * Bd * 512 * 80000 / med2 / 1e9:6.2f} TFLOP/s", flush=True)
        img = ops.dense_pack_split(w2, b2)
        for Bd in sorted({1, 8, 16, F, 64}):
            hid = torch.rand((Bd, 512), device=dev, generator=g)
            med2, best2 = timeit(lambda: ops.dense_layer_split(hid, img, 80000), reps=20)
            by = 4.0 * (80000 * 512 + 2 * 80000 + Bd * (512 + 80000))
            print(f"dense_layer_split batch {Bd:3d} 512->80000 (tensor cores, fp32 accuracy) median {med2 * 1e3:8.1f} us  best "
                  f"{best2 * 1e3:8.1f} us  {by / med2 / 1e6:8.1f} GB/s", flush=True)
        ref = x @ w1.t()
        torch.backends.cuda.matmul.allow_tf32 = False
        med3, _ = timeit(lambda: torch.addmm(b2, torch.relu(torch.addmm(b1, x, w1.t())), w2.t()), reps=20)
        print(f"  (cuBLAS fp32, torch.addmm x2, batch {F}: median {med3 * 1e3:8.1f} us)", flush=True)
    if args.what in ("conv", "all"):
        g = torch.Generator(device=dev).manual_seed(3)
        shapes = [("head 5x(256->64) as 256->320", 256, 320, 3, G, G), ("bev_fusion.0 768->512", 768, 512, 3, G, G),
                  ("bev_fusion.3 512->256", 512, 256, 3, G, G), ("camera_proj.0 512->512 @57x100", 512, 512, 3, 57, 100),
                  ("camera_proj.3 512->256 1x1 @57x100", 512, 256, 1, 57, 100), ("radar_refine 256->256", 256, 256, 3, G, G),
                  ("lidar_upsample.4 128->256", 128, 256, 3, G, G)]
        tot = {"tc": 0.0, "cudnn_bf16": 0.0, "cudnn_f32": 0.0, "flop": 0.0}
        for name, cin, cout, k, H, W in shapes:
            x = torch.randn((F, cin, H, W), device=dev, generator=g)
            w = torch.randn((cout, cin, k, k), device=dev, generator=g) / (cin * k * k) ** 0.5
            b = torch.randn(cout, device=dev, generator=g)
            nhwc = ops.nchw_to_nhwc_bf16([x])
            img = ops.conv_pack(w)
            med, best = timeit(lambda: ops.conv_bn_relu_bf16(nhwc, img, b, cout, k * k), reps=10)
            medl, _ = timeit(lambda: ops.nchw_to_nhwc_bf16([x]), reps=10)
            nxt = torch.empty((F, H, W, cout), dtype=torch.bfloat16, device=dev)     # channels-last bf16 only: what the step runs
            medn, _ = timeit(lambda: ops.conv_bn_relu_bf16(nhwc, img, b, cout, k * k, out_nhwc=nxt, want_nchw=False), reps=10)
            print(f"conv {name:36s} channels-last bf16 output only {medn * 1e3:8.1f} us ({2.0 * F * H * W * cout * cin * k * k / medn / 1e9:7.1f} TFLOP/s)", flush=True)
            xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            wb = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            bb = b.to(torch.bfloat16)
            medc, _ = timeit(lambda: torch.relu_(torch.nn.functional.conv2d(xb, wb, bb, padding=k // 2)), reps=10)
            torch.backends.cudnn.allow_tf32 = False
            medf, _ = timeit(lambda: torch.relu_(torch.nn.functional.conv2d(x, w, b, padding=k // 2)), reps=5)
            fl = 2.0 * F * H * W * cout * cin * k * k
            tot["tc"] += med; tot["cudnn_bf16"] += medc; tot["cudnn_f32"] += medf; tot["flop"] += fl
            print(f"conv {name:36s} tcgen05 {med * 1e3:8.1f} us ({fl / med / 1e9:7.1f} TFLOP/s)  layout pass {medl * 1e3:6.1f} us | "
                  f"cuDNN bf16 NHWC {medc * 1e3:8.1f} us  cuDNN fp32 {medf * 1e3:8.1f} us", flush=True)
        print(f"conv total ({F} frames): tcgen05 {tot['tc']:.3f} ms ({tot['flop'] / tot['tc'] / 1e9:.1f} TFLOP/s)  cuDNN bf16 "
              f"{tot['cudnn_bf16']:.3f} ms  cuDNN fp32 {tot['cudnn_f32']:.3f} ms", flush=True)
    if args.what in ("decode", "all"):
        maps = {k: to(v) for k, v in syn.head_maps(44, F, 10, G, G).items()}
        med, best = timeit(lambda: ops.centernet_decode(maps["heatmap"], maps["offset"], maps["size"], maps["rot"], maps["vel"],
                                                        100, 2.048), reps=20)
        by = F * (4.0 * 10 * G * G + 3600 + 6800)
        print(f"centernet_decode {F}x10x{G}x{G}       median {med * 1e3:8.1f} us  best {best * 1e3:8.1f} us  {by / med / 1e6:8.1f} GB/s", flush=True)
        logits = torch.logit(maps["heatmap"].clamp(1e-6, 1 - 1e-6))
        med, best = timeit(lambda: ops.centernet_decode(logits, maps["offset"], maps["size"], maps["rot"], maps["vel"],
                                                        100, 2.048, heat_is_logit=True), reps=20)
        med2, _ = timeit(lambda: ops.centernet_decode(torch.sigmoid(logits), maps["offset"], maps["size"], maps["rot"], maps["vel"],
                                                      100, 2.048), reps=20)
        print(f"  from logits (sigmoid fused)       median {med * 1e3:8.1f} us  best {best * 1e3:8.1f} us   torch.sigmoid + decode: {med2 * 1e3:8.1f} us", flush=True)


if __name__ == "__main__":
    main()
