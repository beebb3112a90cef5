"""Driver for `compute-sanitizer` (race / memory checking of the kernels).

    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tests/sanitize_run.py [--tc] [--big]

One small invocation of every kernel of libb200bev.so through the C-ABI, on shapes that take each kernel's special
paths (ragged tails, cluster / DSMEM scans, TMA-staged rings, the overflow segment of the staged projection, ties in the
decode).  `--tc` adds the tcgen05 kernels (PointNet MLP global / cell / CTA pairs, the fp32-accuracy split MLP, the
convolution kernels).  Results are also checked loosely against the oracle, so a sanitizer run that "passes" on garbage
is not possible.  One tool per gpurun call (B200_PROFILING.md); logs go to profiles/.
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


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tc", action="store_true", help="include the tcgen05 kernels")
    ap.add_argument("--big", action="store_true", help="add one base.yaml-size frame per kernel")
    args = ap.parse_args()
    print("SANITIZE-RUN-OK:", "; ".join(run_all(args.tc, args.big)))


def run_all(tc: bool = False, big: bool = False):
    """Every kernel once, on shapes that take its special paths; returns the list of kernel groups that ran."""
    import types
    args = types.SimpleNamespace(tc=tc, big=big)
    dev = torch.device("cuda:0")
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    done = []

    # N3 lidar_prepare (cluster scan)
    raw = [syn.raw_sweep(21 + i, n) for i, n in enumerate((1500, 40, 0, 3000))]
    offs = torch.tensor([0] + list(np.cumsum([len(r) for r in raw])), dtype=torch.int64, device=dev)
    prep, cnt = ops.lidar_prepare(to(np.concatenate(raw)), offs, 1024, syn.PC_RANGE)
    for b, r in enumerate(raw):
        ref_p, n_in = orc.lidar_prepare(r, 1024, syn.PC_RANGE)
        assert int(cnt[b]) == n_in and np.array_equal(prep[b].cpu().numpy(), ref_p)
    done.append("lidar_prepare")

    # S1a bin_sort: ranked kernel (clusters, DSMEM scan), the 16-bit-rank path (100x100) and the fallback
    for (B, N, W) in [(2, 777, 50), (1, 9000, 100), (3, 31, 7)] + ([(1, 35000, 50)] if args.big else []):
        pts = syn.lidar_batch(11 + N, B, n_valid=max(N - 9, 1), n_total=N)
        cell, perm, off = ops.bin_sort(to(pts), W, W)
        ref_cell = orc.cell_index(pts, syn.PC_RANGE, W, W)
        assert np.array_equal(cell.cpu().numpy(), ref_cell)
        for b in range(B):
            rp, ro = orc.bin_sort(ref_cell[b], W * W)
            assert np.array_equal(perm[b].cpu().numpy(), rp) and np.array_equal(off[b].cpu().numpy(), ro)
    done.append("bin_sort")

    # S1b fp32 FFMA MLP (global + cell), S1c radar
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    ws, bs = orc.fold_layers(layers)
    blob, dims = ops.pack_mlp_params([torch.from_numpy(w) for w in ws], [torch.from_numpy(b) for b in bs], dev)
    pts = syn.lidar_batch(11, 2, n_valid=300, n_total=333)
    d = to(pts)
    cell, perm, off = ops.bin_sort(d, 50, 50)
    glob, canvas = ops.pointnet_encode(d, blob, dims, perm=perm, offsets=off, n_cells=2500, precision=_lib.F32, tc_params=None)
    assert rel(glob.cpu().numpy(), orc.pointnet_global(pts, layers)) < 1e-5
    assert rel(canvas[1].cpu().numpy(), orc.pointnet_cell_max(pts[1], layers, orc.cell_index(pts, syn.PC_RANGE, 50, 50)[1], 2500)) < 1e-5
    done.append("pointnet_mlp_f32")
    rl = syn.mlp_weights(111, syn.RADAR_DIMS)
    rw, rb = orc.fold_layers(rl)
    rblob, rdims = ops.pack_mlp_params([torch.from_numpy(w) for w in rw], [torch.from_numpy(b) for b in rb], dev)
    fcw, fcb = syn.linear_weights(112, 1280, 256)
    radars = syn.radar_batch(12, 2)
    for method in ("concat", "max", "mean"):
        fused, _ = ops.radar_encode([to(r) for r in radars], rblob, rdims, method, to(fcw), to(fcb))
        assert rel(fused.cpu().numpy(), orc.multi_radar(radars, rl, method, fcw, fcb)[0]) < 1e-5
    done.append("radar_encode")

    # S2: mean, staged resize, projection (staged bands / gather / overflow segments)
    feats = syn.camera_features(13, 2, channels=8, h=28, w=50)
    mean = ops.camera_mean(to(feats))
    assert np.array_equal(mean.cpu().numpy(), orc.camera_mean(feats))
    rs = ops.bilinear_resize(mean, (50, 50))
    assert rel(rs.cpu().numpy(), orc.bilinear_resize(orc.camera_mean(feats), (50, 50))) < 1e-5
    odd = syn.camera_features(14, 1, channels=3, h=7, w=9)
    ops.bilinear_resize(ops.camera_mean(to(odd)), (20, 30))
    K, E = syn.camera_rig()
    rt = orc.project_cells(K, E, (1600.0, 900.0), (28, 50), (50, 50), syn.PC_RANGE)
    for impl in ("staged", "gather"):
        cv = ops.camera_project(to(feats), to(K), to(E), (1600.0, 900.0), (50, 50), impl=impl)
        assert rel(cv[0].cpu().numpy(), orc.camera_project(feats[0], rt, (50, 50))) < 1e-5
    down = np.array([[0.0, -1.0, 0.0], [-1.0, 0.0, 0.0], [0.0, 0.0, -1.0]], dtype=np.float32)
    t = -(down @ np.array([0.0, 0.0, 100.0], dtype=np.float32))
    E6 = np.stack([np.concatenate([down, t[:, None]], axis=1)] * 6).astype(np.float32)     # every camera sees every cell
    ops.camera_project(to(feats), to(K), to(E6), (1600.0, 900.0), (50, 50), impl="staged")
    ops.camera_mean_nhwc_bf16(to(syn.camera_features(15, 1, channels=64, h=12, w=20)))
    done.append("camera_mean / bilinear_resize / camera_project")

    # S3 decode (+ logits), nms, topk, a tie-heavy map
    maps = syn.head_maps(14, 2)
    out = ops.centernet_decode(*[to(maps[k]) for k in ("heatmap", "offset", "size", "rot", "vel")], 100, 2.048)
    ref = orc.decode(maps, score_thresh=0.0)
    for b, r in enumerate(ref):
        n = len(r["scores"])
        assert int(out["count"][b]) == n and np.array_equal(out["scores"][b, :n].cpu().numpy(), r["scores"])
    flat = {k: v.copy() for k, v in maps.items()}
    flat["heatmap"][:] = 0.25                                    # one plateau: every cell survives the NMS, all scores tie
    ops.centernet_decode(*[to(flat[k]) for k in ("heatmap", "offset", "size", "rot", "vel")], 100, 2.048)
    ops.centernet_nms(to(maps["heatmap"]))
    ops.centernet_topk(to(maps["heatmap"]), 37)
    logit = np.log(maps["heatmap"] / (1.0 - maps["heatmap"])).astype(np.float32)
    ops.centernet_decode(to(logit), *[to(maps[k]) for k in ("offset", "size", "rot", "vel")], 100, 0.512, heat_is_logit=True)
    done.append("centernet nms / topk / decode")

    # N2 dense layers (row-stream, stream, rows kernels), layout
    for (B, K_, O) in [(5, 256, 3000), (32, 128, 1000), (3, 70, 33)]:
        w, b = syn.linear_weights(31 + B, K_, O)
        x = syn.global_features(33, B, K_)
        y = ops.dense_layer(to(x), to(w), to(b), relu=True)
        assert rel(y.cpu().numpy(), orc.dense_layer(x, w, b, relu=True)) < 1e-5
    w1, b1 = syn.linear_weights(41, 256, 128)
    w2, b2 = syn.linear_weights(42, 128, 3000)
    gf = syn.global_features(43, 5, 256)
    li = ops.lidar_init(to(gf), to(w1), to(b1), to(w2), to(b2))
    assert rel(li.cpu().numpy(), orc.lidar_init(gf, w1, b1, w2, b2)) < 1e-5
    xin = syn._rng(35).standard_normal((2, 64, 10, 14)).astype(np.float32)
    nhwc = ops.nchw_to_nhwc_bf16([to(xin)])
    ops.nchw_to_nhwc_bf16([to(xin[:, :, :5, :7].copy())])
    done.append("dense layers / layout")

    if args.tc:
        tc = ops.pack_mlp_params_bf16(blob, dims)
        for (B, N) in [(2, 333), (1, 128), (3, 1000)] + ([(1, 35000)] if args.big else []):
            p2 = syn.lidar_batch(50 + N, B, n_valid=max(N - 7, 1), n_total=N)
            d2 = to(p2)
            g16 = ops.pointnet_encode(d2, blob, dims, precision=_lib.BF16_TENSOR, tc_params=tc)
            assert rel(g16.cpu().numpy(), orc.pointnet_global(p2, layers)) < 1e-2
            _, perm2, off2 = ops.bin_sort(d2, 50, 50)
            g2, c2 = ops.pointnet_encode(d2, blob, dims, perm=perm2, offsets=off2, n_cells=2500, precision=_lib.BF16_TENSOR, tc_params=tc)
            assert rel(g2.cpu().numpy(), orc.pointnet_global(p2, layers)) < 1e-2
        done.append("pointnet_mlp_tc (global, cell, pairs)")
        split = ops.pack_mlp_params_split(blob, dims)
        if split is not None:
            for (B, N) in [(2, 333), (1, 700)]:
                p2 = syn.lidar_batch(60 + N, B, n_valid=max(N - 7, 1), n_total=N)
                d2 = to(p2)
                _, perm2, off2 = ops.bin_sort(d2, 50, 50)
                g3, c3 = ops.pointnet_encode(d2, blob, dims, perm=perm2, offsets=off2, n_cells=2500, precision=_lib.F32, tc_params=split)
                assert rel(g3.cpu().numpy(), orc.pointnet_global(p2, layers)) < 1e-5
            done.append("pointnet_mlp_split (fp32 accuracy on tcgen05)")
        sd = syn.fill_state_dict(34, {"w": (96, 64, 3, 3), "b": (96,), "w1": (70, 64, 1, 1)})
        cv16 = ops.conv_bn_relu_bf16(nhwc, ops.conv_pack(to(sd["w"])), to(sd["b"]), 96, 9, relu=True)
        assert rel(cv16.cpu().numpy(), orc.conv_bn_relu(xin, sd["w"], sd["b"])) < 1e-2
        ops.conv_bn_relu_bf16(nhwc, ops.conv_pack(to(sd["w1"])), None, 70, 1, relu=False)
        done.append("conv3x3_tc_halo / conv_tc_ws")
        for (B, K_, O) in [(5, 128, 384), (33, 64, 128)]:
            w, b = syn.linear_weights(51 + B, K_, O)
            x = syn.global_features(53, B, K_)
            y = ops.dense_layer_split(to(x), ops.dense_pack_split(to(w), to(b)), O, relu=True)
            assert rel(y.cpu().numpy(), orc.dense_layer(x, w, b, relu=True)) < 1e-5
        done.append("dense_split_tc (fp32 accuracy on tcgen05)")
    torch.cuda.synchronize()
    return done


if __name__ == "__main__":
    main()
