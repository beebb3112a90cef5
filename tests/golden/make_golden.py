"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules (imported from
/root/reference/src) on seeded synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; tests read the committed .npz files.  Inputs and weights are not
stored — they are regenerated from the seed by bevfusion_multimodal_3d_object_detection_b200.synthetic
and checked against the sha256 stored here; the reference's outputs are stored.
"""
from __future__ import annotations

import contextlib
import io
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/src")

from bevfusion_multimodal_3d_object_detection_b200 import synthetic as syn  # noqa: E402
from oracle import bev_oracle as orc  # noqa: E402
from oracle import torch_port  # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    import centernet_target  # noqa: E402  (reference)
    import encoders  # noqa: E402          (reference)
    import fusion  # noqa: E402            (reference)
    import fusion_detection  # noqa: E402  (reference)

OUT = Path(__file__).resolve().parent
torch.set_grad_enabled(False)


def load_mlp(module, layers):
    """Writes synthetic layer dicts into a reference PointNetLiDAREncoder / RadarEncoder."""
    sd = module.state_dict()
    for i, lay in enumerate(layers, start=1):
        sd[f"conv{i}.weight"] = torch.from_numpy(lay["weight"]).unsqueeze(-1)
        sd[f"conv{i}.bias"] = torch.from_numpy(lay["bias"])
        sd[f"bn{i}.weight"] = torch.from_numpy(lay["bn_weight"])
        sd[f"bn{i}.bias"] = torch.from_numpy(lay["bn_bias"])
        sd[f"bn{i}.running_mean"] = torch.from_numpy(lay["bn_mean"])
        sd[f"bn{i}.running_var"] = torch.from_numpy(lay["bn_var"])
    module.load_state_dict(sd)
    module.eval()


def lidar_cases():
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    enc = encoders.PointNetLiDAREncoder(input_channels=4, feat_dim=1024, use_bn=True)
    load_mlp(enc, layers)
    out = {"weights_digest": syn.digest(*[v for lay in layers for v in lay.values()])}
    # small batch, ragged tail (N not a multiple of the tile sizes)
    pts = syn.lidar_batch(201, 2, n_valid=1900, n_total=2011)
    out["small_digest"] = syn.digest(pts)
    out["small_global"] = enc(torch.from_numpy(pts)).numpy()
    # (B,C,N) input layout is accepted too (src/encoders.py:282)
    assert torch.equal(enc(torch.from_numpy(pts).transpose(1, 2).contiguous()), torch.from_numpy(out["small_global"]))
    # full-size frame, configs[0]
    full = syn.lidar_batch(301, 1)
    out["full_digest"] = syn.digest(full)
    out["full_global"] = enc(torch.from_numpy(full)).numpy()
    # per-point features of the reference -> per-cell max (restatement of the unpinned scatter stage)
    enc.return_point_features = True
    feat = enc(torch.from_numpy(pts))[:, :, :1024]            # src/encoders.py:300-304
    enc.return_point_features = False
    W = H = 50
    cell = orc.cell_index(pts, syn.PC_RANGE, W, H)
    canv = []
    for b in range(pts.shape[0]):
        ok = torch.from_numpy(cell[b] >= 0)
        idx = torch.from_numpy(cell[b].astype(np.int64))[ok]
        c = torch.zeros(W * H, 1024)
        c.scatter_reduce_(0, idx[:, None].expand(-1, 1024), feat[b][ok], reduce="amax", include_self=False)
        canv.append(c.numpy())
    canv = np.stack(canv)
    np.testing.assert_allclose(
        canv, np.stack([orc.pointnet_cell_max(pts[b], layers, cell[b], W * H) for b in range(2)]), rtol=0, atol=2e-5 * canv.max())
    out["small_cell"] = cell
    out["small_canvas_sub"] = canv[:, :, ::16]                # 64 of 1024 channels keeps the fixture small
    np.savez_compressed(OUT / "lidar_encoder.npz", **out)


def radar_cases():
    layers = syn.mlp_weights(111, syn.RADAR_DIMS)
    fcw, fcb = syn.linear_weights(112, 5 * 256, 256)
    out = {"weights_digest": syn.digest(*[v for lay in layers for v in lay.values()], fcw, fcb)}
    for method in ("concat", "max", "mean"):
        with contextlib.redirect_stdout(io.StringIO()):
            enc = encoders.MultiRadarEncoder(input_channels=7, feat_dim=256, num_radars=5, fusion_method=method)
        load_mlp(enc.radar_encoder, layers)
        if method == "concat":
            enc.fusion_fc.weight.copy_(torch.from_numpy(fcw))
            enc.fusion_fc.bias.copy_(torch.from_numpy(fcb))
        enc.eval()
        radars = syn.radar_batch(211, 3)
        out["digest"] = syn.digest(*radars)
        out[f"fused_{method}"] = enc([torch.from_numpy(r) for r in radars]).numpy()
        ragged = [r[:, : 125 - 17 * i] for i, r in enumerate(radars)]
        out[f"ragged_{method}"] = enc([torch.from_numpy(np.ascontiguousarray(r)) for r in ragged]).numpy()
    np.savez_compressed(OUT / "radar_encoder.npz", **out)


def camera_cases():
    out = {}
    for name, (C, h, w, H, W) in {"ref28x50": (16, 28, 50, 50, 50), "hd57x100": (8, 57, 100, 50, 50),
                                  "up7x9": (8, 7, 9, 20, 30)}.items():
        fus = fusion.FlexibleBEVFusion(use_camera=True, use_lidar=False, use_radar=False, camera_channels=C,
                                       bev_h=H, bev_w=W, bev_channels=8)
        fus.eval()
        grabbed = {}
        fus.camera_proj.register_forward_hook(lambda m, i, o: grabbed.update(proj_in=i[0].clone(), proj_out=o.clone()))
        fus.bev_fusion.register_forward_hook(lambda m, i, o: grabbed.update(fuse_in=i[0].clone()))
        feats = syn.camera_features(401, 2, n_cam=6, channels=C, h=h, w=w)
        fus(camera_features=torch.from_numpy(feats))
        out[f"{name}_digest"] = syn.digest(feats)
        out[f"{name}_mean"] = grabbed["proj_in"].numpy()          # camera_features.mean(dim=1), src/fusion.py:234
        out[f"{name}_resize_in"] = grabbed["proj_out"].numpy()    # input of F.interpolate, src/fusion.py:242
        out[f"{name}_resize_out"] = grabbed["fuse_in"].numpy()    # its output (camera-only: the concat is just it)
    # geometric projection: cross-check the numpy restatement against torch's grid_sample
    K, E = syn.camera_rig()
    feats = syn.camera_features(402, 1, n_cam=6, channels=8, h=57, w=100)
    table = orc.project_cells(K, E, (1600.0, 900.0), (57, 100), (50, 50), syn.PC_RANGE)
    ours = orc.camera_project(feats[0], table, (50, 50))
    ref = torch_port.camera_project(torch.from_numpy(feats), torch.from_numpy(table), (50, 50))[0].numpy()
    np.testing.assert_allclose(ours, ref, rtol=0, atol=1e-5 * np.abs(ref).max())
    assert table[:, :, 2].sum() > 1000, "rig sees too few cells"
    out["project_digest"] = syn.digest(feats, K, E)
    out["project_table"] = table
    out["project_canvas_grid_sample"] = ref
    np.savez_compressed(OUT / "camera_bev.npz", **out)


def decode_cases():
    out = {}
    maps = syn.head_maps(501, 3)
    pred = {k: torch.from_numpy(v) for k, v in maps.items()}
    out["digest"] = syn.digest(*maps.values())
    out["nms"] = centernet_target._nms(pred["heatmap"]).numpy()
    assert torch.equal(fusion_detection._nms(pred["heatmap"]), torch.from_numpy(out["nms"]))
    for name, val in zip(("score", "ind", "classes", "ys", "xs"), centernet_target._topk(torch.from_numpy(out["nms"]), K=100)):
        out[f"topk_{name}"] = val.numpy()
    for tag, mod in (("ct", centernet_target), ("fd", fusion_detection)):
        for thr in (0.0, 0.3, 0.999):
            dets = mod.decode_centernet_predictions(pred, score_thresh=thr, max_detections=100)
            for b, d in enumerate(dets):
                for k, v in d.items():
                    out[f"{tag}_thr{thr}_b{b}_{k}"] = v.numpy()
    # sparse map: fewer positive peaks than K in some classes
    sparse = syn.head_maps(502, 2, peak_frac=0.02)
    sp = {k: torch.from_numpy(v) for k, v in sparse.items()}
    out["sparse_digest"] = syn.digest(*sparse.values())
    for b, d in enumerate(centernet_target.decode_centernet_predictions(sp, score_thresh=0.1, max_detections=100)):
        for k, v in d.items():
            out[f"sparse_b{b}_{k}"] = v.numpy()
    # 100x100 grid (stress config), K=100
    big = syn.head_maps(503, 1, H=100, W=100)
    bg = {k: torch.from_numpy(v) for k, v in big.items()}
    out["big_digest"] = syn.digest(*big.values())
    for k, v in fusion_detection.decode_centernet_predictions(bg, score_thresh=0.0, max_detections=100)[0].items():
        out[f"big_{k}"] = v.numpy()
    # hand-made known-answer maps: corner/edge peaks and a plateau
    hm = np.zeros((1, 2, 6, 7), dtype=np.float32)
    hm[0, 0, 0, 0], hm[0, 0, 0, 6], hm[0, 0, 5, 0], hm[0, 0, 5, 6], hm[0, 0, 2, 3] = 0.9, 0.8, 0.7, 0.6, 0.5
    hm[0, 1, 3:5, 2:4] = 0.4                                   # 2x2 plateau: every cell survives _nms
    hm[0, 1, 0, 3] = 0.95
    out["hand_heat"] = hm
    out["hand_nms"] = centernet_target._nms(torch.from_numpy(hm)).numpy()
    for name, val in zip(("score", "ind", "classes", "ys", "xs"), centernet_target._topk(torch.from_numpy(out["hand_nms"]), K=6)):
        out[f"hand_topk_{name}"] = val.numpy()
    try:
        centernet_target._topk(torch.from_numpy(hm), K=43)
        raise AssertionError("expected torch.topk to reject K > H*W")
    except RuntimeError as e:
        out["k_too_large_message"] = np.array(str(e).splitlines()[0])
    np.savez_compressed(OUT / "centernet_decode.npz", **out)


def lidar_prepare_cases():
    """N3: the reference's own NuScenesDataset._load_lidar_points (src/train_detect.py:147-189) on sweeps written
    to .bin files, pad branch and subsample branch.  The subsample draw is np.random.choice without a seed in the
    reference; here numpy's global generator is seeded before each call and the same draw is repeated to store
    the indices next to the output."""
    import tempfile

    import train_detect  # noqa: E402  (reference)

    ds = object.__new__(train_detect.NuScenesDataset)      # no pkl / images needed for this method
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (rows, max_points) in {"pad": (3000, 2600), "pad_exact_empty": (64, 80), "subsample": (5000, 2048)}.items():
            raw = syn.raw_sweep(601 + rows, rows)
            if name == "pad_exact_empty":
                raw[:, 0] = 60.0                            # nothing in range: all zero rows
            path = f"{tmp}/{name}.bin"
            raw.tofile(path)
            ds.max_points = max_points
            np.random.seed(1234)
            got = ds._load_lidar_points({"lidar_path": path}).numpy()
            ref, n_in = orc.lidar_prepare(raw, max_points, syn.PC_RANGE, None)
            out[f"{name}_digest"] = syn.digest(raw)
            out[f"{name}_count"] = np.int32(n_in)
            if n_in >= max_points:
                np.random.seed(1234)
                idx = np.random.choice(n_in, max_points, replace=False)
                ref, _ = orc.lidar_prepare(raw, max_points, syn.PC_RANGE, idx)
                out[f"{name}_indices"] = idx.astype(np.int32)
            assert got.shape == (max_points, 4) and np.array_equal(got, ref), name
            out[f"{name}_out"] = got
    np.savez_compressed(OUT / "lidar_prepare.npz", **out)


def glue_cases():
    """the dense layers of FlexibleBEVFusion (lidar_init, radar_proj) and CenterNetHead's sigmoid in
    front of the decode, all produced by the reference's own modules."""
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        fus = fusion.FlexibleBEVFusion(use_camera=False, use_lidar=True, use_radar=True, lidar_channels=1024,
                                       radar_channels=256, bev_h=50, bev_w=50, bev_channels=256)
    w1, b1 = syn.linear_weights(701, 1024, 512)
    w2, b2 = syn.linear_weights(702, 512, 128 * 25 * 25)
    wr, br = syn.linear_weights(703, 256, 256)
    for lin, (w, b) in ((fus.lidar_init[0], (w1, b1)), (fus.lidar_init[2], (w2, b2)), (fus.radar_proj[0], (wr, br))):
        lin.weight.copy_(torch.from_numpy(w))
        lin.bias.copy_(torch.from_numpy(b))
    fus.eval()
    grabbed = {}
    fus.lidar_init[1].register_forward_hook(lambda m, i, o: grabbed.update(hidden=o.clone()))
    fus.lidar_init.register_forward_hook(lambda m, i, o: grabbed.update(lidar_init=o.clone()))
    fus.radar_proj.register_forward_hook(lambda m, i, o: grabbed.update(radar_proj=o.clone()))
    feats, radar = syn.global_features(704, 3, 1024), syn.global_features(705, 3, 256)
    fus(lidar_features=torch.from_numpy(feats), radar_features=torch.from_numpy(radar))
    out["dense_digest"] = syn.digest(w1, b1, w2, b2, wr, br, feats, radar)
    out["lidar_hidden"] = grabbed["hidden"].numpy()
    full = grabbed["lidar_init"].numpy()
    np.testing.assert_allclose(orc.lidar_init(feats, w1, b1, w2, b2), full, rtol=0, atol=1e-5 * np.abs(full).max())
    out["lidar_init_sub"] = full[:, ::16]                      # 5000 of 80000 columns keeps the fixture small
    out["lidar_init_absmax"] = np.float32(np.abs(full).max())
    out["radar_proj"] = grabbed["radar_proj"].numpy()

    # CenterNetHead.forward (src/fusion.py:869-884) -> decode: the head's raw heat-map output next to what the
    # reference decodes from its sigmoid
    hw = syn.head_weights(711, 32, 16, 10)
    with contextlib.redirect_stdout(io.StringIO()):
        head = fusion.CenterNetHead(in_channels=32, num_classes=10, head_conv=16)
    head.load_state_dict({k: torch.from_numpy(v) for k, v in hw.items()})
    head.eval()
    head.heatmap_head.register_forward_hook(lambda m, i, o: grabbed.update(logits=o.clone()))
    x = syn._rng(712).standard_normal((2, 32, 24, 40)).astype(np.float32)
    pred = head(torch.from_numpy(x))
    out["head_digest"] = syn.digest(x, *hw.values())
    out["head_logits"] = grabbed["logits"].numpy()
    assert torch.equal(torch.sigmoid(grabbed["logits"]), pred["heatmap"])
    for k in ("heatmap", "offset", "size", "rot", "vel"):
        out[f"head_{k}"] = pred[k].numpy()
    nms = centernet_target._nms(pred["heatmap"])
    top = centernet_target._topk(nms, K=60)[0]
    assert (top[:, :-1] > top[:, 1:]).all(), "ties among the winners: pick another seed"
    thr = float(top[0, 30])                                   # about half of the winners of sample 0 pass
    out["head_thresh"] = np.float32(thr)
    for tag, t in (("all", 0.0), ("mid", thr)):
        for b, d in enumerate(fusion_detection.decode_centernet_predictions(pred, score_thresh=t, max_detections=60)):
            for k, v in d.items():
                out[f"head_{tag}_b{b}_{k}"] = v.numpy()
    # the convolution stacks of the fusion module with channel counts the tcgen05 kernel takes (multiples of 64):
    # camera_proj (3x3 + 1x1), F.interpolate, bev_fusion (3x3, 3x3), src/fusion.py:229-248,292-295
    with contextlib.redirect_stdout(io.StringIO()):
        cf = fusion.FlexibleBEVFusion(use_camera=True, use_lidar=False, use_radar=False, camera_channels=64,
                                      bev_h=12, bev_w=20, bev_channels=64)
    sd = syn.fill_state_dict(731, {k: tuple(v.shape) for k, v in cf.state_dict().items()})
    cf.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    cf.eval()
    cam = syn.camera_features(732, 2, n_cam=6, channels=64, h=9, w=14)
    out["stack_digest"] = syn.digest(cam, *[sd[k] for k in sorted(sd)])
    out["stack_out"] = cf(camera_features=torch.from_numpy(cam)).numpy()
    np.savez_compressed(OUT / "bev_glue.npz", **out)


def metrics_cases():
    """N4: the reference's compute_metrics (src/utils_v2.py:94-205) on the reference's own decode outputs."""
    import utils_v2  # noqa: E402  (reference)

    maps = syn.head_maps(501, 3)
    dets = fusion_detection.decode_centernet_predictions({k: torch.from_numpy(v) for k, v in maps.items()},
                                                        score_thresh=0.3, max_detections=100)
    gts = syn.ground_truth_near(801, [{k: v.numpy() for k, v in d.items()} for d in dets])
    m_t = utils_v2.compute_metrics(dets, gts)                                             # torch tensors, as the shim returns
    m_n = utils_v2.compute_metrics([{k: v.numpy() for k, v in d.items()} for d in dets], gts)
    assert m_t == m_n and m_t["mAP"] > 0
    out = {"mAP": np.float64(m_t["mAP"]), "NDS": np.float64(m_t["NDS"]),
           "AP_per_class": np.array([m_t["AP_per_class"][c] for c in orc.CLASS_NAMES], dtype=np.float64),
           "gt_digest": np.array(syn.digest(*[a for g in gts for a in g.values()]))}
    np.savez_compressed(OUT / "metrics.npz", **out)


def chain_cases():
    """The reference's whole inference pass behind the camera backbone — PointNetLiDAREncoder, MultiRadarEncoder,
    FlexibleBEVFusion, CenterNetHead as FlexibleMultiModal3DDetector.forward chains them (src/fusion.py:1113-1137) — and
    eval.py's decode (src/eval.py:58-62), at the base.yaml sizes, with one seeded state_dict under the reference's names.
    This is what the drop-in route must reproduce end to end and what bench.py's step runs."""
    with contextlib.redirect_stdout(io.StringIO()):
        mods = {"lidar_encoder.": encoders.PointNetLiDAREncoder(input_channels=4, feat_dim=1024, use_bn=True),
                "radar_encoder.": encoders.MultiRadarEncoder(input_channels=7, feat_dim=256, num_radars=5, fusion_method="concat"),
                "fusion.": fusion.FlexibleBEVFusion(use_camera=True, use_lidar=True, use_radar=True, camera_channels=512,
                                                    lidar_channels=1024, radar_channels=256, bev_h=50, bev_w=50, bev_channels=256),
                "det_head.": fusion.CenterNetHead(in_channels=256, num_classes=10, head_conv=64)}
    shapes = {pre + k: tuple(v.shape) for pre, m in mods.items() for k, v in m.state_dict().items()}
    sd = syn.detector_state(syn.CHAIN_SEED, shapes)
    for pre, m in mods.items():
        m.load_state_dict({k[len(pre):]: torch.from_numpy(v) for k, v in sd.items() if k.startswith(pre)}, strict=False)
        m.eval()
    lidar, radars, cam = syn.chain_inputs()
    out = {"state_digest": syn.digest(*[sd[k] for k in sorted(sd)]), "input_digest": syn.digest(lidar, *radars, cam),
           "shape_names": np.array(sorted(shapes))}
    for k in sorted(shapes):
        out["shape__" + k] = np.array(shapes[k], dtype=np.int64)
    lf = mods["lidar_encoder."](torch.from_numpy(lidar))
    rf = mods["radar_encoder."]([torch.from_numpy(r) for r in radars])
    bev = mods["fusion."](camera_features=torch.from_numpy(cam), lidar_features=lf, radar_features=rf)
    pred = mods["det_head."](bev)
    out["lidar_feat"], out["radar_feat"] = lf.numpy(), rf.numpy()
    out["bev_sub"] = bev[:, ::8].numpy()
    out["bev_absmax"] = np.float32(bev.abs().max())
    for k, v in pred.items():
        out["pred_" + k] = v.numpy()
    top = fusion_detection._topk(fusion_detection._nms(pred["heatmap"]), K=100)[0]
    assert (top[:, :-1] > top[:, 1:]).all(), "ties among the winners: pick another seed"
    dets = fusion_detection.decode_centernet_predictions(pred, score_thresh=0.0, max_detections=100)
    for b, d in enumerate(dets):
        for k, v in d.items():
            out[f"det_b{b}_{k}"] = v.numpy()
    # the torch port bench.py's reference arm times must be this, op for op
    tsd = {k: torch.from_numpy(v) for k, v in sd.items()}
    pbev, ppred, pdets = torch_port.detector_chain(tsd, torch.from_numpy(cam), torch.from_numpy(lidar),
                                                   [torch.from_numpy(r) for r in radars])
    assert torch.allclose(pbev, bev, rtol=0, atol=1e-6 * float(bev.abs().max())), "port differs from the reference"
    np.savez_compressed(OUT / "detector_chain.npz", **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    only = set(sys.argv[1:])
    for fn in (lidar_cases, radar_cases, camera_cases, decode_cases, lidar_prepare_cases, glue_cases, metrics_cases, chain_cases):
        if only and fn.__name__ not in only:
            continue
        fn()
        print("wrote", fn.__name__)
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)
