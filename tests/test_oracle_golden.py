"""CPU: the oracle (numpy) and the torch-CPU port against the golden vectors produced by the
reference's own modules (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from bevfusion_multimodal_3d_object_detection_b200 import synthetic as syn
from oracle import bev_oracle as orc
from oracle import torch_port
from tests.conftest import max_rel

FP32_TOL = 1e-5  # of max|ref| — north_star's fp32 bound


def test_inputs_regenerate_bit_exact(golden):
    g = golden("lidar_encoder")
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    assert syn.digest(*[v for lay in layers for v in lay.values()]) == str(g["weights_digest"])
    assert syn.digest(syn.lidar_batch(201, 2, n_valid=1900, n_total=2011)) == str(g["small_digest"])
    assert syn.digest(syn.lidar_batch(301, 1)) == str(g["full_digest"])
    assert syn.digest(*syn.radar_batch(211, 3)) == str(golden("radar_encoder")["digest"])
    assert syn.digest(*syn.head_maps(501, 3).values()) == str(golden("centernet_decode")["digest"])


def test_lidar_points_respect_the_reference_filter():
    pts = syn.lidar_points(7)
    v = pts[:34720]
    assert np.all((v[:, 0] > -51.2) & (v[:, 0] < 51.2) & (v[:, 1] > -51.2) & (v[:, 1] < 51.2))
    assert np.all((v[:, 2] > -5.0) & (v[:, 2] < 3.0))            # src/train_detect.py:153-155
    assert not pts[34720:].any()                                  # zero padding, src/train_detect.py:188


def test_lidar_prepare_matches_the_reference_dataset(golden):
    """N3: oracle vs the reference's NuScenesDataset._load_lidar_points (pad branch and seeded subsample branch)."""
    g = golden("lidar_prepare")
    for name, (rows, max_points) in {"pad": (3000, 2600), "pad_exact_empty": (64, 80), "subsample": (5000, 2048)}.items():
        raw = syn.raw_sweep(601 + rows, rows)
        if name == "pad_exact_empty":
            raw[:, 0] = 60.0
        assert syn.digest(raw) == str(g[f"{name}_digest"])
        idx = g.get(f"{name}_indices")
        out, n = orc.lidar_prepare(raw, max_points, syn.PC_RANGE, idx)
        assert n == int(g[f"{name}_count"])
        np.testing.assert_array_equal(out, g[f"{name}_out"])
    # boundary and NaN rows never pass the strict filter
    edge = np.array([[-51.2, 0, 0, 1], [51.2, 0, 0, 1], [0, 51.2, 0, 1], [0, 0, -5.0, 1], [0, 0, 3.0, 1], [np.nan, 0, 0, 1],
                     [0, 0, 0, 7]], dtype=np.float32)
    out, n = orc.lidar_prepare(edge, 3, syn.PC_RANGE)
    assert n == 1 and out[0, 3] == 7 and not out[1:].any()


def test_lidar_global_max(golden):
    g = golden("lidar_encoder")
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    small = syn.lidar_batch(201, 2, n_valid=1900, n_total=2011)
    assert max_rel(orc.pointnet_global(small, layers), g["small_global"]) < FP32_TOL
    tl = torch_port.layers_to_torch(layers)
    assert max_rel(torch_port.shared_mlp_max(torch.from_numpy(small), tl).numpy(), g["small_global"]) < 1e-6
    full = syn.lidar_batch(301, 1)
    assert max_rel(orc.pointnet_global(full, layers), g["full_global"]) < FP32_TOL


def test_zero_padding_rows_take_part_in_the_max():
    """an all-zero row yields non-zero features and must not be masked."""
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    only_pad = np.zeros((1, 3, 4), dtype=np.float32)
    assert orc.pointnet_global(only_pad, layers).max() > 0.0


def test_cell_canvas_restatement(golden):
    g = golden("lidar_encoder")
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    pts = syn.lidar_batch(201, 2, n_valid=1900, n_total=2011)
    cell = orc.cell_index(pts, syn.PC_RANGE, 50, 50)
    np.testing.assert_array_equal(cell, g["small_cell"])
    for b in range(2):
        canvas = orc.pointnet_cell_max(pts[b], layers, cell[b], 2500)
        assert max_rel(canvas[:, ::16], g["small_canvas_sub"][b]) < FP32_TOL
    # zero rows land in the centre cell (W/2, H/2) — SURVEY Q5
    assert np.all(cell[:, 1900:] == 25 * 50 + 25)


def test_cell_index_edges():
    W = H = 50
    vx = np.float32(102.4 / 50)
    pts = np.array([[-51.2, -51.2], [51.2, 0.0], [np.nextafter(np.float32(51.2), np.float32(0)), 0.0],
                    [-51.3, 0.0], [0.0, np.nan], [np.float32(-51.2) + vx, np.float32(-51.2) + vx],
                    [0.0, 0.0]], dtype=np.float32)
    c = orc.cell_index(pts, syn.PC_RANGE, W, H)
    assert c[0] == 0                      # lower corner is inside
    assert c[1] == -1                     # px == W is rejected (src/centernet_target.py:254)
    assert c[2] == 25 * 50 + 49
    assert c[3] == -1 and c[4] == -1      # below range, NaN
    assert c[5] == 51                     # exactly one voxel in -> cell (1,1)
    assert c[6] == 25 * 50 + 25
    perm, off = orc.bin_sort(c, W * H)
    assert off[-1] == 5 and list(perm[-2:]) == [1, 3] or list(perm[5:]) == [1, 3, 4][-2:]
    assert sorted(perm.tolist()) == list(range(7))


def test_bin_sort_is_stable():
    rng = np.random.default_rng(0)
    cell = rng.integers(-1, 12, 500).astype(np.int32)
    perm, off = orc.bin_sort(cell, 12)
    for c in range(12):
        seg = perm[off[c]:off[c + 1]]
        assert np.all(cell[seg] == c) and np.all(np.diff(seg) > 0)
    tail = perm[off[12]:]
    assert np.all(cell[tail] == -1) and np.all(np.diff(tail) > 0)


@pytest.mark.parametrize("method", ["concat", "max", "mean"])
def test_multi_radar(golden, method):
    g = golden("radar_encoder")
    layers = syn.mlp_weights(111, syn.RADAR_DIMS)
    fcw, fcb = syn.linear_weights(112, 5 * 256, 256)
    radars = syn.radar_batch(211, 3)
    fused, stacked = orc.multi_radar(radars, layers, method, fcw, fcb)
    assert stacked.shape == (3, 5, 256)
    assert max_rel(fused, g[f"fused_{method}"]) < FP32_TOL
    ragged = [r[:, : 125 - 17 * i] for i, r in enumerate(radars)]
    assert max_rel(orc.multi_radar(ragged, layers, method, fcw, fcb)[0], g[f"ragged_{method}"]) < FP32_TOL
    if method == "concat":
        tl = torch_port.layers_to_torch(layers)
        got = torch_port.multi_radar([torch.from_numpy(r) for r in radars], tl, torch.from_numpy(fcw), torch.from_numpy(fcb))
        assert max_rel(got.numpy(), g["fused_concat"]) < 1e-6


def test_unknown_radar_fusion_raises():
    with pytest.raises(ValueError, match="Unknown fusion method"):
        orc.multi_radar([np.zeros((1, 2, 7), np.float32)], syn.mlp_weights(1, syn.RADAR_DIMS), "sum")


@pytest.mark.parametrize("name,shape", [("ref28x50", (16, 28, 50, 50, 50)), ("hd57x100", (8, 57, 100, 50, 50)),
                                        ("up7x9", (8, 7, 9, 20, 30))])
def test_camera_mean_and_resize(golden, name, shape):
    g = golden("camera_bev")
    C, h, w, H, W = shape
    feats = syn.camera_features(401, 2, n_cam=6, channels=C, h=h, w=w)
    assert syn.digest(feats) == str(g[f"{name}_digest"])
    # same association as ATen's vectorised outer reduction (bit-exact on the two real shapes); ATen's
    # scalar path for tiny planes differs by 1 ulp in <1% of elements
    assert max_rel(orc.camera_mean(feats), g[f"{name}_mean"]) < 2e-7
    if name != "up7x9":
        np.testing.assert_array_equal(orc.camera_mean(feats), g[f"{name}_mean"])
    got = orc.bilinear_resize(g[f"{name}_resize_in"], (H, W))
    assert max_rel(got, g[f"{name}_resize_out"]) < FP32_TOL
    tp = torch_port.bilinear_resize(torch.from_numpy(g[f"{name}_resize_in"]), (H, W)).numpy()
    np.testing.assert_array_equal(tp, g[f"{name}_resize_out"])


def test_camera_projection_restatement(golden):
    g = golden("camera_bev")
    K, E = syn.camera_rig()
    feats = syn.camera_features(402, 1, n_cam=6, channels=8, h=57, w=100)
    assert syn.digest(feats, K, E) == str(g["project_digest"])
    table = orc.project_cells(K, E, (1600.0, 900.0), (57, 100), (50, 50), syn.PC_RANGE)
    np.testing.assert_array_equal(table, g["project_table"])
    canvas = orc.camera_project(feats[0], table, (50, 50))
    assert max_rel(canvas, g["project_canvas_grid_sample"]) < FP32_TOL
    seen = table[:, :, 2].sum(axis=1)
    assert seen.max() >= 2 and (seen == 0).any()     # overlapping cameras and blind cells both occur


def test_nms_and_topk(golden):
    g = golden("centernet_decode")
    heat = syn.head_maps(501, 3)["heatmap"]
    nms = orc.nms(heat)
    np.testing.assert_array_equal(nms, g["nms"])
    score, ind, cls, ys, xs = orc.topk(nms, 100)
    assert np.all(score > 0)                           # tie-free regime: every winner is a real peak
    np.testing.assert_array_equal(score, g["topk_score"])
    np.testing.assert_array_equal(ind, g["topk_ind"])
    np.testing.assert_array_equal(ys, g["topk_ys"])
    np.testing.assert_array_equal(xs, g["topk_xs"])
    np.testing.assert_array_equal(cls, g["topk_classes"])
    assert not cls.any()                               # labels are identically 0


def test_hand_made_peaks(golden):
    g = golden("centernet_decode")
    nms = orc.nms(g["hand_heat"])
    np.testing.assert_array_equal(nms, g["hand_nms"])
    assert nms[0, 0, 0, 0] == np.float32(0.9) and nms[0, 0, 5, 6] == np.float32(0.6)   # border maxima survive (-inf pad)
    assert np.all(nms[0, 1, 3:5, 2:4] == np.float32(0.4))                                # plateau is kept whole
    score, _, _, ys, xs = orc.topk(nms, 6)
    np.testing.assert_array_equal(score, g["hand_topk_score"])
    # six distinct values above the plateau: positions are pinned by the reference
    np.testing.assert_array_equal(ys, g["hand_topk_ys"])
    np.testing.assert_array_equal(xs, g["hand_topk_xs"])
    # ranks 7 and 8 come from the 2x2 plateau of equal values; torch leaves their order undefined
    # (SURVEY Q4), the oracle and the kernel define it as ascending flat index
    score8, _, _, ys8, xs8 = orc.topk(nms, 8)
    assert np.all(score8[0, 6:] == np.float32(0.4))
    assert [(int(ys8[0, i]), int(xs8[0, i])) for i in (6, 7)] == [(3, 2), (3, 3)]
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        orc.topk(g["hand_heat"], 43)
    assert "selected index k out of range" in str(g["k_too_large_message"])


def _assert_dets(dets, g, prefix, n_samples):
    for b in range(n_samples):
        ref_scores = g[f"{prefix}_b{b}_scores"]
        d = dets[b]
        assert d["scores"].shape == ref_scores.shape
        np.testing.assert_array_equal(d["scores"], ref_scores)
        np.testing.assert_array_equal(d["labels"], g[f"{prefix}_b{b}_labels"])
        assert d["labels"].dtype == np.int64
        np.testing.assert_array_equal(d["velocities"], g[f"{prefix}_b{b}_velocities"])
        ref_boxes = g[f"{prefix}_b{b}_boxes"]
        assert d["boxes"].shape == ref_boxes.shape
        if len(ref_boxes):
            np.testing.assert_array_equal(d["boxes"][:, :6], ref_boxes[:, :6])   # mul/add: exact
            np.testing.assert_allclose(d["boxes"][:, 6], ref_boxes[:, 6], rtol=0, atol=1e-6)  # atan2 libm vs numpy


@pytest.mark.parametrize("tag,voxel", [("ct", 2.048), ("fd", 0.512)])
@pytest.mark.parametrize("thr", [0.0, 0.3, 0.999])
def test_decode_matches_both_reference_copies(golden, tag, voxel, thr):
    g = golden("centernet_decode")
    maps = syn.head_maps(501, 3)
    dets = orc.decode(maps, score_thresh=thr, max_detections=100, voxel_size_m=voxel)
    _assert_dets(dets, g, f"{tag}_thr{thr}", 3)
    if thr == 0.999:
        assert any(len(d["scores"]) == 0 for d in dets) or all(len(d["scores"]) < 100 for d in dets)


def test_decode_sparse_and_big(golden):
    g = golden("centernet_decode")
    sparse = syn.head_maps(502, 2, peak_frac=0.02)
    assert syn.digest(*sparse.values()) == str(g["sparse_digest"])
    _assert_dets(orc.decode(sparse, score_thresh=0.1), g, "sparse", 2)
    big = syn.head_maps(503, 1, H=100, W=100)
    dets = orc.decode(big, score_thresh=0.0, voxel_size_m=0.512)
    g2 = {f"big_b0_{k}": g[f"big_{k}"] for k in ("boxes", "scores", "labels", "velocities")}
    _assert_dets(dets, g2, "big", 1)


def test_torch_port_decode_equals_reference(golden):
    g = golden("centernet_decode")
    maps = {k: torch.from_numpy(v) for k, v in syn.head_maps(501, 3).items()}
    dets = torch_port.decode(maps, score_thresh=0.0)
    for b, d in enumerate(dets):
        for k, v in d.items():
            np.testing.assert_array_equal(v.numpy(), g[f"ct_thr0.0_b{b}_{k}"])


# ------------------------------------------------------------------------------------------------ N1 / N2
def test_dense_layers_match_the_reference_fusion_module(golden):
    """lidar_init (Linear+ReLU+Linear, 80000 outputs) and radar_proj of the reference's FlexibleBEVFusion."""
    g = golden("bev_glue")
    w1, b1 = syn.linear_weights(701, 1024, 512)
    w2, b2 = syn.linear_weights(702, 512, 128 * 25 * 25)
    wr, br = syn.linear_weights(703, 256, 256)
    feats, radar = syn.global_features(704, 3, 1024), syn.global_features(705, 3, 256)
    assert syn.digest(w1, b1, w2, b2, wr, br, feats, radar) == str(g["dense_digest"])
    hidden = orc.dense_layer(feats, w1, b1, relu=True)
    assert max_rel(hidden, g["lidar_hidden"]) < 1e-5
    full = orc.lidar_init(feats, w1, b1, w2, b2)
    assert full.shape == (3, 80000)
    assert np.abs(full[:, ::16] - g["lidar_init_sub"]).max() < 1e-5 * float(g["lidar_init_absmax"])
    assert max_rel(orc.dense_layer(radar, wr, br, relu=True), g["radar_proj"]) < 1e-5


def test_sigmoid_then_decode_matches_the_reference_head(golden):
    """CenterNetHead's raw heat-map output -> sigmoid -> decode, against what the reference decoded."""
    g = golden("bev_glue")
    heat = orc.sigmoid(g["head_logits"])
    assert max_rel(heat, g["head_heatmap"]) < 1e-6
    pred = {"heatmap": g["head_heatmap"], **{k: g[f"head_{k}"] for k in ("offset", "size", "rot", "vel")}}
    for tag, thr in (("all", 0.0), ("mid", float(g["head_thresh"]))):
        dets = orc.decode(pred, score_thresh=thr, max_detections=60, voxel_size_m=0.512)
        _assert_dets(dets, g, f"head_{tag}", 2)
    assert 0 < len(g["head_mid_b0_scores"]) < 60


def test_conv_stacks_match_the_reference_fusion_module(golden):
    """camera mean -> camera_proj (3x3 + 1x1 blocks) -> bilinear resize -> bev_fusion (two 3x3 blocks), restated block by
    block, against the output of the reference's FlexibleBEVFusion with the same state_dict."""
    g = golden("bev_glue")
    shapes = {}
    for name, (ci, co, k) in {"camera_proj.0": (64, 512, 3), "camera_proj.3": (512, 64, 1), "bev_fusion.0": (64, 128, 3),
                              "bev_fusion.3": (128, 64, 3)}.items():
        seq, idx = name.split(".")
        shapes[f"{name}.weight"], shapes[f"{name}.bias"] = (co, ci, k, k), (co,)
        for stat in ("weight", "bias", "running_mean", "running_var"):
            shapes[f"{seq}.{int(idx) + 1}.{stat}"] = (co,)
        shapes[f"{seq}.{int(idx) + 1}.num_batches_tracked"] = ()
    sd = syn.fill_state_dict(731, shapes)
    cam = syn.camera_features(732, 2, n_cam=6, channels=64, h=9, w=14)
    assert syn.digest(cam, *[sd[k] for k in sorted(sd)]) == str(g["stack_digest"])

    def block(x, name, relu=True):
        seq, idx = name.split(".")
        bn = {s: sd[f"{seq}.{int(idx) + 1}.{s}"] for s in ("weight", "bias", "running_mean", "running_var")}
        return orc.conv_bn_relu(x, sd[f"{name}.weight"], sd[f"{name}.bias"], bn, relu)

    x = block(block(orc.camera_mean(cam), "camera_proj.0"), "camera_proj.3")
    x = orc.bilinear_resize(x, (12, 20))
    x = block(block(x, "bev_fusion.0"), "bev_fusion.3")
    assert max_rel(x, g["stack_out"]) < 5e-5      # four fp32 convolutions deep, different summation order


def test_metrics_consumer_restatement_matches_the_reference(golden):
    """N4: mAP / NDS of the reference's compute_metrics on its own decode outputs, restated in numpy."""
    g = golden("metrics")
    dets = orc.decode(syn.head_maps(501, 3), score_thresh=0.3, max_detections=100, voxel_size_m=0.512)
    gts = syn.ground_truth_near(801, dets)
    assert syn.digest(*[a for gt in gts for a in gt.values()]) == str(g["gt_digest"])
    m = orc.compute_metrics(dets, gts)
    assert abs(m["mAP"] - float(g["mAP"])) < 1e-12 and abs(m["NDS"] - float(g["NDS"])) < 1e-6
    np.testing.assert_allclose([m["AP_per_class"][c] for c in orc.CLASS_NAMES], g["AP_per_class"], rtol=0, atol=1e-12)
    assert 0.0 < m["mAP"] < 1.0 and m["AP_per_class"]["car"] > 0.1


@pytest.mark.parametrize("B,C,O,H,W,k", [(2, 5, 7, 6, 9, 3), (1, 8, 3, 1, 1, 3), (3, 4, 4, 5, 2, 1)])
def test_conv_and_dense_restatements_against_aten(B, C, O, H, W, k):
    """The numpy conv block and dense layer against the ATen ops the reference calls (conv2d, batch_norm in eval mode,
    relu, linear), on shapes with edges everywhere."""
    g = np.random.default_rng(B * 100 + C)
    x = g.standard_normal((B, C, H, W)).astype(np.float32)
    w = g.standard_normal((O, C, k, k)).astype(np.float32)
    b = g.standard_normal(O).astype(np.float32)
    bn = {"weight": g.uniform(0.5, 1.5, O).astype(np.float32), "bias": g.standard_normal(O).astype(np.float32),
          "running_mean": g.standard_normal(O).astype(np.float32), "running_var": g.uniform(0.5, 2.0, O).astype(np.float32)}
    t = torch.nn.functional.conv2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), padding=k // 2)
    t = torch.nn.functional.batch_norm(t, torch.from_numpy(bn["running_mean"]), torch.from_numpy(bn["running_var"]),
                                       torch.from_numpy(bn["weight"]), torch.from_numpy(bn["bias"]), training=False, eps=1e-5)
    assert max_rel(orc.conv_bn_relu(x, w, b, bn, relu=True), torch.relu(t).numpy()) < 1e-5
    assert max_rel(orc.conv_bn_relu(x, w, b, None, relu=False),
                   torch.nn.functional.conv2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), padding=k // 2).numpy()) < 1e-5
    xf = x.reshape(B, -1)
    wl = g.standard_normal((11, xf.shape[1])).astype(np.float32)
    bl = g.standard_normal(11).astype(np.float32)
    lin = torch.nn.functional.linear(torch.from_numpy(xf), torch.from_numpy(wl), torch.from_numpy(bl))
    assert max_rel(orc.dense_layer(xf, wl, bl, relu=True), torch.relu(lin).numpy()) < 1e-5
    assert max_rel(orc.sigmoid(xf), torch.sigmoid(torch.from_numpy(xf)).numpy()) < 1e-6


def test_detector_chain_port_and_mirror_names_vs_reference_golden(golden):
    """The whole inference pass behind the camera backbone (src/fusion.py:1113-1137 + src/eval.py:58-62): the torch port
    the reference arm times reproduces what the reference's own modules produced, and the mirror chain has the
    reference's state_dict names and shapes for those sub-modules."""
    import bevfusion_multimodal_3d_object_detection_b200 as b200bev

    g = golden("detector_chain")
    chain = b200bev.BEVDetectorChain()
    shapes = chain.state_shapes()
    assert sorted(shapes) == [str(n) for n in g["shape_names"]]
    for k, shp in shapes.items():
        assert tuple(int(v) for v in g["shape__" + k]) == shp, k
    sd = syn.detector_state(syn.CHAIN_SEED, shapes)
    assert syn.digest(*[sd[k] for k in sorted(sd)]) == str(g["state_digest"])
    lidar, radars, cam = syn.chain_inputs()
    assert syn.digest(lidar, *radars, cam) == str(g["input_digest"])
    tsd = {k: torch.from_numpy(v) for k, v in sd.items()}
    bev, pred, dets = torch_port.detector_chain(tsd, torch.from_numpy(cam), torch.from_numpy(lidar), [torch.from_numpy(r) for r in radars])
    assert max_rel(bev[:, ::8].numpy(), g["bev_sub"]) < 1e-5
    for k in ("heatmap", "offset", "size", "rot", "vel"):
        assert max_rel(pred[k].numpy(), g["pred_" + k]) < 1e-5, k
    for b, d in enumerate(dets):
        assert d["scores"].shape == g[f"det_b{b}_scores"].shape
        np.testing.assert_allclose(d["scores"].numpy(), g[f"det_b{b}_scores"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(d["boxes"].numpy(), g[f"det_b{b}_boxes"], rtol=0, atol=1e-4)
    # training mode of the mirror chain is the plain torch graph (CPU is fine there): loads the same state and runs
    chain.load_state_dict(tsd)
    chain.train()
    out = chain(torch.from_numpy(cam[:1, :, :, :8, :10]), torch.from_numpy(lidar[:1, :64]), [torch.from_numpy(r[:1]) for r in radars])
    assert set(out) == {"heatmap", "offset", "size", "rot", "vel"} and tuple(out["heatmap"].shape) == (1, 10, 50, 50)
