"""CPU tests: the C-ABI library loads and exports exactly what include/b200bev.h declares, argument
validation happens before any CUDA work, and the host-side mirrors keep the reference's interface."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import bevfusion_multimodal_3d_object_detection_b200 as b200bev
from bevfusion_multimodal_3d_object_detection_b200 import _lib, build, encoders, fusion, ops
from bevfusion_multimodal_3d_object_detection_b200 import synthetic as syn
from oracle import bev_oracle as orc
from oracle import torch_port
from tests.conftest import max_rel

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "b200bev.h"


@pytest.fixture(scope="module")
def lib():
    build.build()           # no-op when the in-tree .so is current; nvcc cross-compiles without a GPU
    return _lib.lib()


def declared_symbols():
    return re.findall(r"^B200BEV_API\s+[\w\s\*]+?\b(b200bev_\w+)\s*\(", HEADER.read_text(), flags=re.M)


def test_header_and_binding_declare_the_same_entry_points():
    names = declared_symbols()
    assert len(names) == len(set(names)) >= 15
    assert set(names) == set(_lib.PROTOTYPES)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} missing from libb200bev.so"
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\b(b200bev_\w+)\b", out))
    assert exported == set(declared_symbols())          # nothing undeclared leaks out either


def test_library_is_sm100a_only(lib):
    out = subprocess.run(["cuobjdump", "--list-elf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_version_and_error_strings(lib):
    assert lib.b200bev_abi_version() == _lib.ABI_VERSION
    assert lib.b200bev_error_string(0) == b"ok"
    assert b"k out of range" in lib.b200bev_error_string(_lib.ERR_K_OUT_OF_RANGE)
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        _lib.check(_lib.ERR_K_OUT_OF_RANGE)
    with pytest.raises(_lib.B200BevError):
        _lib.check(_lib.ERR_INVALID_ARGUMENT)


def test_argument_validation_needs_no_gpu(lib):
    null = C.c_void_p(0)
    one = C.c_void_p(16)
    assert lib.b200bev_bin_sort(null, 1, 1, 4, 0, 0, 1, 1, 2, 2, one, one, one, null) == _lib.ERR_INVALID_ARGUMENT
    assert lib.b200bev_bin_sort(one, 1, 1, 4, 0, 0, 0.0, 1, 2, 2, one, one, one, null) == _lib.ERR_INVALID_ARGUMENT
    assert lib.b200bev_centernet_nms(null, 1, 1, 1, 1, one, null) == _lib.ERR_INVALID_ARGUMENT
    dims = (C.c_int32 * 3)(4, 8, 16)
    assert lib.b200bev_pointnet_encode(one, 1, 0, 4, one, dims, 2, null, null, 0, 0, null, one, null, null) == _lib.ERR_INVALID_ARGUMENT
    # K > H*W mirrors torch.topk's error; workspace too small is its own code
    ws = lib.b200bev_centernet_workspace_bytes(1, 2, 43)
    assert ws >= 2 * 43 * 8 + 4
    assert lib.b200bev_centernet_topk(one, 1, 2, 6, 7, 43, one, one, one, one, one, one, ws, null) == _lib.ERR_K_OUT_OF_RANGE
    assert lib.b200bev_centernet_topk(one, 1, 2, 6, 7, 5, one, one, one, one, one, one, 8, null) == _lib.ERR_WORKSPACE
    assert lib.b200bev_camera_project(one, 2, 6, 8, 4, 4, one, one, 3, 1600, 900, 0, 0, 1, 1, 0, 5, 5, one, null, 0, null) \
        == _lib.ERR_INVALID_ARGUMENT                     # T must be 1 or B
    # SURVEY 8f entry points: dense layers and convolution blocks
    assert lib.b200bev_dense_layer(null, 1, 4, one, null, 4, 0, one, null) == _lib.ERR_INVALID_ARGUMENT
    assert lib.b200bev_dense_layer(one, 0, 4, one, null, 4, 0, one, null) == _lib.ERR_INVALID_ARGUMENT
    assert lib.b200bev_lidar_init(one, 1, 4, one, one, 4, one, one, 4, null, one, null) == _lib.ERR_INVALID_ARGUMENT
    assert lib.b200bev_conv_pack_bytes(320, 256, 9) == 3 * 9 * 4 * 16384          # 3 channel tiles x 9 taps x 4 chunks of 16 KB
    assert lib.b200bev_conv_pack_bytes(19, 320, 1) == 5 * 16384
    assert lib.b200bev_conv_pack_bytes(64, 48, 9) == 0 and lib.b200bev_conv_pack_bytes(64, 64, 25) == 0
    assert lib.b200bev_conv_pack_bf16(one, 64, 48, 9, one, 1 << 20, null) == _lib.ERR_UNSUPPORTED
    assert lib.b200bev_conv_pack_bf16(one, 64, 64, 9, one, 16, null) == _lib.ERR_WORKSPACE
    assert lib.b200bev_conv_bn_relu_bf16(one, 1, 4, 4, 48, one, null, 8, 9, 1, one, null) == _lib.ERR_UNSUPPORTED
    assert lib.b200bev_conv_bn_relu_bf16(one, 1, 4, 4, 64, one, null, 8, 9, 1, null, null) == _lib.ERR_INVALID_ARGUMENT
    assert lib.b200bev_conv_bn_relu_bf16_nhwc(one, 1, 4, 4, 64, one, null, 8, 1, 1, one, 12, 8, null, null) == _lib.ERR_INVALID_ARGUMENT  # slice past C_total
    assert lib.b200bev_nchw_to_nhwc_bf16(one, 1, 8, 4, 4, one, 12, 8, null) == _lib.ERR_INVALID_ARGUMENT
    assert lib.b200bev_centernet_decode_logits(null, one, one, one, one, 1, 1, 4, 4, 2, 1.0, 0, 0, 0, 0, one, one, one, one, null, null,
                                               null, one, one, 64, null) == _lib.ERR_INVALID_ARGUMENT


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setenv("B200BEV_LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_handle", None)
    with pytest.raises(ImportError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_cpu_tensors_are_rejected_not_computed():
    enc = encoders.PointNetLiDAREncoder(input_channels=4).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.zeros(1, 8, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        b200bev.decode_centernet_predictions({k: torch.zeros(1, c, 4, 4) for k, c in
                                              (("heatmap", 2), ("offset", 2), ("size", 3), ("rot", 2), ("vel", 2))})


# ------------------------------------------------------------------------------------------------ interface
LIDAR_KEYS = [f"{p}{i}.{s}" for i in range(1, 6) for p, s in
              (("conv", "weight"), ("conv", "bias"), ("bn", "weight"), ("bn", "bias"), ("bn", "running_mean"),
               ("bn", "running_var"), ("bn", "num_batches_tracked"))]


def test_state_dict_layout_matches_the_reference():
    """Names and shapes of SURVEY §8b; 702,528 / 372,608 parameters as the reference reports."""
    lid = encoders.PointNetLiDAREncoder(input_channels=4, feat_dim=1024)
    assert sorted(lid.state_dict()) == sorted(LIDAR_KEYS)
    assert tuple(lid.conv1.weight.shape) == (64, 4, 1) and tuple(lid.conv5.weight.shape) == (1024, 512, 1)
    assert sum(p.numel() for p in lid.parameters()) == 702528
    rad = encoders.MultiRadarEncoder()
    assert tuple(rad.radar_encoder.conv1.weight.shape) == (32, 7, 1)
    assert tuple(rad.fusion_fc.weight.shape) == (256, 1280)
    assert sum(p.numel() for p in rad.parameters()) == 372608
    fus = fusion.FlexibleBEVFusion(bev_h=50, bev_w=50)
    keys = set(fus.state_dict())
    for k in ("camera_proj.0.weight", "camera_proj.4.running_var", "lidar_init.2.weight", "lidar_upsample.5.bias",
              "radar_proj.0.weight", "radar_refine.4.weight", "bev_fusion.3.weight"):
        assert k in keys
    assert tuple(fus.lidar_init[2].weight.shape) == (80000, 512)
    assert sum(p.numel() for p in fus.parameters()) == 50468864


def test_config_driven_construction():
    cfg = {"model": {"lidar_encoder": {"input_channels": 4, "feature_dim": 1024, "mlp_layers": [64, 128, 256, 512, 1024]},
                     "radar_encoder": {"input_channels": 7, "num_radars": 3, "fusion_method": "max"},
                     "bev_fusion": {"bev_h": 50, "bev_w": 50, "bev_channels": 64}, "use_radar": False},
           "dataset": {"point_cloud_range": [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]}}
    assert encoders.PointNetLiDAREncoder(config=cfg).input_channels == 4
    rad = encoders.MultiRadarEncoder(config=cfg)
    assert rad.num_radars == 3 and rad.fusion_method == "max" and not hasattr(rad, "fusion_fc")
    fus = fusion.FlexibleBEVFusion(config=cfg)
    assert (fus.bev_h, fus.bev_channels, fus.use_radar, fus.num_modalities) == (50, 64, False, 2)
    with pytest.raises(FileNotFoundError):
        encoders.load_config("/nonexistent/base.yaml")
    with pytest.raises(AssertionError):
        fusion.FlexibleBEVFusion(use_camera=False, use_lidar=False, use_radar=False)


def _load(module, layers):
    sd = module.state_dict()
    for i, lay in enumerate(layers, start=1):
        sd[f"conv{i}.weight"] = torch.from_numpy(lay["weight"]).unsqueeze(-1)
        sd[f"conv{i}.bias"] = torch.from_numpy(lay["bias"])
        for ours, theirs in (("bn_weight", "weight"), ("bn_bias", "bias"), ("bn_mean", "running_mean"), ("bn_var", "running_var")):
            sd[f"bn{i}.{theirs}"] = torch.from_numpy(lay[ours])
    module.load_state_dict(sd)


def test_batchnorm_folding_matches_the_unfolded_oracle(golden):
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    enc = encoders.PointNetLiDAREncoder(input_channels=4).eval()
    _load(enc, layers)
    ws, bs = [], []
    for conv, bn in encoders._mlp_stages(enc):
        w, b = ops.fold_batchnorm(conv.weight, conv.bias, bn)
        ws.append(w.numpy())
        bs.append(b.numpy())
    ref_w, ref_b = orc.fold_layers(layers)
    for a, b in zip(ws + bs, ref_w + ref_b):
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=0)
    # folded fp32 chain == reference output within the fp32 bound
    pts = syn.lidar_batch(201, 2, n_valid=1900, n_total=2011)
    x = pts.astype(np.float32)
    for w, b in zip(ws, bs):
        x = np.maximum(x @ w.astype(np.float32).T + b.astype(np.float32), 0)
    assert max_rel(x.max(axis=1), golden("lidar_encoder")["small_global"]) < 1e-5
    # blob layout: W^T then bias, layer after layer
    blob, dims = ops.pack_mlp_params([torch.from_numpy(w) for w in ws], [torch.from_numpy(b) for b in bs], torch.device("cpu"))
    assert dims == list(syn.LIDAR_DIMS) and blob.numel() == sum(a * b + b for a, b in zip(dims[:-1], dims[1:]))
    np.testing.assert_array_equal(blob[:256].view(4, 64).numpy(), ws[0].T.astype(np.float32))
    np.testing.assert_array_equal(blob[256:320].numpy(), bs[0].astype(np.float32))


def test_training_mode_uses_the_torch_graph_and_matches_the_reference_ops():
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    enc = encoders.PointNetLiDAREncoder(input_channels=4)
    _load(enc, layers)
    enc.train()
    pts = torch.from_numpy(syn.lidar_batch(201, 2, n_valid=100, n_total=128))
    out = enc(pts)
    assert out.requires_grad and out.shape == (2, 1024)
    out.sum().backward()
    assert enc.conv1.weight.grad is not None
    # (B,C,N) layout is accepted like the reference does
    enc.eval()
    enc.return_point_features = True            # forces the torch path on CPU for an eval-mode comparison
    both = enc(pts)
    assert both.shape == (2, 128, 2048)
    ref = torch_port.shared_mlp_max(pts, torch_port.layers_to_torch(layers))
    # eval BN here uses the running stats that the train-mode pass above has updated, so compare the
    # structure only: global feature is repeated along points in the second half
    assert torch.equal(both[:, 0, 1024:], both[:, 5, 1024:]) and ref.shape == (2, 1024)


def test_weight_cache_follows_parameter_updates():
    enc = encoders.PointNetLiDAREncoder(input_channels=4).eval()
    k0 = encoders._state_key(enc, torch.device("cpu"))
    with torch.no_grad():
        enc.bn3.running_mean.add_(1.0)
    k1 = encoders._state_key(enc, torch.device("cpu"))
    enc.load_state_dict(enc.state_dict())
    k2 = encoders._state_key(enc, torch.device("cpu"))
    assert k0 != k1 and k1 != k2


def test_patch_rebinds_the_reference_names():
    ref_src = Path("/root/reference/src")
    if not ref_src.exists():
        pytest.skip("reference checkout not present (GPU box)")
    code = """
import sys, io, contextlib
sys.path.insert(0, %r); sys.path.insert(0, %r)
with contextlib.redirect_stdout(io.StringIO()):
    import encoders, fusion, centernet_target, fusion_detection, eval as ref_eval
import bevfusion_multimodal_3d_object_detection_b200 as b
orig = encoders.PointNetLiDAREncoder.forward
done = b.patch()
assert set(done) >= {'encoders', 'fusion', 'centernet_target', 'fusion_detection', 'eval'}, done
assert encoders.PointNetLiDAREncoder.forward is not orig
assert ref_eval.decode_centernet_predictions is fusion_detection.decode_centernet_predictions
import torch
enc = encoders.PointNetLiDAREncoder(input_channels=4)
assert enc.train()(torch.zeros(2, 16, 4)).shape == (2, 1024)      # training: reference graph, CPU is fine
try:
    enc.eval()(torch.zeros(2, 16, 4)); raise SystemExit('eval on CPU must raise')
except RuntimeError as e:
    assert 'no CPU fallback' in str(e)
b.unpatch()
assert encoders.PointNetLiDAREncoder.forward is orig
print('PATCH-OK')
""" % (str(ROOT), str(ref_src))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "PATCH-OK" in r.stdout, r.stdout + r.stderr


def test_calibration_plumbing_round_trip():
    """N4: info dicts in the converter's layout (src/data_converter.py:110-117) -> the [R|t] the projection kernel takes."""
    from bevfusion_multimodal_3d_object_detection_b200 import dataset
    from bevfusion_multimodal_3d_object_detection_b200 import synthetic as syn

    def to_quat(R):                                              # rotation matrix -> (w, x, y, z), w >= 0 branch and the others
        t = np.trace(R)
        if t > 0:
            s = np.sqrt(t + 1.0) * 2
            return [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = [0.0] * 4
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
        return q

    K, E = syn.camera_rig()
    # a lidar mounted 0.9 m forward, 1.8 m up, yawed by -90 degrees as on the nuScenes car
    yaw = -np.pi / 2
    R_l = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1.0]])
    t_l = np.array([0.9, 0.0, 1.8])
    info = {"lidar_calibrated_sensor": {"translation": t_l.tolist(), "rotation": to_quat(R_l)}, "cams": {}}
    for name, k, e in zip(dataset.CAMERA_ORDER, K, E):
        R_ec, t_ec = e[:, :3].astype(np.float64), e[:, 3].astype(np.float64)   # ego -> camera
        R_c, t_c = R_ec.T, -R_ec.T @ t_ec                                        # sensor -> ego, what nuScenes stores
        info["cams"][name] = {"filename": "x.jpg", "calibrated_sensor": {
            "translation": t_c.tolist(), "rotation": to_quat(R_c), "camera_intrinsic": k.tolist()}}
    K2, E_ego = dataset.calibration_from_info(info, frame="ego")
    np.testing.assert_array_equal(K2, K)
    np.testing.assert_allclose(E_ego, E, atol=2e-6)
    _, E_lidar = dataset.calibration_from_info(info, frame="lidar")
    p = np.array([3.0, -7.0, 0.5])                                # a point in the lidar frame, through both routes
    for e_l, e_e in zip(E_lidar, E):
        via_ego = e_e[:, :3] @ (R_l @ p + t_l) + e_e[:, 3]
        np.testing.assert_allclose(e_l[:, :3] @ p + e_l[:, 3], via_ego, atol=1e-5)
    Kb, Eb = dataset.calibration_batch([info, info], torch.device("cpu"))
    assert tuple(Kb.shape) == (2, 6, 3, 3) and tuple(Eb.shape) == (2, 6, 3, 4)
    with pytest.raises(ValueError):
        dataset.calibration_from_info(info, frame="world")


def test_centernet_head_mirror_state_dict_and_default_path(golden):
    """The head mirror keeps the reference's sub-module / state_dict names (src/fusion.py:822-854) and, with the bf16
    path off, runs the reference's own layers: its CPU output equals what the reference's CenterNetHead produced."""
    import bevfusion_multimodal_3d_object_detection_b200 as b200bev
    from bevfusion_multimodal_3d_object_detection_b200 import conv_blocks

    g = golden("bev_glue")
    head = b200bev.CenterNetHead(in_channels=32, num_classes=10, head_conv=16)
    want = {f"{n}_head.{i}.{p}" for n in conv_blocks.HEADS for i in (0, 2) for p in ("weight", "bias")}
    assert set(head.state_dict()) == want
    assert abs(float(head.heatmap_head[2].bias.detach()[0]) + 4.59512) < 1e-4                  # prior 0.01, src/fusion.py:865-867
    head.load_state_dict({k: torch.from_numpy(v) for k, v in syn.head_weights(711, 32, 16, 10).items()})
    x = torch.from_numpy(syn._rng(712).standard_normal((2, 32, 24, 40)).astype(np.float32))
    with torch.no_grad():
        pred = head.eval()(x)
    for k in ("heatmap", "offset", "size", "rot", "vel"):
        np.testing.assert_allclose(pred[k].numpy(), g[f"head_{k}"], rtol=0, atol=1e-6 * float(np.abs(g[f"head_{k}"]).max()) + 1e-7)
    assert "heatmap_logits" not in pred
    # what the tcgen05 path accepts: Cin a multiple of 64, 3x3/pad 1 or 1x1, hidden width a multiple of 64
    assert not conv_blocks.head_supported(head)
    assert conv_blocks.head_supported(b200bev.CenterNetHead(in_channels=256, num_classes=10, head_conv=64))
    fus = b200bev.FlexibleBEVFusion(use_camera=True, use_lidar=True, use_radar=True, camera_channels=512, bev_h=50, bev_w=50)
    assert all(conv_blocks.supported(s) for s in (fus.camera_proj, fus.lidar_upsample, fus.radar_refine, fus.bev_fusion))
    odd = b200bev.FlexibleBEVFusion(use_camera=True, use_lidar=False, use_radar=False, camera_channels=16, bev_h=8, bev_w=8,
                                    bev_channels=8)
    assert not conv_blocks.supported(odd.camera_proj)
    with pytest.raises(RuntimeError):                                                 # eval + CPU: no fallback for the kernels
        fus.eval()(lidar_features=torch.zeros(1, 1024))


def test_lidar_start_size_is_the_references_25_unless_asked():
    """The reference hard-codes 25 (src/fusion.py:141): the mirror keeps it for every grid so that the state_dict shapes
    match; `lidar_start_size=` is the explicit extension that lets other grids run with the lidar branch (the reference
    itself raises in torch.cat there, src/fusion.py:292)."""
    for hw in (50, 200):
        f = b200bev.FlexibleBEVFusion(use_camera=False, use_lidar=True, use_radar=False, bev_h=hw, bev_w=hw, bev_channels=8)
        assert f.lidar_start_size == 25 and tuple(f.lidar_init[2].weight.shape) == (128 * 25 * 25, 512)
    f20 = b200bev.FlexibleBEVFusion(use_camera=False, use_lidar=True, use_radar=False, lidar_channels=32, bev_h=20, bev_w=20,
                                    bev_channels=8, lidar_start_size=10)
    assert fusion.lidar_start_size(f20) == 10
    del f20.lidar_start_size                       # the reference's class has no such attribute: recovered from the shapes
    assert fusion.lidar_start_size(f20) == 10
    f20.train()
    out = f20(lidar_features=torch.randn(2, 32))
    assert tuple(out.shape) == (2, 8, 20, 20)


def test_reference_metrics_consumer_accepts_the_shim_format():
    """With the reference checkout present (build container only): utils_v2.compute_metrics takes the list-of-dicts the
    shim's decode returns — torch tensors with the reference's dtypes — and agrees with the numpy restatement."""
    ref_src = Path("/root/reference/src")
    if not (ref_src / "utils_v2.py").exists():
        pytest.skip("reference checkout not present (GPU box)")
    sys.path.insert(0, str(ref_src))
    try:
        import utils_v2
    finally:
        sys.path.remove(str(ref_src))
    from oracle import bev_oracle as orc

    dets = orc.decode(syn.head_maps(501, 3), score_thresh=0.3, max_detections=100, voxel_size_m=0.512)
    gts = syn.ground_truth_near(801, dets)
    shim_format = [{k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in d.items()} for d in dets]
    assert shim_format[0]["labels"].dtype == torch.int64 and shim_format[0]["boxes"].dtype == torch.float32
    got, want = utils_v2.compute_metrics(shim_format, gts), orc.compute_metrics(dets, gts)
    assert abs(got["mAP"] - want["mAP"]) < 1e-12 and abs(got["NDS"] - want["NDS"]) < 1e-6


def _write_radar_pcd(path, n, seed):
    """A nuScenes-style radar sweep: the 18 fields with their mixed types, DATA binary."""
    fields = "x y z dyn_prop id rcs vx vy vx_comp vy_comp is_quality_valid ambig_state x_rms y_rms invalid_state pdh0 vx_rms vy_rms".split()
    types = "F F F I I F F F F F I I I I I I I I".split()
    sizes = "4 4 4 1 2 4 4 4 4 4 1 1 1 1 1 1 1 1".split()
    np_t = {("F", "4"): "<f4", ("I", "1"): "<i1", ("I", "2"): "<i2"}
    dt = np.dtype([(f, np_t[(t, s)]) for f, t, s in zip(fields, types, sizes)])
    g = np.random.default_rng(seed)
    rec = np.zeros(n, dtype=dt)
    for f in fields:
        rec[f] = g.uniform(-50, 50, n).astype(rec[f].dtype) if dt[f].kind == "f" else g.integers(0, 7, n).astype(rec[f].dtype)
    head = (f"# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS {' '.join(fields)}\nSIZE {' '.join(sizes)}\n"
            f"TYPE {' '.join(types)}\nCOUNT {' '.join(['1'] * 18)}\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA binary\n")
    Path(path).write_bytes(head.encode() + rec.tobytes())
    return rec


def test_radar_sweeps_are_read_from_the_files_and_calibration_rides_in_the_batch(tmp_path):
    """N3 / N4: the radar returns come from the .pcd files the converter recorded (the reference feeds randn,
    src/train_detect.py:171-177) and the per-sample calibration survives the collate (the reference drops it, :197-242)."""
    from bevfusion_multimodal_3d_object_detection_b200 import dataset

    info = {"radars": {}, "cams": {}, "lidar_calibrated_sensor": {"translation": [0.9, 0.0, 1.8], "rotation": [1.0, 0.0, 0.0, 0.0]}}
    recs = {}
    for i, name in enumerate(dataset.RADAR_ORDER):
        rel = f"samples/{name}/sweep{i}.pcd"
        (tmp_path / "samples" / name).mkdir(parents=True)
        recs[name] = _write_radar_pcd(tmp_path / rel, (40, 0, 125, 300, 7)[i], 50 + i)
        info["radars"][name] = {"filename": rel}
    K, E = syn.camera_rig()
    for c, cam in enumerate(dataset.CAMERA_ORDER):
        info["cams"][cam] = {"calibrated_sensor": {"translation": [0.1 * c, 0.2, 1.5], "rotation": [1.0, 0.0, 0.0, 0.0],
                                                   "camera_intrinsic": K[c].tolist()}}
    sweep = dataset.read_radar_pcd(tmp_path / info["radars"]["RADAR_FRONT"]["filename"])
    assert sweep.shape == (40, 7) and sweep.dtype == np.float32
    for j, f in enumerate(dataset.RADAR_FIELDS):
        np.testing.assert_array_equal(sweep[:, j], recs["RADAR_FRONT"][f].astype(np.float32))
    radars = dataset.load_radar_points(info, tmp_path, rng=np.random.default_rng(1))
    assert [tuple(r.shape) for r in radars] == [(125, 7)] * 5
    assert not bool(radars[0][40:].any()) and not bool(radars[1].any())                    # zero padding, an empty sweep
    rows300 = {tuple(r) for r in dataset.read_radar_pcd(tmp_path / info["radars"]["RADAR_BACK_LEFT"]["filename"]).tolist()}
    assert all(tuple(r) in rows300 for r in radars[3].tolist())                            # 125 of the 300 returns

    def item(n_obj, token):
        it = {"camera_imgs": torch.zeros(6, 3, 4, 4), "lidar_points": torch.zeros(10, 4), "radar_points": radars,
              "gt_boxes": torch.ones(n_obj, 7), "gt_labels": torch.zeros(n_obj, dtype=torch.long), "gt_velocities": torch.ones(n_obj, 2),
              "token": token}
        return dataset.attach_calibration(it, info)

    batch = dataset.collate_with_calibration([item(3, "a"), item(1, "b")])
    assert tuple(batch["intrinsics"].shape) == (2, 6, 3, 3) and tuple(batch["lidar2cam"].shape) == (2, 6, 3, 4)
    assert len(batch["radar_points"]) == 5 and tuple(batch["radar_points"][0].shape) == (2, 125, 7)
    assert tuple(batch["gt_boxes"].shape) == (2, 3, 7) and batch["gt_labels"][1].tolist() == [0, -1, -1] and batch["tokens"] == ["a", "b"]
    Kb, Eb = dataset.calibration_from_info(info)
    np.testing.assert_array_equal(batch["intrinsics"][1].numpy(), Kb)
    plain = dataset.collate_with_calibration([{k: v for k, v in item(2, "c").items() if k not in ("intrinsics", "lidar2cam")}])
    assert "intrinsics" not in plain
    # the reference's own collate gives the same tensors for the keys it knows (build container only)
    ref_src = Path("/root/reference/src")
    if (ref_src / "train_detect.py").exists():
        code = ("import sys, io, contextlib; sys.path.insert(0, %r)\n"
                "with contextlib.redirect_stdout(io.StringIO()):\n    import train_detect\nprint('COLLATE-IMPORT-OK')") % str(ref_src)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
        if "COLLATE-IMPORT-OK" in r.stdout:
            sys.path.insert(0, str(ref_src))
            try:
                import contextlib
                import io
                with contextlib.redirect_stdout(io.StringIO()):
                    import train_detect
                want = train_detect.collate_fn([item(3, "a"), item(1, "b")])
                for k, v in want.items():
                    if isinstance(v, torch.Tensor):
                        assert torch.equal(batch[k], v), k
            finally:
                sys.path.remove(str(ref_src))
                for m in ("train_detect", "fusion", "encoders", "centernet_target", "utils_v2"):
                    sys.modules.pop(m, None)


def test_constant_image_conv_stack_has_border_classes_only():
    """The radar branch's shortcut (fusion.radar_branch): k padded 3x3 convolutions of a spatially constant image give
    out[y][x] = small[cls(y)][cls(x)] with the stack run on a (2k+1)^2 image — checked here with torch's own convolutions."""
    torch.manual_seed(3)
    seq = torch.nn.Sequential(torch.nn.Conv2d(6, 6, 3, padding=1), torch.nn.BatchNorm2d(6), torch.nn.ReLU(),
                              torch.nn.Conv2d(6, 4, 3, padding=1), torch.nn.BatchNorm2d(4), torch.nn.ReLU()).eval()
    assert fusion._const_image_size(seq, 11, 9) == 5 and fusion._const_image_size(seq, 4, 9) == 0
    assert fusion._const_image_size(torch.nn.Sequential(torch.nn.Conv2d(6, 6, 5, padding=2)), 50, 50) == 0
    r = torch.randn(3, 6)
    with torch.no_grad():
        full = seq(r.view(3, 6, 1, 1).expand(3, 6, 11, 9))
        small = seq(r.view(3, 6, 1, 1).expand(3, 6, 5, 5))
    iy, ix = ops.border_class_index(11, 5), ops.border_class_index(9, 5)
    assert iy == [0, 1, 2, 2, 2, 2, 2, 2, 2, 3, 4] and ix == [0, 1, 2, 2, 2, 2, 2, 3, 4]
    spread = small[:, :, iy][:, :, :, ix]
    assert torch.allclose(spread, full, rtol=0, atol=1e-6)
    assert not torch.allclose(full[:, :, 0], full[:, :, 5])            # the borders do differ from the interior
