"""GPU (-m gpu): the drop-in layer itself — the mirror modules and functions with the reference's signatures
(encoders.py, fusion.py, centernet_decode.py, the same code patch() installs on the reference's classes) —
in eval mode on CUDA, against (1) the outputs the REFERENCE's modules produced for the same state_dict and
inputs (tests/golden) and (2) the plain torch graph of the same module on the same device.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import bevfusion_multimodal_3d_object_detection_b200 as b200bev
from bevfusion_multimodal_3d_object_detection_b200 import encoders as enc_mod
from bevfusion_multimodal_3d_object_detection_b200 import synthetic as syn
from oracle import bev_oracle as orc
from tests.conftest import max_rel

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-5
BF16_TOL = 1e-2


@pytest.fixture(autouse=True)
def _exact_fp32_torch():
    """The torch side of every comparison must be real fp32 (no TF32 in cuDNN / cuBLAS)."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def load_mlp(module, layers):
    """Synthetic layer dicts -> the reference's state_dict names (conv{i}.*, bn{i}.*), as make_golden.py does."""
    sd = module.state_dict()
    for i, lay in enumerate(layers, start=1):
        sd[f"conv{i}.weight"] = torch.from_numpy(lay["weight"]).unsqueeze(-1)
        sd[f"conv{i}.bias"] = torch.from_numpy(lay["bias"])
        sd[f"bn{i}.weight"] = torch.from_numpy(lay["bn_weight"])
        sd[f"bn{i}.bias"] = torch.from_numpy(lay["bn_bias"])
        sd[f"bn{i}.running_mean"] = torch.from_numpy(lay["bn_mean"])
        sd[f"bn{i}.running_var"] = torch.from_numpy(lay["bn_var"])
    module.load_state_dict(sd)
    return module.eval()


def test_lidar_encoder_module_reproduces_the_reference_outputs(cuda, golden):
    g = golden("lidar_encoder")
    enc = load_mlp(b200bev.PointNetLiDAREncoder(input_channels=4, feat_dim=1024, use_bn=True),
                   syn.mlp_weights(101, syn.LIDAR_DIMS)).to(cuda)
    pts = torch.from_numpy(syn.lidar_batch(201, 2, n_valid=1900, n_total=2011)).to(cuda)
    with torch.no_grad():
        out = enc(pts)
        assert out.is_cuda and tuple(out.shape) == (2, 1024)
        assert max_rel(out.cpu().numpy(), g["small_global"]) < FP32_TOL                      # vs the reference itself
        assert torch.equal(enc(pts.transpose(1, 2).contiguous()), out)                       # (B,C,N) layout, src/encoders.py:282
        torch_graph = torch.max(enc_mod._torch_mlp(enc, pts.transpose(1, 2)), 2)[0]          # same module, plain ATen ops
        assert max_rel(out.cpu().numpy(), torch_graph.cpu().numpy()) < FP32_TOL
        full = torch.from_numpy(syn.lidar_batch(301, 1)).to(cuda)
        assert max_rel(enc(full).cpu().numpy(), g["full_global"]) < FP32_TOL
        enc.b200_precision = "bf16"
        assert max_rel(enc(full).cpu().numpy(), g["full_global"]) < BF16_TOL                 # tcgen05 path behind the same call
        enc.b200_precision = None
        canvas = enc.cell_canvas(pts, (50, 50))                                              # north_star S1, extended mode
        assert tuple(canvas.shape) == (2, 1024, 50, 50)
        sub = canvas.permute(0, 2, 3, 1).reshape(2, 2500, 1024)[:, :, ::16]
        assert max_rel(sub.cpu().numpy(), g["small_canvas_sub"]) < FP32_TOL


def test_lidar_encoder_follows_state_dict_updates_and_training_mode(cuda):
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    enc = load_mlp(b200bev.PointNetLiDAREncoder(input_channels=4), layers).to(cuda)
    pts = torch.from_numpy(syn.lidar_batch(205, 2, n_valid=500, n_total=512)).to(cuda)
    with torch.no_grad():
        a = enc(pts)
        other = syn.mlp_weights(177, syn.LIDAR_DIMS)
        load_mlp(enc, other)                                   # load_state_dict -> the folded-weight cache must be rebuilt
        b = enc(pts)
        assert max_rel(b.cpu().numpy(), orc.pointnet_global(pts.cpu().numpy(), other)) < FP32_TOL
        assert not torch.allclose(a, b)
    enc.train()                                                # training: torch graph with batch statistics + autograd
    out = enc(pts.requires_grad_(True))
    out.sum().backward()
    assert pts.grad is not None and out.requires_grad
    with pytest.raises(RuntimeError):
        enc.eval()(pts.detach().cpu())                         # no CPU path


@pytest.mark.parametrize("method", ["concat", "max", "mean"])
def test_multi_radar_module_reproduces_the_reference_outputs(cuda, golden, method):
    g = golden("radar_encoder")
    m = b200bev.MultiRadarEncoder(input_channels=7, feat_dim=256, num_radars=5, fusion_method=method)
    load_mlp(m.radar_encoder, syn.mlp_weights(111, syn.RADAR_DIMS))
    if method == "concat":
        fcw, fcb = syn.linear_weights(112, 5 * 256, 256)
        with torch.no_grad():
            m.fusion_fc.weight.copy_(torch.from_numpy(fcw))
            m.fusion_fc.bias.copy_(torch.from_numpy(fcb))
    m = m.eval().to(cuda)
    radars = [torch.from_numpy(r).to(cuda) for r in syn.radar_batch(211, 3)]
    with torch.no_grad():
        assert max_rel(m(radars).cpu().numpy(), g[f"fused_{method}"]) < FP32_TOL
        ragged = [r[:, : 125 - 17 * i].contiguous() for i, r in enumerate(radars)]
        assert max_rel(m(ragged).cpu().numpy(), g[f"ragged_{method}"]) < FP32_TOL
    m.fusion_method = "bogus"
    with pytest.raises(ValueError, match="Unknown fusion method"):
        m(radars)


@pytest.mark.parametrize("name,shape", [("ref28x50", (16, 28, 50, 50, 50)), ("hd57x100", (8, 57, 100, 50, 50)),
                                        ("up7x9", (8, 7, 9, 20, 30))])
def test_fusion_module_camera_branch_vs_the_same_module_in_torch(cuda, golden, name, shape):
    """FlexibleBEVFusion.forward in eval mode on CUDA: mean and resize on the kernels, convs on cuDNN, against the
    module's own torch graph (the training-mode branch evaluated with eval-mode BatchNorm)."""
    C, h, w, H, W = shape
    torch.manual_seed(11)
    fus = b200bev.FlexibleBEVFusion(use_camera=True, use_lidar=True, use_radar=True, camera_channels=C, lidar_channels=64,
                                    radar_channels=32, bev_h=H, bev_w=W, bev_channels=8)
    if (H, W) != (50, 50):
        fus = b200bev.FlexibleBEVFusion(use_camera=True, use_lidar=False, use_radar=True, camera_channels=C,
                                        radar_channels=32, bev_h=H, bev_w=W, bev_channels=8)   # lidar BEV is fixed at 50x50 (src/fusion.py:141)
    for mod in fus.modules():                                   # BatchNorm statistics away from the identity
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.3)
            mod.running_var.uniform_(0.5, 2.0)
    fus = fus.eval().to(cuda)
    feats = torch.from_numpy(syn.camera_features(401, 2, n_cam=6, channels=C, h=h, w=w)).to(cuda)
    lidar = torch.randn(2, 64, device=cuda) if fus.use_lidar else None
    radar = torch.randn(2, 32, device=cuda)
    with torch.no_grad():
        got = fus(camera_features=feats, lidar_features=lidar, radar_features=radar)
        # the same module through plain torch ops
        cam = F.interpolate(fus.camera_proj(feats.mean(dim=1)), size=(H, W), mode="bilinear", align_corners=False)
        parts = [cam]
        if fus.use_lidar:
            parts.append(fus.lidar_upsample(fus.lidar_init(lidar).view(2, 128, 25, 25)))
        r = fus.radar_proj(radar).view(2, 8, 1, 1).expand(2, 8, H, W)
        parts.append(fus.radar_refine(r))
        ref = fus.bev_fusion(torch.cat(parts, dim=1))
        assert tuple(got.shape) == (2, 8, H, W)
        assert max_rel(got.cpu().numpy(), ref.cpu().numpy()) < FP32_TOL
        # the two kernel steps against what the reference's module fed / produced around F.interpolate
        g = golden("camera_bev")
        mean = b200bev.ops.camera_mean(feats).cpu().numpy()
        assert max_rel(mean, g[f"{name}_mean"]) < 2e-7
        if name != "up7x9":          # ATen's scalar path for tiny planes associates differently (1 ulp in <1 % of elements)
            assert np.array_equal(mean, g[f"{name}_mean"])
        rs = b200bev.ops.bilinear_resize(torch.from_numpy(g[f"{name}_resize_in"]).to(cuda), (H, W))
        assert max_rel(rs.cpu().numpy(), g[f"{name}_resize_out"]) < FP32_TOL
        with pytest.raises(ValueError, match="No modality features provided"):
            fus()
        # 4-D camera input (already averaged) is accepted, src/fusion.py:233-236
        got4 = fus(camera_features=feats.mean(dim=1), lidar_features=lidar, radar_features=radar)
        assert max_rel(got4.cpu().numpy(), ref.cpu().numpy()) < FP32_TOL


@pytest.mark.parametrize("tag,fn,voxel", [("ct", "decode_centernet_predictions", 2.048),
                                          ("fd", "decode_centernet_predictions_fusion_detection", 0.512)])
def test_decode_function_returns_the_reference_lists(cuda, golden, tag, fn, voxel):
    """The list-of-dicts interface (data-dependent lengths, dtypes, the CPU tensors of an empty sample)."""
    g = golden("centernet_decode")
    maps = {k: torch.from_numpy(v).to(cuda) for k, v in syn.head_maps(501, 3).items()}
    decode = getattr(b200bev, fn)
    for thr in (0.0, 0.3, 0.999):
        dets = decode(maps, score_thresh=thr, max_detections=100)
        assert len(dets) == 3
        for b, d in enumerate(dets):
            ref_scores = g[f"{tag}_thr{thr}_b{b}_scores"]
            assert d["labels"].dtype == torch.int64 and d["boxes"].dtype == torch.float32
            np.testing.assert_array_equal(d["scores"].cpu().numpy(), ref_scores)
            np.testing.assert_allclose(d["boxes"].cpu().numpy(), g[f"{tag}_thr{thr}_b{b}_boxes"], rtol=0, atol=1e-5)
            np.testing.assert_array_equal(d["labels"].cpu().numpy(), g[f"{tag}_thr{thr}_b{b}_labels"])
            np.testing.assert_array_equal(d["velocities"].cpu().numpy(), g[f"{tag}_thr{thr}_b{b}_velocities"])
    empty = decode(maps, score_thresh=2.0, max_detections=100)           # nothing passes: CPU tensors
    assert all(d["boxes"].device.type == "cpu" and tuple(d["boxes"].shape) == (0, 7) and d["labels"].dtype == torch.int64
               for d in empty)
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        decode({k: v[:, :, :3, :3].contiguous() for k, v in maps.items()}, max_detections=100)   # K > H*W (Q7)


def test_encode_fuse_head_decode_chain_on_device(cuda):
    """The whole inference chain the pipelines run (src/fusion.py:1090-1137 + decode), mirror modules with kernels
    against the same modules through torch ops: identical boxes."""
    torch.manual_seed(5)
    lidar_enc = load_mlp(b200bev.PointNetLiDAREncoder(input_channels=4), syn.mlp_weights(101, syn.LIDAR_DIMS)).to(cuda)
    radar_enc = b200bev.MultiRadarEncoder(input_channels=7, feat_dim=256, num_radars=5, fusion_method="concat")
    load_mlp(radar_enc.radar_encoder, syn.mlp_weights(111, syn.RADAR_DIMS))
    radar_enc = radar_enc.eval().to(cuda)
    fus = b200bev.FlexibleBEVFusion(use_camera=True, use_lidar=True, use_radar=True, camera_channels=32, bev_h=50, bev_w=50,
                                    bev_channels=16).eval().to(cuda)
    head = torch.nn.ModuleDict({k: torch.nn.Sequential(torch.nn.Conv2d(16, 16, 3, padding=1), torch.nn.ReLU(),
                                                       torch.nn.Conv2d(16, c, 1))
                                for k, c in (("heatmap", 10), ("offset", 2), ("size", 3), ("rot", 2), ("vel", 2))}).eval().to(cuda)
    B = 2
    pts = torch.from_numpy(syn.lidar_batch(901, B, n_valid=3000, n_total=3072)).to(cuda)
    radars = [torch.from_numpy(r).to(cuda) for r in syn.radar_batch(902, B)]
    feats = torch.from_numpy(syn.camera_features(903, B, channels=32, h=28, w=50)).to(cuda)

    def run(kernels: bool):
        with torch.no_grad():
            if kernels:
                lf, rf = lidar_enc(pts), radar_enc(radars)
                bev = fus(camera_features=feats, lidar_features=lf, radar_features=rf)
            else:
                lf = torch.max(enc_mod._torch_mlp(lidar_enc, pts.transpose(1, 2)), 2)[0]
                per = torch.stack([torch.max(enc_mod._torch_mlp(radar_enc.radar_encoder, r.transpose(1, 2)), 2)[0] for r in radars], 1)
                rf = radar_enc.fusion_fc(per.view(B, -1))
                cam = F.interpolate(fus.camera_proj(feats.mean(dim=1)), size=(50, 50), mode="bilinear", align_corners=False)
                li = fus.lidar_upsample(fus.lidar_init(lf).view(B, 128, 25, 25))
                ra = fus.radar_refine(fus.radar_proj(rf).view(B, 16, 1, 1).expand(B, 16, 50, 50))
                bev = fus.bev_fusion(torch.cat([cam, li, ra], dim=1))
            pred = {k: m(bev) for k, m in head.items()}
            pred["heatmap"] = torch.sigmoid(pred["heatmap"])                 # src/fusion.py:871
            return pred

    pk, pt = run(True), run(False)
    for k in pk:
        assert max_rel(pk[k].cpu().numpy(), pt[k].cpu().numpy()) < 1e-4, k   # conv stacks amplify the 1e-5 of the inputs a little
    dets = b200bev.decode_centernet_predictions(pk, score_thresh=0.0, max_detections=50)
    ref = orc.decode({k: v.cpu().numpy() for k, v in pk.items()}, score_thresh=0.0, max_detections=50)
    for d, r in zip(dets, ref):
        np.testing.assert_array_equal(d["scores"].cpu().numpy(), r["scores"])
        np.testing.assert_allclose(d["boxes"].cpu().numpy(), r["boxes"], rtol=0, atol=1e-5)


# ------------------------------------------------------------------------------------------------ SURVEY 8f N1 / N2
def _conv_fusion(cuda):
    fus = b200bev.FlexibleBEVFusion(use_camera=True, use_lidar=False, use_radar=False, camera_channels=64, bev_h=12, bev_w=20,
                                    bev_channels=64)
    sd = syn.fill_state_dict(731, {k: tuple(v.shape) for k, v in fus.state_dict().items()})
    fus.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    return fus.eval().to(cuda)


def test_fusion_conv_stacks_vs_reference_golden_fp32_and_tcgen05(cuda, golden):
    """camera_proj + resize + bev_fusion of the reference's module: the default path (kernels + fp32 cuDNN convs, 1e-5)
    and the opt-in bf16 path (every convolution on the tcgen05 kernel, 1e-2) against the reference's own output."""
    g = golden("bev_glue")
    fus = _conv_fusion(cuda)
    cam = torch.from_numpy(syn.camera_features(732, 2, n_cam=6, channels=64, h=9, w=14)).to(cuda)
    with torch.no_grad():
        out32 = fus(camera_features=cam)
        assert max_rel(out32.cpu().numpy(), g["stack_out"]) < 5 * FP32_TOL        # four conv layers deep
        fus.b200_precision = "bf16"
        out16 = fus(camera_features=cam)
        assert "_b200bev_conv_cache" in fus.bev_fusion.__dict__                   # the tcgen05 path did run
        assert out16.dtype == torch.float32 and tuple(out16.shape) == (2, 64, 12, 20)
        assert max_rel(out16.cpu().numpy(), g["stack_out"]) < BF16_TOL
        # weight updates invalidate the packed images
        fus.bev_fusion[3].weight.mul_(2.0)
        assert not torch.allclose(fus(camera_features=cam), out16)


def test_all_three_branches_on_the_tcgen05_convs(cuda):
    """camera + lidar (lidar_init kernel, upsample kernel) + radar (dense kernel, broadcast) + concat-free bev_fusion."""
    torch.manual_seed(3)
    fus = b200bev.FlexibleBEVFusion(use_camera=True, use_lidar=True, use_radar=True, camera_channels=64, lidar_channels=128,
                                    radar_channels=64, bev_h=50, bev_w=50, bev_channels=64)
    for mod in fus.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.3)
            mod.running_var.uniform_(0.5, 2.0)
    fus = fus.eval().to(cuda)
    cam = torch.from_numpy(syn.camera_features(733, 2, n_cam=6, channels=64, h=28, w=50)).to(cuda)
    lidar = torch.rand(2, 128, device=cuda)
    radar = torch.rand(2, 64, device=cuda)
    with torch.no_grad():
        fus.b200_precision = "f32_cudnn"
        ref = fus(camera_features=cam, lidar_features=lidar, radar_features=radar)      # the module's own fp32 cuDNN convolutions
        fus.b200_precision = None
        acc = fus(camera_features=cam, lidar_features=lidar, radar_features=radar)      # default: fp32 accuracy on tcgen05
        fus.b200_precision = "bf16"
        got = fus(camera_features=cam, lidar_features=lidar, radar_features=radar)
    assert max_rel(acc.cpu().numpy(), ref.cpu().numpy()) < 4 * FP32_TOL                 # four convolution layers deep
    assert max_rel(got.cpu().numpy(), ref.cpu().numpy()) < BF16_TOL


def test_centernet_head_mirror_vs_reference_golden_and_fused_path(cuda, golden):
    g = golden("bev_glue")
    # (1) the reference's own head (32 -> 16 channels: below the tensor-core kernel's granularity -> torch layers)
    head = b200bev.CenterNetHead(in_channels=32, num_classes=10, head_conv=16)
    head.load_state_dict({k: torch.from_numpy(v) for k, v in syn.head_weights(711, 32, 16, 10).items()})
    head = head.eval().to(cuda)
    x = torch.from_numpy(syn._rng(712).standard_normal((2, 32, 24, 40)).astype(np.float32)).to(cuda)
    with torch.no_grad():
        pred = head(x)
    for k in ("heatmap", "offset", "size", "rot", "vel"):
        assert max_rel(pred[k].cpu().numpy(), g[f"head_{k}"]) < FP32_TOL, k
    # (2) 64 -> 5 x 64 channels with the bf16 path on: two tcgen05 launches, sigmoid inside the decode launch
    big = b200bev.CenterNetHead(in_channels=64, num_classes=10, head_conv=64)
    # second-layer weights scaled so that the logits spread over a few units, as a trained head's do: the bf16 bound is
    # relative to max|ref| and the sigmoid pins that at <= 1 while the logit error grows with the logit range
    big.load_state_dict({k: torch.from_numpy(v) for k, v in syn.head_weights(713, 64, 64, 10, out_scale=0.3).items()})
    big = big.eval().to(cuda)
    xb = torch.from_numpy(syn._rng(714).standard_normal((3, 64, 50, 50)).astype(np.float32)).to(cuda)
    with torch.no_grad():
        big.b200_precision = "f32_cudnn"            # the module's own torch layers: the reference on this device
        ref = big(xb)
        big.b200_precision = None                   # the default: fp32 accuracy on the tensor cores
        acc = big(xb)
        big.b200_precision = "bf16"
        got = big(xb)
    from bevfusion_multimodal_3d_object_detection_b200 import conv_blocks
    for k in ("heatmap", "offset", "size", "rot", "vel"):
        assert max_rel(acc[k].cpu().numpy(), ref[k].cpu().numpy()) < 2 * FP32_TOL, k       # two layers deep
    assert set(got) == set(ref) == {"heatmap", "offset", "size", "rot", "vel"}          # the reference's five keys, nothing else
    assert conv_blocks.logits_of(got["heatmap"]) is not None and conv_blocks.logits_of(ref["heatmap"]) is None
    for k in ("heatmap", "offset", "size", "rot", "vel"):
        assert got[k].is_contiguous() and max_rel(got[k].cpu().numpy(), ref[k].cpu().numpy()) < BF16_TOL, k
    dets = b200bev.decode_centernet_predictions(got, score_thresh=0.0, max_detections=50)
    plain = {**got, "heatmap": got["heatmap"].clone()}                                  # a new tensor: no note, sigmoid already applied
    want = b200bev.decode_centernet_predictions(plain, score_thresh=0.0, max_detections=50)
    for d, w in zip(dets, want):
        assert torch.equal(d["scores"], w["scores"]) and torch.equal(d["boxes"], w["boxes"])
    # a caller that edits the heat map in place (masking, temperature, flip-TTA averaging) is decoded on what it left there
    got["heatmap"].mul_(0.5)
    assert conv_blocks.logits_of(got["heatmap"]) is None
    halved = b200bev.decode_centernet_predictions(got, score_thresh=0.0, max_detections=50)
    for d, w in zip(halved, want):
        assert torch.equal(d["scores"], w["scores"] * 0.5)


def test_decode_outputs_feed_the_metrics_consumer(cuda, golden):
    """N4: the shim's list of per-sample dicts (device tensors) goes into compute_metrics as the reference's
    own decode output does — `.cpu().numpy()` on every field, src/utils_v2.py:126-133 — and gives the reference's mAP / NDS."""
    g = golden("metrics")
    maps = {k: torch.from_numpy(v).to(cuda) for k, v in syn.head_maps(501, 3).items()}
    dets = b200bev.decode_centernet_predictions_fusion_detection(maps, score_thresh=0.3, max_detections=100)
    as_numpy = [{k: v.cpu().numpy() for k, v in d.items()} for d in dets]           # what compute_metrics does with tensors
    gts = syn.ground_truth_near(801, as_numpy)
    assert syn.digest(*[a for gt in gts for a in gt.values()]) == str(g["gt_digest"])
    m = orc.compute_metrics(as_numpy, gts)
    assert abs(m["mAP"] - float(g["mAP"])) < 1e-9 and abs(m["NDS"] - float(g["NDS"])) < 1e-6


def test_parallel_branches_of_the_fused_path_change_nothing(cuda):
    """The lidar and radar branches of the fused bf16 path run on side streams next to the camera branch
    (runtime.BranchStreams); the result is the serial one, bit for bit, call after call."""
    torch.manual_seed(5)
    fus = b200bev.FlexibleBEVFusion(use_camera=True, use_lidar=True, use_radar=True, camera_channels=64, lidar_channels=128,
                                    radar_channels=64, bev_h=50, bev_w=50, bev_channels=64).eval().to(cuda)
    fus.b200_precision = "bf16"
    with torch.no_grad():
        for rep in range(6):
            cam = torch.from_numpy(syn.camera_features(900 + rep, 2, n_cam=6, channels=64, h=28, w=50)).to(cuda)
            lidar, radar = torch.rand(2, 128, device=cuda), torch.rand(2, 64, device=cuda)
            outs = []
            for parallel in (True, False, True, True):
                fus.b200_parallel_branches = parallel
                outs.append(fus(camera_features=cam, lidar_features=lidar, radar_features=radar).clone())
            assert all(torch.equal(o, outs[1]) for o in outs), rep


def test_head_starts_from_the_channels_last_twin_of_the_fusion_output(cuda):
    """bf16 path: bev_fusion's last convolution also writes its result channels-last bf16 and the returned tensor remembers it;
    the head then skips its layout pass.  Same outputs as from a copy of the tensor (no note), and an edited tensor loses the note."""
    from bevfusion_multimodal_3d_object_detection_b200 import conv_blocks

    torch.manual_seed(11)
    fus = b200bev.FlexibleBEVFusion(use_camera=True, use_lidar=True, use_radar=False, camera_channels=64, lidar_channels=128,
                                    bev_h=50, bev_w=50, bev_channels=64).eval().to(cuda)
    head = b200bev.CenterNetHead(in_channels=64, num_classes=10, head_conv=64)
    head.load_state_dict({k: torch.from_numpy(v) for k, v in syn.head_weights(717, 64, 64, 10, out_scale=0.3).items()})
    head = head.eval().to(cuda)
    fus.b200_precision = head.b200_precision = "bf16"
    cam = torch.from_numpy(syn.camera_features(735, 2, n_cam=6, channels=64, h=28, w=50)).to(cuda)
    with torch.no_grad():
        bev = fus(camera_features=cam, lidar_features=torch.rand(2, 128, device=cuda))
        twin = conv_blocks.nhwc_of(bev)
        assert twin is not None and torch.equal(twin, bev.permute(0, 2, 3, 1).to(torch.bfloat16))
        with_note = head(bev)
        without = head(bev.clone())
        for k in with_note:
            assert torch.equal(with_note[k], without[k]), k
        bev.mul_(2.0)                                    # edited in place: the note no longer describes the tensor
        assert conv_blocks.nhwc_of(bev) is None


def test_graphed_step_replays_the_chain_bit_for_bit(cuda):
    """runtime.GraphedStep: fusion -> head -> fixed-size decode captured in one CUDA graph; new inputs are copied into
    the captured tensors, a replay gives what the eager calls give."""
    from bevfusion_multimodal_3d_object_detection_b200 import conv_blocks, ops, runtime

    torch.manual_seed(9)
    fus = b200bev.FlexibleBEVFusion(use_camera=True, use_lidar=True, use_radar=True, camera_channels=64, lidar_channels=128,
                                    radar_channels=64, bev_h=50, bev_w=50, bev_channels=64).eval().to(cuda)
    head = b200bev.CenterNetHead(in_channels=64, num_classes=10, head_conv=64)
    head.load_state_dict({k: torch.from_numpy(v) for k, v in syn.head_weights(713, 64, 64, 10, out_scale=0.3).items()})
    head = head.eval().to(cuda)
    fus.b200_precision = head.b200_precision = "bf16"
    cam = torch.from_numpy(syn.camera_features(733, 2, n_cam=6, channels=64, h=28, w=50)).to(cuda)
    lidar, radar = torch.rand(2, 128, device=cuda), torch.rand(2, 64, device=cuda)

    def step():
        pred = head(fus(camera_features=cam, lidar_features=lidar, radar_features=radar))
        return ops.centernet_decode(conv_blocks.logits_of(pred["heatmap"]), pred["offset"], pred["size"], pred["rot"], pred["vel"], 40, 2.048,
                                    heat_is_logit=True)

    graphed = runtime.GraphedStep(step, cuda)
    for seed in (1, 2):
        cam.copy_(torch.from_numpy(syn.camera_features(740 + seed, 2, n_cam=6, channels=64, h=28, w=50)).to(cuda))
        lidar.copy_(torch.rand(2, 128, device=cuda))
        with torch.no_grad():
            eager = {k: v.clone() for k, v in step().items()}
        out = graphed.replay()
        torch.cuda.synchronize()
        for k in ("scores", "boxes", "velocities", "ys", "xs", "count"):
            assert torch.equal(out[k], eager[k]), k


# ------------------------------------------------------------------------------------------------ the whole chain
def _chain(cuda, precision):
    chain = b200bev.BEVDetectorChain(precision=precision)
    sd = syn.detector_state(syn.CHAIN_SEED, chain.state_shapes())
    chain.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return chain.eval().to(cuda)


@pytest.mark.parametrize("precision,tol", [("f32", FP32_TOL), ("bf16", BF16_TOL)])
def test_detector_chain_reproduces_the_reference_end_to_end(cuda, golden, precision, tol):
    """Encoders -> FlexibleBEVFusion -> CenterNetHead -> decode on the kernels, with the reference's state_dict names,
    against what the REFERENCE's own modules produced for the same state and inputs (tests/golden/detector_chain.npz,
    src/fusion.py:1113-1137 + src/eval.py:58-62)."""
    g = golden("detector_chain")
    chain = _chain(cuda, precision)
    lidar, radars, cam = syn.chain_inputs()
    assert syn.digest(lidar, *radars, cam) == str(g["input_digest"])
    args = (torch.from_numpy(cam).to(cuda), torch.from_numpy(lidar).to(cuda), [torch.from_numpy(r).to(cuda) for r in radars])
    with torch.no_grad():
        lf = chain.lidar_encoder(args[1])
        rf = chain.radar_encoder(args[2])
        bev = chain.fusion(camera_features=args[0], lidar_features=lf, radar_features=rf)
        pred = chain(*args)
        dets = chain.detect(*args, score_thresh=0.0, max_detections=100)
        fixed = chain.detect_fixed(*args, score_thresh=0.0, max_detections=100)
    assert max_rel(lf.cpu().numpy(), g["lidar_feat"]) < tol
    assert max_rel(rf.cpu().numpy(), g["radar_feat"]) < FP32_TOL                # the radar MLP is fp32 in both settings
    # The per-kernel bounds (1e-5 fp32, 1e-2 bf16 of max|ref|, every kernel against its own fp32 inputs) are in
    # test_gpu_parity.py.  This is the whole chain, 12 layers deep: roundings compound, so the bound on the fused BEV
    # features is 4x (fp32) / 2x (bf16) the per-kernel one and on the head outputs — three more bf16 layers and a sigmoid
    # whose output is pinned below 1 while the logit error grows with the logit range — 4x / 5x.  Measured on B200: fp32
    # 2e-6 .. 1e-5, bf16 bev 1.1e-2, heat map 3.2e-2.
    deep_bev, deep_head = (4 * tol, 4 * tol) if precision == "f32" else (2 * tol, 5 * tol)
    assert max_rel(bev[:, ::8].cpu().numpy(), g["bev_sub"]) < deep_bev
    for k in ("heatmap", "offset", "size", "rot", "vel"):
        assert max_rel(pred[k].cpu().numpy(), g["pred_" + k]) < deep_head, k
    assert fixed["count"].tolist() == [len(d["scores"]) for d in dets]
    if precision == "f32":
        for b, d in enumerate(dets):
            n = len(g[f"det_b{b}_scores"])
            assert len(d["scores"]) == n
            np.testing.assert_allclose(d["scores"].cpu().numpy(), g[f"det_b{b}_scores"], rtol=0, atol=2e-5)
            # same winners in the same order wherever the reference's neighbouring scores are further apart than the tolerance
            gap = np.abs(np.diff(g[f"det_b{b}_scores"]))
            stable = np.concatenate([[True], gap > 1e-4]) & np.concatenate([gap > 1e-4, [True]])
            np.testing.assert_allclose(d["boxes"].cpu().numpy()[stable], g[f"det_b{b}_boxes"][stable], rtol=0, atol=2e-3)
            assert d["labels"].dtype == torch.int64 and not bool(d["labels"].any())


def test_eval_mode_under_autograd_stays_differentiable(cuda):
    """Eval mode with autograd recording (frozen-backbone fine-tuning, saliency): the kernels return tensors without a
    grad_fn, so such calls take the modules' torch graph on the GPU — as the reference's eval-mode modules behave — and
    the same call under no_grad runs the kernels and agrees."""
    chain = b200bev.BEVDetectorChain(camera_channels=64, bev_channels=64)
    sd = syn.detector_state(77, chain.state_shapes(), head_in=64)
    chain.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    chain = chain.eval().to(cuda)
    lidar, radars, _ = syn.chain_inputs(frames=1, points=700)
    cam = torch.from_numpy(syn.camera_features(5, 1, channels=64, h=12, w=20)).to(cuda)
    pts = torch.from_numpy(lidar).to(cuda).requires_grad_(True)
    rad = [torch.from_numpy(r).to(cuda) for r in radars]
    pred = chain(cam, pts, rad)                                  # grad mode is on, parameters and `pts` require grad
    assert pred["heatmap"].requires_grad
    pred["heatmap"].sum().backward()
    assert pts.grad is not None and float(pts.grad.abs().sum()) > 0
    assert chain.fusion.bev_fusion[0].weight.grad is not None
    with torch.no_grad():
        fast = chain(cam, pts, rad)
    assert not fast["heatmap"].requires_grad
    for k in pred:
        assert max_rel(fast[k].cpu().numpy(), pred[k].detach().cpu().numpy()) < 4 * FP32_TOL, k
    # parameters frozen and inputs without grad: nothing to differentiate -> kernels even with grad mode on
    for p in chain.parameters():
        p.requires_grad_(False)
    b200bev.invalidate_cache(chain)
    frozen = chain(cam, pts.detach(), rad)
    assert not frozen["heatmap"].requires_grad and torch.equal(frozen["heatmap"], fast["heatmap"])


def test_lidar_init_on_the_tensor_cores_follows_the_module(cuda):
    """`lidar_init` of the fusion module: the second layer runs from its split-fp16 image on the tensor cores (every batch
    size), equals the module's own torch layers at the fp32 bound, follows weight updates, and `b200_dense_tc = False` keeps
    the FFMA kernels on torch's own weight."""
    from bevfusion_multimodal_3d_object_detection_b200 import _lib
    from bevfusion_multimodal_3d_object_detection_b200.fusion import lidar_init_dense
    torch.manual_seed(5)
    fus = b200bev.FlexibleBEVFusion(use_camera=False, use_lidar=True, use_radar=False, lidar_channels=1024, bev_h=50, bev_w=50,
                                    bev_channels=256).eval().to(cuda)
    _lib.enable_call_counting()
    with torch.no_grad():
        for B in (1, 8, 32, 70):
            x = torch.rand(B, 1024, device=cuda)
            ref = fus.lidar_init(x)
            _lib.reset_call_counts()
            got = lidar_init_dense(fus, x)
            assert _lib._call_counts.get("b200bev_lidar_init_split", 0) == 1 and "b200bev_lidar_init" not in _lib._call_counts
            assert float((got - ref).abs().max()) < FP32_TOL * float(ref.abs().max())
        x = torch.rand(32, 1024, device=cuda)
        before = lidar_init_dense(fus, x)
        fus.lidar_init[2].weight.mul_(2.0)                      # an in-place update: the image is rebuilt
        fus.lidar_init[2].bias.zero_()
        after = lidar_init_dense(fus, x)
        ref = fus.lidar_init(x)
        assert float((after - ref).abs().max()) < FP32_TOL * float(ref.abs().max())
        assert float((after - before).abs().max()) > 0.1 * float(ref.abs().max())
        fus.b200_dense_tc = False
        _lib.reset_call_counts()
        ffma = lidar_init_dense(fus, x)
        assert _lib._call_counts.get("b200bev_lidar_init", 0) == 1 and "b200bev_lidar_init_split" not in _lib._call_counts
        assert float((ffma - ref).abs().max()) < FP32_TOL * float(ref.abs().max())


def test_cache_invalidation_after_data_writes(cuda):
    """`.data` writes do not bump a tensor's version counter: `invalidate_cache` is the documented call after them."""
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    enc = load_mlp(b200bev.PointNetLiDAREncoder(input_channels=4), layers).to(cuda)
    pts = torch.from_numpy(syn.lidar_batch(205, 1, n_valid=500, n_total=512)).to(cuda)
    with torch.no_grad():
        a = enc(pts)
        enc.bn5.bias.data.add_(0.75)                            # invisible to the version counter
        b200bev.invalidate_cache(enc)
        b = enc(pts)
        assert bool((b >= a).all()) and float((b - a).max()) > 0.7     # the maxima moved up by the bias shift (0.75, before the ReLU)
        with torch.no_grad():
            enc.bn5.bias.add_(-0.75)                            # an ordinary in-place op IS seen
        assert torch.allclose(enc(pts), a, atol=1e-5)


@pytest.mark.parametrize("H,W,s,C", [(50, 50, 5, 256), (7, 9, 5, 24), (100, 100, 3, 64), (5, 5, 5, 8)])
def test_border_expand_kernel(cuda, H, W, s, C):
    """out[y][x] = small[cls(y)][cls(x)] — both output layouts, against torch indexing (bit-exact: it is a copy)."""
    from bevfusion_multimodal_3d_object_detection_b200 import ops

    small = torch.randn(3, C, s, s, device=cuda)
    iy = torch.tensor(ops.border_class_index(H, s), device=cuda)
    ix = torch.tensor(ops.border_class_index(W, s), device=cuda)
    want = small[:, :, iy][:, :, :, ix]
    assert torch.equal(ops.border_expand(small, (H, W)), want)
    if C % 8 == 0:
        cat = torch.zeros(3, H, W, C + 16, dtype=torch.bfloat16, device=cuda)
        assert ops.border_expand(small, (H, W), out_nhwc=cat, c_offset=8, want_nchw=False) is None
        assert torch.equal(cat[..., 8:8 + C], want.permute(0, 2, 3, 1).to(torch.bfloat16))
        assert not bool(cat[..., :8].any()) and not bool(cat[..., 8 + C:].any())


@pytest.mark.parametrize("precision,tol", [("f32", FP32_TOL), ("bf16", BF16_TOL)])
def test_radar_branch_shortcut_equals_the_full_size_stack(cuda, precision, tol):
    """fusion.radar_branch (5 x 5 image + border classes) against the module's own layers on the broadcast (B,C,H,W) image."""
    from bevfusion_multimodal_3d_object_detection_b200 import fusion as fus_mod

    fus = b200bev.FlexibleBEVFusion(use_camera=False, use_lidar=False, use_radar=True, radar_channels=64, bev_h=50, bev_w=40,
                                    bev_channels=64)
    sd = syn.fill_state_dict(61, {k: tuple(v.shape) for k, v in fus.state_dict().items()})
    fus.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    fus = fus.eval().to(cuda)
    fus.b200_precision = precision
    feat = torch.from_numpy(syn.global_features(62, 3, 64)).to(cuda)
    with torch.no_grad():
        got = fus_mod.radar_branch(fus, feat)
        r = fus.radar_proj(feat).view(3, 64, 1, 1).expand(3, 64, 50, 40)
        want = fus.radar_refine(r)
        assert max_rel(got.cpu().numpy(), want.cpu().numpy()) < tol
        whole = fus(radar_features=feat)
        assert max_rel(whole.cpu().numpy(), fus.bev_fusion(want).cpu().numpy()) < 2 * tol
