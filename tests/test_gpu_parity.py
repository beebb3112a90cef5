"""GPU parity tests (-m gpu): every kernel, called through the C-ABI (ops -> ctypes -> libb200bev.so),
against the numpy oracle on the same seeded inputs and against the golden vectors the reference
produced.  Tolerances: indices / cells / permutations bit-exact; fp32 features max|d| <= 1e-5 * max|ref|
(north_star's bound in the form SURVEY §7 derives); bf16 tensor path 1e-2.
"""
import numpy as np
import pytest
import torch

from bevfusion_multimodal_3d_object_detection_b200 import _lib, ops
from bevfusion_multimodal_3d_object_detection_b200 import synthetic as syn
from oracle import bev_oracle as orc
from tests.conftest import max_rel

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-5
BF16_TOL = 1e-2


def dev_t(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


def packed(layers, cuda):
    ws, bs = orc.fold_layers(layers)
    return ops.pack_mlp_params([torch.from_numpy(w) for w in ws], [torch.from_numpy(b) for b in bs], cuda)


def test_library_is_loaded_and_device_is_blackwell(cuda):
    import ctypes as C

    sm, major, minor = C.c_int(), C.c_int(), C.c_int()
    _lib.check(_lib.lib().b200bev_device_info(C.byref(sm), C.byref(major), C.byref(minor)))
    assert major.value == 10 and sm.value >= 100


# ------------------------------------------------------------------------------------------------ N3
def test_lidar_prepare_vs_reference_golden(cuda, golden):
    """Range filter + pad / subsample, against what the reference's own dataset code produced."""
    g = golden("lidar_prepare")
    cases = {"pad": (3000, 2600), "pad_exact_empty": (64, 80), "subsample": (5000, 2048)}
    for name, (rows, max_points) in cases.items():
        raw = syn.raw_sweep(601 + rows, rows)
        if name == "pad_exact_empty":
            raw[:, 0] = 60.0
        off = torch.tensor([0, rows], dtype=torch.int64, device=cuda)
        sel = None
        if f"{name}_indices" in g:
            sel = torch.from_numpy(g[f"{name}_indices"]).to(cuda).view(1, -1)
        out, count = ops.lidar_prepare(dev_t(raw, cuda), off, max_points, syn.PC_RANGE, select=sel)
        assert int(count[0]) == int(g[f"{name}_count"])
        np.testing.assert_array_equal(out[0].cpu().numpy(), g[f"{name}_out"])       # bit-exact: it is a copy


@pytest.mark.parametrize("rows,C,max_points", [([35211, 0, 1, 40000, 777], 4, 35000), ([3000, 2999], 5, 1024),
                                                ([300000, 280000], 4, 300000)])
def test_lidar_prepare_ragged_batches(cuda, rows, C, max_points):
    """Whole batches with empty, tiny and over-full frames, 4- and 5-channel rows, against the oracle."""
    sweeps = [syn.raw_sweep(700 + i, r, channels=C) for i, r in enumerate(rows)]
    raw = np.concatenate(sweeps, axis=0) if sum(rows) else np.zeros((0, C), np.float32)
    off = torch.tensor([0] + list(np.cumsum(rows)), dtype=torch.int64, device=cuda)
    out, count = ops.lidar_prepare(dev_t(raw, cuda), off, max_points, syn.PC_RANGE, max_frame_rows=max(rows))
    for b, s in enumerate(sweeps):
        ref, n = orc.lidar_prepare(s, max_points, syn.PC_RANGE)
        assert int(count[b]) == n
        np.testing.assert_array_equal(out[b].cpu().numpy(), ref)
    # the subsample branch with a caller-supplied draw, for the frames that overflow
    counts = count.tolist()
    over = [b for b, n in enumerate(counts) if n >= max_points]
    if over:
        rng = np.random.default_rng(5)
        sel = np.full((len(rows), max_points), -1, dtype=np.int32)
        for b in over:
            sel[b] = rng.choice(counts[b], max_points, replace=False)
        got, _ = ops.lidar_prepare(dev_t(raw, cuda), off, max_points, syn.PC_RANGE, select=dev_t(sel, cuda), max_frame_rows=max(rows))
        for b in over:
            ref, _ = orc.lidar_prepare(sweeps[b], max_points, syn.PC_RANGE, sel[b])
            np.testing.assert_array_equal(got[b].cpu().numpy(), ref)


@pytest.mark.parametrize("rows,C,max_points,W", [([35211, 0, 1, 40000, 777], 4, 35000, 50), ([3000, 2999, 10], 5, 1024, 7),
                                                  ([300000, 280000], 4, 300000, 100), ([5000], 4, 2048, 50)])
def test_lidar_prepare_bin_sort_in_one_launch(cuda, rows, C, max_points, W):
    """SURVEY 8f N3 as worded: range filter + compaction + padding fused into the bin-and-sort launch — bit-identical to
    lidar_prepare followed by bin_sort, and to the oracle."""
    sweeps = [syn.raw_sweep(900 + i, r, channels=C) for i, r in enumerate(rows)]
    raw = dev_t(np.concatenate(sweeps, axis=0) if sum(rows) else np.zeros((0, C), np.float32), cuda)
    off = torch.tensor([0] + list(np.cumsum(rows)), dtype=torch.int64, device=cuda)
    pts, count, cell, perm, offs = ops.lidar_prepare_bin_sort(raw, off, max_points, W, W, syn.PC_RANGE, max_frame_rows=max(rows))
    want_pts, want_count = ops.lidar_prepare(raw, off, max_points, syn.PC_RANGE, max_frame_rows=max(rows))
    want_cell, want_perm, want_offs = ops.bin_sort(want_pts, W, W, syn.PC_RANGE)
    for got, want in ((pts, want_pts), (count, want_count), (cell, want_cell), (perm, want_perm), (offs, want_offs)):
        assert torch.equal(got, want)
    for b, sw in enumerate(sweeps[:2]):
        ref, n = orc.lidar_prepare(sw, max_points, syn.PC_RANGE)
        assert int(count[b]) == n
        np.testing.assert_array_equal(pts[b].cpu().numpy(), ref)
        rc = orc.cell_index(ref[None], syn.PC_RANGE, W, W)[0]
        np.testing.assert_array_equal(cell[b].cpu().numpy(), rc)
        rp, ro = orc.bin_sort(rc, W * W)
        np.testing.assert_array_equal(perm[b].cpu().numpy(), rp)
        np.testing.assert_array_equal(offs[b].cpu().numpy(), ro)


def test_prepare_lidar_batch_feeds_the_encoder_path(cuda, tmp_path):
    """dataset.prepare_lidar_batch: .bin files -> filtered/padded batch -> bin_sort accepts every kept point."""
    from bevfusion_multimodal_3d_object_detection_b200 import dataset

    paths = []
    for i, r in enumerate((5000, 1200)):
        p = tmp_path / f"sweep{i}.bin"
        syn.raw_sweep(800 + i, r).tofile(p)
        paths.append(p)
    pts, count = dataset.prepare_lidar_batch(paths, cuda, max_points=2048, rng=np.random.default_rng(3))
    assert tuple(pts.shape) == (2, 2048, 4) and count.tolist()[1] < 2048 <= count.tolist()[0]
    cell, _, off = ops.bin_sort(pts, 50, 50)
    n1 = int(count[1])
    assert bool((cell[0] >= 0).all()) and bool((cell[1, :n1] >= 0).all())     # in-range points are all in the grid
    assert not bool(pts[1, n1:].any())                                         # zero padding
    ref_rows = {tuple(r) for r in orc.lidar_prepare(dataset.read_sweep(paths[0]), 10**6, syn.PC_RANGE)[0][: int(count[0])].tolist()}
    assert all(tuple(r) in ref_rows for r in pts[0].cpu().numpy().tolist())     # a subset of the in-range points


# ------------------------------------------------------------------------------------------------ S1a
@pytest.mark.parametrize("B,N,W,H", [(1, 35000, 50, 50), (3, 2011, 50, 50), (2, 777, 7, 5), (2, 40000, 100, 100),
                                     (1, 31, 50, 50), (1, 300000, 100, 100)])
def test_bin_sort_bit_exact(cuda, B, N, W, H):
    pts = syn.lidar_batch(900 + N, B, n_valid=max(N - N // 100 - 1, 1), n_total=N)
    # push some points outside the grid / onto edges / to NaN
    pts[:, 3::97, 0] = 60.0
    pts[:, 5::101, 1] = -51.2
    pts[:, 7::103, 0] = 51.2
    pts[:, 11::211, 1] = np.nan
    cell, perm, off = ops.bin_sort(dev_t(pts, cuda), W, H)
    ref_cell = orc.cell_index(pts, syn.PC_RANGE, W, H)
    np.testing.assert_array_equal(cell.cpu().numpy(), ref_cell)
    for b in range(B):
        rp, ro = orc.bin_sort(ref_cell[b], W * H)
        np.testing.assert_array_equal(off[b].cpu().numpy(), ro)
        np.testing.assert_array_equal(perm[b].cpu().numpy(), rp)


def test_bin_sort_all_points_in_one_cell_and_generic_channels(cuda):
    pts = np.zeros((2, 5000, 5), dtype=np.float32)            # 5-channel rows: scalar load path
    pts[1, :, 0] = np.linspace(-60, 60, 5000)
    cell, perm, off = ops.bin_sort(dev_t(pts, cuda), 50, 50)
    ref_cell = orc.cell_index(pts, syn.PC_RANGE, 50, 50)
    np.testing.assert_array_equal(cell.cpu().numpy(), ref_cell)
    for b in range(2):
        rp, ro = orc.bin_sort(ref_cell[b], 2500)
        np.testing.assert_array_equal(perm[b].cpu().numpy(), rp)
        np.testing.assert_array_equal(off[b].cpu().numpy(), ro)


def test_bin_sort_round_trip_properties_full_size(cuda):
    """Size-independent properties at BASELINE scale: perm is a permutation, cells are sorted, stable."""
    pts = dev_t(syn.lidar_batch(950, 4), cuda)
    cell, perm, off = ops.bin_sort(pts, 50, 50)
    for b in range(4):
        p = perm[b].long()
        assert torch.equal(torch.sort(p)[0], torch.arange(35000, device=cuda))
        c = cell[b][p]
        n_in = int(off[b, -1])
        assert bool((c[:n_in] >= 0).all()) and bool((c[n_in:] == -1).all())
        assert bool((c[1:n_in] >= c[: n_in - 1]).all())
        same = c[1:n_in] == c[: n_in - 1]
        assert bool((p[1:n_in][same] > p[: n_in - 1][same]).all())       # stability
        assert torch.equal(torch.bincount(c[:n_in].long(), minlength=2500), (off[b, 1:] - off[b, :-1]).long())


# ------------------------------------------------------------------------------------------------ S1b
def test_lidar_global_max_vs_golden_and_oracle(cuda, golden):
    g = golden("lidar_encoder")
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    blob, dims = packed(layers, cuda)
    small = syn.lidar_batch(201, 2, n_valid=1900, n_total=2011)
    got = ops.pointnet_encode(dev_t(small, cuda), blob, dims).cpu().numpy()
    assert max_rel(got, g["small_global"]) < FP32_TOL
    assert max_rel(got, orc.pointnet_global(small, layers)) < FP32_TOL
    full = syn.lidar_batch(301, 1)
    got = ops.pointnet_encode(dev_t(full, cuda), blob, dims).cpu().numpy()
    assert max_rel(got, g["full_global"]) < FP32_TOL


@pytest.mark.parametrize("B,N", [(1, 1), (1, 63), (2, 64), (3, 65), (5, 1000)])
def test_lidar_global_ragged_sizes(cuda, B, N):
    layers = syn.mlp_weights(102, syn.LIDAR_DIMS)
    blob, dims = packed(layers, cuda)
    pts = syn.lidar_batch(600 + N, B, n_valid=max(N - 2, 1), n_total=N)
    got = ops.pointnet_encode(dev_t(pts, cuda), blob, dims).cpu().numpy()
    assert max_rel(got, orc.pointnet_global(pts, layers)) < FP32_TOL


def test_lidar_without_batchnorm_and_other_widths(cuda):
    dims = (5, 32, 48, 136)
    layers = syn.mlp_weights(103, dims, use_bn=False)
    blob, d = packed(layers, cuda)
    pts = syn.lidar_batch(610, 2, n_valid=300, n_total=333, channels=5)
    got = ops.pointnet_encode(dev_t(pts, cuda), blob, d).cpu().numpy()
    assert max_rel(got, orc.pointnet_global(pts, layers)) < FP32_TOL


def test_lidar_cell_canvas(cuda, golden):
    g = golden("lidar_encoder")
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    blob, dims = packed(layers, cuda)
    pts = syn.lidar_batch(201, 2, n_valid=1900, n_total=2011)
    pts[:, 50:60, 0] = 70.0                                     # a few points outside the grid are dropped
    d = dev_t(pts, cuda)
    cell, perm, off = ops.bin_sort(d, 50, 50)
    glob, canvas = ops.pointnet_encode(d, blob, dims, perm=perm, offsets=off, n_cells=2500)
    canvas = canvas.cpu().numpy()
    # one pass, both outputs: the global max still sees the out-of-grid points (reference semantics)
    assert max_rel(glob.cpu().numpy(), orc.pointnet_global(pts, layers)) < FP32_TOL
    ref_cell = orc.cell_index(pts, syn.PC_RANGE, 50, 50)
    for b in range(2):
        ref = orc.pointnet_cell_max(pts[b], layers, ref_cell[b], 2500)
        assert max_rel(canvas[b], ref) < FP32_TOL
        empty = np.bincount(ref_cell[b][ref_cell[b] >= 0], minlength=2500) == 0
        assert not canvas[b][empty].any()                      # empty cells are exactly zero
    # unmodified input against the canvas the reference's per-point features give (golden)
    pts0 = syn.lidar_batch(201, 2, n_valid=1900, n_total=2011)
    d0 = dev_t(pts0, cuda)
    _, perm0, off0 = ops.bin_sort(d0, 50, 50)
    c0 = ops.pointnet_encode(d0, blob, dims, perm=perm0, offsets=off0, n_cells=2500, want_global=False).cpu().numpy()
    assert max_rel(c0[:, :, ::16], g["small_canvas_sub"]) < FP32_TOL


def test_global_max_equals_max_over_canvas_full_size(cuda):
    """Property at full size: with every point in the grid, max over cells == global max."""
    layers = syn.mlp_weights(104, syn.LIDAR_DIMS)
    blob, dims = packed(layers, cuda)
    d = dev_t(syn.lidar_batch(620, 2), cuda)
    glob = ops.pointnet_encode(d, blob, dims)
    _, perm, off = ops.bin_sort(d, 50, 50)
    glob2, canvas = ops.pointnet_encode(d, blob, dims, perm=perm, offsets=off, n_cells=2500)
    assert torch.equal(canvas.max(dim=1)[0], glob) and torch.equal(glob2, glob)


# ------------------------------------------------------------------------------------------------ S1c
@pytest.mark.parametrize("method", ["concat", "max", "mean"])
def test_multi_radar(cuda, golden, method):
    g = golden("radar_encoder")
    layers = syn.mlp_weights(111, syn.RADAR_DIMS)
    fcw, fcb = syn.linear_weights(112, 5 * 256, 256)
    blob, dims = packed(layers, cuda)
    radars = syn.radar_batch(211, 3)
    fused, stacked = ops.radar_encode([dev_t(r, cuda) for r in radars], blob, dims, method, dev_t(fcw, cuda), dev_t(fcb, cuda))
    ref_fused, ref_stacked = orc.multi_radar(radars, layers, method, fcw, fcb)
    assert max_rel(stacked.cpu().numpy(), ref_stacked) < FP32_TOL
    assert max_rel(fused.cpu().numpy(), g[f"fused_{method}"]) < FP32_TOL
    ragged = [np.ascontiguousarray(r[:, : 125 - 17 * i]) for i, r in enumerate(radars)]
    fused, _ = ops.radar_encode([dev_t(r, cuda) for r in ragged], blob, dims, method, dev_t(fcw, cuda), dev_t(fcb, cuda))
    assert max_rel(fused.cpu().numpy(), g[f"ragged_{method}"]) < FP32_TOL


# ------------------------------------------------------------------------------------------------ S2
@pytest.mark.parametrize("name,shape", [("ref28x50", (16, 28, 50, 50, 50)), ("hd57x100", (8, 57, 100, 50, 50)),
                                        ("up7x9", (8, 7, 9, 20, 30))])
def test_camera_mean_and_resize(cuda, golden, name, shape):
    g = golden("camera_bev")
    C, h, w, H, W = shape
    feats = syn.camera_features(401, 2, n_cam=6, channels=C, h=h, w=w)
    mean = ops.camera_mean(dev_t(feats, cuda)).cpu().numpy()
    np.testing.assert_array_equal(mean, orc.camera_mean(feats))          # same association: bit-exact
    assert max_rel(mean, g[f"{name}_mean"]) < 2e-7
    out = ops.bilinear_resize(dev_t(g[f"{name}_resize_in"], cuda), (H, W)).cpu().numpy()
    assert max_rel(out, g[f"{name}_resize_out"]) < FP32_TOL


@pytest.mark.parametrize("B,h,w,C,H,W,C_total,c_off", [(2, 57, 100, 64, 50, 50, 192, 64),     # camera_proj's map into a slice of bev_fusion's input
                                                         (3, 25, 25, 128, 50, 50, 128, 0),      # the lidar branch's x2 upsample
                                                         (1, 7, 5, 8, 3, 11, 24, 16)])          # down in y, up in x, one channel group
def test_bilinear_resize_channels_last_bf16(cuda, B, h, w, C, H, W, C_total, c_off):
    """b200bev_bilinear_resize_nhwc_bf16: the resize of src/fusion.py:242-247 on channels-last bf16, between two convolutions of
    the bf16 path.  Bit-exact against the fp32 kernel applied to the same bf16 values and rounded to bf16 (same op order);
    the channels of `out` outside the slice stay untouched."""
    rng = np.random.default_rng(B * 100 + C)
    x = torch.from_numpy(rng.standard_normal((B, C, h, w)).astype(np.float32)).to(cuda).to(torch.bfloat16)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    out = torch.full((B, H, W, C_total), 7.0, dtype=torch.bfloat16, device=cuda)
    ops.bilinear_resize_nhwc_bf16(x_nhwc, (H, W), out=out, c_offset=c_off)
    ref = ops.bilinear_resize(x.float().contiguous(), (H, W)).to(torch.bfloat16).permute(0, 2, 3, 1)
    assert torch.equal(out[..., c_off:c_off + C], ref)
    assert bool((out[..., :c_off] == 7.0).all()) and bool((out[..., c_off + C:] == 7.0).all())
    # and against torch's own interpolate, at the bf16 bound
    tref = torch.nn.functional.interpolate(x.float(), size=(H, W), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    assert max_rel(out[..., c_off:c_off + C].float().cpu().numpy(), tref.cpu().numpy()) < BF16_TOL
    alone = ops.bilinear_resize_nhwc_bf16(x_nhwc, (H, W))
    assert torch.equal(alone, ref.contiguous())


def test_camera_mean_odd_sizes(cuda):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 5, 3, 7, 11)).astype(np.float32)          # inner not a multiple of 4, 5 cameras
    np.testing.assert_array_equal(ops.camera_mean(dev_t(x, cuda)).cpu().numpy(), orc.camera_mean(x))
    y = rng.standard_normal((3, 4, 8, 4, 4)).astype(np.float32)           # vector path, n_cam != 6
    np.testing.assert_array_equal(ops.camera_mean(dev_t(y, cuda)).cpu().numpy(), orc.camera_mean(y))


@pytest.fixture(params=["staged", "gather"])
def project_impl(request):
    """Both camera_project kernels (TMA-staged bands / plain global gather) must give the same bits."""
    return request.param


def test_camera_projection(cuda, golden, project_impl):
    g = golden("camera_bev")
    K, E = syn.camera_rig()
    feats = syn.camera_features(402, 1, n_cam=6, channels=8, h=57, w=100)
    canvas, table = ops.camera_project(dev_t(feats, cuda), dev_t(K, cuda), dev_t(E, cuda), (1600.0, 900.0), (50, 50),
                                       return_table=True, impl=project_impl)
    np.testing.assert_array_equal(table[0].cpu().numpy(), g["project_table"])   # (u, v, valid) bit-exact
    assert max_rel(canvas[0].cpu().numpy(), g["project_canvas_grid_sample"]) < FP32_TOL
    assert max_rel(canvas[0].cpu().numpy(), orc.camera_project(feats[0], g["project_table"], (50, 50))) < 1e-6


def test_camera_projection_per_sample_rigs_and_big_grid(cuda, project_impl):
    K, E = syn.camera_rig()
    E2 = E.copy()
    E2[:, :, 3] += np.float32(0.25)
    Ks, Es = np.stack([K, K]), np.stack([E, E2])
    feats = syn.camera_features(403, 2, n_cam=6, channels=5, h=28, w=50)
    canvas, table = ops.camera_project(dev_t(feats, cuda), dev_t(Ks, cuda), dev_t(Es, cuda), (1600.0, 900.0), (100, 100),
                                       return_table=True, impl=project_impl)
    for b in range(2):
        t = orc.project_cells(Ks[b], Es[b], (1600.0, 900.0), (28, 50), (100, 100), syn.PC_RANGE)
        np.testing.assert_array_equal(table[b].cpu().numpy(), t)
        assert max_rel(canvas[b].cpu().numpy(), orc.camera_project(feats[b], t, (100, 100))) < 1e-6


def test_camera_projection_staged_equals_gather_full_size(cuda):
    """BASELINE-size features (6x512x57x100, 3 frames, ragged channel tail): the staged kernel and the
    gather kernel agree bit for bit, and a blind camera / an all-blind rig give zeros, not garbage."""
    K, E = syn.camera_rig()
    g = torch.Generator(device=cuda).manual_seed(9)
    feats = torch.relu(torch.randn((3, 6, 510, 57, 100), device=cuda, generator=g))
    Kd, Ed = dev_t(K, cuda), dev_t(E, cuda)
    outs = {}
    for impl in ("staged", "gather"):
        outs[impl] = ops.camera_project(feats, Kd, Ed, (1600.0, 900.0), (50, 50), impl=impl)
        outs[impl + "_big"] = ops.camera_project(feats[:1, :, :37].contiguous(), Kd, Ed, (1600.0, 900.0), (100, 100), impl=impl)
    assert torch.equal(outs["staged"], outs["gather"])
    assert torch.equal(outs["staged_big"], outs["gather_big"])
    assert float(outs["staged"].abs().max()) > 0
    # camera 3 looks at the sky (pitch it up by 90 degrees): nothing on the ground plane projects into it
    E_sky = E.copy()
    E_sky[3, :, :3] = E[3, [2, 0, 1], :3]
    E_far = E.copy()
    E_far[:, 2, 3] = -1.0e6                      # every cell ends up behind every camera
    for rig in (E_sky, E_far):
        res = {}
        for impl in ("staged", "gather"):
            res[impl] = ops.camera_project(feats[:2, :, :9].contiguous(), Kd, dev_t(rig, cuda), (1600.0, 900.0), (50, 50), impl=impl)
        assert torch.equal(res["staged"], res["gather"])
    assert float(res["staged"].abs().max()) == 0.0


def test_camera_projection_overlapping_cameras(cuda):
    """Means over 3 and 6 cameras (not powers of two: the IEEE-division branch) and a rig whose cameras all
    see most of the grid, which overflows the staged kernel's shared-memory table and takes its
    gather-from-global segment path.  All against the oracle, and staged == gather bit for bit."""
    K, E = syn.camera_rig()
    E3 = E.copy()
    E3[1] = E[0]
    E3[2] = E[0]                                  # the forward wedge is seen by cameras 0, 1 and 2
    down = np.array([[0.0, -1.0, 0.0], [-1.0, 0.0, 0.0], [0.0, 0.0, -1.0]], dtype=np.float32)   # looking straight down
    t = -(down @ np.array([0.0, 0.0, 100.0], dtype=np.float32))
    E6 = np.stack([np.concatenate([down, t[:, None]], axis=1)] * 6).astype(np.float32)
    feats = syn.camera_features(404, 2, n_cam=6, channels=7, h=57, w=100)
    d = dev_t(feats, cuda)
    for rig, bev, min_overlap in ((E3, (50, 50), 3), (E6, (50, 50), 6), (E6, (100, 100), 6)):
        table = orc.project_cells(K, rig, (1600.0, 900.0), (57, 100), bev, syn.PC_RANGE)
        assert int(table[:, :, 2].sum(axis=1).max()) == min_overlap
        res = {}
        for impl in ("staged", "gather"):
            res[impl] = ops.camera_project(d, dev_t(K, cuda), dev_t(rig, cuda), (1600.0, 900.0), bev, impl=impl)
        assert torch.equal(res["staged"], res["gather"])
        for b in range(2):
            assert max_rel(res["staged"][b].cpu().numpy(), orc.camera_project(feats[b], table, bev)) < 1e-6


# ------------------------------------------------------------------------------------------------ S3
def test_nms_and_topk_vs_golden(cuda, golden):
    g = golden("centernet_decode")
    heat = syn.head_maps(501, 3)["heatmap"]
    nms = ops.centernet_nms(dev_t(heat, cuda))
    np.testing.assert_array_equal(nms.cpu().numpy(), g["nms"])
    score, ind, cls, ys, xs = ops.centernet_topk(nms, 100)
    np.testing.assert_array_equal(score.cpu().numpy(), g["topk_score"])
    np.testing.assert_array_equal(ind.cpu().numpy(), g["topk_ind"])
    np.testing.assert_array_equal(ys.cpu().numpy(), g["topk_ys"])
    np.testing.assert_array_equal(xs.cpu().numpy(), g["topk_xs"])
    np.testing.assert_array_equal(cls.cpu().numpy(), g["topk_classes"])
    assert ind.dtype == torch.int64 and cls.dtype == torch.int64


def test_hand_made_peaks_and_tie_rule(cuda, golden):
    g = golden("centernet_decode")
    nms = ops.centernet_nms(dev_t(g["hand_heat"], cuda))
    np.testing.assert_array_equal(nms.cpu().numpy(), g["hand_nms"])
    score, _, _, ys, xs = ops.centernet_topk(nms, 6)
    np.testing.assert_array_equal(score.cpu().numpy(), g["hand_topk_score"])
    np.testing.assert_array_equal(ys.cpu().numpy(), g["hand_topk_ys"])
    np.testing.assert_array_equal(xs.cpu().numpy(), g["hand_topk_xs"])
    ref = orc.topk(g["hand_nms"], 8)
    got = ops.centernet_topk(nms, 8)
    for a, b in zip(got, ref):
        np.testing.assert_array_equal(a.cpu().numpy(), b)           # includes the plateau ties: index-ascending
    with pytest.raises(RuntimeError, match="selected index k out of range"):
        ops.centernet_topk(dev_t(g["hand_heat"], cuda), 43)


def test_topk_on_plateaus_and_negative_scores(cuda):
    rng = np.random.default_rng(5)
    s = rng.integers(-3, 4, (2, 3, 9, 13)).astype(np.float32)           # massive ties, negatives, zeros
    got = ops.centernet_topk(dev_t(s, cuda), 17)
    for a, b in zip(got, orc.topk(s, 17)):
        np.testing.assert_array_equal(a.cpu().numpy(), b)


def _check_decode(out, ref_dets):
    counts = out["count"].cpu().numpy()
    for b, ref in enumerate(ref_dets):
        n = len(ref["scores"])
        assert counts[b] == n
        np.testing.assert_array_equal(out["scores"][b, :n].cpu().numpy(), ref["scores"])
        np.testing.assert_array_equal(out["labels"][b, :n].cpu().numpy(), ref["labels"])
        np.testing.assert_array_equal(out["velocities"][b, :n].cpu().numpy(), ref["velocities"])
        boxes = out["boxes"][b, :n].cpu().numpy()
        np.testing.assert_array_equal(boxes[:, :6], ref["boxes"][:, :6])
        np.testing.assert_allclose(boxes[:, 6], ref["boxes"][:, 6], rtol=0, atol=1e-6)


@pytest.mark.parametrize("tag,voxel", [("ct", 2.048), ("fd", 0.512)])
@pytest.mark.parametrize("thr", [0.0, 0.3, 0.999])
def test_fused_decode_vs_reference_golden(cuda, golden, tag, voxel, thr):
    g = golden("centernet_decode")
    maps = syn.head_maps(501, 3)
    out = ops.centernet_decode(*[dev_t(maps[k], cuda) for k in ("heatmap", "offset", "size", "rot", "vel")], 100, voxel,
                               score_thresh=thr)
    ref = [{k: g[f"{tag}_thr{thr}_b{b}_{k}"] for k in ("boxes", "scores", "labels", "velocities")} for b in range(3)]
    _check_decode(out, ref)
    np.testing.assert_array_equal(out["ys"].cpu().numpy(), g["topk_ys"])
    np.testing.assert_array_equal(out["xs"].cpu().numpy(), g["topk_xs"])
    np.testing.assert_array_equal(out["ind"].cpu().numpy(), g["topk_ind"])


def test_fused_decode_sparse_big_and_batch32(cuda, golden):
    g = golden("centernet_decode")
    sparse = syn.head_maps(502, 2, peak_frac=0.02)
    out = ops.centernet_decode(*[dev_t(sparse[k], cuda) for k in ("heatmap", "offset", "size", "rot", "vel")], 100, 2.048,
                               score_thresh=0.1)
    _check_decode(out, [{k: g[f"sparse_b{b}_{k}"] for k in ("boxes", "scores", "labels", "velocities")} for b in range(2)])
    big = syn.head_maps(503, 1, H=100, W=100)
    out = ops.centernet_decode(*[dev_t(big[k], cuda) for k in ("heatmap", "offset", "size", "rot", "vel")], 100, 0.512)
    _check_decode(out, [{k: g[f"big_{k}"] for k in ("boxes", "scores", "labels", "velocities")}])
    maps = syn.head_maps(504, 32)
    out = ops.centernet_decode(*[dev_t(maps[k], cuda) for k in ("heatmap", "offset", "size", "rot", "vel")], 100, 2.048)
    _check_decode(out, orc.decode(maps, score_thresh=0.0))
    # idempotence of the ticket counters: a second call on the same workspace pattern gives the same answer
    out2 = ops.centernet_decode(*[dev_t(maps[k], cuda) for k in ("heatmap", "offset", "size", "rot", "vel")], 100, 2.048)
    assert torch.equal(out["boxes"], out2["boxes"]) and torch.equal(out["count"], out2["count"])


def test_decode_small_k_and_odd_grid(cuda):
    maps = syn.head_maps(505, 2, classes=3, H=9, W=13)
    for K in (1, 5, 64):
        out = ops.centernet_decode(*[dev_t(maps[k], cuda) for k in ("heatmap", "offset", "size", "rot", "vel")], K, 2.048,
                                   score_thresh=0.2)
        _check_decode(out, orc.decode(maps, score_thresh=0.2, max_detections=K))


# ------------------------------------------------------------------------------------------------ S1b, tcgen05 path
def _tc(cuda, layers):
    blob, dims = packed(layers, cuda)
    return blob, dims, ops.pack_mlp_params_bf16(blob, dims)


@pytest.mark.parametrize("B,N", [(1, 1), (1, 128), (2, 129), (2, 2011), (5, 1000), (3, 35000)])
def test_tensor_core_global_max(cuda, B, N):
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    blob, dims, tc = _tc(cuda, layers)
    pts = syn.lidar_batch(700 + N, B, n_valid=max(N - N // 50 - 1, 1), n_total=N)
    got = ops.pointnet_encode(dev_t(pts, cuda), blob, dims, precision=_lib.BF16_TENSOR, tc_params=tc).cpu().numpy()
    ref = orc.pointnet_global(pts, layers)
    assert max_rel(got, ref) < BF16_TOL
    # the fp32 kernel on the same input is the tighter cross-check of everything but the rounding
    f32 = ops.pointnet_encode(dev_t(pts, cuda), blob, dims).cpu().numpy()
    assert max_rel(got, f32) < BF16_TOL


def test_tensor_core_is_deterministic_and_matches_golden(cuda, golden):
    g = golden("lidar_encoder")
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    blob, dims, tc = _tc(cuda, layers)
    d = dev_t(syn.lidar_batch(301, 1), cuda)
    a = ops.pointnet_encode(d, blob, dims, precision=_lib.BF16_TENSOR, tc_params=tc)
    b = ops.pointnet_encode(d, blob, dims, precision=_lib.BF16_TENSOR, tc_params=tc)
    assert torch.equal(a, b)
    assert max_rel(a.cpu().numpy(), g["full_global"]) < BF16_TOL


def test_tensor_core_rejects_other_widths(cuda):
    layers = syn.mlp_weights(103, (5, 32, 48, 136), use_bn=False)
    blob, dims = packed(layers, cuda)
    with pytest.raises(_lib.B200BevError):
        ops.pack_mlp_params_bf16(blob, dims)


@pytest.mark.parametrize("B,N,W", [(2, 2011, 50), (1, 35000, 50), (2, 5000, 100)])
def test_tensor_core_cell_canvas(cuda, B, N, W):
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    blob, dims, tc = _tc(cuda, layers)
    pts = syn.lidar_batch(800 + N, B, n_valid=max(N - N // 50 - 1, 1), n_total=N)
    pts[:, 7::53, 0] = 75.0                                    # out-of-grid points: global max only
    d = dev_t(pts, cuda)
    _, perm, off = ops.bin_sort(d, W, W)
    glob, canvas = ops.pointnet_encode(d, blob, dims, perm=perm, offsets=off, n_cells=W * W, precision=_lib.BF16_TENSOR,
                                       tc_params=tc)
    g32, c32 = ops.pointnet_encode(d, blob, dims, perm=perm, offsets=off, n_cells=W * W)
    scale = float(c32.max())
    assert float((canvas - c32).abs().max()) < BF16_TOL * scale
    assert float((glob - g32).abs().max()) < BF16_TOL * float(g32.max())
    empty_cells = (off[:, 1:] - off[:, :-1]) == 0
    assert not bool(canvas[empty_cells].any())                  # untouched cells stay exactly zero
    ref_cell = orc.cell_index(pts, syn.PC_RANGE, W, W)
    for b in range(B):                                          # every batch element against the numpy oracle
        ref = orc.pointnet_cell_max(pts[b], layers, ref_cell[b], W * W)
        assert max_rel(canvas[b].cpu().numpy(), ref) < BF16_TOL, b
    only = ops.pointnet_encode(d, blob, dims, perm=perm, offsets=off, n_cells=W * W, precision=_lib.BF16_TENSOR, tc_params=tc,
                               want_global=False)
    assert torch.equal(only, canvas)                            # deterministic, with or without the global output


# ------------------------------------------------------------------------------------------------ S1b, fp32 accuracy on tcgen05
def _split(cuda, layers):
    blob, dims = packed(layers, cuda)
    img = ops.pack_mlp_params_split(blob, dims)
    assert img is not None
    return blob, dims, img


@pytest.mark.parametrize("B,N", [(1, 1), (1, 255), (2, 256), (2, 257), (3, 2011), (5, 1000), (2, 35000)])
def test_split_tensor_core_global_max_is_fp32_accurate(cuda, B, N):
    """The f32 precision on the tensor cores (three fp16 products per fp32 product): 1e-5 of max|ref| against the numpy
    oracle, and against the FFMA kernel."""
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    blob, dims, img = _split(cuda, layers)
    pts = syn.lidar_batch(900 + N, B, n_valid=max(N - N // 50 - 1, 1), n_total=N)
    d = dev_t(pts, cuda)
    got = ops.pointnet_encode(d, blob, dims, precision=_lib.F32, tc_params=img)
    assert max_rel(got.cpu().numpy(), orc.pointnet_global(pts, layers)) < FP32_TOL
    ffma = ops.pointnet_encode(d, blob, dims)
    assert max_rel(got.cpu().numpy(), ffma.cpu().numpy()) < FP32_TOL
    assert torch.equal(got, ops.pointnet_encode(d, blob, dims, precision=_lib.F32, tc_params=img))     # deterministic


def test_split_tensor_core_vs_reference_golden(cuda, golden):
    g = golden("lidar_encoder")
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    blob, dims, img = _split(cuda, layers)
    full = dev_t(syn.lidar_batch(301, 1), cuda)
    assert max_rel(ops.pointnet_encode(full, blob, dims, precision=_lib.F32, tc_params=img).cpu().numpy(), g["full_global"]) < FP32_TOL
    small = syn.lidar_batch(201, 2, n_valid=1900, n_total=2011)
    d = dev_t(small, cuda)
    assert max_rel(ops.pointnet_encode(d, blob, dims, precision=_lib.F32, tc_params=img).cpu().numpy(), g["small_global"]) < FP32_TOL
    cell, perm, off = ops.bin_sort(d, 50, 50)
    np.testing.assert_array_equal(cell.cpu().numpy(), g["small_cell"])
    glob, canvas = ops.pointnet_encode(d, blob, dims, perm=perm, offsets=off, n_cells=2500, precision=_lib.F32, tc_params=img)
    assert max_rel(glob.cpu().numpy(), g["small_global"]) < FP32_TOL
    assert max_rel(canvas[:, :, ::16].cpu().numpy(), g["small_canvas_sub"]) < FP32_TOL


@pytest.mark.parametrize("scale", [1e-4, 1.0, 3e3])
def test_split_tensor_core_is_scale_free(cuda, scale):
    """fp16 operands, fp32 range: a network whose activations sit far from unit scale (1e-8 .. 1e9 across the layers) keeps
    the 1e-5 bound — the per-channel weight scales and the per-layer activation scales are exact powers of two taken from
    the data.  ReLU is positively homogeneous, so scaling layer i's weights by f_i (and every bias by the cumulative
    factor) gives an exactly scaled copy of the unit-scale network."""
    layers = syn.mlp_weights(131, syn.LIDAR_DIMS, use_bn=False)
    factors = [scale, 1.0, scale, 1.0 / scale, 1.0 if scale == 1.0 else 7.0]
    cum = 1.0
    for lay, f in zip(layers, factors):
        cum *= f
        lay["weight"] = (lay["weight"].astype(np.float64) * f).astype(np.float32)
        lay["bias"] = (lay["bias"].astype(np.float64) * cum).astype(np.float32)
    blob, dims, img = _split(cuda, layers)
    pts = syn.lidar_batch(77, 2, n_valid=1500, n_total=1536)
    got = ops.pointnet_encode(dev_t(pts, cuda), blob, dims, precision=_lib.F32, tc_params=img).cpu().numpy()
    ref = orc.pointnet_global(pts, layers)
    assert np.isfinite(got).all() and float(np.abs(ref).max()) > 0 and max_rel(got, ref) < FP32_TOL


@pytest.mark.parametrize("B,N,W", [(2, 2011, 50), (1, 35000, 50), (2, 5000, 100), (3, 300, 7)])
def test_split_tensor_core_cell_canvas(cuda, B, N, W):
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    blob, dims, img = _split(cuda, layers)
    pts = syn.lidar_batch(800 + N, B, n_valid=max(N - N // 50 - 1, 1), n_total=N)
    pts[:, 7::53, 0] = 75.0                                    # out-of-grid points: global max only
    d = dev_t(pts, cuda)
    _, perm, off = ops.bin_sort(d, W, W)
    glob, canvas = ops.pointnet_encode(d, blob, dims, perm=perm, offsets=off, n_cells=W * W, precision=_lib.F32, tc_params=img)
    g32, c32 = ops.pointnet_encode(d, blob, dims, perm=perm, offsets=off, n_cells=W * W)
    assert float((canvas - c32).abs().max()) < FP32_TOL * float(c32.max())
    assert float((glob - g32).abs().max()) < FP32_TOL * float(g32.max())
    empty_cells = (off[:, 1:] - off[:, :-1]) == 0
    assert not bool(canvas[empty_cells].any())
    ref_cell = orc.cell_index(pts, syn.PC_RANGE, W, W)
    for b in range(B):
        assert max_rel(canvas[b].cpu().numpy(), orc.pointnet_cell_max(pts[b], layers, ref_cell[b], W * W)) < FP32_TOL, b
    only = ops.pointnet_encode(d, blob, dims, perm=perm, offsets=off, n_cells=W * W, precision=_lib.F32, tc_params=img, want_global=False)
    assert torch.equal(only, canvas)


def test_split_tensor_core_runs_in_passes_when_the_workspace_is_small(cuda):
    """The C-ABI takes the scratch from the caller: with room for one frame only the batch runs frame by frame, same bits."""
    import ctypes as C

    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    blob, dims, img = _split(cuda, layers)
    B, N = 3, 700
    d = dev_t(syn.lidar_batch(55, B, n_valid=690, n_total=N), cuda)
    want = ops.pointnet_encode(d, blob, dims, precision=_lib.F32, tc_params=img)
    lib = _lib.lib()
    one = lib.b200bev_pointnet_split_workspace_bytes(1, N)
    assert one < lib.b200bev_pointnet_split_workspace_bytes(B, N)
    ws = torch.empty(one, dtype=torch.uint8, device=cuda)
    out = torch.empty((B, 1024), dtype=torch.float32, device=cuda)
    dd = (C.c_int32 * 6)(*dims)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.b200bev_pointnet_encode_split(C.c_void_p(d.data_ptr()), B, N, 4, dd, 5, None, None, 0, C.c_void_p(img.data_ptr()),
                                           C.c_void_p(out.data_ptr()), None, C.c_void_p(ws.data_ptr()), one, st)
    assert rc == 0 and torch.equal(out, want)
    assert lib.b200bev_pointnet_encode_split(C.c_void_p(d.data_ptr()), B, N, 4, dd, 5, None, None, 0, C.c_void_p(img.data_ptr()),
                                             C.c_void_p(out.data_ptr()), None, C.c_void_p(ws.data_ptr()), 4096, st) == _lib.ERR_WORKSPACE


def _oracle_cells(pts_b, layers, cell_b, cells):
    """Oracle canvas rows of the chosen cells only: the MLP runs on the points that fall into them."""
    keep = np.isin(cell_b, cells)
    remap = np.searchsorted(cells, cell_b[keep]).astype(np.int32)
    return orc.pointnet_cell_max(pts_b[keep], layers, remap, len(cells))


@pytest.mark.parametrize("precision", ["bf16", "f32"])
def test_cell_canvas_at_the_stress_shape(cuda, precision):
    """BASELINE configs[4] shapes: 4 frames x 300,000 points (10 sweeps), 100x100 grid.  bin_sort takes its "ranks in the
    upper half of the cell word" path here and the cell-mode MLP kernels consume its output.  Every frame is checked:
    against the numpy oracle on sampled cells (dense centre cells, sparse rim cells, the cell of the zero-padding rows),
    and — the bf16 kernel — against the fp32 kernel on ALL cells."""
    B, N, W = 4, 300000, 100
    layers = syn.mlp_weights(101, syn.LIDAR_DIMS)
    blob, dims, tc = _tc(cuda, layers)
    pts = syn.lidar_batch(1300, B, n_valid=N - 2000, n_total=N)          # 2,000 zero rows: they land in one cell
    pts[:, 11::97, 1] = -80.0                                            # out-of-grid points: global max only
    d = dev_t(pts, cuda)
    cell, perm, off = ops.bin_sort(d, W, W)
    ref_cell = orc.cell_index(pts, syn.PC_RANGE, W, W)
    np.testing.assert_array_equal(cell.cpu().numpy(), ref_cell)
    g32, c32 = ops.pointnet_encode(d, blob, dims, perm=perm, offsets=off, n_cells=W * W)          # FFMA kernel
    if precision == "bf16":
        glob, canvas = ops.pointnet_encode(d, blob, dims, perm=perm, offsets=off, n_cells=W * W, precision=_lib.BF16_TENSOR, tc_params=tc)
        tol = BF16_TOL
    else:
        img = ops.pack_mlp_params_split(blob, dims)
        glob, canvas = ops.pointnet_encode(d, blob, dims, perm=perm, offsets=off, n_cells=W * W, precision=_lib.F32, tc_params=img)
        tol = FP32_TOL
    assert float((canvas - c32).abs().max()) < tol * float(c32.max())          # ALL cells of all frames against the FFMA kernel
    assert float((glob - g32).abs().max()) < tol * float(g32.max())
    counts = (off[:, 1:] - off[:, :-1]).cpu().numpy()
    assert int(counts.max()) > 1000 and int(off[:, -1].max()) < N        # a cell far over 255 points; some points out of grid
    rng = np.random.default_rng(77)
    for b in range(B):
        assert not bool(canvas[b][torch.from_numpy(counts[b] == 0).to(cuda)].any())       # empty cells stay exactly zero
        occupied = np.flatnonzero(counts[b])
        order = occupied[np.argsort(counts[b][occupied])]
        cells = np.unique(np.concatenate([order[:8], order[-6:], rng.choice(occupied, 24, replace=False),
                                          [int(ref_cell[b, N - 1])]]))   # sparsest, densest, random, the zero-row cell
        ref = _oracle_cells(pts[b], layers, ref_cell[b], cells)
        got = canvas[b][torch.from_numpy(cells).to(cuda)].cpu().numpy()
        scale = float(c32[b].max())
        assert float(np.abs(got - ref).max()) < tol * scale, b
    ref_g = np.stack([orc.pointnet_global(pts[b:b + 1, ::37], layers)[0] for b in range(B)])   # a lower bound of the maxima
    assert bool((glob.cpu().numpy() >= ref_g * (1 - 2 * tol) - tol).all())


# ------------------------------------------------------------------------------------------------ randomised shapes
@pytest.mark.parametrize("seed", range(6))
def test_random_shapes_bin_sort_prepare_decode(cuda, seed):
    """Shapes drawn at random (odd sizes, tiny and large grids, K near H*W): integer outputs stay bit-exact."""
    rng = np.random.default_rng(1000 + seed)
    # bin_sort
    B, N = int(rng.integers(1, 5)), int(rng.integers(1, 9000))
    W, H = int(rng.integers(1, 120)), int(rng.integers(1, 120))
    pts = np.zeros((B, N, 4), dtype=np.float32)
    pts[..., 0] = rng.uniform(-60, 60, (B, N))
    pts[..., 1] = rng.uniform(-60, 60, (B, N))
    cell, perm, off = ops.bin_sort(dev_t(pts, cuda), W, H)
    ref_cell = orc.cell_index(pts, syn.PC_RANGE, W, H)
    np.testing.assert_array_equal(cell.cpu().numpy(), ref_cell)
    for b in range(B):
        rp, ro = orc.bin_sort(ref_cell[b], W * H)
        np.testing.assert_array_equal(perm[b].cpu().numpy(), rp)
        np.testing.assert_array_equal(off[b].cpu().numpy(), ro)
    # lidar_prepare
    rows = [int(rng.integers(0, 6000)) for _ in range(int(rng.integers(1, 5)))]
    C = int(rng.integers(3, 7))
    max_points = int(rng.integers(1, 5000))
    sweeps = [syn.raw_sweep(2000 + seed * 10 + i, r, channels=C) for i, r in enumerate(rows)]
    raw = np.concatenate(sweeps, axis=0) if sum(rows) else np.zeros((1, C), np.float32)[:0]
    offs = torch.tensor([0] + list(np.cumsum(rows)), dtype=torch.int64, device=cuda)
    if raw.shape[0]:
        out, count = ops.lidar_prepare(dev_t(raw, cuda), offs, max_points, syn.PC_RANGE, max_frame_rows=max(max(rows), 1))
        for b, s in enumerate(sweeps):
            ref, n = orc.lidar_prepare(s, max_points, syn.PC_RANGE)
            assert int(count[b]) == n
            np.testing.assert_array_equal(out[b].cpu().numpy(), ref)
    # decode: odd grids, few classes, K close to H*W
    Hh, Ww, Cc = int(rng.integers(3, 40)), int(rng.integers(3, 40)), int(rng.integers(1, 12))
    K = int(rng.integers(1, min(Hh * Ww, 300) + 1))
    maps = syn.head_maps(3000 + seed, 2, Cc, Hh, Ww)
    got = ops.centernet_decode(*[dev_t(maps[k], cuda) for k in ("heatmap", "offset", "size", "rot", "vel")], K, 2.048)
    ref = orc.decode(maps, score_thresh=0.0, max_detections=K)
    cnt = got["count"].cpu().numpy()
    for b, r in enumerate(ref):
        n = len(r["scores"])
        assert cnt[b] == n
        np.testing.assert_array_equal(got["scores"][b, :n].cpu().numpy(), r["scores"])
        np.testing.assert_allclose(got["boxes"][b, :n].cpu().numpy(), r["boxes"], rtol=0, atol=1e-5)


# ------------------------------------------------------------------------------------------------ N2: dense layers
@pytest.mark.parametrize("B,K,O,relu,bias", [
    (32, 512, 80000, False, True),      # lidar_init.2 at the bench batch: streaming kernel, 32-row batch tile
    (3, 512, 80000, False, True),       # small batch: 8-row tile
    (16, 512, 5000, True, True),        # 16-row tile, last CTA step ragged
    (40, 256, 2048, True, False),       # two batch tiles, no bias
    (5, 128, 1100, False, True),        # the smallest K the streaming kernel takes; rows not a multiple of 64
    (32, 1024, 512, True, True),        # lidar_init.0: one warp per row
    (7, 100, 33, False, True),          # any-shape form
    (1, 256, 256, True, True),          # radar_proj at batch 1
])
def test_dense_layer_vs_oracle(cuda, B, K, O, relu, bias):
    g = np.random.default_rng(B * 1000 + O)
    x = np.abs(g.standard_normal((B, K))).astype(np.float32)
    w, b = syn.linear_weights(B + K + O, K, O)
    got = ops.dense_layer(dev_t(x, cuda), dev_t(w, cuda), dev_t(b, cuda) if bias else None, relu=relu)
    ref = orc.dense_layer(x, w, b if bias else None, relu=relu)
    assert tuple(got.shape) == (B, O)
    assert max_rel(got.cpu().numpy(), ref) < FP32_TOL


def test_lidar_init_vs_reference_golden(cuda, golden):
    """Both layers in one call against what the reference's FlexibleBEVFusion.lidar_init produced."""
    g = golden("bev_glue")
    w1, b1 = syn.linear_weights(701, 1024, 512)
    w2, b2 = syn.linear_weights(702, 512, 128 * 25 * 25)
    feats = syn.global_features(704, 3, 1024)
    out, hid = ops.lidar_init(dev_t(feats, cuda), dev_t(w1, cuda), dev_t(b1, cuda), dev_t(w2, cuda), dev_t(b2, cuda),
                              return_hidden=True)
    assert max_rel(hid.cpu().numpy(), g["lidar_hidden"]) < FP32_TOL
    assert np.abs(out.cpu().numpy()[:, ::16] - g["lidar_init_sub"]).max() < FP32_TOL * float(g["lidar_init_absmax"])
    wr, br = syn.linear_weights(703, 256, 256)
    radar = syn.global_features(705, 3, 256)
    rp = ops.dense_layer(dev_t(radar, cuda), dev_t(wr, cuda), dev_t(br, cuda), relu=True)
    assert max_rel(rp.cpu().numpy(), g["radar_proj"]) < FP32_TOL


@pytest.mark.parametrize("B,K,O,relu,bias", [
    (32, 512, 80000, False, True),      # lidar_init.2 at the bench batch
    (9, 512, 80000, False, True),       # the smallest batch the module sends here: 16-row operand, 7 padding rows
    (64, 512, 12800, True, True),       # 64-row operand, ring of 3
    (70, 64, 128, True, False),         # two launches (64 + 6 rows), one tile, one k block, no bias
    (17, 1024, 640, False, True),       # K = 1024 (lidar_init.0's shape class), 5 tiles
    (1, 256, 256, True, True),          # one row
    (32, 512, 128 * 50 * 50, False, True),   # the stress configuration's lidar_init.2 (lidar_start_size = 50): 655 MB image
])
def test_dense_layer_split_vs_oracle(cuda, B, K, O, relu, bias):
    """The tensor-core dense layer (three fp16 products per fp32 product) against the float64 oracle at the fp32 bound."""
    g = np.random.default_rng(B * 1000 + O)
    x = g.standard_normal((B, K)).astype(np.float32)            # both signs: the scale comes from max|x|
    x[0, :3] = (0.0, 1e-30, -1e-30)
    w, b = syn.linear_weights(B + K + O, K, O)
    w[5] = 0.0                                                  # an all-zero row keeps scale 1
    w[7] *= 1e4                                                 # rows of very different magnitude: the scale is per row
    img = ops.dense_pack_split(dev_t(w, cuda), dev_t(b, cuda) if bias else None)
    assert img is not None and img.numel() == O * K * 4 + 8 * O
    got = ops.dense_layer_split(dev_t(x, cuda), img, O, relu=relu).cpu().numpy()
    assert got.shape == (B, O)
    ref = orc.dense_layer(x, w, b if bias else None, relu=relu)
    keep = np.ones(O, bool)
    keep[7] = False                                             # the big row has its own maximum
    assert max_rel(got[:, keep], ref[:, keep]) < FP32_TOL
    assert max_rel(got[:, 7], ref[:, 7]) < FP32_TOL
    again = ops.dense_layer_split(dev_t(x, cuda), img, O, relu=relu).cpu().numpy()
    np.testing.assert_array_equal(got, again)                   # no atomics, no split-K: the same bits every run


def test_dense_split_shapes_without_a_tensor_core_form(cuda):
    w, b = syn.linear_weights(1, 100, 33)
    assert ops.dense_pack_split(dev_t(w, cuda), dev_t(b, cuda)) is None
    w, b = syn.linear_weights(2, 64, 100)
    assert ops.dense_pack_split(dev_t(w, cuda), dev_t(b, cuda)) is None


def test_lidar_init_split_vs_reference_golden(cuda, golden):
    """lidar_init with its second layer on the tensor cores against what the reference's FlexibleBEVFusion.lidar_init
    produced (3 frames), and against the FFMA kernels on a 32-frame batch."""
    g = golden("bev_glue")
    w1, b1 = syn.linear_weights(701, 1024, 512)
    w2, b2 = syn.linear_weights(702, 512, 128 * 25 * 25)
    feats = syn.global_features(704, 3, 1024)
    d = [dev_t(t, cuda) for t in (w1, b1, w2, b2)]
    img = ops.dense_pack_split(d[2], d[3])
    out = ops.lidar_init_split(dev_t(feats, cuda), d[0], d[1], img, 128 * 25 * 25)
    assert np.abs(out.cpu().numpy()[:, ::16] - g["lidar_init_sub"]).max() < FP32_TOL * float(g["lidar_init_absmax"])
    feats32 = dev_t(syn.global_features(706, 32, 1024), cuda)
    a = ops.lidar_init_split(feats32, d[0], d[1], img, 128 * 25 * 25)
    ref = ops.lidar_init(feats32, *d)
    assert float((a - ref).abs().max()) < FP32_TOL * float(ref.abs().max())


# ------------------------------------------------------------------------------------------------ N1: sigmoid fused into the decode
def test_decode_from_logits_vs_reference_head_golden(cuda, golden):
    g = golden("bev_glue")
    maps = {k: dev_t(g[f"head_{k}"], cuda) for k in ("offset", "size", "rot", "vel")}
    logits = dev_t(g["head_logits"], cuda)
    for tag, thr in (("all", 0.0), ("mid", float(g["head_thresh"]))):
        fused = ops.centernet_decode(logits, maps["offset"], maps["size"], maps["rot"], maps["vel"], 60, 0.512,
                                     score_thresh=thr, heat_is_logit=True)
        # the same launch fed with torch.sigmoid of the same device: identical bits
        two_step = ops.centernet_decode(torch.sigmoid(logits), maps["offset"], maps["size"], maps["rot"], maps["vel"], 60,
                                        0.512, score_thresh=thr)
        for k in ("scores", "boxes", "velocities", "ys", "xs", "ind", "count"):
            assert torch.equal(fused[k], two_step[k]), k
        # and the reference's own decode of its CPU sigmoid: same winners in the same order, scores to an ulp
        for b in range(2):
            n = int(fused["count"][b])
            ref_scores = g[f"head_{tag}_b{b}_scores"]
            assert abs(n - len(ref_scores)) <= (1 if tag == "mid" else 0)     # the threshold IS a score: an ulp decides it
            n = min(n, len(ref_scores))
            np.testing.assert_allclose(fused["scores"][b, :n].cpu().numpy(), ref_scores[:n], rtol=3e-7, atol=0)
            np.testing.assert_allclose(fused["boxes"][b, :n].cpu().numpy(), g[f"head_{tag}_b{b}_boxes"][:n], rtol=0, atol=1e-5)
            np.testing.assert_array_equal(fused["velocities"][b, :n].cpu().numpy(), g[f"head_{tag}_b{b}_velocities"][:n])


# ------------------------------------------------------------------------------------------------ N1: convolution blocks on tcgen05
@pytest.mark.parametrize("B,H,W,Cin,Cout,k,relu", [
    (2, 12, 20, 64, 64, 3, True),        # two pixel tiles, the second ragged
    (3, 50, 50, 128, 320, 3, True),      # the batched head conv: partial channel tile, 32-pixel groups straddling frames
    (2, 25, 25, 512, 256, 1, True),      # 1x1 (camera_proj's second block): HW = 625 is not a multiple of 4 -> scalar stores
    (1, 7, 9, 64, 19, 1, False),         # tiny: one ragged tile, no ReLU (the head's second layer)
    (4, 50, 50, 256, 256, 3, False),     # more tiles than one wave of the ring phases
    (2, 57, 100, 64, 64, 3, True),       # camera_proj's image: tiles of 256 flat indices = 2.5 image rows, starting mid-row
    (1, 9, 150, 64, 32, 3, True),        # a wide map: the pixel block only leaves room for tiles of 176 flat indices
    (1, 3, 300, 64, 32, 3, True),        # an image row wider than a tile: the per-tap kernel takes the 3x3 too
    (5, 5, 6, 64, 130, 3, True),         # whole frames smaller than a tile; a 2-channel last channel tile
    (3, 1, 1, 64, 1, 3, False),          # one pixel, one output channel: every tap but the centre is padding
    (1, 1, 7, 128, 5, 1, True),          # a single image row through the 1x1 form
    (8, 57, 100, 512, 256, 1, True),     # camera_proj's 1x1 at its own size: several pixel tiles per CTA, both co tiles
    (4, 50, 50, 256, 192, 1, False),     # 1x1 with a half-empty second co tile
    (6, 40, 40, 192, 128, 1, True),      # 1x1, three k chunks, one co tile
])
def test_conv_bn_relu_tcgen05(cuda, B, H, W, Cin, Cout, k, relu):
    g = np.random.default_rng(Cin * 7 + Cout)
    x = g.standard_normal((B, Cin, H, W)).astype(np.float32)
    w = (g.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k)).astype(np.float32)
    b = g.standard_normal(Cout).astype(np.float32)
    xd, wd, bd = dev_t(x, cuda), dev_t(w, cuda), dev_t(b, cuda)
    # the layout kernel: split the channels in two parts to exercise the concat form
    half = Cin // 2
    nhwc = ops.nchw_to_nhwc_bf16([xd[:, :half].contiguous(), xd[:, half:].contiguous()])
    assert torch.equal(nhwc, xd.permute(0, 2, 3, 1).to(torch.bfloat16))
    img = ops.conv_pack(wd)
    got = ops.conv_bn_relu_bf16(nhwc, img, bd, Cout, k * k, relu=relu)
    assert tuple(got.shape) == (B, Cout, H, W)
    # (a) the same bf16-rounded operands in float64: only the accumulation order differs
    xr = xd.to(torch.bfloat16).double()
    wr = wd.to(torch.bfloat16).double()
    ref = torch.nn.functional.conv2d(xr, wr, bd.double(), padding=k // 2)
    ref = torch.relu(ref) if relu else ref
    assert max_rel(got.cpu().numpy(), ref.cpu().numpy()) < 2e-5
    # the channels-last bf16 output (the next convolution's input), written into a slice of a wider tensor
    wide = torch.full((B, H, W, Cout + 6), 7.0, dtype=torch.bfloat16, device=cuda)
    both = ops.conv_bn_relu_bf16(nhwc, img, bd, Cout, k * k, relu=relu, out_nhwc=wide, c_offset=4)
    assert torch.equal(both, got)
    assert torch.equal(wide[..., 4:4 + Cout], got.permute(0, 2, 3, 1).to(torch.bfloat16))
    assert bool((wide[..., :4] == 7).all()) and bool((wide[..., 4 + Cout:] == 7).all())
    only = ops.conv_bn_relu_bf16(nhwc, img, bd, Cout, k * k, relu=relu, out_nhwc=torch.empty_like(wide), c_offset=0, want_nchw=False)
    assert only.dtype == torch.bfloat16 and torch.equal(only[..., :Cout], wide[..., 4:4 + Cout])
    # (b) the fp32 convolution the reference runs: the bf16 tolerance of north_star
    ref32 = torch.nn.functional.conv2d(xd.double(), wd.double(), bd.double(), padding=k // 2)
    ref32 = torch.relu(ref32) if relu else ref32
    assert max_rel(got.cpu().numpy(), ref32.cpu().numpy()) < BF16_TOL


CONV_SHAPES = [
    (2, 12, 20, 64, 64, 3, True), (3, 50, 50, 128, 320, 3, True), (2, 25, 25, 512, 256, 1, True), (1, 7, 9, 64, 19, 1, False),
    (4, 50, 50, 256, 256, 3, False), (2, 57, 100, 64, 64, 3, True), (1, 9, 150, 64, 32, 3, True), (1, 3, 300, 64, 32, 3, True), (5, 5, 6, 64, 130, 3, True),
    (3, 1, 1, 64, 1, 3, False), (1, 1, 7, 128, 5, 1, True), (2, 50, 50, 768, 512, 3, True),
]


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,relu", CONV_SHAPES)
@pytest.mark.parametrize("scale", [1.0, 3e-4, 2e3])
def test_conv_bn_relu_fp32_accuracy_on_tcgen05(cuda, B, H, W, Cin, Cout, k, relu, scale):
    """The split-fp16 mode of both convolution kernels (three fp16 products per fp32 product): 1e-5 of max|ref| against the
    float64 convolution, for inputs and weights far from unit scale too (the operand scales are exact powers of two taken
    from the data)."""
    if scale != 1.0 and (B, H, W) not in ((2, 12, 20), (2, 25, 25), (5, 5, 6)):
        pytest.skip("scale sweep on three shapes")
    g = np.random.default_rng(Cin * 7 + Cout)
    x = (g.standard_normal((B, Cin, H, W)) * scale).astype(np.float32)
    w = (g.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k) / scale * 3.0).astype(np.float32)
    w[Cout // 2] *= np.float32(1e-3)                     # an output channel with tiny weights: the scales are per channel
    b = g.standard_normal(Cout).astype(np.float32)
    xd, wd, bd = dev_t(x, cuda), dev_t(w, cuda), dev_t(b, cuda)
    half = Cin // 2 // 8 * 8 or Cin
    parts = [xd[:, :half].contiguous(), xd[:, half:].contiguous()] if half < Cin else [xd]
    x_split, stat = ops.nchw_to_nhwc_split(parts)
    assert tuple(x_split.shape) == (B, H, W, 2 * Cin) and x_split.dtype == torch.float16
    assert float(stat.view(torch.float32)) == float(np.abs(x).max())
    S = 2.0 ** (14 - np.frexp(np.abs(x).max())[1])
    back = (x_split[..., :Cin].double() + x_split[..., Cin:].double()) / S          # hi + lo reproduces x to ~2^-22
    assert max_rel(back.cpu().numpy(), xd.permute(0, 2, 3, 1).cpu().numpy()) < 2e-6
    got = ops.conv_bn_relu_split(x_split, stat, ops.conv_pack_split(wd), bd, Cout, k * k, relu=relu)
    ref = torch.nn.functional.conv2d(xd.double(), wd.double(), bd.double(), padding=k // 2)
    ref = torch.relu(ref) if relu else ref
    assert tuple(got.shape) == (B, Cout, H, W) and max_rel(got.cpu().numpy(), ref.cpu().numpy()) < FP32_TOL
    # per-channel check: the tiny-weight channel is as accurate relative to ITS OWN maximum (its scale is its own)
    c = Cout // 2
    own = float(ref[:, c].abs().max())
    if own > 0:
        assert float((got[:, c].double() - ref[:, c]).abs().max()) < 3 * FP32_TOL * max(own, float(bd[c].abs()))


def test_camera_mean_channels_last_bf16_is_the_rounded_mean(cuda):
    """The fused mean + layout kernel gives exactly bf16(camera_mean): same summation order, IEEE divide."""
    for B, C, h, w, n_cam in ((2, 64, 8, 14, 6), (1, 72, 6, 10, 6), (3, 512, 8, 8, 6), (2, 64, 4, 6, 3)):
        feats = dev_t(syn.camera_features(77 + C, B, n_cam=n_cam, channels=C, h=h, w=w), cuda)
        got = ops.camera_mean_nhwc_bf16(feats)
        want = ops.camera_mean(feats).permute(0, 2, 3, 1).to(torch.bfloat16)
        assert tuple(got.shape) == (B, h, w, C) and torch.equal(got, want)
    with pytest.raises(_lib.B200BevError):
        ops.camera_mean_nhwc_bf16(dev_t(syn.camera_features(5, 1, n_cam=2, channels=64, h=3, w=5), cuda))   # H*W % 4 != 0
    wide = torch.zeros((2, 4, 4, 24), dtype=torch.bfloat16, device=cuda)
    x = torch.rand((2, 8, 4, 4), device=cuda)
    ops.nchw_to_nhwc_bf16([x], out=wide, c_offset=8)
    assert torch.equal(wide[..., 8:16], x.permute(0, 2, 3, 1).to(torch.bfloat16)) and float(wide[..., :8].abs().max()) == 0.0
