"""CPU oracle for the BEV encode + decode hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module, and only as the checker.  The product path (bevfusion_multimodal_3d_object_detection_b200)
never imports it and has no CPU fallback.

Plain numpy float32 restatement of the reference's arithmetic, function by function; every
function cites the reference lines it follows (paths relative to the reference checkout).

Parity pinning: the reference has no tests or golden vectors for this path.  The
functions marked [pinned] are checked in tests/test_oracle_golden.py against outputs of the
reference's own modules, imported unmodified from /root/reference/src by tests/golden/make_golden.py
and committed as tests/golden/*.npz.  The functions marked [unpinned] have no counterpart in the
reference (cell binning, per-cell scatter-max, geometric camera projection — SURVEY §0): they state
the conventions the reference uses elsewhere and are cross-checked against torch library ops
(scatter_reduce 'amax', grid_sample) in the golden generator, but the reference itself cannot pin
them.  PARITY UNPINNED applies to exactly those three.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32
BN_EPS = F32(1e-5)
NEAR_PLANE = F32(0.1)


# ------------------------------------------------------------------------------------------------
# S1 — shared MLP + max
# ------------------------------------------------------------------------------------------------
def shared_mlp(points: np.ndarray, layers: Sequence[Dict[str, np.ndarray]]) -> np.ndarray:
    """[pinned] Per-point features (N, C_out): x = relu(bn(conv1d_k1(x))) per layer, BatchNorm in eval
    mode — PointNetLiDAREncoder.forward src/encoders.py:289-295, RadarEncoder.forward src/encoders.py:549-552.
    A k=1 Conv1d is a matmul with the (C_out, C_in) weight (src/encoders.py:252-256)."""
    x = np.asarray(points, dtype=F32)
    for lay in layers:
        x = x @ lay["weight"].T.astype(F32) + lay["bias"].astype(F32)
        if "bn_var" in lay:
            inv = F32(1.0) / np.sqrt(lay["bn_var"].astype(F32) + BN_EPS)
            x = (x - lay["bn_mean"].astype(F32)) * inv * lay["bn_weight"].astype(F32) + lay["bn_bias"].astype(F32)
        x = np.maximum(x, F32(0.0))
    return x.astype(F32)


def pointnet_global(points: np.ndarray, layers) -> np.ndarray:
    """[pinned] (B,N,C) -> (B,C_out): torch.max(x, 2)[0] over ALL rows, zero padding included
    (src/encoders.py:298)."""
    return np.stack([shared_mlp(p, layers).max(axis=0) for p in points])


def fold_layers(layers) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """BN-folded (W', b') in float64 — the transformation the product applies on the host; the oracle
    itself never uses it for checking (it evaluates the unfolded form above)."""
    ws, bs = [], []
    for lay in layers:
        w = lay["weight"].astype(np.float64)
        b = lay["bias"].astype(np.float64)
        if "bn_var" in lay:
            s = lay["bn_weight"].astype(np.float64) / np.sqrt(lay["bn_var"].astype(np.float64) + 1e-5)
            w = w * s[:, None]
            b = (b - lay["bn_mean"].astype(np.float64)) * s + lay["bn_bias"].astype(np.float64)
        ws.append(w)
        bs.append(b)
    return ws, bs


def multi_radar(radar_list: Sequence[np.ndarray], layers, fusion: str = "concat",
                fc_weight: Optional[np.ndarray] = None, fc_bias: Optional[np.ndarray] = None):
    """[pinned] MultiRadarEncoder.forward src/encoders.py:641-659: shared encoder per radar, stack to
    (B,R,F), then concat->Linear / max / mean.  Returns (fused (B,F), stacked (B,R,F))."""
    stacked = np.stack([pointnet_global(r, layers) for r in radar_list], axis=1)  # src/encoders.py:647
    B = stacked.shape[0]
    if fusion == "concat":
        fused = stacked.reshape(B, -1) @ fc_weight.T.astype(F32) + fc_bias.astype(F32)  # :652-653
    elif fusion == "max":
        fused = stacked.max(axis=1)  # :655
    elif fusion == "mean":
        fused = stacked.mean(axis=1, dtype=F32)  # :657
    else:
        raise ValueError(f"Unknown fusion method: {fusion}")  # :659
    return fused.astype(F32), stacked


# ------------------------------------------------------------------------------------------------
# S1 — cell binning, sort, per-cell max   [unpinned: conventions only]
# ------------------------------------------------------------------------------------------------
def voxel_size(pc_range, W: int, H: int) -> Tuple[np.float32, np.float32]:
    """voxel = (max - min) / cells, src/centernet_target.py:222-224 (float64), rounded once to fp32."""
    x_min, y_min, _, x_max, y_max, _ = pc_range
    return F32((float(x_max) - float(x_min)) / W), F32((float(y_max) - float(y_min)) / H)


def lidar_prepare(raw: np.ndarray, max_points: int, pc_range, indices: Optional[np.ndarray] = None) -> Tuple[np.ndarray, int]:
    """[pinned] One sweep (M, C) -> ((max_points, C), n_in_range): NuScenesDataset._load_lidar_points and
    _pad_or_subsample, src/train_detect.py:152-158,181-189.  Strict open-interval mask on x, y, z; points[mask];
    zero rows up to max_points, or points[indices] when N >= max_points — `indices` being the draw the reference
    takes from np.random.choice(N, max_points, replace=False) (None: the first max_points, for want of a draw)."""
    p = np.asarray(raw, dtype=F32)
    with np.errstate(invalid="ignore"):
        mask = ((p[:, 0] > F32(pc_range[0])) & (p[:, 0] < F32(pc_range[3])) & (p[:, 1] > F32(pc_range[1])) &
                (p[:, 1] < F32(pc_range[4])) & (p[:, 2] > F32(pc_range[2])) & (p[:, 2] < F32(pc_range[5])))
    kept = p[mask]
    n = kept.shape[0]
    if n >= max_points:
        out = kept[indices] if indices is not None else kept[:max_points]
    else:
        out = np.concatenate([kept, np.zeros((max_points - n, p.shape[1]), dtype=F32)], axis=0)
    return out.astype(F32), n


def cell_index(points: np.ndarray, pc_range, W: int, H: int) -> np.ndarray:
    """[unpinned] (..., C) -> int32 cell ids iy*W+ix, -1 outside.  px=(x-x_min)/voxel; reject px<0 or
    px>=W; ix=int(px); flat=iy*W+ix — src/centernet_target.py:250-257,285, evaluated in fp32."""
    vx, vy = voxel_size(pc_range, W, H)
    x = np.asarray(points[..., 0], dtype=F32)
    y = np.asarray(points[..., 1], dtype=F32)
    with np.errstate(invalid="ignore"):
        px = (x - F32(pc_range[0])) / vx
        py = (y - F32(pc_range[1])) / vy
        ok = (px >= 0) & (px < F32(W)) & (py >= 0) & (py < F32(H))
    ix = np.where(ok, px, 0).astype(np.int32)
    iy = np.where(ok, py, 0).astype(np.int32)
    return np.where(ok, iy * np.int32(W) + ix, np.int32(-1)).astype(np.int32)


def bin_sort(cell: np.ndarray, n_cells: int) -> Tuple[np.ndarray, np.ndarray]:
    """[unpinned] cell (N,) -> (perm (N,), offsets (n_cells+1,)): stable counting sort by cell, the
    out-of-grid points (cell -1) last, in index order."""
    key = np.where(cell < 0, n_cells, cell)
    perm = np.argsort(key, kind="stable").astype(np.int32)
    counts = np.bincount(key, minlength=n_cells + 1)[:n_cells]
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return perm, offsets


def pointnet_cell_max(points: np.ndarray, layers, cell: np.ndarray, n_cells: int) -> np.ndarray:
    """[unpinned] (N,C) -> canvas (n_cells, C_out): per-cell max of the reference's per-point features
    (the first C_out columns of return_point_features, src/encoders.py:300-304), 0 for empty cells —
    what torch.scatter_reduce_('amax', include_self=False) gives on a zero canvas."""
    feat = shared_mlp(points, layers)
    canvas = np.zeros((n_cells, feat.shape[1]), dtype=F32)
    ok = cell >= 0
    np.maximum.at(canvas, cell[ok], feat[ok])
    return canvas


# ------------------------------------------------------------------------------------------------
# S2 — camera -> BEV
# ------------------------------------------------------------------------------------------------
def camera_mean(feats: np.ndarray) -> np.ndarray:
    """[pinned] camera_features.mean(dim=1), src/fusion.py:233-234: fp32 sum in camera order, / n_cam."""
    s = feats[:, 0].astype(F32).copy()
    for c in range(1, feats.shape[1]):
        s += feats[:, c]
    return (s / F32(feats.shape[1])).astype(F32)


def _resize_axis(out_size: int, in_size: int):
    """aten area_pixel_compute_source_index + guard_index_and_lambda, align_corners=False."""
    scale = F32(in_size) / F32(out_size)
    dst = np.arange(out_size, dtype=F32)
    src = scale * (dst + F32(0.5)) - F32(0.5)
    src = np.maximum(src, F32(0.0))
    i0 = np.minimum(np.floor(src).astype(np.int64), in_size - 1)
    i1 = np.minimum(i0 + 1, in_size - 1)
    l1 = np.clip(src - i0.astype(F32), F32(0.0), F32(1.0)).astype(F32)
    l0 = (F32(1.0) - l1).astype(F32)
    return i0, i1, l0, l1


def bilinear_resize(x: np.ndarray, size: Tuple[int, int]) -> np.ndarray:
    """[pinned] F.interpolate(x, size, mode='bilinear', align_corners=False), src/fusion.py:242-247.
    x (B,C,h,w) -> (B,C,H,W)."""
    H, W = size
    y0, y1, ly0, ly1 = _resize_axis(H, x.shape[2])
    x0, x1, lx0, lx1 = _resize_axis(W, x.shape[3])
    x = x.astype(F32)
    top = x[:, :, y0][:, :, :, x0] * lx0 + x[:, :, y0][:, :, :, x1] * lx1
    bot = x[:, :, y1][:, :, :, x0] * lx0 + x[:, :, y1][:, :, :, x1] * lx1
    return (top * ly0[:, None] + bot * ly1[:, None]).astype(F32)


def project_cells(intrinsics: np.ndarray, ego2cam: np.ndarray, img_size, feat_size, bev_size,
                  pc_range, z_plane: float = 0.0) -> np.ndarray:
    """[unpinned] Table (H*W, n_cam, 3) of (u, v, valid) in feature-map coordinates for the centres of
    the BEV cells: p_cam = R p_ego + t, pinhole K, valid = in front of the near plane and inside the
    image; pixel -> feature coordinate by u = U*(w/img_w) - 0.5 (grid_sample align_corners=False).
    Calibration layout: translation/rotation/camera_intrinsic per camera, src/data_converter.py:110-117.
    One correctly-rounded fp32 operation at a time, in the order the kernel uses."""
    img_w, img_h = F32(img_size[0]), F32(img_size[1])
    h, w = feat_size
    H, W = bev_size
    vx, vy = voxel_size(pc_range, W, H)
    ix = np.tile(np.arange(W, dtype=F32), H)
    iy = np.repeat(np.arange(H, dtype=F32), W)
    X = F32(pc_range[0]) + (ix + F32(0.5)) * vx
    Y = F32(pc_range[1]) + (iy + F32(0.5)) * vy
    Z = F32(z_plane)
    K = intrinsics.astype(F32)
    E = ego2cam.astype(F32)
    n_cam = K.shape[0]
    out = np.zeros((H * W, n_cam, 3), dtype=F32)
    sx, sy = F32(w) / img_w, F32(h) / img_h
    for c in range(n_cam):
        pc = []
        for r in range(3):
            s = (E[c, r, 0] * X + E[c, r, 1] * Y) + E[c, r, 2] * Z
            pc.append((s + E[c, r, 3]).astype(F32))
        front = pc[2] > NEAR_PLANE
        zs = np.where(front, pc[2], F32(1.0))
        xn, yn = pc[0] / zs, pc[1] / zs
        U = (K[c, 0, 0] * xn + K[c, 0, 1] * yn) + K[c, 0, 2]
        V = (K[c, 1, 0] * xn + K[c, 1, 1] * yn) + K[c, 1, 2]
        valid = front & (U >= 0) & (U < img_w) & (V >= 0) & (V < img_h)
        out[:, c, 0] = U * sx - F32(0.5)
        out[:, c, 1] = V * sy - F32(0.5)
        out[:, c, 2] = valid
    return out


def camera_project(feats: np.ndarray, table: np.ndarray, bev_size) -> np.ndarray:
    """[unpinned] feats (n_cam,C,h,w) + table from project_cells -> canvas (C,H,W): bilinear sample with
    zeros outside the map, mean over the cameras that see the cell, 0 where none does."""
    n_cam, C, h, w = feats.shape
    H, W = bev_size
    acc = np.zeros((H * W, C), dtype=F32)
    cnt = np.zeros(H * W, dtype=F32)
    for c in range(n_cam):
        u, v, valid = table[:, c, 0], table[:, c, 1], table[:, c, 2] > 0
        fx, fy = np.floor(u), np.floor(v)
        x0, y0 = fx.astype(np.int64), fy.astype(np.int64)
        ax, ay = (u - fx).astype(F32), (v - fy).astype(F32)
        bx, by = ((fx + F32(1.0)) - u).astype(F32), ((fy + F32(1.0)) - v).astype(F32)
        val = np.zeros((H * W, C), dtype=F32)
        for dy, dx, wgt in ((0, 0, bx * by), (0, 1, ax * by), (1, 0, bx * ay), (1, 1, ax * ay)):
            xx, yy = x0 + dx, y0 + dy
            ok = valid & (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
            tap = feats[c][:, np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)].T  # (HW, C)
            val = val + np.where(ok, wgt, F32(0.0))[:, None] * tap
        acc = acc + np.where(valid[:, None], val, F32(0.0))
        cnt = cnt + valid.astype(F32)
    out = acc / np.maximum(cnt, F32(1.0))[:, None]
    return out.T.reshape(C, H, W).astype(F32)


# ------------------------------------------------------------------------------------------------
# N2 — dense layers around the canvas
# ------------------------------------------------------------------------------------------------
def dense_layer(x: np.ndarray, weight: np.ndarray, bias: Optional[np.ndarray] = None, relu: bool = False) -> np.ndarray:
    """[pinned] nn.Linear (+ nn.ReLU): x @ W^T + b — FlexibleBEVFusion.radar_proj src/fusion.py:170-173 and each layer of
    lidar_init src/fusion.py:144-148."""
    y = np.asarray(x, dtype=F32) @ np.asarray(weight, dtype=F32).T
    if bias is not None:
        y = y + np.asarray(bias, dtype=F32)
    return np.maximum(y, F32(0.0)) if relu else y.astype(F32)


def lidar_init(feats: np.ndarray, w1, b1, w2, b2) -> np.ndarray:
    """[pinned] Linear + ReLU + Linear, src/fusion.py:144-148 (applied at :258)."""
    return dense_layer(dense_layer(feats, w1, b1, relu=True), w2, b2)


def conv_bn_relu(x: np.ndarray, weight: np.ndarray, bias: Optional[np.ndarray], bn: Optional[Dict[str, np.ndarray]] = None,
                 relu: bool = True) -> np.ndarray:
    """[pinned] nn.Conv2d(k=3, padding=1 | k=1) -> nn.BatchNorm2d (eval) -> nn.ReLU, the block every stack of
    FlexibleBEVFusion is made of (camera_proj src/fusion.py:126-133, bev_fusion :199-207) and CenterNetHead's layers
    (:822-854).  x (B,C,H,W), weight (O,C,k,k); bn = dict(weight, bias, running_mean, running_var) or None."""
    x = np.asarray(x, dtype=F32)
    B, C, H, W = x.shape
    O, _, k, _ = weight.shape
    pad = k // 2
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    y = np.zeros((B, O, H, W), dtype=F32)
    for ky in range(k):
        for kx in range(k):
            y += np.einsum("oc,bchw->bohw", weight[:, :, ky, kx].astype(F32), xp[:, :, ky:ky + H, kx:kx + W], optimize=True).astype(F32)
    if bias is not None:
        y = y + bias.astype(F32)[None, :, None, None]
    if bn is not None:
        inv = F32(1.0) / np.sqrt(bn["running_var"].astype(F32) + BN_EPS)
        y = (y - bn["running_mean"].astype(F32)[None, :, None, None]) * (inv * bn["weight"].astype(F32))[None, :, None, None] \
            + bn["bias"].astype(F32)[None, :, None, None]
    return np.maximum(y, F32(0.0)) if relu else y.astype(F32)


def sigmoid(x: np.ndarray) -> np.ndarray:
    """[pinned] torch.sigmoid of the heat-map head, src/fusion.py:870-871, in fp32."""
    x = np.asarray(x, dtype=F32)
    return (F32(1.0) / (F32(1.0) + np.exp(-x, dtype=F32))).astype(F32)


# ------------------------------------------------------------------------------------------------
# S3 — CenterNet decode
# ------------------------------------------------------------------------------------------------
def nms(heat: np.ndarray) -> np.ndarray:
    """[pinned] _nms src/centernet_target.py:416-421 (= src/fusion_detection.py:784-789):
    hmax = max_pool2d(heat, 3, stride 1, pad 1) with -inf padding; heat * (hmax == heat)."""
    B, C, H, W = heat.shape
    pad = np.full((B, C, H + 2, W + 2), -np.inf, dtype=F32)
    pad[:, :, 1:-1, 1:-1] = heat
    hmax = heat.copy()
    for dy in range(3):
        for dx in range(3):
            hmax = np.maximum(hmax, pad[:, :, dy:dy + H, dx:dx + W])
    return (heat * (hmax == heat).astype(F32)).astype(F32)


def _topk_desc(values: np.ndarray, K: int) -> Tuple[np.ndarray, np.ndarray]:
    """Top-K along the last axis, ties by ascending index (torch.topk leaves tie order undefined)."""
    if K > values.shape[-1]:
        raise RuntimeError("selected index k out of range")  # torch.topk, SURVEY Q7
    order = np.argsort(-values, axis=-1, kind="stable")[..., :K]
    return np.take_along_axis(values, order, axis=-1), order


def topk(scores: np.ndarray, K: int):
    """[pinned on tie-free positive entries] _topk src/centernet_target.py:424-452 (= fusion_detection.py:792-820).
    Returns (topk_score (B,K) f32, topk_ind, topk_classes, topk_ys, topk_xs (B,K) i64)."""
    B, C, H, W = scores.shape
    s1, i1 = _topk_desc(scores.reshape(B, C, -1), K)      # :429-432
    cls1 = i1 // (H * W)                                  # :434 — always 0
    i1 = i1 % (H * W)                                     # :435
    ys1, xs1 = i1 // W, i1 % W                            # :436-437
    s2, ind = _topk_desc(s1.reshape(B, -1), K)            # :440-441
    gather = lambda a: np.take_along_axis(a.reshape(B, -1), ind, axis=1)
    return s2.astype(F32), ind.astype(np.int64), gather(cls1).astype(np.int64), gather(ys1).astype(np.int64), \
        gather(xs1).astype(np.int64)


def decode(pred: Dict[str, np.ndarray], score_thresh: float = 0.3, max_detections: int = 100,
           voxel_size_m: float = 2.048, pc_origin=(-51.2, -51.2)) -> List[Dict[str, np.ndarray]]:
    """[pinned] decode_centernet_predictions src/centernet_target.py:326-413 (voxel 2.048, :389) and its
    copy src/fusion_detection.py:695-781 (voxel 0.512, :757).  One dict per sample."""
    heat = nms(pred["heatmap"].astype(F32))                                   # :351
    scores, _, classes, ys, xs = topk(heat, max_detections)                   # :354
    out = []
    for b in range(heat.shape[0]):                                            # :358
        m = scores[b] > F32(score_thresh)                                     # :360
        if m.sum() == 0:                                                      # :362-369
            out.append({"boxes": np.zeros((0, 7), F32), "scores": np.zeros(0, F32),
                        "labels": np.zeros(0, np.int64), "velocities": np.zeros((0, 2), F32)})
            continue
        by, bx = ys[b][m], xs[b][m]
        off = pred["offset"][b][:, by, bx].T.astype(F32)                      # :377
        size = pred["size"][b][:, by, bx].T.astype(F32)                       # :378
        rot = pred["rot"][b][:, by, bx].T.astype(F32)                         # :379
        vel = pred["vel"][b][:, by, bx].T.astype(F32)                         # :380
        cx = bx.astype(F32) + off[:, 0]                                       # :383
        cy = by.astype(F32) + off[:, 1]                                       # :384
        wx = cx * F32(voxel_size_m) + F32(pc_origin[0])                       # :392
        wy = cy * F32(voxel_size_m) + F32(pc_origin[1])                       # :393
        wz = np.zeros_like(wx) - F32(1.0)                                     # :394
        yaw = np.arctan2(rot[:, 0], rot[:, 1]).astype(F32)                    # :397
        boxes = np.stack([wx, wy, wz, size[:, 0], size[:, 1], size[:, 2], yaw], axis=1).astype(F32)  # :400-404
        out.append({"boxes": boxes, "scores": scores[b][m], "labels": classes[b][m], "velocities": vel})
    return out


# ------------------------------------------------------------------------------------------------
# N4 — the consumer of the decode outputs: mAP / NDS
# ------------------------------------------------------------------------------------------------
CLASS_NAMES = ("car", "truck", "bus", "trailer", "construction_vehicle", "pedestrian", "motorcycle", "bicycle",
               "traffic_cone", "barrier")


def _greedy_assign(dist: np.ndarray, scores: np.ndarray, threshold: float) -> List[Tuple[int, int]]:
    """Predictions in descending score order each take the nearest still-free ground-truth box if it is within
    `threshold` metres (centre distance) — the matching rule of src/utils_v2.py:14-37 and :52-71.  Returns
    [(rank in score order, prediction index, gt index or -1)]."""
    order = np.argsort(-scores)
    free = np.ones(dist.shape[1], dtype=bool)
    out = []
    for rank, p in enumerate(order):
        g = -1
        if free.any():
            d = np.where(free, dist[p], np.inf)
            best = int(np.argmin(d))
            if d[best] <= threshold:
                g = best
                free[best] = False
        out.append((rank, int(p), g))
    return out


def compute_metrics(predictions: Sequence[Dict[str, np.ndarray]], ground_truths: Sequence[Dict[str, np.ndarray]],
                    threshold: float = 2.0) -> Dict:
    """[pinned] utils_v2.compute_metrics, src/utils_v2.py:94-205: per sample and class an 11-point interpolated AP over the
    greedy centre-distance matching (:42-88), mean over samples then classes = mAP; translation / scale / orientation
    errors of the matched pairs; NDS = mean(5 mAP, 1 - min(mATE/4, 1), 1 - min(mASE, 1), 1 - min(mAOE/pi, 1)).
    Inputs are the shim's per-sample dicts (boxes (n,7), scores, labels) as numpy arrays."""
    n_cls = len(CLASS_NAMES)
    aps = [[] for _ in range(n_cls)]
    ate, ase, aoe = [], [], []
    for pred, gt in zip(predictions, ground_truths):
        pb, ps, pl = (np.asarray(pred[k]) for k in ("boxes", "scores", "labels"))
        gb, gl = np.asarray(gt["boxes"]), np.asarray(gt["labels"])
        keep = gl >= 0
        gb, gl = gb[keep], gl[keep]
        if len(gb) == 0 and len(pb) == 0:
            continue
        for c in range(n_cls):
            cp, cs, cg = pb[pl == c], ps[pl == c], gb[gl == c]
            if len(cp) == 0 and len(cg) == 0:
                continue
            if len(cp) == 0 or len(cg) == 0:
                aps[c].append(0.0)
                continue
            dist = np.sqrt(((cp[:, None, :2] - cg[None, :, :2]) ** 2).sum(axis=2))
            assign = _greedy_assign(dist, cs, threshold)
            hit = np.array([1.0 if g >= 0 else 0.0 for _, _, g in assign])
            tp, fp = np.cumsum(hit), np.cumsum(1.0 - hit)
            recall, precision = tp / len(cg), tp / (tp + fp + 1e-10)
            ap = 0.0
            for t in np.linspace(0, 1, 11):
                ok = precision[recall >= t]
                ap += (ok.max() if len(ok) else 0) / 11.0
            aps[c].append(ap)
            for _, p, g in assign:
                if g < 0:
                    continue
                a, b = cp[p], cg[g]
                ate.append(np.linalg.norm(a[:2] - b[:2]))
                ase.append(np.mean(np.abs(a[3:6] - b[3:6]) / (b[3:6] + 1e-6)))
                d = a[6] - b[6]
                aoe.append(abs(np.arctan2(np.sin(d), np.cos(d))))
    class_ap = [float(np.mean(v)) if v else 0.0 for v in aps]
    m_ap = float(np.mean(class_ap))
    m_ate, m_ase, m_aoe = (float(np.mean(v)) if v else 1.0 for v in (ate, ase, aoe))
    nds = float(np.mean([5 * m_ap, 1 - min(m_ate / 4.0, 1.0), 1 - min(m_ase, 1.0), 1 - min(m_aoe / np.pi, 1.0)]))
    return {"mAP": m_ap, "NDS": nds, "AP_per_class": dict(zip(CLASS_NAMES, class_ap))}
