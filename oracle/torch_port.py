"""Torch-CPU port of the reference's hot-path stages — TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

The reference (/root/reference) is Python + PyTorch and cannot travel to the GPU box, so the
`cpu_baseline` and `--impl reference` legs of bench.py time this port instead: the same ATen
operators, in the same order, on the same tensor layouts as the reference's forward functions
(F.conv1d k=1 -> F.batch_norm(eval) -> relu -> torch.max; mean -> F.interpolate; max_pool2d ->
== -> * -> topk -> gather -> per-sample python loop).  tests/test_oracle_golden.py checks it against
the golden vectors made from the real reference, so "port" here means op-for-op equal, verified.

Only tests/ and bench.py (cpu_baseline / --impl reference) may import this module.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F


def _t(a) -> torch.Tensor:
    return a if isinstance(a, torch.Tensor) else torch.from_numpy(a)


def layers_to_torch(layers) -> List[Dict[str, torch.Tensor]]:
    return [{k: _t(v) for k, v in lay.items()} for lay in layers]


@torch.no_grad()
def shared_mlp_max(points: torch.Tensor, layers: Sequence[Dict[str, torch.Tensor]]) -> torch.Tensor:
    """(B,N,C) -> (B,C_out).  src/encoders.py:282-298 (lidar), :542-555 (radar)."""
    x = points.transpose(1, 2)                                   # :284 — (B,C,N) view, as the reference feeds conv1d
    for lay in layers:
        x = F.conv1d(x, lay["weight"].unsqueeze(-1), lay["bias"])     # :289 nn.Conv1d(k=1)
        if "bn_var" in lay:
            x = F.batch_norm(x, lay["bn_mean"], lay["bn_var"], lay["bn_weight"], lay["bn_bias"],
                             training=False, eps=1e-5)                # eval-mode BatchNorm1d
        x = F.relu(x)
    return torch.max(x, 2)[0]                                    # :298


@torch.no_grad()
def multi_radar(radar_list: Sequence[torch.Tensor], layers, fc_weight, fc_bias) -> torch.Tensor:
    """src/encoders.py:641-653 (fusion_method 'concat')."""
    feats = torch.stack([shared_mlp_max(r, layers) for r in radar_list], dim=1)
    return F.linear(feats.view(feats.shape[0], -1), fc_weight, fc_bias)


@torch.no_grad()
def camera_mean(feats: torch.Tensor) -> torch.Tensor:
    return feats.mean(dim=1)                                     # src/fusion.py:234


@torch.no_grad()
def bilinear_resize(x: torch.Tensor, size) -> torch.Tensor:
    return F.interpolate(x, size=tuple(size), mode="bilinear", align_corners=False)  # src/fusion.py:242-247


@torch.no_grad()
def cell_index_and_sort(points: torch.Tensor, pc_range, W: int, H: int):
    """Torch statement of the binning conventions (src/centernet_target.py:250-257,285) + stable sort."""
    vx = (pc_range[3] - pc_range[0]) / W
    vy = (pc_range[4] - pc_range[1]) / H
    px = (points[..., 0] - pc_range[0]) / torch.tensor(vx, dtype=torch.float32)
    py = (points[..., 1] - pc_range[1]) / torch.tensor(vy, dtype=torch.float32)
    ok = (px >= 0) & (px < W) & (py >= 0) & (py < H)
    cell = torch.where(ok, py.to(torch.int32) * W + px.to(torch.int32), torch.full_like(px, -1, dtype=torch.int32))
    key = torch.where(cell < 0, torch.full_like(cell, H * W), cell)
    perm = torch.sort(key, dim=-1, stable=True)[1].to(torch.int32)
    return cell, perm


@torch.no_grad()
def camera_project(feats: torch.Tensor, table: torch.Tensor, bev_size) -> torch.Tensor:
    """grid_sample restatement of the geometric gather: feats (B,n_cam,C,h,w), table (H*W,n_cam,3) of
    feature-map (u,v,valid) -> (B,C,H,W).  padding_mode='zeros', align_corners=False."""
    B, n_cam, C, h, w = feats.shape
    H, W = bev_size
    acc = torch.zeros((B, C, H, W), dtype=torch.float32)
    cnt = torch.zeros((1, 1, H, W), dtype=torch.float32)
    for c in range(n_cam):
        u, v, valid = table[:, c, 0], table[:, c, 1], table[:, c, 2]
        gx = (2.0 * u + 1.0) / w - 1.0                          # inverse of grid_sample's unnormalisation
        gy = (2.0 * v + 1.0) / h - 1.0
        grid = torch.stack([gx, gy], dim=-1).view(1, H, W, 2).expand(B, H, W, 2)
        s = F.grid_sample(feats[:, c], grid, mode="bilinear", padding_mode="zeros", align_corners=False)
        m = valid.view(1, 1, H, W)
        acc = acc + s * m
        cnt = cnt + m
    return acc / cnt.clamp(min=1.0)


@torch.no_grad()
def nms(heat: torch.Tensor) -> torch.Tensor:
    hmax = F.max_pool2d(heat, 3, stride=1, padding=1)            # src/centernet_target.py:419
    return heat * (hmax == heat).float()                         # :420-421


@torch.no_grad()
def topk(scores: torch.Tensor, K: int):
    B, C, H, W = scores.shape
    s1, i1 = torch.topk(scores.view(B, C, -1), K, dim=2)         # src/centernet_target.py:432
    cls1 = i1 // (H * W)
    i1 = i1 % (H * W)
    ys1, xs1 = i1 // W, i1 % W
    s2, ind = torch.topk(s1.view(B, -1), K, dim=1)               # :441
    g = lambda a: torch.gather(a.reshape(B, -1), 1, ind)
    return s2, ind, g(cls1), g(ys1), g(xs1)


@torch.no_grad()
def decode(pred: Dict[str, torch.Tensor], score_thresh: float = 0.3, max_detections: int = 100,
           voxel_size: float = 2.048, origin=(-51.2, -51.2)) -> List[Dict[str, torch.Tensor]]:
    """src/centernet_target.py:342-413, per-sample python loop included (it is part of what is timed)."""
    heat = nms(pred["heatmap"])
    scores, _, classes, ys, xs = topk(heat, max_detections)
    out = []
    for b in range(heat.shape[0]):
        m = scores[b] > score_thresh
        if m.sum() == 0:
            out.append({"boxes": torch.zeros(0, 7), "scores": torch.zeros(0),
                        "labels": torch.zeros(0, dtype=torch.long), "velocities": torch.zeros(0, 2)})
            continue
        by, bx = ys[b][m], xs[b][m]
        off = pred["offset"][b, :, by, bx].T
        size = pred["size"][b, :, by, bx].T
        rot = pred["rot"][b, :, by, bx].T
        vel = pred["vel"][b, :, by, bx].T
        wx = (bx.float() + off[:, 0]) * voxel_size + origin[0]
        wy = (by.float() + off[:, 1]) * voxel_size + origin[1]
        wz = torch.zeros_like(wx) - 1.0
        yaw = torch.atan2(rot[:, 0], rot[:, 1])
        out.append({"boxes": torch.stack([wx, wy, wz, size[:, 0], size[:, 1], size[:, 2], yaw], dim=1),
                    "scores": scores[b][m], "labels": classes[b][m], "velocities": vel})
    return out


# ------------------------------------------------------------------------------------------------
# The rest of the inference pass around the hot-path stages: FlexibleBEVFusion.forward,
# CenterNetHead.forward — the same ATen ops in the reference's order, reading parameters from a state_dict with the
# reference's names.  bench.py's reference arm times this next to the stages above; tests pin it to
# tests/golden/detector_chain.npz, which the reference's own modules produced.
# ------------------------------------------------------------------------------------------------
def _conv_bn_relu(x: torch.Tensor, sd: Dict[str, torch.Tensor], conv: str, bn: str, pad: int) -> torch.Tensor:
    x = F.conv2d(x, sd[conv + ".weight"], sd[conv + ".bias"], padding=pad)
    x = F.batch_norm(x, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                     training=False, eps=1e-5)
    return F.relu(x)


@torch.no_grad()
def fusion_forward(sd: Dict[str, torch.Tensor], camera_features=None, lidar_features=None, radar_features=None,
                   bev_hw=(50, 50), prefix: str = "fusion.") -> torch.Tensor:
    """FlexibleBEVFusion.forward, src/fusion.py:209-297 (eval mode)."""
    p = prefix
    parts = []
    if camera_features is not None:
        cam = camera_features.mean(dim=1) if camera_features.dim() == 5 else camera_features          # :233-236
        cam = _conv_bn_relu(cam, sd, p + "camera_proj.0", p + "camera_proj.1", 1)                     # :239
        cam = _conv_bn_relu(cam, sd, p + "camera_proj.3", p + "camera_proj.4", 0)
        parts.append(F.interpolate(cam, size=tuple(bev_hw), mode="bilinear", align_corners=False))    # :242-247
    if lidar_features is not None:
        B = lidar_features.shape[0]
        flat = F.linear(F.relu(F.linear(lidar_features, sd[p + "lidar_init.0.weight"], sd[p + "lidar_init.0.bias"])),
                        sd[p + "lidar_init.2.weight"], sd[p + "lidar_init.2.bias"])                   # :258
        hidden = sd[p + "lidar_upsample.0.weight"].shape[1]
        s = int(round((flat.shape[1] // hidden) ** 0.5))
        x = flat.view(B, hidden, s, s)                                                                # :259
        x = _conv_bn_relu(x, sd, p + "lidar_upsample.0", p + "lidar_upsample.1", 1)                   # :262
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
        parts.append(_conv_bn_relu(x, sd, p + "lidar_upsample.4", p + "lidar_upsample.5", 1))
    if radar_features is not None:
        B = radar_features.shape[0]
        r = F.relu(F.linear(radar_features, sd[p + "radar_proj.0.weight"], sd[p + "radar_proj.0.bias"]))   # :274
        r = r.view(B, -1, 1, 1).expand(B, r.shape[1], bev_hw[0], bev_hw[1])                                # :277-278
        r = _conv_bn_relu(r, sd, p + "radar_refine.0", p + "radar_refine.1", 1)                            # :281
        parts.append(_conv_bn_relu(r, sd, p + "radar_refine.3", p + "radar_refine.4", 1))
    if not parts:
        raise ValueError("No modality features provided")                                            # :289
    x = torch.cat(parts, dim=1)                                                                      # :292
    x = _conv_bn_relu(x, sd, p + "bev_fusion.0", p + "bev_fusion.1", 1)                               # :295
    return _conv_bn_relu(x, sd, p + "bev_fusion.3", p + "bev_fusion.4", 1)


@torch.no_grad()
def head_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, prefix: str = "det_head.") -> Dict[str, torch.Tensor]:
    """CenterNetHead.forward, src/fusion.py:869-884."""
    out = {}
    for name in ("heatmap", "offset", "size", "rot", "vel"):
        h = F.relu(F.conv2d(x, sd[f"{prefix}{name}_head.0.weight"], sd[f"{prefix}{name}_head.0.bias"], padding=1))
        out[name] = F.conv2d(h, sd[f"{prefix}{name}_head.2.weight"], sd[f"{prefix}{name}_head.2.bias"])
    out["heatmap"] = torch.sigmoid(out["heatmap"])
    return out


def mlp_layers_from_state(sd: Dict[str, torch.Tensor], prefix: str) -> List[Dict[str, torch.Tensor]]:
    """conv{i}/bn{i} entries of a PointNet-style encoder -> the layer dicts shared_mlp_max takes."""
    layers, i = [], 1
    while f"{prefix}conv{i}.weight" in sd:
        layers.append({"weight": sd[f"{prefix}conv{i}.weight"].squeeze(-1), "bias": sd[f"{prefix}conv{i}.bias"],
                       "bn_weight": sd[f"{prefix}bn{i}.weight"], "bn_bias": sd[f"{prefix}bn{i}.bias"],
                       "bn_mean": sd[f"{prefix}bn{i}.running_mean"], "bn_var": sd[f"{prefix}bn{i}.running_var"]})
        i += 1
    return layers


@torch.no_grad()
def detector_chain(sd: Dict[str, torch.Tensor], camera_features=None, lidar_points=None, radar_list=None, bev_hw=(50, 50),
                   score_thresh: float = 0.0, max_detections: int = 100, voxel_size: float = 0.512):
    """FlexibleMultiModal3DDetector.forward without the camera backbone (src/fusion.py:1113-1137: encoders -> fusion ->
    det_head) followed by decode_centernet_predictions as eval.py calls it (src/eval.py:58-62; the fusion_detection copy,
    0.512 m per cell).  Returns (fused BEV features, head dict, detections)."""
    lidar_feat = radar_feat = None
    if lidar_points is not None:
        lidar_feat = shared_mlp_max(lidar_points, mlp_layers_from_state(sd, "lidar_encoder."))
    if radar_list is not None:
        radar_feat = multi_radar(radar_list, mlp_layers_from_state(sd, "radar_encoder.radar_encoder."),
                                 sd["radar_encoder.fusion_fc.weight"], sd["radar_encoder.fusion_fc.bias"])
    bev = fusion_forward(sd, camera_features, lidar_feat, radar_feat, bev_hw)
    pred = head_forward(sd, bev)
    return bev, pred, decode(pred, score_thresh, max_detections, voxel_size)


@torch.no_grad()
def cell_canvas(points: torch.Tensor, layers, cell: torch.Tensor, n_cells: int, with_global: bool = False):
    """north_star S1 on the CPU: per-point features (what the reference exposes with return_point_features=True,
    src/encoders.py:300-304) scatter-max'ed into the (B, n_cells, C) canvas; one frame at a time (the per-point
    tensor of a frame is 143 MB at 35,000 x 1024)."""
    out, glob = [], []
    for b in range(points.shape[0]):
        x = points[b:b + 1].transpose(1, 2)
        for lay in layers:
            x = F.conv1d(x, lay["weight"].unsqueeze(-1), lay["bias"])
            if "bn_var" in lay:
                x = F.batch_norm(x, lay["bn_mean"], lay["bn_var"], lay["bn_weight"], lay["bn_bias"], training=False, eps=1e-5)
            x = F.relu(x)
        glob.append(torch.max(x, 2)[0][0])                                # src/encoders.py:298 — the reference's output
        feat = x[0].t()                                                   # (N, C)
        ok = cell[b] >= 0
        canvas = torch.zeros((n_cells, feat.shape[1]), dtype=feat.dtype)
        idx = cell[b][ok].long().unsqueeze(1).expand(-1, feat.shape[1])
        canvas.scatter_reduce_(0, idx, feat[ok], reduce="amax", include_self=True)   # features are >= 0 (ReLU): zeros = empty
        out.append(canvas)
    return (torch.stack(out), torch.stack(glob)) if with_global else torch.stack(out)
