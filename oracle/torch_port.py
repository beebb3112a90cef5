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
