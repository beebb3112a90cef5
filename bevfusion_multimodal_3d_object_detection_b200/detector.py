"""The inference chain behind the camera backbone, with the reference's sub-module names.

`FlexibleMultiModal3DDetector.forward` (src/fusion.py:1090-1137) is

    camera_encoder -> lidar_encoder -> radar_encoder -> fusion -> det_head            [-> decode, in the caller]

The ResNet camera backbone is outside this package's path (dense cuDNN, fed as *features* by
BASELINE.json's configs), so `BEVDetectorChain` is that forward from the camera FEATURES on: same attribute names
(`lidar_encoder`, `radar_encoder`, `fusion`, `det_head`), hence the same state_dict keys as the reference's detector for
those four sub-modules — `load_state_dict(ckpt['model_state_dict'], strict=False)` of a reference checkpoint fills it
(src/eval.py:208-210 loads with strict=False too).  It is what bench.py steps and what the chain tests check against
tests/golden/detector_chain.npz; with the reference's own detector, `patch()` gives the same code path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import conv_blocks, ops
from .centernet_decode import FUSION_DETECTION_VOXEL, GROUND_Z, PC_ORIGIN, decode_centernet_predictions
from .encoders import MultiRadarEncoder, PointNetLiDAREncoder
from .fusion import CenterNetHead, FlexibleBEVFusion


class BEVDetectorChain(nn.Module):
    def __init__(self, use_camera: bool = True, use_lidar: bool = True, use_radar: bool = True, camera_channels: int = 512,
                 lidar_input_channels: int = 4, radar_input_channels: int = 7, num_radars: int = 5, bev_h: int = 50,
                 bev_w: int = 50, bev_channels: int = 256, num_classes: int = 10, head_conv: int = 64,
                 lidar_start_size: Optional[int] = None, precision: Optional[str] = None):
        super().__init__()
        self.use_camera, self.use_lidar, self.use_radar = use_camera, use_lidar, use_radar
        if use_lidar:
            self.lidar_encoder = PointNetLiDAREncoder(input_channels=lidar_input_channels, feat_dim=1024)
        if use_radar:
            self.radar_encoder = MultiRadarEncoder(input_channels=radar_input_channels, feat_dim=256, num_radars=num_radars,
                                                   fusion_method="concat")
        self.fusion = FlexibleBEVFusion(use_camera=use_camera, use_lidar=use_lidar, use_radar=use_radar,
                                        camera_channels=camera_channels, lidar_channels=1024, radar_channels=256, bev_h=bev_h,
                                        bev_w=bev_w, bev_channels=bev_channels, lidar_start_size=lidar_start_size)
        self.det_head = CenterNetHead(in_channels=bev_channels, num_classes=num_classes, head_conv=head_conv)
        self.set_precision(precision)

    def set_precision(self, precision: Optional[str]) -> "BEVDetectorChain":
        """None | "f32" | "bf16" on every sub-module (None: B200BEV_PRECISION, else f32)."""
        for m in self.children():
            m.b200_precision = precision
            if isinstance(m, MultiRadarEncoder):
                m.radar_encoder.b200_precision = precision
        return self

    def state_shapes(self) -> Dict[str, Tuple[int, ...]]:
        return {k: tuple(v.shape) for k, v in self.state_dict().items()}

    def forward(self, camera_features: Optional[torch.Tensor] = None, lidar_points: Optional[torch.Tensor] = None,
                radar_points: Optional[Sequence[torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """src/fusion.py:1109-1137 with `camera_features` in place of `camera_encoder(camera_imgs)`."""
        cam = camera_features if self.use_camera else None
        lidar_feat = self.lidar_encoder(lidar_points) if self.use_lidar and lidar_points is not None else None
        radar_feat = self.radar_encoder(radar_points) if self.use_radar and radar_points is not None else None
        return self.det_head(self.fusion(camera_features=cam, lidar_features=lidar_feat, radar_features=radar_feat))

    def detect(self, camera_features=None, lidar_points=None, radar_points=None, score_thresh: float = 0.3,
               max_detections: int = 100, voxel_size: float = FUSION_DETECTION_VOXEL) -> List[Dict[str, torch.Tensor]]:
        """forward + decode_centernet_predictions as eval.py calls it (src/eval.py:53-62): a list of per-sample dicts."""
        pred = self.forward(camera_features, lidar_points, radar_points)
        return decode_centernet_predictions(pred, score_thresh, max_detections, voxel_size)

    def detect_fixed(self, camera_features=None, lidar_points=None, radar_points=None, score_thresh: float = 0.0,
                     max_detections: int = 100, voxel_size: float = FUSION_DETECTION_VOXEL) -> Dict[str, torch.Tensor]:
        """forward + the decode kernel's fixed-size outputs (boxes (B,K,7), scores, labels, velocities, count): no host
        sync, so the whole call can be captured in a CUDA graph (runtime.GraphedStep)."""
        pred = self.forward(camera_features, lidar_points, radar_points)
        logits = conv_blocks.logits_of(pred["heatmap"])
        return ops.centernet_decode(pred["heatmap"] if logits is None else logits, pred["offset"], pred["size"], pred["rot"],
                                    pred["vel"], max_detections, voxel_size, PC_ORIGIN, GROUND_Z, score_thresh,
                                    heat_is_logit=logits is not None)
