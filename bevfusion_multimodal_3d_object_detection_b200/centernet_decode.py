"""CenterNet decode with the reference's function signatures, running on one fused kernel.

Mirrors `decode_centernet_predictions`, `_nms`, `_topk` of src/centernet_target.py:326-452 and of
their copy src/fusion_detection.py:695-820 (which differs only in voxel_size: 0.512 instead of
2.048).  The reference's quirks are kept on purpose: labels are always 0 (Q1), `indices`
index the C*K candidate list (Q2), a sample without detections yields CPU tensors (Q6), and K > H*W
raises the RuntimeError torch.topk raises (Q7).  Ties, which torch leaves undefined, are broken by
ascending index (Q4).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from . import conv_blocks, ops

CENTERNET_TARGET_VOXEL = 2.048    # src/centernet_target.py:389
FUSION_DETECTION_VOXEL = 0.512    # src/fusion_detection.py:757
PC_ORIGIN = (-51.2, -51.2)        # pc_range[0:2], hard-coded at src/centernet_target.py:390
GROUND_Z = -1.0                   # world_z, src/centernet_target.py:394


def _nms(heat: torch.Tensor, kernel: int = 3) -> torch.Tensor:
    """heat * (max_pool2d(heat, 3, 1, 1) == heat)."""
    if kernel != 3:
        raise NotImplementedError("the b200bev NMS kernel implements the 3x3 window the reference uses")
    return ops.centernet_nms(heat)


def _topk(scores: torch.Tensor, K: int = 100) -> Tuple[torch.Tensor, ...]:
    """(topk_score, topk_ind, topk_classes, topk_ys, topk_xs), each (B,K)."""
    return ops.centernet_topk(scores, K)


def decode_centernet_predictions(predictions: Dict[str, torch.Tensor], score_thresh: float = 0.3,
                                 max_detections: int = 100,
                                 voxel_size: float = CENTERNET_TARGET_VOXEL) -> List[Dict[str, torch.Tensor]]:
    """List (one dict per sample) of boxes (n,7) [x,y,z,w,l,h,yaw], scores (n,), labels (n,) int64,
    velocities (n,2); n <= max_detections is the number of winners with score > score_thresh."""
    # The bf16 head notes its raw output on the heat-map tensor it returns: if `predictions['heatmap']` still IS that
    # tensor, untouched, the sigmoid runs inside the decode launch (bit-identical to torch.sigmoid + decode on the
    # device).  Any other tensor under 'heatmap' is decoded as it is, as the reference does (src/centernet_target.py:345).
    logits = conv_blocks.logits_of(predictions["heatmap"])
    out = ops.centernet_decode(predictions["heatmap"] if logits is None else logits, predictions["offset"], predictions["size"],
                               predictions["rot"], predictions["vel"], max_detections, voxel_size,
                               PC_ORIGIN, GROUND_Z, score_thresh, heat_is_logit=logits is not None)
    counts = out["count"].tolist()                      # the one host sync of the call
    detections = []
    for b, n in enumerate(counts):
        if n == 0:
            # CPU tensors regardless of the input device, as in the reference (src/centernet_target.py:362-369)
            detections.append({"boxes": torch.zeros(0, 7), "scores": torch.zeros(0),
                               "labels": torch.zeros(0, dtype=torch.long), "velocities": torch.zeros(0, 2)})
            continue
        detections.append({"boxes": out["boxes"][b, :n], "scores": out["scores"][b, :n],
                           "labels": out["labels"][b, :n], "velocities": out["velocities"][b, :n]})
    return detections


def decode_centernet_predictions_fusion_detection(predictions, score_thresh: float = 0.3, max_detections: int = 100):
    """The copy eval.py / inference.py import (src/fusion_detection.py:695): 0.512 m per cell."""
    return decode_centernet_predictions(predictions, score_thresh, max_detections, FUSION_DETECTION_VOXEL)
