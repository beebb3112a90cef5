"""b200bev — B200-native BEV encode + decode hot path behind the interface of
meg89/bevfusion_multimodal_3d_object_detection.

    libb200bev.so (csrc/, C-ABI in include/b200bev.h)   hand-written sm_100a kernels
    ops                                                 torch-tensor front end of the C-ABI
    encoders / fusion / centernet_decode                the reference's module & function signatures
    patch()                                             rebinds them onto the imported reference modules
"""
from . import _lib  # noqa: F401
from .centernet_decode import (_nms, _topk, decode_centernet_predictions,  # noqa: F401
                               decode_centernet_predictions_fusion_detection)
from .encoders import MultiRadarEncoder, PointNetLiDAREncoder, RadarEncoder, load_config  # noqa: F401
from .fusion import CenterNetHead, FlexibleBEVFusion  # noqa: F401
from .detector import BEVDetectorChain  # noqa: F401
from .weight_cache import invalidate_cache  # noqa: F401
from .patch import patch, unpatch  # noqa: F401

__all__ = [
    "PointNetLiDAREncoder", "RadarEncoder", "MultiRadarEncoder", "FlexibleBEVFusion", "CenterNetHead",
    "decode_centernet_predictions", "decode_centernet_predictions_fusion_detection", "_nms", "_topk",
    "load_config", "patch", "unpatch", "BEVDetectorChain", "invalidate_cache",
]
