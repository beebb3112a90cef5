"""Drop-in route: rebind the hot-path callables of the ALREADY-IMPORTED reference modules.

The reference's pipelines (train_detect.py, eval.py, inference.py) import their model code by module
name from their own src/ directory:

    from fusion import create_detector                      -> builds encoders.* and FlexibleBEVFusion
    from centernet_target import decode_centernet_predictions   (train_detect.py)
    from fusion_detection import decode_centernet_predictions   (eval.py, inference.py)

`patch()` leaves constructors, parameters, state_dict names and signatures untouched and swaps only

    encoders.PointNetLiDAREncoder.forward     (src/encoders.py:271)
    encoders.MultiRadarEncoder.forward        (src/encoders.py:628)
    fusion.FlexibleBEVFusion.forward          (src/fusion.py:209)
    fusion.CenterNetHead.forward              (src/fusion.py:869)
    centernet_target.{_nms,_topk,decode_centernet_predictions}    (src/centernet_target.py:326-452)
    fusion_detection.{_nms,_topk,decode_centernet_predictions}    (src/fusion_detection.py:695-820)

plus the names the pipeline scripts already bound with `from ... import ...`.  Usage, from the
reference's src/ directory:

    import bevfusion_multimodal_3d_object_detection_b200 as b200bev
    import eval as ref_eval            # or train_detect / inference
    b200bev.patch()
    ref_eval.main(...)

The library is loaded at patch time, so a missing libb200bev.so fails here, loudly, not mid-run.
"""
from __future__ import annotations

import functools
import importlib
import sys
from typing import Dict, List, Tuple

from . import _lib, centernet_decode, conv_blocks, encoders, fusion

_saved: List[Tuple[object, str, object]] = []


def _swap(owner, name: str, new) -> None:
    _saved.append((owner, name, getattr(owner, name)))
    setattr(owner, name, new)


def _module(name: str):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:
        return None


def patch(precision: str = None) -> Dict[str, List[str]]:
    """Installs the kernels behind the reference's names. Returns {module: [patched names]}.

    precision: None (each module's own `b200_precision` attribute, else `B200BEV_PRECISION`, else "f32"), "f32" or
    "bf16".  A value given here is written to `b200_precision` of EVERY patched module the first time its forward runs —
    LiDAR encoder, radar encoder, fusion module and detection head alike — unless the module already carries its own
    setting, so `patch("bf16")` moves the MLP and all convolution blocks together and `patch("f32")` wins over the
    environment variable."""
    _lib.lib()  # fail now if the CUDA extension is missing
    if precision is not None and precision not in encoders.PRECISIONS:
        raise ValueError(f"unknown precision {precision!r}; choose one of {sorted(encoders.PRECISIONS)}")
    if _saved:
        return {}

    def with_precision(fn):
        """The reference's method signature in front of `fn(self, ...)`, stamping the patch-wide precision on first use."""
        @functools.wraps(fn)
        def forward(self, *args, **kwargs):
            if precision is not None and getattr(self, "b200_precision", None) is None:
                self.b200_precision = precision
            return fn(self, *args, **kwargs)
        return forward

    done: Dict[str, List[str]] = {}
    enc_mod, fus_mod = _module("encoders"), _module("fusion")
    if enc_mod is not None and hasattr(enc_mod, "PointNetLiDAREncoder") and enc_mod is not encoders:
        _swap(enc_mod.PointNetLiDAREncoder, "forward", with_precision(lambda self, x: encoders.lidar_forward(self, x)))
        _swap(enc_mod.MultiRadarEncoder, "forward",
              with_precision(lambda self, radar_list: encoders.multi_radar_forward(self, radar_list)))
        done["encoders"] = ["PointNetLiDAREncoder.forward", "MultiRadarEncoder.forward"]
    if fus_mod is not None and hasattr(fus_mod, "FlexibleBEVFusion") and fus_mod is not fusion:
        _swap(fus_mod.FlexibleBEVFusion, "forward",
              with_precision(lambda self, camera_features=None, lidar_features=None, radar_features=None:
                             fusion.fusion_forward(self, camera_features, lidar_features, radar_features)))
        done["fusion"] = ["FlexibleBEVFusion.forward"]
        if hasattr(fus_mod, "CenterNetHead"):
            _swap(fus_mod.CenterNetHead, "forward", with_precision(lambda self, x: conv_blocks.head_forward(self, x)))
            done["fusion"].append("CenterNetHead.forward")
    variants = {"centernet_target": centernet_decode.CENTERNET_TARGET_VOXEL,
                "fusion_detection": centernet_decode.FUSION_DETECTION_VOXEL}
    for mod_name, voxel in variants.items():
        mod = _module(mod_name)
        if mod is None or not hasattr(mod, "decode_centernet_predictions"):
            continue
        original = mod.decode_centernet_predictions
        decode = functools.partial(_decode_with_voxel, voxel)
        functools.update_wrapper(decode, original)
        _swap(mod, "decode_centernet_predictions", decode)
        _swap(mod, "_nms", centernet_decode._nms)
        _swap(mod, "_topk", centernet_decode._topk)
        done[mod_name] = ["decode_centernet_predictions", "_nms", "_topk"]
        # scripts that did `from <mod> import decode_centernet_predictions` hold their own binding
        for script in ("train_detect", "eval", "inference"):
            s = sys.modules.get(script)
            if s is not None and getattr(s, "decode_centernet_predictions", None) is original:
                _swap(s, "decode_centernet_predictions", decode)
                done.setdefault(script, []).append("decode_centernet_predictions")
    return done


def _decode_with_voxel(voxel, predictions, score_thresh: float = 0.3, max_detections: int = 100):
    return centernet_decode.decode_centernet_predictions(predictions, score_thresh, max_detections, voxel)


def unpatch() -> None:
    while _saved:
        owner, name, old = _saved.pop()
        setattr(owner, name, old)
