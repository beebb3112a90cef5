"""ctypes binding of libb200bev.so — one prototype per entry point of include/b200bev.h.

There is no fallback: if the library is missing or fails to load, ``lib()`` raises.  The library is
built in-tree by ``build.py`` (``python -m bevfusion_multimodal_3d_object_detection_b200.build``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "_native" / "libb200bev.so"

ABI_VERSION = 2
OK = 0
ERR_INVALID_ARGUMENT = 1
ERR_UNSUPPORTED = 2
ERR_K_OUT_OF_RANGE = 3
ERR_WORKSPACE = 4
ERR_CUDA = 1000

F32 = 0
BF16_TENSOR = 1
RADAR_FUSION = {"concat": 0, "max": 1, "mean": 2}
PROJECT_IMPL = {"auto": 0, "staged": 1, "gather": 2}

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_z = C.c_size_t

# name -> (restype, argtypes); mirrors include/b200bev.h declaration by declaration
PROTOTYPES = {
    "b200bev_abi_version": (_i, []),
    "b200bev_error_string": (C.c_char_p, [_i]),
    "b200bev_device_info": (_i, [C.POINTER(_i)] * 3),
    "b200bev_lidar_prepare_workspace_bytes": (_z, [_i, C.c_int64, _i]),
    "b200bev_lidar_prepare": (_i, [_p, _p, _i, _i, C.c_int64, C.POINTER(C.c_float), _i, _p, _p, _p, _p, _z, _p]),
    "b200bev_bin_sort": (_i, [_p, _i, _i, _i, _f, _f, _f, _f, _i, _i, _p, _p, _p, _p]),
    "b200bev_lidar_prepare_bin_sort": (_i, [_p, _p, _i, _i, C.c_int64, C.POINTER(C.c_float), _i, _f, _f, _i, _i, _p, _p, _p, _p, _p, _p]),
    "b200bev_pointnet_encode": (_i, [_p, _i, _i, _i, _p, C.POINTER(C.c_int32), _i, _p, _p, _i, _i, _p, _p, _p, _p]),
    "b200bev_pointnet_pack_bf16_bytes": (_z, [C.POINTER(C.c_int32), _i]),
    "b200bev_pointnet_pack_bf16": (_i, [_p, C.POINTER(C.c_int32), _i, _p, _z, _p]),
    "b200bev_pointnet_pack_split_bytes": (_z, [C.POINTER(C.c_int32), _i]),
    "b200bev_pointnet_pack_split": (_i, [_p, C.POINTER(C.c_int32), _i, _p, _z, _p]),
    "b200bev_pointnet_split_workspace_bytes": (_z, [_i, _i]),
    "b200bev_pointnet_encode_split": (_i, [_p, _i, _i, _i, C.POINTER(C.c_int32), _i, _p, _p, _i, _p, _p, _p, _p, _z, _p]),
    "b200bev_radar_encode": (_i, [C.POINTER(_p), C.POINTER(C.c_int32), _i, _i, _i, _p, C.POINTER(C.c_int32), _i,
                                  _i, _p, _p, _p, _p, _p]),
    "b200bev_camera_mean": (_i, [_p, _i, _i, C.c_int64, _p, _p]),
    "b200bev_bilinear_resize": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _p]),
    "b200bev_camera_project": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _i, _f, _f, _f, _f, _f, _f, _f, _i, _i, _p, _p, _i, _p]),
    "b200bev_centernet_nms": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "b200bev_centernet_workspace_bytes": (_z, [_i, _i, _i]),
    "b200bev_centernet_topk": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _z, _p]),
    "b200bev_centernet_decode": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _f, _f, _f,
                                      _p, _p, _p, _p, _p, _p, _p, _p, _p, _z, _p]),
    "b200bev_centernet_decode_logits": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _f, _f, _f, _f,
                                             _p, _p, _p, _p, _p, _p, _p, _p, _p, _z, _p]),
    "b200bev_dense_layer": (_i, [_p, _i, _i, _p, _p, _i, _i, _p, _p]),
    "b200bev_lidar_init": (_i, [_p, _i, _i, _p, _p, _i, _p, _p, _i, _p, _p, _p]),
    "b200bev_dense_pack_split_bytes": (_z, [_i, _i]),
    "b200bev_dense_pack_split": (_i, [_p, _p, _i, _i, _p, _z, _p]),
    "b200bev_dense_layer_split": (_i, [_p, _i, _i, _p, _i, _i, _p, _p]),
    "b200bev_lidar_init_split": (_i, [_p, _i, _i, _p, _p, _i, _p, _i, _p, _p, _p]),
    "b200bev_conv_pack_bytes": (_z, [_i, _i, _i]),
    "b200bev_conv_pack_bf16": (_i, [_p, _i, _i, _i, _p, _z, _p]),
    "b200bev_nchw_to_nhwc_bf16": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _p]),
    "b200bev_bilinear_resize_nhwc_bf16": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _i, _i, _p]),
    "b200bev_camera_mean_nhwc_bf16": (_i, [_p, _i, _i, _i, _i, _i, _p, _i, _i, _p]),
    "b200bev_conv_bn_relu_bf16": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _i, _i, _p, _p]),
    "b200bev_absmax": (_i, [_p, C.c_int64, _p, _p]),
    "b200bev_nchw_to_nhwc_split": (_i, [_p, _i, _i, _i, _i, _p, _i, _i, _p, _p]),
    "b200bev_conv_pack_split_bytes": (_z, [_i, _i, _i]),
    "b200bev_conv_pack_split": (_i, [_p, _i, _i, _i, _p, _z, _p]),
    "b200bev_conv_bn_relu_split": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _i, _i, _i, _p, _p]),
    "b200bev_border_expand": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _i, _i, _p]),
    "b200bev_conv_bn_relu_bf16_nhwc": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _i, _i, _p, _i, _i, _p, _p]),
}

_lock = threading.Lock()
_handle = None


class B200BevError(RuntimeError):
    """Non-zero status from libb200bev.so."""

    def __init__(self, status: int, message: str):
        super().__init__(f"libb200bev: {message} (status {status})")
        self.status = status


def lib() -> C.CDLL:
    """Loads libb200bev.so once and installs the prototypes. Raises if the library is absent."""
    global _handle
    if _handle is not None:
        return _handle
    with _lock:
        if _handle is not None:
            return _handle
        path = Path(os.environ.get("B200BEV_LIB", LIB_PATH))
        if not path.exists():
            raise ImportError(
                f"{path} not found: the CUDA extension is not built. Run "
                "`python -m bevfusion_multimodal_3d_object_detection_b200.build` (needs nvcc); "
                "there is no CPU or PyTorch fallback for the hot path.")
        h = C.CDLL(str(path))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(h, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        got = h.b200bev_abi_version()
        if got != ABI_VERSION:
            raise ImportError(f"libb200bev ABI version {got}, binding expects {ABI_VERSION}")
        _handle = h
    return _handle


# ---- optional call accounting (bench.py: `gpu_launches`) -------------------------------------------------------------
# Kernels one call of an entry point launches (memsets not counted); entries absent here launch one, entries mapped to 0 none.
KERNELS_PER_CALL = {
    "b200bev_abi_version": 0, "b200bev_error_string": 0, "b200bev_device_info": 0, "b200bev_lidar_prepare_workspace_bytes": 0,
    "b200bev_pointnet_pack_bf16_bytes": 0, "b200bev_centernet_workspace_bytes": 0, "b200bev_conv_pack_bytes": 0,
    "b200bev_pointnet_pack_split_bytes": 0, "b200bev_pointnet_split_workspace_bytes": 0, "b200bev_conv_pack_split_bytes": 0,
    "b200bev_conv_pack_split": 2, "b200bev_dense_pack_split_bytes": 0, "b200bev_lidar_init_split": 2,
    "b200bev_radar_encode": 2, "b200bev_lidar_init": 2, "b200bev_lidar_prepare": 1,
    "b200bev_pointnet_encode_split": 5,       # per pass: layer 1 + four GEMM launches (one pass up to ~1M points)
    "b200bev_pointnet_pack_split": 2,
}
_call_counts = None


def enable_call_counting() -> None:
    """Wraps every entry point of the loaded library in a counter (reset_call_counts / kernel_launches)."""
    global _call_counts
    if _call_counts is not None:
        return
    h = lib()
    _call_counts = {}
    for name in PROTOTYPES:
        fn = getattr(h, name)

        def counted(*args, _fn=fn, _name=name):
            _call_counts[_name] = _call_counts.get(_name, 0) + 1
            return _fn(*args)

        setattr(h, name, counted)


def reset_call_counts() -> None:
    if _call_counts is not None:
        _call_counts.clear()


def kernel_launches() -> int:
    """Kernels launched by the entry-point calls since the last reset (KERNELS_PER_CALL per call; the fp32-accuracy
    tensor-core MLP reports its own count through `pointnet_split_launches`)."""
    if _call_counts is None:
        return 0
    return sum(n * KERNELS_PER_CALL.get(name, 1) for name, n in _call_counts.items())


def check(status: int) -> None:
    """Maps a status code to the exception the reference would raise at the same point."""
    if status == OK:
        return
    msg = lib().b200bev_error_string(status).decode()
    if status == ERR_K_OUT_OF_RANGE:
        # torch.topk: "RuntimeError: selected index k out of range" 
        raise RuntimeError("selected index k out of range")
    raise B200BevError(status, msg)
