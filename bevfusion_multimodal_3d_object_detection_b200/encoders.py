"""Point-cloud encoders with the reference's interface, running on the b200bev kernels.

Mirrors `PointNetLiDAREncoder` (src/encoders.py:191-306), `RadarEncoder` (:458-557) and
`MultiRadarEncoder` (:560-661) of the reference: same constructor arguments (direct kwargs | config
dict | config_path), same attribute and state_dict names (conv1..5 / bn1..5, radar_encoder.*,
fusion_fc.*), same input layouts, same exceptions.  What differs is the forward pass in eval mode on
a CUDA tensor: one fused kernel (shared MLP with BatchNorm folded + max) instead of 15 ATen ops.

    eval + CUDA   -> libb200bev (fp32 FFMA, or bf16 tcgen05 when precision='bf16')
    training      -> the plain torch graph (BatchNorm needs batch statistics and autograd)
    eval + CPU    -> RuntimeError: there is no CPU fallback

The same functions (`lidar_forward`, `multi_radar_forward`) are what `patch()` installs on the
reference's own classes, so both routes execute identical code.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
import yaml

from . import _lib, ops
from .weight_cache import invalidate_cache, mark_dirty, state_token, wants_autograd  # noqa: F401

# MLP precision per name; "f32_cudnn" only differs for the convolution stacks (conv_blocks.conv_mode)
PRECISIONS = {"f32": _lib.F32, "fp32": _lib.F32, "bf16": _lib.BF16_TENSOR, "f32_cudnn": _lib.F32}


def load_config(config_path: str = "configs/base.yaml") -> Dict:
    """yaml.safe_load of the single config file; FileNotFoundError if absent (src/encoders.py:16-33)."""
    path = Path(config_path)
    if not path.exists():
        raise FileNotFoundError(f"Config file not found: {path}")
    with open(path, "r") as f:
        return yaml.safe_load(f)


def default_precision() -> str:
    return os.environ.get("B200BEV_PRECISION", "f32").lower()


# ------------------------------------------------------------------------------------------------
# folded-weight cache shared by the mirror classes and the patched reference classes
# ------------------------------------------------------------------------------------------------
def _mlp_stages(module: nn.Module) -> List[Tuple[nn.Conv1d, nn.Module]]:
    stages = []
    i = 1
    while hasattr(module, f"conv{i}"):
        stages.append((getattr(module, f"conv{i}"), getattr(module, f"bn{i}")))
        i += 1
    return stages


def _state_key(module: nn.Module, device: torch.device):
    """Cache key of a module's packed weights (weight_cache.state_token: hooks + version counters, no per-call walk)."""
    return state_token(module, device)


def packed_params(module: nn.Module, device: torch.device, want_bf16: bool = False, precision: Optional[int] = None):
    """(blob, dims, tc_blob|None) for a PointNet-style module; rebuilt only when a parameter or a
    BatchNorm statistic changed (load_state_dict hook, training-mode forward, tensor version counters; see
    weight_cache.py — `invalidate_cache(module)` after writes through `.data`).
    tc_blob is the tensor-core weight image of the asked precision: the bf16 stage image for BF16_TENSOR, the split-fp16
    image of the fp32-accuracy tensor-core path for F32 (None when the layer widths are not the ones that path takes —
    the FFMA kernel then runs)."""
    if precision is None:
        precision = _lib.BF16_TENSOR if want_bf16 else None
    key = _state_key(module, device)
    cache = module.__dict__.get("_b200bev_cache")
    if cache is None or cache["key"] != key:
        ws, bs = [], []
        for conv, bn in _mlp_stages(module):
            w, b = ops.fold_batchnorm(conv.weight, conv.bias, bn)
            ws.append(w)
            bs.append(b)
        blob, dims = ops.pack_mlp_params(ws, bs, device)
        cache = {"key": key, "blob": blob, "dims": dims, "tc": {}}
        module.__dict__["_b200bev_cache"] = cache
    if precision is None:
        return cache["blob"], cache["dims"], None
    if precision not in cache["tc"]:
        if precision == _lib.BF16_TENSOR:
            cache["tc"][precision] = ops.pack_mlp_params_bf16(cache["blob"], cache["dims"])
        else:
            cache["tc"][precision] = ops.pack_mlp_params_split(cache["blob"], cache["dims"])     # None if unsupported widths
    return cache["blob"], cache["dims"], cache["tc"][precision]


def _as_bnc(x: torch.Tensor, channels: int) -> torch.Tensor:
    """Accepts (B,N,C) or (B,C,N) exactly as the reference does (src/encoders.py:282-284): a 3-D input
    whose last dim equals input_channels is (B,N,C); anything else is taken as (B,C,N)."""
    if x.dim() != 3:
        raise ValueError(f"expected a 3-D point tensor, got shape {tuple(x.shape)}")
    if x.shape[2] == channels:
        return x
    return x.transpose(1, 2)


def _torch_mlp(module: nn.Module, x_bcn: torch.Tensor) -> torch.Tensor:
    for conv, bn in _mlp_stages(module):
        x_bcn = F.relu(bn(conv(x_bcn)))
    return x_bcn


def _precision_of(module: nn.Module) -> int:
    name = getattr(module, "b200_precision", None) or default_precision()
    if name not in PRECISIONS:
        raise ValueError(f"unknown precision {name!r}; choose one of {sorted(PRECISIONS)}")
    return PRECISIONS[name]


def lidar_forward(module: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """PointNetLiDAREncoder.forward (src/encoders.py:271-306)."""
    if module.training:
        mark_dirty(module)                                   # an optimizer step follows: rebuild at the next eval forward
    if module.training or getattr(module, "return_point_features", False) \
            or (x.is_cuda and wants_autograd(module, x)):
        # training graph; the per-point output nothing in the pipelines enables (src/encoders.py:239); or an eval-mode
        # call that autograd records (the kernels return tensors without a grad_fn)
        xb = _as_bnc(x, module.input_channels).transpose(1, 2)
        feat = _torch_mlp(module, xb)
        glob = torch.max(feat, 2)[0]
        if getattr(module, "return_point_features", False):
            both = torch.cat([feat, glob.unsqueeze(2).expand(-1, -1, feat.shape[2])], dim=1)
            return both.transpose(1, 2)
        return glob
    pts = _as_bnc(x, module.input_channels)
    prec = _precision_of(module)
    blob, dims, tc = packed_params(module, pts.device, precision=prec)
    return ops.pointnet_encode(pts, blob, dims, precision=prec, tc_params=tc)


def lidar_cell_canvas(module: nn.Module, x: torch.Tensor, bev_size: Tuple[int, int],
                      pc_range: Sequence[float] = ops.DEFAULT_PC_RANGE) -> torch.Tensor:
    """Extended mode (north_star S1): bin-and-sort + per-cell scatter-max.  Returns the canvas as a
    (B, C_out, H, W) tensor in channels_last memory format (a view of the kernel's (B,H*W,C) output)."""
    if module.training:
        raise RuntimeError("the per-cell canvas is an inference-only kernel path (eval mode)")
    pts = _as_bnc(x, module.input_channels)
    H, W = int(bev_size[0]), int(bev_size[1])
    _, perm, offsets = ops.bin_sort(pts, W, H, pc_range)
    prec = _precision_of(module)
    blob, dims, tc = packed_params(module, pts.device, precision=prec)
    canvas = ops.pointnet_encode(pts, blob, dims, perm=perm, offsets=offsets, n_cells=H * W, precision=prec,
                                 tc_params=tc, want_global=False)
    return canvas.view(pts.shape[0], H, W, dims[-1]).permute(0, 3, 1, 2)


def multi_radar_forward(module: nn.Module, radar_list: Sequence[torch.Tensor]) -> torch.Tensor:
    """MultiRadarEncoder.forward (src/encoders.py:628-661)."""
    if module.fusion_method not in _lib.RADAR_FUSION:
        raise ValueError(f"Unknown fusion method: {module.fusion_method}")
    enc = module.radar_encoder
    if module.training:
        mark_dirty(enc)
    if module.training or (len(radar_list) > 0 and radar_list[0].is_cuda and wants_autograd(module, list(radar_list))):
        feats = torch.stack([torch.max(_torch_mlp(enc, _as_bnc(r, enc.input_channels).transpose(1, 2)), 2)[0]
                             for r in radar_list], dim=1)
        if module.fusion_method == "concat":
            return module.fusion_fc(feats.view(feats.shape[0], -1))
        return torch.max(feats, dim=1)[0] if module.fusion_method == "max" else torch.mean(feats, dim=1)
    radars = [_as_bnc(r, enc.input_channels) for r in radar_list]
    blob, dims, _ = packed_params(enc, radars[0].device)
    fc = getattr(module, "fusion_fc", None) if module.fusion_method == "concat" else None
    fused, _ = ops.radar_encode(radars, blob, dims, module.fusion_method,
                                None if fc is None else fc.weight.detach(), None if fc is None else fc.bias.detach())
    return fused


# ------------------------------------------------------------------------------------------------
# mirror classes
# ------------------------------------------------------------------------------------------------
def _build_mlp(module: nn.Module, widths: Sequence[int], use_bn: bool) -> None:
    for i, (c_in, c_out) in enumerate(zip(widths[:-1], widths[1:]), start=1):
        setattr(module, f"conv{i}", nn.Conv1d(c_in, c_out, 1))
        setattr(module, f"bn{i}", nn.BatchNorm1d(c_out) if use_bn else nn.Identity())


class PointNetLiDAREncoder(nn.Module):
    """(B,N,C)|(B,C,N) -> (B,feat_dim).  Constructor contract of src/encoders.py:208-269."""

    def __init__(self, input_channels: Optional[int] = None, feat_dim: Optional[int] = None,
                 use_bn: Optional[bool] = None, return_point_features: Optional[bool] = None,
                 config: Optional[Dict] = None, config_path: Optional[str] = None):
        super().__init__()
        mlp_layers = [64, 128, 256, 512, 1024]
        if config is not None or config_path is not None:
            cfg = (config if config is not None else load_config(config_path)).get("model", {}).get("lidar_encoder", {})
            self.input_channels = cfg.get("input_channels", 5)
            self.feat_dim = cfg.get("feature_dim", 1024)
            use_bn = cfg.get("use_batch_norm", True)
            self.return_point_features = False
            mlp_layers = cfg.get("mlp_layers", mlp_layers)
        else:
            self.input_channels = 5 if input_channels is None else input_channels
            self.feat_dim = 1024 if feat_dim is None else feat_dim
            use_bn = True if use_bn is None else use_bn
            self.return_point_features = bool(return_point_features)
        _build_mlp(self, [self.input_channels, *mlp_layers[:5]], use_bn)
        self.b200_precision: Optional[str] = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return lidar_forward(self, x)

    def cell_canvas(self, x: torch.Tensor, bev_size: Tuple[int, int],
                    pc_range: Sequence[float] = ops.DEFAULT_PC_RANGE) -> torch.Tensor:
        return lidar_cell_canvas(self, x, bev_size, pc_range)


class RadarEncoder(nn.Module):
    """(B,N,C)|(B,C,N) -> (B,feat_dim).  Constructor contract of src/encoders.py:476-529."""

    def __init__(self, input_channels: Optional[int] = None, feat_dim: Optional[int] = None,
                 use_bn: Optional[bool] = None, config: Optional[Dict] = None, config_path: Optional[str] = None):
        super().__init__()
        mlp_layers = [32, 64, 128, 256]
        if config is not None or config_path is not None:
            cfg = (config if config is not None else load_config(config_path)).get("model", {}).get("radar_encoder", {})
            self.input_channels = cfg.get("input_channels", 7)
            self.feat_dim = cfg.get("feature_dim", 256)
            use_bn = cfg.get("use_batch_norm", True)
            mlp_layers = cfg.get("mlp_layers", mlp_layers)
        else:
            self.input_channels = 7 if input_channels is None else input_channels
            self.feat_dim = 256 if feat_dim is None else feat_dim
            use_bn = True if use_bn is None else use_bn
        _build_mlp(self, [self.input_channels, *mlp_layers[:4]], use_bn)
        self.b200_precision: Optional[str] = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            mark_dirty(self)
        if self.training or (x.is_cuda and wants_autograd(self, x)):
            return torch.max(_torch_mlp(self, _as_bnc(x, self.input_channels).transpose(1, 2)), 2)[0]
        pts = _as_bnc(x, self.input_channels)
        blob, dims, _ = packed_params(self, pts.device)
        return ops.pointnet_encode(pts, blob, dims)


class MultiRadarEncoder(nn.Module):
    """List of (B,N_i,C) -> (B,feat_dim).  Constructor contract of src/encoders.py:575-626."""

    def __init__(self, input_channels: Optional[int] = None, feat_dim: Optional[int] = None,
                 num_radars: Optional[int] = None, fusion_method: Optional[str] = None,
                 config: Optional[Dict] = None, config_path: Optional[str] = None):
        super().__init__()
        if config is not None or config_path is not None:
            if config is None:
                config = load_config(config_path)
            cfg = config.get("model", {}).get("radar_encoder", {})
            input_channels = cfg.get("input_channels", 7)
            self.feat_dim = cfg.get("feature_dim", 256)
            self.num_radars = cfg.get("num_radars", 5)
            self.fusion_method = cfg.get("fusion_method", "concat")
        else:
            input_channels = 7 if input_channels is None else input_channels
            self.feat_dim = 256 if feat_dim is None else feat_dim
            self.num_radars = 5 if num_radars is None else num_radars
            self.fusion_method = "concat" if fusion_method is None else fusion_method
        self.radar_encoder = RadarEncoder(input_channels=input_channels, feat_dim=self.feat_dim, config=config)
        if self.fusion_method == "concat":
            self.fusion_fc = nn.Linear(self.feat_dim * self.num_radars, self.feat_dim)
        self.output_dim = self.feat_dim

    def forward(self, radar_list: List[torch.Tensor]) -> torch.Tensor:
        return multi_radar_forward(self, radar_list)
