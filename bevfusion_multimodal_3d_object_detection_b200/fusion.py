"""BEV fusion module with the reference's interface; the camera->BEV step runs on b200bev kernels.

Mirrors `FlexibleBEVFusion` (src/fusion.py:46-327): same constructor contract, same sub-module and
state_dict names (camera_proj, lidar_init, lidar_upsample, radar_proj, radar_refine, bev_fusion),
same forward signature and exceptions.  The hot-path piece — src/fusion.py:229-248 — is

    mean over the 6 cameras  ->  [camera_proj convs, cuDNN]  ->  bilinear resize to (bev_h, bev_w)

and in eval mode on CUDA the first and the last step are `b200bev_camera_mean` and
`b200bev_bilinear_resize`; the dense layers (`lidar_init`, `radar_proj`) run on `b200bev_lidar_init` /
`b200bev_dense_layer`.  The convolution blocks between them are the reference's own fp32 cuDNN layers by
default (parity 1e-5) and the tcgen05 convolution kernel when `b200_precision = "bf16"` (parity 1e-2; `conv_blocks.py`,
`_fused_bf16_path` below).  `project_cameras()` is the geometric form north_star
describes: calibrated projection of the BEV cell centres and a one-pass gather of all cameras.
"""
from __future__ import annotations

import contextlib
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import conv_blocks, ops, runtime
from .encoders import load_config
from .weight_cache import mark_dirty, state_token, wants_autograd


def _conv_bn_relu(c_in: int, c_out: int, k: int) -> List[nn.Module]:
    return [nn.Conv2d(c_in, c_out, k, padding=k // 2), nn.BatchNorm2d(c_out), nn.ReLU(inplace=True)]


def lidar_start_size(module: nn.Module) -> int:
    """Side of the square map `lidar_init` produces.  The reference keeps it in a local (`start_size = 25`,
    src/fusion.py:141), so on the reference's own class it is recovered from the layer shapes."""
    s = getattr(module, "lidar_start_size", None)
    if s is None:
        hidden = module.lidar_upsample[0].in_channels
        s = math.isqrt(module.lidar_init[2].out_features // hidden)
    return int(s)


def camera_branch(module: nn.Module, camera_features: torch.Tensor, torch_graph: bool = False) -> torch.Tensor:
    """src/fusion.py:233-247 with the mean and the resize on the b200bev kernels (eval, CUDA)."""
    size = (module.bev_h, module.bev_w)
    if torch_graph:
        cam = camera_features.mean(dim=1) if camera_features.dim() == 5 else camera_features
        return F.interpolate(module.camera_proj(cam), size=size, mode="bilinear", align_corners=False)
    cam = ops.camera_mean(camera_features) if camera_features.dim() == 5 else camera_features
    return ops.bilinear_resize(_stack(module, module.camera_proj, [cam]), size)


def _stack(module: nn.Module, seq: nn.Sequential, parts) -> torch.Tensor:
    """A conv/BN/ReLU stack of the fusion module in eval mode on the tcgen05 convolution kernels: fp32
    accuracy by default (split fp16 operands, parity 1e-5), bf16 operands when the bf16 path is enabled (parity 1e-2); the
    module's own torch layers for "f32_cudnn" or shapes the kernels do not take."""
    mode = conv_blocks.conv_mode(module)
    if mode != "torch" and conv_blocks.supported(seq) and sum(int(p.shape[1]) for p in parts) % 64 == 0:
        return conv_blocks.run(seq, parts) if mode == "bf16" else conv_blocks.run_split(seq, parts)
    return seq(parts[0] if len(parts) == 1 else torch.cat(parts, dim=1))


def _const_image_size(seq: nn.Sequential, H: int, W: int) -> int:
    """Side of the small image on which a conv stack over a spatially constant input can run (2k+1 for k 3x3 convolutions),
    or 0 when the stack is not of that kind or the grid is too small for the shortcut."""
    k = 0
    for layer in seq:
        if isinstance(layer, nn.Conv2d):
            if layer.kernel_size == (3, 3) and layer.padding == (1, 1) and layer.stride == (1, 1) and layer.dilation == (1, 1) \
                    and layer.padding_mode == "zeros":
                k += 1
            elif layer.kernel_size != (1, 1) or layer.stride != (1, 1):
                return 0
        elif not isinstance(layer, (nn.BatchNorm2d, nn.ReLU)):
            return 0
    s = 2 * k + 1
    return s if (k > 0 and H >= s and W >= s) else 0


def radar_branch(module: nn.Module, radar_features: torch.Tensor, out_nhwc: Optional[torch.Tensor] = None,
                 c_offset: int = 0) -> Optional[torch.Tensor]:
    """src/fusion.py:274-281 in eval mode: radar_proj -> view(B,C,1,1).expand(B,C,H,W) -> radar_refine.  The expanded image
    is constant over space, so the two 3x3 convolutions run on a 5 x 5 image and `b200bev_border_expand` spreads the 25
    distinct pixels over the grid (1/100 of the convolution work at 50 x 50, and no 82 MB broadcast copy)."""
    B, c, H, W = radar_features.shape[0], module.bev_channels, module.bev_h, module.bev_w
    r = ops.dense_layer(radar_features, module.radar_proj[0].weight, module.radar_proj[0].bias, relu=True)      # :274
    s = _const_image_size(module.radar_refine, H, W)
    if s:
        small = _stack(module, module.radar_refine, [r.view(B, c, 1, 1).expand(B, c, s, s).contiguous()])
        return ops.border_expand(small, (H, W), out_nhwc=out_nhwc, c_offset=c_offset, want_nchw=out_nhwc is None)
    full = r.view(B, c, 1, 1).expand(B, c, H, W)                                                               # :277-278
    if out_nhwc is not None:
        conv_blocks.run(module.radar_refine, [full], out_nhwc=out_nhwc, c_offset=c_offset)
        return None
    return _stack(module, module.radar_refine, [full])                                                         # :281


DENSE_TC_MIN_BATCH = 1      # measured on B200 (512 -> 80000): 33 us at batch 1, 39 at 32 against 41 and 98 for the FFMA kernels


def lidar_init_dense(module: nn.Module, lidar_features: torch.Tensor) -> torch.Tensor:
    """`module.lidar_init(lidar_features)` (src/fusion.py:144-148, :258) on the dense kernels.  The 164 MB second layer runs
    on the tensor cores at fp32 accuracy (`b200bev_lidar_init_split`) from a split-fp16 image of the weight — as large as the
    weight itself, cached beside the module like the convolutions' stage images (weight_cache: rebuilt when a parameter
    changes).  Shapes without a tensor-core form and `module.b200_dense_tc = False` (no second copy of the weight on the
    device) take the FFMA kernels, which read torch's own weight."""
    l0, l2 = module.lidar_init[0], module.lidar_init[2]
    B = int(lidar_features.shape[0])
    if B >= DENSE_TC_MIN_BATCH and getattr(module, "b200_dense_tc", True) and lidar_features.is_cuda:
        seq = module.lidar_init
        key = state_token(seq, lidar_features.device)
        caches = seq.__dict__.setdefault("_b200bev_conv_cache", {})
        cache = caches.get("dense")
        if cache is None or cache["key"] != key:
            cache = {"key": key, "image": ops.dense_pack_split(l2.weight.to(lidar_features.device),
                                                                None if l2.bias is None else l2.bias.to(lidar_features.device))}
            caches["dense"] = cache
        if cache["image"] is not None:
            return ops.lidar_init_split(lidar_features, l0.weight, l0.bias, cache["image"], l2.out_features)
    return ops.lidar_init(lidar_features, l0.weight, l0.bias, l2.weight, l2.bias)


def lidar_branch(module: nn.Module, lidar_features: torch.Tensor, torch_graph: bool = False) -> torch.Tensor:
    """src/fusion.py:258-262: lidar_init (two dense layers, the second a 164 MB weight) -> (B,128,s,s) ->
    conv+BN+ReLU -> x2 bilinear upsample -> conv+BN+ReLU.  Eval mode on CUDA: the dense layers run on
    `b200bev_lidar_init` and the upsample on `b200bev_bilinear_resize`; the two convolutions stay cuDNN."""
    B = lidar_features.shape[0]
    s = lidar_start_size(module)
    hidden = module.lidar_init[2].out_features // (s * s)
    if torch_graph:
        return module.lidar_upsample(module.lidar_init(lidar_features).view(B, hidden, s, s))
    x = lidar_init_dense(module, lidar_features).view(B, hidden, s, s)
    up = module.lidar_upsample
    mode = conv_blocks.conv_mode(module)
    if mode != "torch" and conv_blocks.supported(up) and hidden % 64 == 0:
        return conv_blocks.run(up, [x]) if mode == "bf16" else conv_blocks.run_split(up, [x])
    for i, layer in enumerate(up):
        if isinstance(layer, nn.Upsample):
            # scale_factor=2, align_corners=False: source coordinate (i+0.5)/2-0.5, what the size-based resize computes
            x = ops.bilinear_resize(x, (int(x.shape[2] * layer.scale_factor), int(x.shape[3] * layer.scale_factor)))
        else:
            x = layer(x)
    return x


def _fused_bf16_path(module: nn.Module, camera_features, lidar_features, radar_features) -> Optional[torch.Tensor]:
    """The whole eval-mode forward with the bf16 path on: every convolution on the tcgen05 kernel, and no
    tensor between two kernels in any layout but the one its consumer reads — the camera mean leaves its kernel as
    camera_proj's channels-last bf16 input, the lidar and radar stacks write their slices of bev_fusion's concatenated
    input from their last convolution's epilogue (torch.cat of src/fusion.py:292 never happens).  Returns None when a
    stack has a shape the kernel does not take; the caller then runs the layer-by-layer path."""
    c = module.bev_channels
    use = [module.use_camera and camera_features is not None, module.use_lidar and lidar_features is not None,
           module.use_radar and radar_features is not None]
    stacks = [module.camera_proj if use[0] else None, module.lidar_upsample if use[1] else None,
              module.radar_refine if use[2] else None, module.bev_fusion]
    if not any(use) or c % 64 != 0 or not all(conv_blocks.supported(s) for s in stacks if s is not None):
        return None
    first = next(t for t, u in zip((camera_features, lidar_features, radar_features), use) if u)
    if not first.is_cuda:
        return None
    if use[0] and camera_features.shape[-3] % 64 != 0:
        return None
    if use[1] and (module.lidar_init[2].out_features // lidar_start_size(module) ** 2) % 64 != 0:
        return None
    B, H, W = first.shape[0], module.bev_h, module.bev_w
    cat = torch.empty((B, H, W, c * sum(use)), dtype=torch.bfloat16, device=first.device)
    off_cam, off_lidar, off_radar = 0, c * use[0], c * (use[0] + use[1])
    # The branches are independent and write disjoint channel slices of `cat`: the lidar and radar branches — a dozen small,
    # latency-bound kernels — go to side streams next to the camera branch (parallel branches of the CUDA graph when the step
    # is captured); bev_fusion waits for all of them.
    fork = runtime.BranchStreams(first.device) if (_parallel_branches(module, B * H * W) and sum(use) > 1) else None
    if use[1]:
        with (fork.fork(0) if fork and use[0] else contextlib.nullcontext()):
            s0 = lidar_start_size(module)
            hidden = module.lidar_init[2].out_features // (s0 * s0)
            x = lidar_init_dense(module, lidar_features).view(B, hidden, s0, s0)
            conv_blocks.run(module.lidar_upsample, [x], out_nhwc=cat, c_offset=off_lidar)    # :258-262
            del x
    if use[2]:
        with (fork.fork(1) if fork and (use[0] or use[1]) else contextlib.nullcontext()):
            radar_branch(module, radar_features, out_nhwc=cat, c_offset=off_radar)           # :274-281
    if use[0]:
        hw = camera_features.shape[-2] * camera_features.shape[-1]
        fh, fw = int(camera_features.shape[-2]), int(camera_features.shape[-1])
        proj = torch.empty((B, fh, fw, c), dtype=torch.bfloat16, device=first.device)       # camera_proj's output, channels-last
        if camera_features.dim() == 5 and hw % 4 == 0:
            conv_blocks.run(module.camera_proj, nhwc=ops.camera_mean_nhwc_bf16(camera_features), out_nhwc=proj)
        else:
            cam = ops.camera_mean(camera_features) if camera_features.dim() == 5 else camera_features
            conv_blocks.run(module.camera_proj, [cam], out_nhwc=proj)
        ops.bilinear_resize_nhwc_bf16(proj, (H, W), out=cat, c_offset=off_cam)               # src/fusion.py:242-247
    if fork:
        fork.join()
    return conv_blocks.run(module.bev_fusion, nhwc=cat, note_nhwc=True)                      # :292-295


PARALLEL_MAX_PIXELS = 32 * 50 * 50


def _parallel_branches(module: nn.Module, pixels: int) -> bool:
    """Side streams for the independent branches of the fused path.  `module.b200_parallel_branches` = True / False forces
    them on / off; unset, they are used up to 32 frames of 50 x 50 cells per call: measured on B200, the parallel branches
    gain 2.5 % at that size (the small kernels of the lidar and radar branches hide next to the camera branch) and lose
    1-2 % at 64 frames or at 100 x 100 cells, where every kernel fills the machine and two of them at once only contend."""
    flag = getattr(module, "b200_parallel_branches", None)
    return bool(flag) if flag is not None else pixels <= PARALLEL_MAX_PIXELS


def _torch_graph_wanted(module: nn.Module, *feats) -> bool:
    """Training mode (BatchNorm batch statistics + autograd), or an eval-mode call on CUDA that autograd
    records: both run the module's own torch layers on the input's device.  Eval mode on a CPU tensor still raises in
    the kernels' front end — there is no CPU fallback."""
    if module.training:
        for sub in module.children():
            mark_dirty(sub)
        return True
    live = [f for f in feats if f is not None]
    return bool(live) and live[0].is_cuda and wants_autograd(module, *live)


def fusion_forward(module: nn.Module, camera_features=None, lidar_features=None, radar_features=None) -> torch.Tensor:
    """FlexibleBEVFusion.forward (src/fusion.py:209-297)."""
    torch_graph = _torch_graph_wanted(module, camera_features, lidar_features, radar_features)
    if not torch_graph and conv_blocks.wants_bf16(module):
        out = _fused_bf16_path(module, camera_features, lidar_features, radar_features)
        if out is not None:
            return out
    use = [module.use_camera and camera_features is not None, module.use_lidar and lidar_features is not None,
           module.use_radar and radar_features is not None]
    if not any(use):
        raise ValueError("No modality features provided")              # :289
    first = next(t for t, u in zip((camera_features, lidar_features, radar_features), use) if u)
    B = first.shape[0]
    # kernel path on CUDA: the lidar and radar branches next to the camera branch (side streams, see _fused_bf16_path); their
    # results are read on the caller's stream after the join
    fork = None
    if not torch_graph and first.is_cuda and sum(use) > 1 and _parallel_branches(module, B * module.bev_h * module.bev_w):
        fork = runtime.BranchStreams(first.device)
    side = (lambda i: fork.fork(i)) if fork else (lambda i: contextlib.nullcontext())
    cam = lidar = radar = None
    if use[1]:
        with (side(0) if use[0] else contextlib.nullcontext()):
            lidar = lidar_branch(module, lidar_features, torch_graph)
    if use[2]:
        with (side(1) if (use[0] or use[1]) else contextlib.nullcontext()):
            if torch_graph:
                r = module.radar_proj(radar_features).view(B, module.bev_channels, 1, 1)
                radar = module.radar_refine(r.expand(B, module.bev_channels, module.bev_h, module.bev_w))       # :274-281
            else:
                radar = radar_branch(module, radar_features)
    if use[0]:
        cam = camera_branch(module, camera_features, torch_graph)
    if fork:
        fork.join()
    parts = [p for p in (cam, lidar, radar) if p is not None]          # the reference's order: camera, lidar, radar
    if torch_graph:
        return module.bev_fusion(torch.cat(parts, dim=1))              # :292-295
    return _stack(module, module.bev_fusion, parts)                    # the concat happens inside the layout kernel


class FlexibleBEVFusion(nn.Module):
    """Constructor contract of src/fusion.py:62-207."""

    def __init__(self, use_camera: Optional[bool] = None, use_lidar: Optional[bool] = None,
                 use_radar: Optional[bool] = None, camera_channels: Optional[int] = None,
                 lidar_channels: Optional[int] = None, radar_channels: Optional[int] = None,
                 bev_h: Optional[int] = None, bev_w: Optional[int] = None, bev_channels: Optional[int] = None,
                 pc_range: Optional[List[float]] = None, config: Optional[Dict] = None,
                 config_path: Optional[str] = None, lidar_start_size: Optional[int] = None):
        super().__init__()
        pick = lambda explicit, fallback: fallback if explicit is None else explicit
        if config is not None or config_path is not None:
            if config is None:
                config = load_config(config_path)
            model = config.get("model", {})
            bev = model.get("bev_fusion", {})
            data = config.get("dataset", {})
            self.use_camera = pick(use_camera, model.get("use_camera", True))
            self.use_lidar = pick(use_lidar, model.get("use_lidar", True))
            self.use_radar = pick(use_radar, model.get("use_radar", True))
            camera_channels = pick(camera_channels, model.get("camera_encoder", {}).get("output_channels", 512))
            lidar_channels = pick(lidar_channels, model.get("lidar_encoder", {}).get("feature_dim", 1024))
            radar_channels = pick(radar_channels, model.get("radar_encoder", {}).get("feature_dim", 256))
            self.bev_h = pick(bev_h, bev.get("bev_h", data.get("bev_h", 200)))
            self.bev_w = pick(bev_w, bev.get("bev_w", data.get("bev_w", 200)))
            self.bev_channels = pick(bev_channels, bev.get("bev_channels", 256))
            self.pc_range = pick(pc_range, data.get("point_cloud_range", list(ops.DEFAULT_PC_RANGE)))
        else:
            self.use_camera, self.use_lidar, self.use_radar = pick(use_camera, True), pick(use_lidar, True), pick(use_radar, True)
            camera_channels, lidar_channels, radar_channels = pick(camera_channels, 512), pick(lidar_channels, 1024), pick(radar_channels, 256)
            self.bev_h, self.bev_w, self.bev_channels = pick(bev_h, 200), pick(bev_w, 200), pick(bev_channels, 256)
            self.pc_range = pick(pc_range, list(ops.DEFAULT_PC_RANGE))
        self.num_modalities = sum([self.use_camera, self.use_lidar, self.use_radar])
        assert self.num_modalities > 0, "At least one modality must be enabled"
        c = self.bev_channels
        if self.use_camera:
            self.camera_proj = nn.Sequential(*_conv_bn_relu(camera_channels, 512, 3), *_conv_bn_relu(512, c, 1))
        if self.use_lidar:
            # The reference hard-codes 25 (src/fusion.py:141): its lidar BEV is always 50x50 and torch.cat raises for any
            # other grid.  25 is the default here too, so the state_dict shapes are the reference's for every
            # grid.  `lidar_start_size=` is this package's explicit extension: e.g. 50 lets the
            # 2x-resolution configuration (BASELINE configs[4], 100x100) run with the lidar branch.
            self.lidar_start_size = 25 if lidar_start_size is None else int(lidar_start_size)
            self.lidar_init = nn.Sequential(nn.Linear(lidar_channels, 512), nn.ReLU(inplace=True),
                                            nn.Linear(512, 128 * self.lidar_start_size ** 2))
            self.lidar_upsample = nn.Sequential(
                *_conv_bn_relu(128, 128, 3), nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False),
                *_conv_bn_relu(128, c, 3))
        if self.use_radar:
            self.radar_proj = nn.Sequential(nn.Linear(radar_channels, c), nn.ReLU(inplace=True))
            self.radar_refine = nn.Sequential(*_conv_bn_relu(c, c, 3), *_conv_bn_relu(c, c, 3))
        self.bev_fusion = nn.Sequential(*_conv_bn_relu(c * self.num_modalities, c * 2, 3), *_conv_bn_relu(c * 2, c, 3))

    def forward(self, camera_features: Optional[torch.Tensor] = None, lidar_features: Optional[torch.Tensor] = None,
                radar_features: Optional[torch.Tensor] = None) -> torch.Tensor:
        return fusion_forward(self, camera_features, lidar_features, radar_features)

    def project_cameras(self, camera_features: torch.Tensor, intrinsics: torch.Tensor, ego2cam: torch.Tensor,
                        img_size: Tuple[float, float] = (1600.0, 900.0), z_plane: float = 0.0) -> torch.Tensor:
        """(B,n_cam,C,h,w) -> (B,C,bev_h,bev_w) by calibrated projection (calibration layout of
        src/data_converter.py:110-117: per-camera intrinsic 3x3 and ego->camera [R|t] 3x4)."""
        return ops.camera_project(camera_features, intrinsics, ego2cam, img_size, (self.bev_h, self.bev_w),
                                  self.pc_range, z_plane)

    def get_config_str(self) -> str:
        names = [n for n, on in (("camera", self.use_camera), ("lidar", self.use_lidar), ("radar", self.use_radar)) if on]
        return "+".join(names)


class CenterNetHead(nn.Module):
    """Constructor contract, sub-module and state_dict names of src/fusion.py:788-867; forward :869-884."""

    def __init__(self, in_channels: Optional[int] = None, num_classes: Optional[int] = None, head_conv: Optional[int] = None,
                 config: Optional[Dict] = None, config_path: Optional[str] = None):
        super().__init__()
        if config is not None or config_path is not None:
            if config is None:
                config = load_config(config_path)
            cfg = config.get("model", {}).get("centernet_head", {})
            in_channels = cfg.get("in_channels", 256) if in_channels is None else in_channels
            num_classes = config.get("dataset", {}).get("num_classes", 10) if num_classes is None else num_classes
            head_conv = cfg.get("head_conv", 64) if head_conv is None else head_conv
        in_channels = 256 if in_channels is None else in_channels
        self.num_classes = 10 if num_classes is None else num_classes
        head_conv = 64 if head_conv is None else head_conv
        for name, n_out in zip(conv_blocks.HEADS, (self.num_classes, 2, 3, 2, 2)):
            setattr(self, f"{name}_head", nn.Sequential(nn.Conv2d(in_channels, head_conv, 3, padding=1, bias=True),
                                                        nn.ReLU(inplace=True), nn.Conv2d(head_conv, n_out, 1, bias=True)))
        for m in self.modules():                                  # src/fusion.py:856-867
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, std=0.001)
                nn.init.constant_(m.bias, 0)
        nn.init.constant_(self.heatmap_head[-1].bias, -math.log((1 - 0.01) / 0.01))

    def forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        return conv_blocks.head_forward(self, x)
