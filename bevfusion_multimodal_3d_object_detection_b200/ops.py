"""torch-tensor front end of the C-ABI: checks arguments, allocates outputs, passes raw device
pointers and the current CUDA stream to libb200bev.so.  PyTorch is plumbing here (memory, streams);
every op below runs a hand-written kernel or raises — nothing falls back to torch or to the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib

DEFAULT_PC_RANGE = (-51.2, -51.2, -5.0, 51.2, 51.2, 3.0)  # configs/base.yaml:48, src/centernet_target.py:390


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream(dev: torch.device):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _need_cuda(t: torch.Tensor, name: str, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(
            f"{name}: tensor is on {t.device}; the b200bev hot path runs on CUDA only (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _i32(values: Sequence[int]):
    return (C.c_int32 * len(values))(*[int(v) for v in values])


def voxel_size(pc_range: Sequence[float], W: int, H: int) -> Tuple[float, float]:
    """voxel = (max - min) / cells in float64, rounded once to fp32 (src/centernet_target.py:222-224)."""
    x_min, y_min, _, x_max, y_max, _ = pc_range
    return (float(x_max) - float(x_min)) / W, (float(y_max) - float(y_min)) / H


# ------------------------------------------------------------------------------------------------
# N3: the step in front of S1
# ------------------------------------------------------------------------------------------------
def lidar_prepare(raw: torch.Tensor, frame_offsets: torch.Tensor, max_points: int,
                  pc_range: Sequence[float] = DEFAULT_PC_RANGE, select: Optional[torch.Tensor] = None,
                  max_frame_rows: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Range filter + pad / subsample of a batch of raw sweeps (src/train_detect.py:147-189).

    raw (total_rows, C) f32 with the frames back to back, frame_offsets (B+1) i64 -> (points (B, max_points, C),
    count (B) i32).  `select` (B, max_points) i32 picks output rows by their index among the in-range points
    (the reference's np.random.choice draw); without it the in-range points keep their file order."""
    raw = _need_cuda(raw, "raw")
    frame_offsets = _need_cuda(frame_offsets, "frame_offsets", torch.int64)
    if raw.dim() != 2 or raw.shape[1] < 3:
        raise ValueError("raw must be (total_rows, C) with C >= 3")
    if frame_offsets.dim() != 1 or frame_offsets.numel() < 2:
        raise ValueError("frame_offsets must be (B + 1,)")
    B = frame_offsets.numel() - 1
    Cc = int(raw.shape[1])
    if max_frame_rows is None:
        max_frame_rows = int(raw.shape[0])     # a bound that needs no host sync
    dev = raw.device
    if select is not None:
        select = _need_cuda(select, "select", torch.int32)
        if tuple(select.shape) != (B, max_points):
            raise ValueError(f"select must be {(B, max_points)}, got {tuple(select.shape)}")
    n_ws = _lib.lib().b200bev_lidar_prepare_workspace_bytes(B, max_frame_rows, 1 if select is not None else 0)
    ws = torch.empty(max(n_ws, 16), dtype=torch.uint8, device=dev)
    out = torch.empty((B, max_points, Cc), dtype=torch.float32, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    rng = (C.c_float * 6)(*[float(v) for v in pc_range])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b200bev_lidar_prepare(_ptr(raw), _ptr(frame_offsets), B, Cc, max_frame_rows, rng, max_points,
                                                    _ptr(select), _ptr(out), _ptr(count), _ptr(ws), ws.numel(), _stream(dev)))
    return out, count


# ------------------------------------------------------------------------------------------------
# S1a
# ------------------------------------------------------------------------------------------------
def bin_sort(points: torch.Tensor, W: int, H: int,
             pc_range: Sequence[float] = DEFAULT_PC_RANGE) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """points (B,N,C) -> cell (B,N) i32, perm (B,N) i32, offsets (B,H*W+1) i32."""
    points = _need_cuda(points, "points")
    if points.dim() != 3:
        raise ValueError("points must be (B, N, C)")
    B, N, Cc = points.shape
    dev = points.device
    cell = torch.empty((B, N), dtype=torch.int32, device=dev)
    perm = torch.empty((B, N), dtype=torch.int32, device=dev)
    offsets = torch.empty((B, H * W + 1), dtype=torch.int32, device=dev)
    vx, vy = voxel_size(pc_range, W, H)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b200bev_bin_sort(_ptr(points), B, N, Cc, pc_range[0], pc_range[1], vx, vy, W, H,
                                               _ptr(cell), _ptr(perm), _ptr(offsets), _stream(dev)))
    return cell, perm, offsets


def lidar_prepare_bin_sort(raw: torch.Tensor, frame_offsets: torch.Tensor, max_points: int, W: int, H: int,
                           pc_range: Sequence[float] = DEFAULT_PC_RANGE, max_frame_rows: Optional[int] = None):
    """lidar_prepare (without `select`) + bin_sort in ONE launch: raw sweeps -> (points (B,max_points,C), count (B) i32,
    cell (B,max_points) i32, perm (B,max_points) i32, offsets (B,H*W+1) i32), bit-identical to the two calls."""
    raw = _need_cuda(raw, "raw")
    frame_offsets = _need_cuda(frame_offsets, "frame_offsets", torch.int64)
    if raw.dim() != 2 or raw.shape[1] < 3:
        raise ValueError("raw must be (total_rows, C) with C >= 3")
    if frame_offsets.dim() != 1 or frame_offsets.numel() < 2:
        raise ValueError("frame_offsets must be (B + 1,)")
    B, Cc, dev = frame_offsets.numel() - 1, int(raw.shape[1]), raw.device
    if max_frame_rows is None:
        max_frame_rows = int(raw.shape[0])
    points = torch.empty((B, max_points, Cc), dtype=torch.float32, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    cell = torch.empty((B, max_points), dtype=torch.int32, device=dev)
    perm = torch.empty((B, max_points), dtype=torch.int32, device=dev)
    offsets = torch.empty((B, H * W + 1), dtype=torch.int32, device=dev)
    vx, vy = voxel_size(pc_range, W, H)
    rng = (C.c_float * 6)(*[float(v) for v in pc_range])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b200bev_lidar_prepare_bin_sort(_ptr(raw), _ptr(frame_offsets), B, Cc, max_frame_rows, rng, max_points, vx, vy,
                                                             W, H, _ptr(points), _ptr(count), _ptr(cell), _ptr(perm), _ptr(offsets),
                                                             _stream(dev)))
    return points, count, cell, perm, offsets


# ------------------------------------------------------------------------------------------------
# S1b / S1c
# ------------------------------------------------------------------------------------------------
def fold_batchnorm(weight: torch.Tensor, bias: Optional[torch.Tensor], bn) -> Tuple[torch.Tensor, torch.Tensor]:
    """Folds an eval-mode BatchNorm1d (or Identity) into a k=1 Conv1d: returns (W', b') in float64.

    y = g*(Wx + b - mean)/sqrt(var + eps) + beta  ==  (W*s) x + ((b - mean)*s + beta),  s = g/sqrt(var+eps)
    (src/encoders.py:289-295 with bn in eval mode).
    """
    w = weight.detach().double().reshape(weight.shape[0], -1)
    b = torch.zeros(w.shape[0], dtype=torch.float64, device=w.device) if bias is None else bias.detach().double()
    if isinstance(bn, torch.nn.modules.batchnorm._BatchNorm):
        var = bn.running_var.detach().double()
        mean = bn.running_mean.detach().double()
        g = torch.ones_like(var) if bn.weight is None else bn.weight.detach().double()
        beta = torch.zeros_like(var) if bn.bias is None else bn.bias.detach().double()
        s = g / torch.sqrt(var + bn.eps)
        w = w * s[:, None]
        b = (b - mean) * s + beta
    return w, b


def pack_mlp_params(weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
                    device: torch.device) -> Tuple[torch.Tensor, List[int]]:
    """[(C_out,C_in)], [(C_out)] -> (blob fp32 on device, dims). Blob layout: W_l^T then b_l per layer."""
    dims = [int(weights[0].shape[1])] + [int(w.shape[0]) for w in weights]
    parts = []
    for w, b in zip(weights, biases):
        parts.append(w.t().contiguous().reshape(-1))
        parts.append(b.reshape(-1))
    blob = torch.cat([p.to(torch.float64) for p in parts]).to(torch.float32).to(device).contiguous()
    return blob, dims


def pack_mlp_params_bf16(params: torch.Tensor, dims: Sequence[int]) -> torch.Tensor:
    """Device-side re-tiling of the fp32 blob into the tcgen05 stage image (uint8 tensor)."""
    params = _need_cuda(params, "params")
    d = _i32(dims)
    n_bytes = _lib.lib().b200bev_pointnet_pack_bf16_bytes(d, len(dims) - 1)
    if n_bytes == 0:
        raise _lib.B200BevError(_lib.ERR_UNSUPPORTED, f"tensor-core path does not support layer widths {list(dims)}")
    out = torch.empty(n_bytes, dtype=torch.uint8, device=params.device)
    with torch.cuda.device(params.device):
        _lib.check(_lib.lib().b200bev_pointnet_pack_bf16(_ptr(params), d, len(dims) - 1, _ptr(out), n_bytes,
                                                         _stream(params.device)))
    return out


def pack_mlp_params_split(params: torch.Tensor, dims: Sequence[int]) -> Optional[torch.Tensor]:
    """Weight image of the fp32-accuracy tensor-core MLP (split fp16 operands, b200bev_pointnet_pack_split), or None when
    the layer widths are not the C-64-128-256-512-1024 that kernel takes (the FFMA kernel then runs)."""
    params = _need_cuda(params, "params")
    d = _i32(dims)
    n_bytes = _lib.lib().b200bev_pointnet_pack_split_bytes(d, len(dims) - 1)
    if n_bytes == 0:
        return None
    out = torch.empty(n_bytes, dtype=torch.uint8, device=params.device)
    with torch.cuda.device(params.device):
        _lib.check(_lib.lib().b200bev_pointnet_pack_split(_ptr(params), d, len(dims) - 1, _ptr(out), n_bytes, _stream(params.device)))
    return out


def pointnet_encode(points: torch.Tensor, params: torch.Tensor, dims: Sequence[int],
                    perm: Optional[torch.Tensor] = None, offsets: Optional[torch.Tensor] = None,
                    n_cells: int = 0, precision: int = _lib.F32, tc_params: Optional[torch.Tensor] = None,
                    want_global: bool = True, want_canvas: Optional[bool] = None):
    """Fused shared-MLP + max.  Returns the global maxima (B, C_out) and/or the per-cell canvas
    (B, n_cells, C_out): one tensor if one was asked for, the pair (global, canvas) if both.
    The canvas needs perm/offsets from bin_sort; by default it is produced whenever they are given.
    precision / tc_params: F32 + None -> FFMA kernel (any layer widths); F32 + the image of pack_mlp_params_split -> the
    same fp32 accuracy on the tensor cores; BF16_TENSOR + the image of pack_mlp_params_bf16 -> bf16 tcgen05 (1e-2)."""
    points = _need_cuda(points, "points")
    params = _need_cuda(params, "params")
    if points.dim() != 3:
        raise ValueError("points must be (B, N, C)")
    B, N, Cc = points.shape
    if Cc != dims[0]:
        raise ValueError(f"points have {Cc} channels, the MLP expects {dims[0]}")
    if want_canvas is None:
        want_canvas = perm is not None
    if want_canvas and perm is None:
        raise ValueError("the per-cell canvas needs perm/offsets from bin_sort")
    if not (want_global or want_canvas):
        raise ValueError("nothing to compute")
    dev = points.device
    c_out = int(dims[-1])
    if perm is not None:
        perm = _need_cuda(perm, "perm", torch.int32)
        offsets = _need_cuda(offsets, "offsets", torch.int32)
    glob = torch.empty((B, c_out), dtype=torch.float32, device=dev) if want_global else None
    canvas = torch.empty((B, n_cells, c_out), dtype=torch.float32, device=dev) if want_canvas else None
    if precision == _lib.F32 and tc_params is not None:
        # fp32 accuracy on the tensor cores: the split-fp16 image of pack_mlp_params_split was handed in
        tc_params = _need_cuda(tc_params, "tc_params", torch.uint8)
        ws = torch.empty(_lib.lib().b200bev_pointnet_split_workspace_bytes(B, N), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().b200bev_pointnet_encode_split(
                _ptr(points), B, N, Cc, _i32(dims), len(dims) - 1, _ptr(perm), _ptr(offsets), n_cells, _ptr(tc_params),
                _ptr(glob), _ptr(canvas), _ptr(ws), ws.numel(), _stream(dev)))
        if want_global and want_canvas:
            return glob, canvas
        return glob if want_global else canvas
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b200bev_pointnet_encode(
            _ptr(points), B, N, Cc, _ptr(params), _i32(dims), len(dims) - 1, _ptr(perm), _ptr(offsets), n_cells,
            precision, _ptr(tc_params), _ptr(glob), _ptr(canvas), _stream(dev)))
    if want_global and want_canvas:
        return glob, canvas
    return glob if want_global else canvas


def radar_encode(radar_list: Sequence[torch.Tensor], params: torch.Tensor, dims: Sequence[int], fusion: str,
                 fc_weight: Optional[torch.Tensor], fc_bias: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """List of R tensors (B,N_r,C) -> (fused (B,F), per-radar maxima (B,R,F))."""
    if fusion not in _lib.RADAR_FUSION:
        raise ValueError(f"Unknown fusion method: {fusion}")  # src/encoders.py:659
    radars = [_need_cuda(r, f"radar_list[{i}]") for i, r in enumerate(radar_list)]
    if not radars:
        raise ValueError("radar_list is empty")
    params = _need_cuda(params, "params")
    B, _, Cc = radars[0].shape
    for r in radars:
        if r.dim() != 3 or r.shape[0] != B or r.shape[2] != Cc:
            raise ValueError("every radar tensor must be (B, N_r, C) with the same B and C")
    R, F = len(radars), int(dims[-1])
    dev = radars[0].device
    per_radar = torch.empty((B, R, F), dtype=torch.float32, device=dev)
    out = torch.empty((B, F), dtype=torch.float32, device=dev)
    if fusion == "concat":
        fc_weight = _need_cuda(fc_weight, "fc_weight")
        fc_bias = _need_cuda(fc_bias, "fc_bias")
        if tuple(fc_weight.shape) != (F, R * F):
            # the reference's Linear would fail the same way
            raise RuntimeError(f"fusion_fc expects {fc_weight.shape[1] // F} radars, got {R}")
    ptrs = (C.c_void_p * R)(*[r.data_ptr() for r in radars])
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b200bev_radar_encode(
            ptrs, _i32([r.shape[1] for r in radars]), R, B, Cc, _ptr(params), _i32(dims), len(dims) - 1,
            _lib.RADAR_FUSION[fusion], _ptr(fc_weight), _ptr(fc_bias), _ptr(per_radar), _ptr(out), _stream(dev)))
    return out, per_radar


# ------------------------------------------------------------------------------------------------
# S2
# ------------------------------------------------------------------------------------------------
def camera_mean(feats: torch.Tensor) -> torch.Tensor:
    """(B, n_cam, C, h, w) -> (B, C, h, w): camera_features.mean(dim=1) (src/fusion.py:234)."""
    feats = _need_cuda(feats, "camera_features")
    if feats.dim() < 3:
        raise ValueError("camera_features must be (B, n_cam, ...)")
    B, n_cam = feats.shape[:2]
    out = torch.empty((B, *feats.shape[2:]), dtype=torch.float32, device=feats.device)
    inner = out[0].numel()
    with torch.cuda.device(feats.device):
        _lib.check(_lib.lib().b200bev_camera_mean(_ptr(feats), B, n_cam, inner, _ptr(out), _stream(feats.device)))
    return out


def bilinear_resize(x: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """F.interpolate(x, size=size, mode='bilinear', align_corners=False) (src/fusion.py:242-247)."""
    x = _need_cuda(x, "input")
    if x.dim() != 4:
        raise ValueError("input must be (B, C, h, w)")
    B, Cc, h, w = x.shape
    H, W = int(size[0]), int(size[1])
    out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().b200bev_bilinear_resize(_ptr(x), B, Cc, h, w, _ptr(out), H, W, _stream(x.device)))
    return out


def camera_project(feats: torch.Tensor, intrinsics: torch.Tensor, ego2cam: torch.Tensor,
                   img_size: Tuple[float, float], bev_size: Tuple[int, int],
                   pc_range: Sequence[float] = DEFAULT_PC_RANGE, z_plane: float = 0.0,
                   return_table: bool = False, impl: str = "auto"):
    """Geometric camera->BEV gather. feats (B,n_cam,C,h,w); intrinsics (T,n_cam,3,3); ego2cam (T,n_cam,3,4);
    img_size = (img_w, img_h) in pixels; bev_size = (H, W). Returns (B,C,H,W) [, table (T,H*W,n_cam,3)].
    impl: "auto", or "staged" / "gather" to name one of the two kernels (same bits; used by the parity tests)."""
    if impl not in _lib.PROJECT_IMPL:
        raise ValueError(f"impl must be one of {sorted(_lib.PROJECT_IMPL)}")
    feats = _need_cuda(feats, "camera_features")
    intrinsics = _need_cuda(intrinsics, "intrinsics")
    ego2cam = _need_cuda(ego2cam, "ego2cam")
    if feats.dim() != 5:
        raise ValueError("camera_features must be (B, n_cam, C, h, w)")
    B, n_cam, Cc, h, w = feats.shape
    if intrinsics.dim() == 3:
        intrinsics = intrinsics.unsqueeze(0).contiguous()
    if ego2cam.dim() == 3:
        ego2cam = ego2cam.unsqueeze(0).contiguous()
    T = intrinsics.shape[0]
    if tuple(intrinsics.shape) != (T, n_cam, 3, 3) or tuple(ego2cam.shape) != (T, n_cam, 3, 4):
        raise ValueError("intrinsics must be (T,n_cam,3,3) and ego2cam (T,n_cam,3,4)")
    H, W = int(bev_size[0]), int(bev_size[1])
    vx, vy = voxel_size(pc_range, W, H)
    dev = feats.device
    out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=dev)
    table = torch.empty((T, H * W, n_cam, 3), dtype=torch.float32, device=dev) if return_table else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b200bev_camera_project(
            _ptr(feats), B, n_cam, Cc, h, w, _ptr(intrinsics), _ptr(ego2cam), T, float(img_size[0]), float(img_size[1]),
            pc_range[0], pc_range[1], vx, vy, z_plane, W, H, _ptr(out), _ptr(table), _lib.PROJECT_IMPL[impl], _stream(dev)))
    return (out, table) if return_table else out


# ------------------------------------------------------------------------------------------------
# N2: dense layers around the canvas
# ------------------------------------------------------------------------------------------------
def dense_layer(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False) -> torch.Tensor:
    """act(x @ weight.T + bias): nn.Linear (+ ReLU) on a small batch, e.g. radar_proj (src/fusion.py:170-173).
    x (B,K), weight (O,K) in torch's nn.Linear layout, bias (O) -> (B,O)."""
    x = _need_cuda(x, "input")
    weight = _need_cuda(weight.detach(), "weight")
    bias = None if bias is None else _need_cuda(bias.detach(), "bias")
    if x.dim() != 2 or weight.dim() != 2 or weight.shape[1] != x.shape[1]:
        raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({tuple(x.shape)} and {tuple(weight.t().shape)})")
    B, K = x.shape
    O = int(weight.shape[0])
    out = torch.empty((B, O), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().b200bev_dense_layer(_ptr(x), B, K, _ptr(weight), _ptr(bias), O, 1 if relu else 0, _ptr(out),
                                                  _stream(x.device)))
    return out


def lidar_init(feats: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor,
               return_hidden: bool = False):
    """Linear + ReLU + Linear of FlexibleBEVFusion.lidar_init (src/fusion.py:144-148): (B,K) -> (B,O)."""
    feats = _need_cuda(feats, "lidar_features")
    w1, b1, w2, b2 = (_need_cuda(t.detach(), n) for t, n in ((w1, "w1"), (b1, "b1"), (w2, "w2"), (b2, "b2")))
    B, K = feats.shape
    hidden, O = int(w1.shape[0]), int(w2.shape[0])
    if w1.shape[1] != K or w2.shape[1] != hidden:
        raise RuntimeError(f"lidar_init: shapes do not chain: x {tuple(feats.shape)}, w1 {tuple(w1.shape)}, w2 {tuple(w2.shape)}")
    hid = torch.empty((B, hidden), dtype=torch.float32, device=feats.device)
    out = torch.empty((B, O), dtype=torch.float32, device=feats.device)
    with torch.cuda.device(feats.device):
        _lib.check(_lib.lib().b200bev_lidar_init(_ptr(feats), B, K, _ptr(w1), _ptr(b1), hidden, _ptr(w2), _ptr(b2), O,
                                                 _ptr(hid), _ptr(out), _stream(feats.device)))
    return (out, hid) if return_hidden else out


def dense_pack_split(weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """(O,K) fp32 nn.Linear weight (+ bias) on the device -> the split-fp16 stage image of b200bev_dense_layer_split
    (uint8 tensor of the weight's own size), or None when the shape has no tensor-core form (O % 128, K % 64)."""
    weight = _need_cuda(weight.detach(), "weight")
    O, K = (int(v) for v in weight.shape)
    n = _lib.lib().b200bev_dense_pack_split_bytes(O, K)
    if n == 0:
        return None
    if bias is not None:
        bias = _need_cuda(bias.detach(), "bias")
    img = torch.empty(n, dtype=torch.uint8, device=weight.device)
    with torch.cuda.device(weight.device):
        _lib.check(_lib.lib().b200bev_dense_pack_split(_ptr(weight), _ptr(bias) if bias is not None else None, O, K,
                                                       _ptr(img), n, _stream(weight.device)))
    return img


def dense_layer_split(x: torch.Tensor, image: torch.Tensor, out_features: int, relu: bool = False) -> torch.Tensor:
    """act(x W^T + b) from a `dense_pack_split` image: fp32 accuracy (1e-5) on the tensor cores."""
    x = _need_cuda(x, "x")
    B, K = (int(v) for v in x.shape)
    out = torch.empty((B, out_features), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().b200bev_dense_layer_split(_ptr(x), B, K, _ptr(image), int(out_features), int(bool(relu)),
                                                        _ptr(out), _stream(x.device)))
    return out


def lidar_init_split(feats: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, image2: torch.Tensor, out_features: int) -> torch.Tensor:
    """`lidar_init` with the second layer read from its `dense_pack_split` image (b200bev_lidar_init_split)."""
    feats = _need_cuda(feats, "lidar_features")
    w1, b1 = _need_cuda(w1.detach(), "w1"), _need_cuda(b1.detach(), "b1")
    B, K = (int(v) for v in feats.shape)
    hidden = int(w1.shape[0])
    if w1.shape[1] != K:
        raise RuntimeError(f"lidar_init: shapes do not chain: x {tuple(feats.shape)}, w1 {tuple(w1.shape)}")
    hid = torch.empty((B, hidden), dtype=torch.float32, device=feats.device)
    out = torch.empty((B, int(out_features)), dtype=torch.float32, device=feats.device)
    with torch.cuda.device(feats.device):
        _lib.check(_lib.lib().b200bev_lidar_init_split(_ptr(feats), B, K, _ptr(w1), _ptr(b1), hidden, _ptr(image2),
                                                       int(out_features), _ptr(hid), _ptr(out), _stream(feats.device)))
    return out


# ------------------------------------------------------------------------------------------------
# N1: convolution blocks on tcgen05 (bf16, parity 1e-2)
# ------------------------------------------------------------------------------------------------
def fold_conv_bn(conv: torch.nn.Conv2d, bn: Optional[torch.nn.Module]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Eval-mode BatchNorm2d folded into the convolution in float64 -> (weight (Cout,Cin,kh,kw), bias (Cout)) fp32."""
    w, b = fold_batchnorm(conv.weight, conv.bias, bn)
    return w.reshape(conv.weight.shape).float().contiguous(), b.float().contiguous()


def conv_pack(weight: torch.Tensor) -> torch.Tensor:
    """(Cout,Cin,kh,kw) fp32 on the device -> the swizzled bf16 stage image of b200bev_conv_bn_relu_bf16 (uint8 tensor)."""
    weight = _need_cuda(weight, "weight")
    Cout, Cin, kh, kw = weight.shape
    taps = kh * kw
    n = _lib.lib().b200bev_conv_pack_bytes(Cout, Cin, taps)
    if n == 0:
        raise _lib.B200BevError(_lib.ERR_UNSUPPORTED, f"conv {tuple(weight.shape)}: needs Cin % 64 == 0 and a 3x3 or 1x1 kernel")
    img = torch.empty(n, dtype=torch.uint8, device=weight.device)
    with torch.cuda.device(weight.device):
        _lib.check(_lib.lib().b200bev_conv_pack_bf16(_ptr(weight), Cout, Cin, taps, _ptr(img), n, _stream(weight.device)))
    return img


def nchw_to_nhwc_bf16(parts: Sequence[torch.Tensor], out: Optional[torch.Tensor] = None, c_offset: int = 0) -> torch.Tensor:
    """[(B,C_i,H,W) fp32] -> (B,H,W,sum C_i) bf16 channels-last: layout change, cast and torch.cat in one pass per part.
    `out`: an existing (B,H,W,C_total) bf16 tensor; the parts go to its channels from `c_offset` on."""
    parts = [_need_cuda(p, f"parts[{i}]") for i, p in enumerate(parts)]
    B, _, H, W = parts[0].shape
    c_sum = sum(int(p.shape[1]) for p in parts)
    if out is None:
        out = torch.empty((B, H, W, c_sum), dtype=torch.bfloat16, device=parts[0].device)
        c_offset = 0
    elif out.dtype != torch.bfloat16 or not out.is_contiguous() or tuple(out.shape[:3]) != (B, H, W) or c_offset + c_sum > out.shape[3]:
        raise ValueError(f"out must be a contiguous (B,H,W,C) bf16 tensor with room for {c_sum} channels at offset {c_offset}")
    c_total = int(out.shape[3])
    off = c_offset
    with torch.cuda.device(out.device):
        for p in parts:
            if p.shape[0] != B or tuple(p.shape[2:]) != (H, W):
                raise RuntimeError("Sizes of tensors must match except in dimension 1")      # what torch.cat says
            _lib.check(_lib.lib().b200bev_nchw_to_nhwc_bf16(_ptr(p), B, int(p.shape[1]), H, W, _ptr(out), c_total, off,
                                                            _stream(out.device)))
            off += int(p.shape[1])
    return out


def bilinear_resize_nhwc_bf16(x: torch.Tensor, size: Tuple[int, int], out: Optional[torch.Tensor] = None, c_offset: int = 0) -> torch.Tensor:
    """(B,h,w,C) bf16 channels-last -> (B,H,W,C) bf16, or channels [c_offset, c_offset + C) of `out` (B,H,W,C_total) bf16:
    F.interpolate(mode='bilinear', align_corners=False) (src/fusion.py:242-247) between two convolutions of the bf16 path."""
    x = _need_cuda(x, "x", torch.bfloat16)
    if x.dim() != 4:
        raise ValueError("x must be (B, h, w, C) channels-last")
    B, h, w, Cc = (int(v) for v in x.shape)
    H, W = int(size[0]), int(size[1])
    if out is None:
        out = torch.empty((B, H, W, Cc), dtype=torch.bfloat16, device=x.device)
        c_offset = 0
    elif out.dtype != torch.bfloat16 or not out.is_contiguous() or tuple(out.shape[:3]) != (B, H, W) or c_offset + Cc > out.shape[3]:
        raise ValueError(f"out must be a contiguous (B,H,W,C) bf16 tensor with room for {Cc} channels at offset {c_offset}")
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().b200bev_bilinear_resize_nhwc_bf16(_ptr(x), B, h, w, Cc, _ptr(out), H, W, int(out.shape[3]), c_offset,
                                                                _stream(x.device)))
    return out


def camera_mean_nhwc_bf16(feats: torch.Tensor) -> torch.Tensor:
    """(B,n_cam,C,h,w) fp32 -> (B,h,w,C) bf16: camera_features.mean(dim=1) (src/fusion.py:234) delivered as the
    channels-last input of camera_proj's first convolution."""
    feats = _need_cuda(feats, "camera_features")
    if feats.dim() != 5:
        raise ValueError("camera_features must be (B, n_cam, C, h, w)")
    B, n_cam, Cc, h, w = feats.shape
    out = torch.empty((B, h, w, Cc), dtype=torch.bfloat16, device=feats.device)
    with torch.cuda.device(feats.device):
        _lib.check(_lib.lib().b200bev_camera_mean_nhwc_bf16(_ptr(feats), B, n_cam, Cc, h, w, _ptr(out), Cc, 0, _stream(feats.device)))
    return out


def conv_bn_relu_bf16(x_nhwc: torch.Tensor, image: torch.Tensor, bias: Optional[torch.Tensor], c_out: int, taps: int,
                      relu: bool = True, out_nhwc: Optional[torch.Tensor] = None, c_offset: int = 0, want_nchw: bool = True):
    """(B,H,W,Cin) bf16 -> (B,Cout,H,W) fp32: 3x3 (padding 1) or 1x1 convolution + folded BatchNorm + ReLU on tcgen05.
    out_nhwc: a (B,H,W,C_total) bf16 tensor whose channels [c_offset, c_offset+Cout) also receive the result — the input
    layout of the next convolution; with want_nchw=False only that is written and it is what the call returns."""
    x_nhwc = _need_cuda(x_nhwc, "input", torch.bfloat16)
    image = _need_cuda(image, "weight_image", torch.uint8)
    bias = None if bias is None else _need_cuda(bias, "bias")
    B, H, W, Cin = x_nhwc.shape
    if image.numel() != _lib.lib().b200bev_conv_pack_bytes(c_out, Cin, taps):
        raise ValueError("weight image does not belong to this (Cout, Cin, taps)")
    dev = x_nhwc.device
    out = torch.empty((B, c_out, H, W), dtype=torch.float32, device=dev) if want_nchw or out_nhwc is None else None
    with torch.cuda.device(dev):
        if out_nhwc is None:
            _lib.check(_lib.lib().b200bev_conv_bn_relu_bf16(_ptr(x_nhwc), B, H, W, Cin, _ptr(image), _ptr(bias), c_out, taps,
                                                            1 if relu else 0, _ptr(out), _stream(dev)))
            return out
        if out_nhwc.dtype != torch.bfloat16 or not out_nhwc.is_contiguous() or tuple(out_nhwc.shape[:3]) != (B, H, W):
            raise ValueError(f"out_nhwc must be a contiguous (B,H,W,C) bf16 tensor matching {(B, H, W)}")
        _lib.check(_lib.lib().b200bev_conv_bn_relu_bf16_nhwc(_ptr(x_nhwc), B, H, W, Cin, _ptr(image), _ptr(bias), c_out, taps,
                                                             1 if relu else 0, _ptr(out_nhwc), int(out_nhwc.shape[3]), c_offset,
                                                             _ptr(out), _stream(dev)))
    return out if out is not None else out_nhwc


# ---- the same blocks at fp32 accuracy (three fp16 tensor-core products per fp32 product, parity 1e-5) --------------------
def conv_pack_split(weight: torch.Tensor) -> torch.Tensor:
    """(Cout,Cin,kh,kw) fp32 on the device -> the split-fp16 stage image of b200bev_conv_bn_relu_split (uint8 tensor)."""
    weight = _need_cuda(weight, "weight")
    Cout, Cin, kh, kw = weight.shape
    taps = kh * kw
    n = _lib.lib().b200bev_conv_pack_split_bytes(Cout, Cin, taps)
    if n == 0:
        raise _lib.B200BevError(_lib.ERR_UNSUPPORTED, f"conv {tuple(weight.shape)}: needs Cin % 64 == 0 and a 3x3 or 1x1 kernel")
    img = torch.empty(n, dtype=torch.uint8, device=weight.device)
    with torch.cuda.device(weight.device):
        _lib.check(_lib.lib().b200bev_conv_pack_split(_ptr(weight), Cout, Cin, taps, _ptr(img), n, _stream(weight.device)))
    return img


def nchw_to_nhwc_split(parts: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """[(B,C_i,H,W) fp32] -> ((B,H,W,2*sum C_i) fp16 [hi | lo], stat): max|x| over all parts (one read), then layout change,
    power-of-two scaling, hi/lo split and torch.cat in one pass per part."""
    parts = [_need_cuda(p, f"parts[{i}]") for i, p in enumerate(parts)]
    B, _, H, W = parts[0].shape
    c_sum = sum(int(p.shape[1]) for p in parts)
    dev = parts[0].device
    stat = torch.zeros(1, dtype=torch.int32, device=dev)
    out = torch.empty((B, H, W, 2 * c_sum), dtype=torch.float16, device=dev)
    with torch.cuda.device(dev):
        for p in parts:
            if p.shape[0] != B or tuple(p.shape[2:]) != (H, W):
                raise RuntimeError("Sizes of tensors must match except in dimension 1")      # what torch.cat says
            _lib.check(_lib.lib().b200bev_absmax(_ptr(p), p.numel(), _ptr(stat), _stream(dev)))
        off = 0
        for p in parts:
            _lib.check(_lib.lib().b200bev_nchw_to_nhwc_split(_ptr(p), B, int(p.shape[1]), H, W, _ptr(out), c_sum, off, _ptr(stat),
                                                            _stream(dev)))
            off += int(p.shape[1])
    return out, stat


def conv_bn_relu_split(x_split: torch.Tensor, stat: torch.Tensor, image: torch.Tensor, bias: Optional[torch.Tensor], c_out: int,
                       taps: int, relu: bool = True) -> torch.Tensor:
    """(B,H,W,2*Cin) fp16 [hi | lo] -> (B,Cout,H,W) fp32: 3x3 (padding 1) or 1x1 convolution + folded BatchNorm + ReLU at
    fp32 accuracy on tcgen05 (b200bev_conv_bn_relu_split)."""
    x_split = _need_cuda(x_split, "input", torch.float16)
    image = _need_cuda(image, "weight_image", torch.uint8)
    stat = _need_cuda(stat, "stat", torch.int32)
    bias = None if bias is None else _need_cuda(bias, "bias")
    B, H, W, C2 = x_split.shape
    Cin = C2 // 2
    if image.numel() != _lib.lib().b200bev_conv_pack_split_bytes(c_out, Cin, taps):
        raise ValueError("weight image does not belong to this (Cout, Cin, taps)")
    out = torch.empty((B, c_out, H, W), dtype=torch.float32, device=x_split.device)
    with torch.cuda.device(x_split.device):
        _lib.check(_lib.lib().b200bev_conv_bn_relu_split(_ptr(x_split), _ptr(stat), B, H, W, Cin, _ptr(image), _ptr(bias), c_out, taps,
                                                         1 if relu else 0, _ptr(out), _stream(x_split.device)))
    return out


def border_class_index(n: int, s: int) -> List[int]:
    """cls(i) for i in range(n): which row (column) of the s x s image a row (column) of an n-wide image equals when
    k = s // 2 padded 3x3 convolutions ran over a spatially constant input."""
    k = s // 2
    return [i if i < k else (s - (n - i) if i >= n - k else k) for i in range(n)]


def border_expand(small: torch.Tensor, size: Tuple[int, int], out_nhwc: Optional[torch.Tensor] = None, c_offset: int = 0,
                  want_nchw: bool = True) -> Optional[torch.Tensor]:
    """(B,C,s,s) -> (B,C,H,W): out[y][x] = small[cls(y)][cls(x)] (b200bev_border_expand), optionally (also) into channels
    [c_offset, c_offset+C) of a channels-last bf16 tensor."""
    small = _need_cuda(small, "small")
    B, Cc, s, s2 = small.shape
    H, W = int(size[0]), int(size[1])
    if s != s2 or s % 2 == 0 or H < s or W < s:
        raise ValueError(f"border_expand: small must be (B,C,s,s) with odd s <= H, W; got {tuple(small.shape)} for {(H, W)}")
    out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=small.device) if want_nchw or out_nhwc is None else None
    c_total = 0
    if out_nhwc is not None:
        if out_nhwc.dtype != torch.bfloat16 or not out_nhwc.is_contiguous() or tuple(out_nhwc.shape[:3]) != (B, H, W):
            raise ValueError(f"out_nhwc must be a contiguous (B,H,W,C) bf16 tensor matching {(B, H, W)}")
        c_total = int(out_nhwc.shape[3])
    with torch.cuda.device(small.device):
        _lib.check(_lib.lib().b200bev_border_expand(_ptr(small), B, Cc, s, H, W, _ptr(out), _ptr(out_nhwc), c_total, c_offset,
                                                    _stream(small.device)))
    return out


# ------------------------------------------------------------------------------------------------
# S3
# ------------------------------------------------------------------------------------------------
def centernet_nms(heat: torch.Tensor) -> torch.Tensor:
    """_nms(heat, kernel=3) (src/centernet_target.py:416-421)."""
    heat = _need_cuda(heat, "heat")
    if heat.dim() != 4:
        raise ValueError("heat must be (B, C, H, W)")
    B, Cc, H, W = heat.shape
    out = torch.empty_like(heat)
    with torch.cuda.device(heat.device):
        _lib.check(_lib.lib().b200bev_centernet_nms(_ptr(heat), B, Cc, H, W, _ptr(out), _stream(heat.device)))
    return out


def _workspace(B: int, Cc: int, K: int, dev: torch.device) -> torch.Tensor:
    n = _lib.lib().b200bev_centernet_workspace_bytes(B, Cc, K)
    return torch.empty(max(n, 16), dtype=torch.uint8, device=dev)


def centernet_topk(scores: torch.Tensor, K: int):
    """_topk(scores, K) (src/centernet_target.py:424-452): (score, ind, classes, ys, xs), each (B,K)."""
    scores = _need_cuda(scores, "scores")
    B, Cc, H, W = scores.shape
    dev = scores.device
    ws = _workspace(B, Cc, K, dev)
    top = torch.empty((B, K), dtype=torch.float32, device=dev)
    ind, cls, ys, xs = (torch.empty((B, K), dtype=torch.int64, device=dev) for _ in range(4))
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b200bev_centernet_topk(_ptr(scores), B, Cc, H, W, K, _ptr(top), _ptr(ind), _ptr(cls),
                                                     _ptr(ys), _ptr(xs), _ptr(ws), ws.numel(), _stream(dev)))
    return top, ind, cls, ys, xs


def centernet_decode(heatmap: torch.Tensor, offset: torch.Tensor, size: torch.Tensor, rot: torch.Tensor,
                     vel: torch.Tensor, K: int, voxel: float, origin: Tuple[float, float] = (-51.2, -51.2),
                     z_value: float = -1.0, score_thresh: float = 0.0, heat_is_logit: bool = False):
    """Fused NMS + top-K + gather + box assembly. Fixed-size device outputs:
    dict(boxes (B,K,7), scores (B,K), labels (B,K) i64, velocities (B,K,2), ys, xs, ind (B,K) i64, count (B) i32).
    heat_is_logit: `heatmap` is the head's raw output and torch.sigmoid (src/fusion.py:871) is applied inside the kernel."""
    heatmap = _need_cuda(heatmap, "heatmap")
    B, Cc, H, W = heatmap.shape
    maps = {"offset": (offset, 2), "size": (size, 3), "rot": (rot, 2), "vel": (vel, 2)}
    fixed = {}
    for name, (t, ch) in maps.items():
        t = _need_cuda(t, name)
        if tuple(t.shape) != (B, ch, H, W):
            raise ValueError(f"{name} must be {(B, ch, H, W)}, got {tuple(t.shape)}")
        fixed[name] = t
    dev = heatmap.device
    ws = _workspace(B, Cc, K, dev)
    o = {
        "boxes": torch.empty((B, K, 7), dtype=torch.float32, device=dev),
        "scores": torch.empty((B, K), dtype=torch.float32, device=dev),
        "labels": torch.empty((B, K), dtype=torch.int64, device=dev),
        "velocities": torch.empty((B, K, 2), dtype=torch.float32, device=dev),
        "ys": torch.empty((B, K), dtype=torch.int64, device=dev),
        "xs": torch.empty((B, K), dtype=torch.int64, device=dev),
        "ind": torch.empty((B, K), dtype=torch.int64, device=dev),
        "count": torch.empty((B,), dtype=torch.int32, device=dev),
    }
    with torch.cuda.device(dev):
        entry = _lib.lib().b200bev_centernet_decode_logits if heat_is_logit else _lib.lib().b200bev_centernet_decode
        _lib.check(entry(
            _ptr(heatmap), _ptr(fixed["offset"]), _ptr(fixed["size"]), _ptr(fixed["rot"]), _ptr(fixed["vel"]),
            B, Cc, H, W, K, voxel, origin[0], origin[1], z_value, score_thresh,
            _ptr(o["boxes"]), _ptr(o["scores"]), _ptr(o["labels"]), _ptr(o["velocities"]),
            _ptr(o["ys"]), _ptr(o["xs"]), _ptr(o["ind"]), _ptr(o["count"]), _ptr(ws), ws.numel(), _stream(dev)))
    return o
