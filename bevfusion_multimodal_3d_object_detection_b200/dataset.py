"""Input side of S1: the LiDAR preprocessing of the reference's dataset on the GPU.

Mirrors `NuScenesDataset._load_lidar_points` / `_pad_or_subsample` (src/train_detect.py:147-189): read
the float32 .bin sweep as (-1, 4), keep the points strictly inside the point-cloud range, pad with
zero rows to `max_points` or draw `max_points` of them without replacement.  The reference does this per
sample in numpy inside a DataLoader worker (a CPU process); this module does it for a whole batch in
one kernel launch, after one host->device copy of the raw sweeps.  It is an additional entry point —
`patch()` does not rebind the dataset class, whose workers must stay CUDA-free.
"""
from __future__ import annotations

from pathlib import Path
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import ops

DEFAULT_MAX_POINTS = 35000   # NuScenesDataset falls back to this (src/train_detect.py:57-63)


def read_sweep(path: Union[str, Path], channels: int = 4) -> np.ndarray:
    """np.fromfile(path, float32).reshape(-1, channels) (src/train_detect.py:151)."""
    return np.fromfile(str(path), dtype=np.float32).reshape(-1, channels)


def prepare_lidar_batch(sweeps: Sequence[Union[np.ndarray, torch.Tensor, str, Path]], device: torch.device,
                        max_points: int = DEFAULT_MAX_POINTS, pc_range: Sequence[float] = ops.DEFAULT_PC_RANGE,
                        rng: Optional[np.random.Generator] = None,
                        subsample: str = "random") -> Tuple[torch.Tensor, torch.Tensor]:
    """Raw sweeps (arrays (M_i, C) or .bin paths) -> (points (B, max_points, C) on `device`, count (B) i32).

    A frame with at least `max_points` points in range is subsampled as the reference does
    (np.random.choice(N, max_points, replace=False), src/train_detect.py:184-186) when subsample ==
    "random" — that needs the counts on the host (one sync) and a second, gathering launch for those
    frames only; subsample == "first" keeps the first max_points in file order with no host round trip."""
    if subsample not in ("random", "first"):
        raise ValueError("subsample must be 'random' or 'first'")
    arrays: List[torch.Tensor] = []
    for s in sweeps:
        if isinstance(s, (str, Path)):
            s = read_sweep(s)
        t = torch.from_numpy(np.ascontiguousarray(s, dtype=np.float32)) if isinstance(s, np.ndarray) else s.to(torch.float32)
        if t.dim() != 2:
            raise ValueError("every sweep must be (M, C)")
        arrays.append(t)
    if not arrays:
        raise ValueError("no sweeps given")
    channels = arrays[0].shape[1]
    if any(t.shape[1] != channels for t in arrays):
        raise ValueError("all sweeps must have the same number of channels")
    rows = [int(t.shape[0]) for t in arrays]
    offsets = torch.tensor([0] + list(np.cumsum(rows)), dtype=torch.int64)
    host = torch.cat([t.cpu() for t in arrays], dim=0) if len(arrays) > 1 else arrays[0].cpu()
    raw = host.pin_memory().to(device, non_blocking=True) if device.type == "cuda" else host
    off_d = offsets.to(device)
    points, count = ops.lidar_prepare(raw, off_d, max_points, pc_range, max_frame_rows=max(rows))
    if subsample == "first":
        return points, count
    counts = count.tolist()                                   # the one host sync
    over = [b for b, n in enumerate(counts) if n >= max_points]
    if over:
        rng = rng if rng is not None else np.random.default_rng()
        select = torch.full((len(arrays), max_points), -1, dtype=torch.int32)
        for b in over:
            select[b] = torch.from_numpy(rng.choice(counts[b], max_points, replace=False).astype(np.int32))
        gathered, _ = ops.lidar_prepare(raw, off_d, max_points, pc_range, select=select.to(device), max_frame_rows=max(rows))
        idx = torch.tensor(over, device=device)
        points[idx] = gathered[idx]
    return points, count


# ------------------------------------------------------------------------------------------------
# N4: calibration plumbing for the geometric camera projection
# ------------------------------------------------------------------------------------------------
CAMERA_ORDER = ("CAM_FRONT", "CAM_FRONT_RIGHT", "CAM_FRONT_LEFT", "CAM_BACK", "CAM_BACK_LEFT", "CAM_BACK_RIGHT")


def quaternion_to_matrix(q: Sequence[float]) -> np.ndarray:
    """nuScenes quaternion (w, x, y, z) -> 3x3 rotation, float64."""
    w, x, y, z = (float(v) for v in q)
    n = np.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def calibration_from_info(info: dict, cameras: Sequence[str] = CAMERA_ORDER, frame: str = "lidar") -> Tuple[np.ndarray, np.ndarray]:
    """One sample's info dict, as ConfigDrivenNuScenesConverter writes it (src/data_converter.py:110-117,
    145-152), -> (intrinsics (n_cam,3,3) f32, ego2cam (n_cam,3,4) f32) in the layout b200bev_camera_project takes.

    `calibrated_sensor` holds sensor->ego transforms (translation, rotation quaternion w,x,y,z).  The BEV grid of
    the reference lives in the LiDAR frame (boxes are moved there, src/data_converter.py:237-247), so with
    frame == "lidar" the returned [R|t] maps LiDAR-frame points to each camera: p_cam = R_c^T (R_l p + t_l - t_c).
    frame == "ego" maps ego-frame points.  NuScenesDataset never reads these fields; this is the
    plumbing the geometric projection needs on real data."""
    if frame not in ("lidar", "ego"):
        raise ValueError("frame must be 'lidar' or 'ego'")
    if frame == "lidar":
        lc = info["lidar_calibrated_sensor"]
        R_l, t_l = quaternion_to_matrix(lc["rotation"]), np.asarray(lc["translation"], dtype=np.float64)
    else:
        R_l, t_l = np.eye(3), np.zeros(3)
    Ks, Es = [], []
    for cam in cameras:
        cs = info["cams"][cam]["calibrated_sensor"]
        R_c, t_c = quaternion_to_matrix(cs["rotation"]), np.asarray(cs["translation"], dtype=np.float64)
        R = R_c.T @ R_l
        t = R_c.T @ (t_l - t_c)
        Ks.append(np.asarray(cs["camera_intrinsic"], dtype=np.float64).reshape(3, 3))
        Es.append(np.concatenate([R, t[:, None]], axis=1))
    return np.stack(Ks).astype(np.float32), np.stack(Es).astype(np.float32)


def calibration_batch(infos: Sequence[dict], device: torch.device, **kw) -> Tuple[torch.Tensor, torch.Tensor]:
    """(intrinsics (B,n_cam,3,3), ego2cam (B,n_cam,3,4)) on `device` for a batch of info dicts (T = B rigs)."""
    pairs = [calibration_from_info(i, **kw) for i in infos]
    K = torch.from_numpy(np.stack([p[0] for p in pairs])).to(device)
    E = torch.from_numpy(np.stack([p[1] for p in pairs])).to(device)
    return K, E
