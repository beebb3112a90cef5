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
