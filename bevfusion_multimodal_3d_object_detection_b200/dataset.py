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


# ------------------------------------------------------------------------------------------------
# N3: radar sweeps read from the files (the reference feeds noise: src/train_detect.py:171-177)
# ------------------------------------------------------------------------------------------------
RADAR_ORDER = ("RADAR_FRONT", "RADAR_FRONT_LEFT", "RADAR_FRONT_RIGHT", "RADAR_BACK_LEFT", "RADAR_BACK_RIGHT")   # src/train_detect.py:163-164
# The model takes 7 channels per radar return (base.yaml:199) and the reference never says which — its loader draws
# randn(125, 7).  Of the 18 fields of a nuScenes radar .pcd these are the geometric and kinematic ones: position,
# ego-motion-compensated velocity, radar cross-section, dynamic property.
RADAR_FIELDS = ("x", "y", "z", "vx_comp", "vy_comp", "rcs", "dyn_prop")
DEFAULT_MAX_RADAR_POINTS = 125     # base.yaml:60, src/train_detect.py:63
_PCD_TYPES = {("F", 4): "<f4", ("F", 8): "<f8", ("I", 1): "<i1", ("I", 2): "<i2", ("I", 4): "<i4", ("I", 8): "<i8",
              ("U", 1): "<u1", ("U", 2): "<u2", ("U", 4): "<u4", ("U", 8): "<u8"}


def read_radar_pcd(path: Union[str, Path], fields: Sequence[str] = RADAR_FIELDS) -> np.ndarray:
    """A nuScenes radar sweep (.pcd, version 0.7, `DATA binary`: one packed record of 18 mixed-type fields per return) ->
    (n, len(fields)) float32 in the sensor frame.  Only the header keywords nuScenes writes are understood."""
    raw = Path(path).read_bytes()
    header, pos = {}, 0
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        key, _, rest = line.partition(" ")
        header[key.upper()] = rest.split()
        if key.upper() == "DATA":
            break
    if header["DATA"] != ["binary"]:
        raise ValueError(f"{path}: only `DATA binary` radar sweeps are supported, got {header['DATA']}")
    names = header["FIELDS"]
    counts = [int(c) for c in header.get("COUNT", ["1"] * len(names))]
    dtype = np.dtype([(n, _PCD_TYPES[(t, int(s))], (c,) if c > 1 else ()) for n, t, s, c in
                      zip(names, header["TYPE"], header["SIZE"], counts)])
    n_points = int(header["POINTS"][0]) if "POINTS" in header else int(header["WIDTH"][0]) * int(header.get("HEIGHT", ["1"])[0])
    body = raw[pos:pos + n_points * dtype.itemsize]
    if len(body) != n_points * dtype.itemsize:
        raise ValueError(f"{path}: {n_points} records of {dtype.itemsize} bytes announced, {len(body)} bytes present")
    rec = np.frombuffer(body, dtype=dtype, count=n_points)
    missing = [f for f in fields if f not in rec.dtype.names]
    if missing:
        raise KeyError(f"{path}: no field(s) {missing}; the file has {list(rec.dtype.names)}")
    return np.stack([rec[f].astype(np.float32) for f in fields], axis=1) if n_points else np.zeros((0, len(fields)), np.float32)


def pad_or_subsample(points: np.ndarray, max_points: int, rng: Optional[np.random.Generator] = None) -> np.ndarray:
    """NuScenesDataset._pad_or_subsample (src/train_detect.py:181-189): zero rows up to max_points, or max_points of the rows
    drawn without replacement."""
    n = points.shape[0]
    if n >= max_points:
        rng = rng if rng is not None else np.random.default_rng()
        return points[rng.choice(n, max_points, replace=False)]
    return np.concatenate([points, np.zeros((max_points - n, points.shape[1]), dtype=points.dtype)], axis=0)


def load_radar_points(info: dict, data_root: Union[str, Path], max_points: int = DEFAULT_MAX_RADAR_POINTS,
                      rng: Optional[np.random.Generator] = None, fields: Sequence[str] = RADAR_FIELDS) -> List[torch.Tensor]:
    """What NuScenesDataset._load_radar_points (src/train_detect.py:160-179) returns — five (max_points, 7) tensors in the
    reference's radar order — with the returns READ from `info['radars'][name]['filename']` (the converter stores it,
    src/data_converter.py:129-135) instead of drawn from randn."""
    root = Path(data_root)
    return [torch.from_numpy(pad_or_subsample(read_radar_pcd(root / info["radars"][name]["filename"], fields), max_points, rng))
            for name in RADAR_ORDER]


# ------------------------------------------------------------------------------------------------
# N4: calibration carried through the batch (the reference's collate drops it: src/train_detect.py:197-242)
# ------------------------------------------------------------------------------------------------
def attach_calibration(item: dict, info: dict, frame: str = "lidar") -> dict:
    """Adds `intrinsics` (6,3,3) and `lidar2cam` (6,3,4) to one dataset item (the dict NuScenesDataset.__getitem__ returns,
    src/train_detect.py:84-121), from the sample's info dict (src/data_converter.py:110-117).  Use it in a thin Dataset
    wrapper: `item = base[i]; attach_calibration(item, base.infos[i])`."""
    K, E = calibration_from_info(info, frame=frame)
    item["intrinsics"], item["lidar2cam"] = torch.from_numpy(K), torch.from_numpy(E)
    return item


def collate_with_calibration(batch: Sequence[dict]) -> dict:
    """collate_fn of the reference (src/train_detect.py:197-242: images and points stacked, radar as 5 stacked tensors,
    ground truth padded to the batch maximum with labels -1) plus the per-sample calibration, stacked to
    `intrinsics` (B,6,3,3) and `lidar2cam` (B,6,3,4): the T = B rig inputs of FlexibleBEVFusion.project_cameras /
    b200bev_camera_project.  Items without calibration collate exactly as the reference's do."""
    n_radar = len(batch[0]["radar_points"])
    out = {
        "camera_imgs": torch.stack([b["camera_imgs"] for b in batch]),
        "lidar_points": torch.stack([b["lidar_points"] for b in batch]),
        "radar_points": [torch.stack([b["radar_points"][r] for b in batch]) for r in range(n_radar)],
    }
    max_objs = max(len(b["gt_boxes"]) for b in batch)
    boxes, labels, vels = [], [], []
    for b in batch:
        pad = max_objs - len(b["gt_boxes"])
        boxes.append(torch.cat([b["gt_boxes"], torch.zeros(pad, 7)], dim=0) if pad else b["gt_boxes"])
        labels.append(torch.cat([b["gt_labels"], torch.full((pad,), -1, dtype=torch.long)], dim=0) if pad else b["gt_labels"])
        vels.append(torch.cat([b["gt_velocities"], torch.zeros(pad, 2)], dim=0) if pad else b["gt_velocities"])
    out.update(gt_boxes=torch.stack(boxes), gt_labels=torch.stack(labels), gt_velocities=torch.stack(vels),
               tokens=[b["token"] for b in batch])
    if all("intrinsics" in b and "lidar2cam" in b for b in batch):
        out["intrinsics"] = torch.stack([b["intrinsics"] for b in batch])
        out["lidar2cam"] = torch.stack([b["lidar2cam"] for b in batch])
    return out
