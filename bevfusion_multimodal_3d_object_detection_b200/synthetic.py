"""Seeded synthetic nuScenes-shaped inputs and weights — numpy only, so the same
bytes come out in the build container, on the GPU box and inside the golden-vector generator.

Every generator takes an integer seed and uses numpy's PCG64 stream; `digest()` gives a sha256 that
the golden fixtures store, so a changed RNG stream is detected instead of silently changing inputs.
"""
from __future__ import annotations

import hashlib
from typing import Dict, List, Sequence, Tuple

import numpy as np

PC_RANGE = (-51.2, -51.2, -5.0, 51.2, 51.2, 3.0)   # configs/base.yaml:48
LIDAR_DIMS = (4, 64, 128, 256, 512, 1024)           # src/encoders.py:252-256 with input_channels=4 (base.yaml:179)
RADAR_DIMS = (7, 32, 64, 128, 256)                  # src/encoders.py:515-518
BN_EPS = 1e-5                                       # torch.nn.BatchNorm1d default


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(int(seed)))


def digest(*arrays: np.ndarray) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


def lidar_points(seed: int, n_valid: int = 34720, n_total: int = 35000, channels: int = 4,
                 sigma: float = 20.0) -> np.ndarray:
    """(n_total, channels) f32: n_valid points strictly inside the range filter of
    src/train_detect.py:153-155, then zero rows (the padding of src/train_detect.py:181-189)."""
    g = _rng(seed)
    lo = np.array(PC_RANGE[:3], dtype=np.float32)
    hi = np.array(PC_RANGE[3:], dtype=np.float32)
    xy = np.empty((n_valid, 2), dtype=np.float32)
    todo = np.arange(n_valid)
    while todo.size:
        cand = (g.standard_normal((todo.size, 2)) * sigma).astype(np.float32)
        ok = np.all((cand > lo[:2]) & (cand < hi[:2]), axis=1)
        xy[todo[ok]] = cand[ok]
        todo = todo[~ok]
    z = g.uniform(-5.0, 3.0, n_valid).astype(np.float32)
    z = np.clip(z, np.nextafter(lo[2], np.float32(0)), np.nextafter(hi[2], np.float32(0)))
    pts = np.zeros((n_total, channels), dtype=np.float32)
    pts[:n_valid, 0:2] = xy
    pts[:n_valid, 2] = z
    if channels > 3:
        pts[:n_valid, 3] = g.uniform(0.0, 255.0, n_valid).astype(np.float32)
    for c in range(4, channels):
        pts[:n_valid, c] = g.integers(0, 32, n_valid).astype(np.float32)
    return pts


def raw_sweep(seed: int, n_rows: int, channels: int = 4) -> np.ndarray:
    """A raw sweep as it sits in a nuScenes .bin file before NuScenesDataset._load_lidar_points filters it
    (src/train_detect.py:151): about a fifth of the rows fall outside the range, a few sit exactly on a
    boundary (the filter is strict) and a few are NaN."""
    g = _rng(seed)
    p = np.empty((n_rows, channels), dtype=np.float32)
    p[:, 0] = (28.0 * g.standard_normal(n_rows)).astype(np.float32)
    p[:, 1] = (28.0 * g.standard_normal(n_rows)).astype(np.float32)
    p[:, 2] = g.uniform(-5.6, 3.6, n_rows).astype(np.float32)
    for c in range(3, channels):
        p[:, c] = g.uniform(0.0, 255.0, n_rows).astype(np.float32)
    edges = np.float32([-51.2, 51.2])
    p[3::997, 0] = edges[0]
    p[5::991, 0] = edges[1]
    p[7::983, 1] = edges[1]
    p[11::977, 2] = np.float32(-5.0)
    p[13::971, 2] = np.float32(3.0)
    p[17::1999, 1] = np.nan
    return p


def lidar_batch(seed: int, batch: int, **kw) -> np.ndarray:
    return np.stack([lidar_points(seed + i, **kw) for i in range(batch)])


def radar_batch(seed: int, batch: int, n_radars: int = 5, n_points: int = 125, channels: int = 7) -> List[np.ndarray]:
    """n_radars arrays (batch, n_points, channels) of N(0,1) — what the reference's dataset feeds
    (np.random.randn, src/train_detect.py:175)."""
    g = _rng(seed)
    return [g.standard_normal((batch, n_points, channels)).astype(np.float32) for _ in range(n_radars)]


def camera_features(seed: int, batch: int, n_cam: int = 6, channels: int = 512, h: int = 57, w: int = 100) -> np.ndarray:
    """relu(N(0,1)) feature maps (batch, n_cam, channels, h, w): 57x100 is ResNet-18 stride 16 on 900x1600."""
    g = _rng(seed)
    x = g.standard_normal((batch, n_cam, channels, h, w), dtype=np.float32)
    np.maximum(x, 0.0, out=x)
    return x


def mlp_weights(seed: int, dims: Sequence[int], use_bn: bool = True) -> List[Dict[str, np.ndarray]]:
    """Per layer: conv weight (C_out,C_in) / bias U(+-1/sqrt(C_in)) and perturbed BatchNorm statistics
    (default stats make BN an identity and would hide folding bugs — SURVEY §7 step 0)."""
    g = _rng(seed)
    layers = []
    for c_in, c_out in zip(dims[:-1], dims[1:]):
        bound = 1.0 / np.sqrt(c_in)
        layer = {
            "weight": g.uniform(-bound, bound, (c_out, c_in)).astype(np.float32),
            "bias": g.uniform(-bound, bound, c_out).astype(np.float32),
        }
        if use_bn:
            layer.update(
                bn_weight=g.uniform(0.5, 1.5, c_out).astype(np.float32),
                bn_bias=(g.standard_normal(c_out) * 0.2).astype(np.float32),
                bn_mean=(g.standard_normal(c_out) * 0.3).astype(np.float32),
                bn_var=g.uniform(0.5, 2.0, c_out).astype(np.float32),
            )
        layers.append(layer)
    return layers


def fold_mlp(layers: Sequence[Dict[str, np.ndarray]]) -> Tuple[List[np.ndarray], List[np.ndarray]]:
    """Eval-mode BatchNorm folded into the k=1 convolutions, in float64 (the numpy twin of
    ops.fold_batchnorm): W' = W*s, b' = (b - mean)*s + beta with s = gamma / sqrt(var + eps)."""
    ws, bs = [], []
    for lay in layers:
        w = lay["weight"].astype(np.float64)
        b = lay["bias"].astype(np.float64)
        if "bn_var" in lay:
            s = lay["bn_weight"].astype(np.float64) / np.sqrt(lay["bn_var"].astype(np.float64) + float(BN_EPS))
            w = w * s[:, None]
            b = (b - lay["bn_mean"].astype(np.float64)) * s + lay["bn_bias"].astype(np.float64)
        ws.append(w)
        bs.append(b)
    return ws, bs


def linear_weights(seed: int, c_in: int, c_out: int) -> Tuple[np.ndarray, np.ndarray]:
    g = _rng(seed)
    bound = 1.0 / np.sqrt(c_in)
    return (g.uniform(-bound, bound, (c_out, c_in)).astype(np.float32),
            g.uniform(-bound, bound, c_out).astype(np.float32))


def global_features(seed: int, batch: int, channels: int) -> np.ndarray:
    """Encoder outputs as the fusion module receives them: (B, C) f32, non-negative (they follow a ReLU and a max)."""
    return np.abs(_rng(seed).standard_normal((batch, channels))).astype(np.float32)


def head_weights(seed: int, in_channels: int, head_conv: int, classes: int, out_scale: float = 1.0) -> Dict[str, np.ndarray]:
    """CenterNetHead parameters (src/fusion.py:822-854) by state_dict name, with weights large enough that the heat-map
    logits spread (the reference's init, std 0.001, leaves every logit within 1e-3 of the -4.595 bias)."""
    g = _rng(seed)
    out = {}
    for name, n_out in (("heatmap", classes), ("offset", 2), ("size", 3), ("rot", 2), ("vel", 2)):
        out[f"{name}_head.0.weight"] = (g.standard_normal((head_conv, in_channels, 3, 3)) * 0.08).astype(np.float32)
        out[f"{name}_head.0.bias"] = (g.standard_normal(head_conv) * 0.1).astype(np.float32)
        out[f"{name}_head.2.weight"] = (g.standard_normal((n_out, head_conv, 1, 1)) * (0.3 * out_scale)).astype(np.float32)
        out[f"{name}_head.2.bias"] = (g.standard_normal(n_out) * 0.1).astype(np.float32)
    out["heatmap_head.2.bias"] = out["heatmap_head.2.bias"] - np.float32(2.0)
    return out


def fill_state_dict(seed: int, shapes: Dict[str, Tuple[int, ...]]) -> Dict[str, np.ndarray]:
    """Seeded values for every float entry of a state_dict, by name and shape: conv / linear weights N(0, 1/fan_in),
    BatchNorm weight U(0.5,1.5), running_var U(0.5,2), everything else 1-D N(0, 0.2^2) — statistics away from the
    identity, the same on every machine (torch's own initialisers depend on construction order)."""
    g = _rng(seed)
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        if name.endswith("num_batches_tracked"):
            continue
        if len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            out[name] = (g.standard_normal(shape) / np.sqrt(fan_in)).astype(np.float32)
        elif name.endswith("running_var"):
            out[name] = g.uniform(0.5, 2.0, shape).astype(np.float32)
        elif name.endswith("weight"):
            out[name] = g.uniform(0.5, 1.5, shape).astype(np.float32)
        else:
            out[name] = (g.standard_normal(shape) * 0.2).astype(np.float32)
    return out


DETECTOR_PREFIXES = ("lidar_encoder.", "radar_encoder.", "fusion.", "det_head.")


def detector_state(seed: int, shapes: Dict[str, Tuple[int, ...]], head_in: int = 256, head_conv: int = 64, classes: int = 10):
    """Seeded parameters of the whole inference chain by the reference's state_dict names (`lidar_encoder.*`,
    `radar_encoder.*`, `fusion.*`, `det_head.*`; `shapes` = name -> shape, e.g. from the modules' own state_dict()):
    fill_state_dict for the encoders and the fusion module, head_weights (logits that spread) for `det_head.*`."""
    body = {k: v for k, v in shapes.items() if not k.startswith("det_head.")}
    out = fill_state_dict(seed, body)
    if any(k.startswith("det_head.") for k in shapes):
        for k, v in head_weights(seed + 1, head_in, head_conv, classes, out_scale=0.3).items():
            out["det_head." + k] = v
    return out


def head_maps(seed: int, batch: int, classes: int = 10, H: int = 50, W: int = 50, peak_frac: float = 1.0):
    """CenterNet head outputs with pairwise-distinct heat-map values (tie-free top-K):
    heatmap is a random permutation of an evenly spaced grid in (0,1) per sample, so no two cells of a
    sample are equal.  `peak_frac` < 1 scales all but the top `peak_frac` share of cells by 1e-3 to imitate a sparse map.
    Returns dict(heatmap (B,C,H,W), offset (B,2,H,W), size (B,3,H,W), rot (B,2,H,W), vel (B,2,H,W))."""
    g = _rng(seed)
    n = classes * H * W
    heat = np.empty((batch, n), dtype=np.float32)
    base = ((np.arange(n, dtype=np.float64) + 1.0) / (n + 1.0)).astype(np.float32)
    for b in range(batch):
        heat[b] = base[g.permutation(n)]
    if peak_frac < 1.0:
        heat = np.where(heat > np.float32(1.0 - peak_frac), heat, heat * np.float32(1e-3))
    return {
        "heatmap": heat.reshape(batch, classes, H, W),
        "offset": g.uniform(0.0, 1.0, (batch, 2, H, W)).astype(np.float32),
        "size": g.uniform(0.5, 5.0, (batch, 3, H, W)).astype(np.float32),
        "rot": g.standard_normal((batch, 2, H, W)).astype(np.float32),
        "vel": g.standard_normal((batch, 2, H, W)).astype(np.float32),
    }


def ground_truth_near(seed: int, detections: Sequence[Dict[str, np.ndarray]], n_from_pred: int = 12, n_random: int = 6):
    """Ground-truth dicts for a metrics check: the first `n_from_pred` detected boxes of each sample, displaced by up to
    ~1.5 m and resized a little, plus random boxes; mostly class 0 (the only label the reference's decode emits,
    SURVEY Q1), a few other classes and two invalid (-1) rows."""
    g = _rng(seed)
    out = []
    for d in detections:
        # rounded to the millimetre: the ground truth must not depend on the last bits of whoever decoded (atan2 differs by
        # an ulp between numpy, torch-CPU and the device)
        base = np.round(np.asarray(d["boxes"], dtype=np.float64)[:n_from_pred], 3).astype(np.float32)
        base[:, :2] += g.uniform(-1.1, 1.1, (len(base), 2)).astype(np.float32)
        base[:, 3:6] *= g.uniform(0.8, 1.25, (len(base), 3)).astype(np.float32)
        base[:, 6] += g.uniform(-0.4, 0.4, len(base)).astype(np.float32)
        rnd = np.concatenate([g.uniform(-50, 50, (n_random, 2)), np.full((n_random, 1), -1.0), g.uniform(0.5, 5, (n_random, 3)),
                              g.uniform(-3.1, 3.1, (n_random, 1))], axis=1).astype(np.float32)
        boxes = np.concatenate([base, rnd])
        labels = np.zeros(len(boxes), dtype=np.int64)
        labels[-n_random:-n_random + 2] = (3, 5)
        labels[-2:] = -1
        out.append({"boxes": boxes, "labels": labels})
    return out


def camera_rig(img_w: float = 1600.0, img_h: float = 900.0) -> Tuple[np.ndarray, np.ndarray]:
    """Six pinhole cameras, nuScenes-like: intrinsics (6,3,3) and ego->camera [R|t] (6,3,4), f32.
    Ego frame: x forward, y left, z up (the lidar-frame convention of src/data_converter.py:237-247);
    camera frame: x right, y down, z forward.  Yaws follow the nuScenes ring
    (FRONT, FRONT_RIGHT, FRONT_LEFT, BACK, BACK_LEFT, BACK_RIGHT)."""
    K = np.array([[1266.0, 0.0, img_w / 2.0 + 16.0], [0.0, 1266.0, img_h / 2.0 + 41.0], [0.0, 0.0, 1.0]])
    yaws = np.deg2rad([0.0, -55.0, 55.0, 180.0, 110.0, -110.0])
    pos = np.array([[1.5, 0.0, 1.5], [1.5, -0.5, 1.5], [1.5, 0.5, 1.5], [0.0, 0.0, 1.5], [1.0, 0.5, 1.5], [1.0, -0.5, 1.5]])
    # camera axes expressed in a yaw-0 ego frame: right = -y, down = -z, forward = +x
    base = np.array([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]])
    Ks, Es = [], []
    for yaw, p in zip(yaws, pos):
        c, s = np.cos(yaw), np.sin(yaw)
        rot_z = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])  # camera heading in ego frame
        R = base @ rot_z.T                                               # ego -> camera
        t = -R @ p
        Ks.append(K)
        Es.append(np.concatenate([R, t[:, None]], axis=1))
    return np.stack(Ks).astype(np.float32), np.stack(Es).astype(np.float32)


# The whole-chain case of tests/golden/detector_chain.npz (made by tests/golden/make_golden.py: chain_cases)
CHAIN_SEED, CHAIN_FRAMES, CHAIN_POINTS = 901, 2, 3000


def chain_inputs(seed: int = CHAIN_SEED, frames: int = CHAIN_FRAMES, points: int = CHAIN_POINTS, feat_hw=(28, 50)):
    """(lidar (F,N,4), 5 radars (F,125,7), camera features (F,6,512,h,w)) of the whole-chain case."""
    lidar = lidar_batch(seed + 10, frames, n_valid=points - 40, n_total=points)
    radars = radar_batch(seed + 11, frames)
    cam = camera_features(seed + 12, frames, n_cam=6, channels=512, h=feat_hw[0], w=feat_hw[1])
    return lidar, radars, cam
