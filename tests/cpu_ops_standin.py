"""TEST INFRASTRUCTURE: torch-CPU stand-ins for the functions of `ops` the patched forwards call.

`patch()` puts `encoders.lidar_forward`, `fusion.fusion_forward`, `conv_blocks.head_forward` and the decode wrapper behind
the names of the REAL reference classes.  The kernels need a GPU and the reference cannot travel to the GPU box, so the
one place both exist is the build container — without a GPU.  `install()` swaps the kernel front end for plain torch
ops with the same signatures and return shapes, so that the whole patched forward (attribute names of the reference's
classes, weight folding and packing, reshapes, branch selection, the decode's list-of-dicts slicing) executes on the
reference's own `create_detector(...)` model and can be compared with the unpatched model.  What this does NOT test is
the kernels themselves — `-m gpu` does, through the mirror classes that share these forward functions.

Used by tests/test_patch_on_reference.py only.  Never imported by the package.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from oracle import bev_oracle as orc


def _unpack(params: torch.Tensor, dims: Sequence[int]):
    ws, bs, off = [], [], 0
    for c_in, c_out in zip(dims[:-1], dims[1:]):
        ws.append(params[off:off + c_in * c_out].view(c_in, c_out))      # W^T
        off += c_in * c_out
        bs.append(params[off:off + c_out])
        off += c_out
    assert off == params.numel()
    return ws, bs


def _mlp(points: torch.Tensor, params: torch.Tensor, dims) -> torch.Tensor:
    x = points
    for wt, b in zip(*_unpack(params, dims)):
        x = torch.relu(x @ wt + b)
    return x


def pointnet_encode(points, params, dims, perm=None, offsets=None, n_cells=0, precision=0, tc_params=None,
                    want_global=True, want_canvas=None):
    assert perm is None, "the stand-in covers the reference's global-max mode"
    return _mlp(points, params, dims).max(dim=1)[0]


def pack_mlp_params_split(params, dims):
    return None          # "layer widths not taken by the tensor-core path": the caller then runs pointnet_encode without an image


def radar_encode(radar_list, params, dims, fusion, fc_weight, fc_bias):
    per = torch.stack([_mlp(r, params, dims).max(dim=1)[0] for r in radar_list], dim=1)
    if fusion == "concat":
        out = F.linear(per.reshape(per.shape[0], -1), fc_weight, fc_bias)
    else:
        out = per.max(dim=1)[0] if fusion == "max" else per.mean(dim=1)
    return out, per


def camera_mean(feats):
    return feats.mean(dim=1)


def bilinear_resize(x, size):
    return F.interpolate(x, size=tuple(size), mode="bilinear", align_corners=False)


def dense_layer(x, weight, bias=None, relu=False):
    y = F.linear(x, weight, bias)
    return torch.relu(y) if relu else y


def lidar_init(feats, w1, b1, w2, b2, return_hidden=False):
    hid = torch.relu(F.linear(feats, w1, b1))
    out = F.linear(hid, w2, b2)
    return (out, hid) if return_hidden else out


def conv_pack_split(weight):
    return weight                                   # the "image" of the stand-in is the folded fp32 weight itself


def nchw_to_nhwc_split(parts):
    return (parts[0] if len(parts) == 1 else torch.cat(list(parts), dim=1)), None      # concat only; no layout, no scale


def conv_bn_relu_split(x, stat, image, bias, c_out, taps, relu=True):
    y = F.conv2d(x, image, bias, padding=image.shape[-1] // 2)
    return torch.relu(y) if relu else y


def border_expand(small, size, out_nhwc=None, c_offset=0, want_nchw=True):
    s = small.shape[-1]
    k = s // 2
    cls = lambda n: [i if i < k else (s - (n - i) if i >= n - k else k) for i in range(n)]
    return small[:, :, cls(size[0])][:, :, :, cls(size[1])]


def centernet_nms(heat):
    return torch.from_numpy(orc.nms(heat.numpy()))


def centernet_topk(scores, K):
    if K > scores.shape[2] * scores.shape[3]:
        raise RuntimeError("selected index k out of range")
    return tuple(torch.from_numpy(np.ascontiguousarray(a)) for a in orc.topk(scores.numpy(), K))


def centernet_decode(heatmap, offset, size, rot, vel, K, voxel, origin=(-51.2, -51.2), z_value=-1.0, score_thresh=0.0,
                     heat_is_logit=False):
    """Fixed-size outputs + count, filled from the numpy oracle's per-sample lists."""
    assert not heat_is_logit and z_value == -1.0
    B = heatmap.shape[0]
    dets = orc.decode({"heatmap": heatmap.numpy(), "offset": offset.numpy(), "size": size.numpy(), "rot": rot.numpy(),
                       "vel": vel.numpy()}, score_thresh=score_thresh, max_detections=K, voxel_size_m=voxel, pc_origin=origin)
    o = {"boxes": torch.zeros(B, K, 7), "scores": torch.zeros(B, K), "labels": torch.zeros(B, K, dtype=torch.int64),
         "velocities": torch.zeros(B, K, 2), "count": torch.zeros(B, dtype=torch.int32)}
    for b, d in enumerate(dets):
        n = len(d["scores"])
        o["count"][b] = n
        for k in ("boxes", "scores", "labels", "velocities"):
            o[k][b, :n] = torch.from_numpy(np.ascontiguousarray(d[k]))
    return o


STAND_INS = ("pointnet_encode", "pack_mlp_params_split", "radar_encode", "camera_mean", "bilinear_resize", "dense_layer", "lidar_init",
             "conv_pack_split", "nchw_to_nhwc_split", "conv_bn_relu_split", "border_expand", "centernet_nms", "centernet_topk", "centernet_decode")


CALLS = {name: 0 for name in STAND_INS}


def _counted(name, fn):
    def wrapper(*args, **kwargs):
        CALLS[name] += 1
        return fn(*args, **kwargs)
    return wrapper


def install(ops_module) -> List[str]:
    """Replaces the kernel front-end functions of `ops_module` in place (each call is counted in CALLS); returns the names."""
    g = globals()
    for name in STAND_INS:
        setattr(ops_module, name, _counted(name, g[name]))
    return list(STAND_INS)
