"""Host side of SURVEY 8f N1: the Conv2d + BatchNorm2d + ReLU stacks of the reference's fusion module and detection
head on the tcgen05 convolution kernels.

Three modes, chosen by `module.b200_precision` (else `B200BEV_PRECISION`, else "f32"):

    "f32"        fp32 accuracy on the tensor cores (`b200bev_conv_bn_relu_split`: three fp16 products per fp32 product, exact
                 power-of-two operand scales; parity 1e-5 of max|ref|) — the default
    "bf16"       bf16 operands (`b200bev_conv_bn_relu_bf16`, parity 1e-2, north_star's bf16 bound); activations stay
                 channels-last bf16 between convolutions
    "f32_cudnn"  the reference's own torch layers (cuDNN; fp32, or TF32 if torch's flags allow it) between the kernels

Eval mode only — BatchNorm is folded with its running statistics.  Folded and packed weights are a cache beside the
module, rebuilt when the parameters change (weight_cache.py).  Stacks with shapes the kernels do not take (input channels
not a multiple of 64, other kernel sizes) run on the torch layers.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from .encoders import _state_key, default_precision
from .weight_cache import mark_dirty, wants_autograd

HEADS = ("heatmap", "offset", "size", "rot", "vel")       # CenterNetHead's sub-modules, src/fusion.py:822-854


def conv_mode(module: nn.Module) -> str:
    """"bf16" | "split" (fp32 accuracy on tcgen05) | "torch" (the module's own layers)."""
    name = getattr(module, "b200_precision", None) or default_precision()
    if name == "bf16":
        return "bf16"
    return "torch" if name in ("f32_cudnn", "fp32_cudnn", "torch") else "split"


def wants_bf16(module: nn.Module) -> bool:
    return conv_mode(module) == "bf16"


def _conv_ok(conv: nn.Conv2d) -> bool:
    k = conv.kernel_size
    return (k in ((3, 3), (1, 1)) and conv.stride == (1, 1) and conv.dilation == (1, 1) and conv.groups == 1
            and conv.padding == (k[0] // 2, k[1] // 2) and conv.in_channels % 64 == 0 and conv.padding_mode == "zeros")


def supported(seq: nn.Sequential) -> bool:
    """True when every layer of the stack is a Conv2d the kernel takes, a BatchNorm2d, a ReLU or a bilinear Upsample."""
    for layer in seq:
        if isinstance(layer, nn.Conv2d):
            if not _conv_ok(layer):
                return False
        elif isinstance(layer, nn.Upsample):
            if layer.mode != "bilinear" or layer.align_corners:
                return False
        elif not isinstance(layer, (nn.BatchNorm2d, nn.ReLU)):
            return False
    return True


def _plan(seq: nn.Sequential, device: torch.device, mode: str = "bf16") -> List[Dict]:
    key = (_state_key(seq, device), mode)
    caches = seq.__dict__.setdefault("_b200bev_conv_cache", {})
    cache = caches.get(mode)
    if cache is not None and cache["key"] == key:
        return cache["steps"]
    pack = ops.conv_pack if mode == "bf16" else ops.conv_pack_split
    layers = list(seq)
    steps: List[Dict] = []
    i = 0
    while i < len(layers):
        layer = layers[i]
        if isinstance(layer, nn.Conv2d):
            bn = layers[i + 1] if i + 1 < len(layers) and isinstance(layers[i + 1], nn.BatchNorm2d) else None
            j = i + (2 if bn is not None else 1)
            relu = j < len(layers) and isinstance(layers[j], nn.ReLU)
            w, b = ops.fold_conv_bn(layer, bn)
            steps.append({"kind": "conv", "image": pack(w.to(device)), "bias": b.to(device),
                          "c_out": layer.out_channels, "taps": layer.kernel_size[0] * layer.kernel_size[1], "relu": relu})
            i = j + (1 if relu else 0)
        elif isinstance(layer, nn.Upsample):
            steps.append({"kind": "upsample", "scale": layer.scale_factor})
            i += 1
        else:
            raise RuntimeError(f"conv_blocks: unexpected layer {type(layer).__name__} (check supported() first)")
    caches[mode] = {"key": key, "steps": steps}
    return steps


def run_split(seq: nn.Sequential, parts: Sequence[torch.Tensor]) -> torch.Tensor:
    """`seq(torch.cat(parts, dim=1))` at fp32 accuracy on the tensor cores: per convolution one pass that finds max|x|,
    one that lays the input out channels-last as scaled fp16 hi / lo halves (the concat happens there), and the split
    convolution itself, which writes the fp32 NCHW tensor the next layer — or the reference's next module — reads."""
    parts = list(parts)
    steps = _plan(seq, parts[0].device, "split")
    for step in steps:
        if step["kind"] == "conv":
            x_split, stat = ops.nchw_to_nhwc_split(parts)
            parts = [ops.conv_bn_relu_split(x_split, stat, step["image"], step["bias"], step["c_out"], step["taps"], step["relu"])]
        else:
            x = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
            s = step["scale"]
            sy, sx = (s, s) if not isinstance(s, (tuple, list)) else s
            parts = [ops.bilinear_resize(x, (int(x.shape[2] * sy), int(x.shape[3] * sx)))]
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)


def run(seq: nn.Sequential, parts: Optional[Sequence[torch.Tensor]] = None, nhwc: Optional[torch.Tensor] = None,
        out_nhwc: Optional[torch.Tensor] = None, c_offset: int = 0, note_nhwc: bool = False) -> Optional[torch.Tensor]:
    """`seq(torch.cat(parts, dim=1))` for a conv/BN/ReLU(/Upsample) stack: (B,C_i,H,W) fp32 parts -> (B,C_out,H',W') fp32.
    Between two convolutions the activations stay channels-last bf16 (written by the first launch's epilogue).
    nhwc: the input already as a (B,H,W,C) bf16 tensor (instead of `parts`).  out_nhwc / c_offset: the stack's last
    convolution writes channels [c_offset, ...) of this channels-last bf16 tensor — the concatenated input of the next
    stack — and nothing else; returns None.  note_nhwc: the last convolution writes the fp32 NCHW result AND its
    channels-last bf16 form in the same launch, and the returned tensor remembers the latter (`nhwc_of`): a following
    `CenterNetHead` on the bf16 path starts from it instead of running a layout pass."""
    parts = list(parts) if parts is not None else []
    device = nhwc.device if nhwc is not None else parts[0].device
    steps = _plan(seq, device, "bf16")
    if out_nhwc is not None and steps[-1]["kind"] != "conv":
        raise ValueError("out_nhwc needs a stack that ends in a convolution block")
    for k, step in enumerate(steps):
        if step["kind"] == "conv":
            if nhwc is None:
                nhwc = ops.nchw_to_nhwc_bf16(parts)      # layout + cast + concat in one pass per part
            last = k + 1 == len(steps)
            # the next consumer reads channels-last bf16: a convolution, or an upsample that a convolution follows
            chained = not last and (steps[k + 1]["kind"] == "conv" or (k + 2 < len(steps) and steps[k + 2]["kind"] == "conv"))
            if chained or (last and out_nhwc is not None):
                B, H, W, _ = nhwc.shape
                nxt, off = (out_nhwc, c_offset) if last else (torch.empty((B, H, W, step["c_out"]), dtype=torch.bfloat16, device=device), 0)
                ops.conv_bn_relu_bf16(nhwc, step["image"], step["bias"], step["c_out"], step["taps"], step["relu"],
                                      out_nhwc=nxt, c_offset=off, want_nchw=False)
                nhwc, parts = nxt, []
            elif last and note_nhwc:
                B, H, W, _ = nhwc.shape
                twin = torch.empty((B, H, W, step["c_out"]), dtype=torch.bfloat16, device=device)
                res = ops.conv_bn_relu_bf16(nhwc, step["image"], step["bias"], step["c_out"], step["taps"], step["relu"], out_nhwc=twin)
                parts, nhwc = [attach_nhwc(res, twin)], None
            else:
                parts = [ops.conv_bn_relu_bf16(nhwc, step["image"], step["bias"], step["c_out"], step["taps"], step["relu"])]
                nhwc = None
        else:
            s = step["scale"]
            sy, sx = (s, s) if not isinstance(s, (tuple, list)) else s
            if nhwc is not None and k + 1 < len(steps):      # between two convolutions: stay channels-last bf16
                nhwc = ops.bilinear_resize_nhwc_bf16(nhwc, (int(nhwc.shape[1] * sy), int(nhwc.shape[2] * sx)))
                continue
            x = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
            parts = [ops.bilinear_resize(x, (int(x.shape[2] * sy), int(x.shape[3] * sx)))]
    if out_nhwc is not None:
        return None
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)


# ------------------------------------------------------------------------------------------------
# CenterNetHead: five 3x3 convs as one, five 1x1 convs as one block-diagonal conv, sigmoid left to the decode kernel
# ------------------------------------------------------------------------------------------------
def head_supported(head: nn.Module) -> bool:
    try:
        subs = [getattr(head, f"{n}_head") for n in HEADS]
    except AttributeError:
        return False
    hidden = sum(s[0].out_channels for s in subs)
    return all(len(s) == 3 and _conv_ok(s[0]) and s[2].kernel_size == (1, 1) for s in subs) and hidden % 64 == 0


def _head_plan(head: nn.Module, device: torch.device, mode: str = "bf16") -> Dict:
    key = (_state_key(head, device), mode)
    caches = head.__dict__.setdefault("_b200bev_conv_cache", {})
    cache = caches.get(mode)
    if cache is not None and cache["key"] == key:
        return cache
    pack = ops.conv_pack if mode == "bf16" else ops.conv_pack_split
    subs = [getattr(head, f"{n}_head") for n in HEADS]
    w1 = torch.cat([s[0].weight.detach() for s in subs], dim=0).float().to(device).contiguous()   # (5*hc, Cin, 3, 3)
    b1 = torch.cat([s[0].bias.detach() for s in subs], dim=0).float().to(device).contiguous()
    hidden = int(w1.shape[0])
    outs = [s[2].out_channels for s in subs]
    w2 = torch.zeros((sum(outs), hidden, 1, 1), dtype=torch.float32, device=device)
    r = c = 0
    for s, n in zip(subs, outs):
        hc = s[0].out_channels
        w2[r:r + n, c:c + hc] = s[2].weight.detach().float().to(device)
        r += n
        c += hc
    b2 = torch.cat([s[2].bias.detach() for s in subs], dim=0).float().to(device).contiguous()
    cache = {"key": key, "img1": pack(w1), "b1": b1, "hidden": hidden, "img2": pack(w2), "b2": b2, "outs": outs}
    caches[mode] = cache
    return cache


def head_forward(head: nn.Module, x: torch.Tensor) -> Dict[str, torch.Tensor]:
    """CenterNetHead.forward (src/fusion.py:869-884).  Eval mode with the bf16 path enabled: two tcgen05 launches for the
    ten convolutions; the returned heat map remembers the raw output it is the sigmoid of (`attach_logits`), which
    `decode_centernet_predictions` feeds to the decode kernel so that the sigmoid is not a separate pass there."""
    if head.training:
        mark_dirty(head)
    mode = conv_mode(head)
    if head.training or not x.is_cuda or mode == "torch" or not head_supported(head) or x.shape[1] % 64 != 0 \
            or wants_autograd(head, x):
        pred = {n: getattr(head, f"{n}_head")(x) for n in HEADS}
        pred["heatmap"] = torch.sigmoid(pred["heatmap"])
        return pred
    p = _head_plan(head, x.device, mode)
    B, _, H, W = x.shape
    if mode == "bf16":
        hid = torch.empty((B, H, W, p["hidden"]), dtype=torch.bfloat16, device=x.device)      # stays channels-last bf16
        x_nhwc = nhwc_of(x)      # left by the fusion module's last convolution, if `x` is still its output
        if x_nhwc is None:
            x_nhwc = ops.nchw_to_nhwc_bf16([x])
        ops.conv_bn_relu_bf16(x_nhwc, p["img1"], p["b1"], p["hidden"], 9, relu=True, out_nhwc=hid, want_nchw=False)
        both = ops.conv_bn_relu_bf16(hid, p["img2"], p["b2"], sum(p["outs"]), 1, relu=False)
    else:       # fp32 accuracy: the same two launches on split fp16 operands, fp32 tensors in between
        x_split, stat = ops.nchw_to_nhwc_split([x])
        hid = ops.conv_bn_relu_split(x_split, stat, p["img1"], p["b1"], p["hidden"], 9, relu=True)
        h_split, h_stat = ops.nchw_to_nhwc_split([hid])
        both = ops.conv_bn_relu_split(h_split, h_stat, p["img2"], p["b2"], sum(p["outs"]), 1, relu=False)
    # the five heads as contiguous tensors: one fused copy launch (split_with_sizes_copy) instead of five
    parts = torch.split_with_sizes_copy(both, list(p["outs"]), dim=1)
    pred = dict(zip(HEADS, parts))
    logits = pred["heatmap"]
    pred["heatmap"] = attach_logits(torch.sigmoid(logits), logits)
    return pred


_LOGITS_ATTR = "_b200bev_logits"
_NHWC_ATTR = "_b200bev_nhwc"


def attach_nhwc(x: torch.Tensor, twin: torch.Tensor) -> torch.Tensor:
    """Remembers, ON an fp32 NCHW activation, its channels-last bf16 form written by the same convolution launch (the same
    note mechanism as `attach_logits`: an edited or replaced tensor carries no valid note)."""
    try:
        setattr(x, _NHWC_ATTR, (twin, x._version))
    except RuntimeError:
        pass
    return x


def nhwc_of(x: torch.Tensor) -> Optional[torch.Tensor]:
    """The channels-last bf16 form of `x` if `x` is still, bit for bit, the tensor the convolution wrote; else None."""
    note = getattr(x, _NHWC_ATTR, None)
    if note is None:
        return None
    twin, version = note
    try:
        same = x._version == version
    except RuntimeError:
        same = False
    B, Cc, H, W = x.shape
    return twin if same and tuple(twin.shape) == (B, H, W, Cc) and twin.device == x.device else None


def attach_logits(heatmap: torch.Tensor, logits: torch.Tensor) -> torch.Tensor:
    """Remembers, ON the heat-map tensor the head returns, the raw output it is the sigmoid of.  The dict keeps exactly
    the reference's five keys (src/fusion.py:877-883); a caller that replaces or edits `predictions['heatmap']`
    (flip-TTA averaging, masking, temperature scaling) hands the decode a tensor without the note, or with a newer
    version counter, and the decode then uses that tensor as it is — as the reference would."""
    try:
        setattr(heatmap, _LOGITS_ATTR, (logits, heatmap._version))
    except RuntimeError:       # inference tensors carry no version counter: no shortcut for them
        pass
    return heatmap


def logits_of(heatmap: torch.Tensor) -> Optional[torch.Tensor]:
    """The head's raw heat-map output if `heatmap` is still, bit for bit, the sigmoid the head computed from it; else None."""
    note = getattr(heatmap, _LOGITS_ATTR, None)
    if note is None:
        return None
    logits, version = note
    try:
        same = heatmap._version == version
    except RuntimeError:
        same = False
    return logits if same and logits.shape == heatmap.shape and logits.device == heatmap.device else None
