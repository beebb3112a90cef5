"""Validity of the folded / packed weight caches the eval-mode kernels read.

The kernels do not read a module's `nn.Parameter`s: BatchNorm is folded, weights are re-tiled into tensor-core stage
images, head convolutions are concatenated.  Those images are a cache beside the module (SURVEY 8b "Ownership / state")
and must be rebuilt whenever a parameter or a BatchNorm statistic changes.  A change is seen through

  * a `load_state_dict` post-hook (covers `assign=True`, which replaces the parameter objects),
  * a forward in training mode (an optimizer step usually follows): the next eval forward rebuilds,
  * the tensors' version counters, read on every eval forward — one attribute read per tensor over a list collected
    once, instead of walking `module.parameters()` / `module.buffers()` per call; in-place ops (`copy_`, `add_`,
    optimizer steps, EMA swaps written with in-place ops) bump them,
  * `invalidate_cache(module)`, the explicit call for what none of the above can see: writes through `.data`
    (`p.data.copy_()`, `p.data.normal_()` do NOT bump the version counter), a parameter object replaced by hand
    (`m.conv1.weight = nn.Parameter(...)`), `module.half()`.

Tensors created under `torch.inference_mode()` have no version counter; for them only the hooks and the explicit call
apply (documented in INTEGRATION.md).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn as nn

_STATE = "_b200bev_state"
_CACHES = ("_b200bev_cache", "_b200bev_conv_cache")


def _collect(module: nn.Module) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
    params = list(module.parameters())
    return params + list(module.buffers()), params


def _state(module: nn.Module) -> Dict:
    st = module.__dict__.get(_STATE)
    if st is None:
        st = {"epoch": 0, "dirty": True, "tensors": [], "params": []}
        module.__dict__[_STATE] = st
        # after load_state_dict (strict or not, assign or copy) every cache derived from the old values is stale
        module.register_load_state_dict_post_hook(lambda m, incompatible: mark_dirty(m))
    return st


def mark_dirty(module: nn.Module) -> None:
    """Cheap flag: the next eval-mode forward re-collects the tensors and rebuilds the packed weights."""
    _state(module)["dirty"] = True


def invalidate_cache(module: nn.Module) -> None:
    """Drops every packed-weight cache of `module` and of its sub-modules.  Call it after changing parameters or BatchNorm
    statistics in a way autograd's version counters do not see: `p.data.<op>_()`, replacing a parameter object,
    `module.half()`, tensors made under `torch.inference_mode()`."""
    for m in module.modules():
        for name in _CACHES:
            m.__dict__.pop(name, None)
        if _STATE in m.__dict__:
            m.__dict__[_STATE]["dirty"] = True


def _versions(tensors: List[torch.Tensor]):
    try:
        return tuple(t._version for t in tensors)
    except RuntimeError:          # "Inference tensors do not track version counter"
        return None


def state_token(module: nn.Module, device: torch.device):
    """A hashable value that changes whenever the packed weights of `module` must be rebuilt for `device`."""
    st = _state(module)
    if st["dirty"]:
        st["tensors"], st["params"] = _collect(module)
        st["epoch"] += 1
        st["dirty"] = False
    return (str(device), st["epoch"], _versions(st["tensors"]))


def wants_autograd(module: nn.Module, *inputs) -> bool:
    """True when an eval-mode forward must stay differentiable: autograd is recording and an input or a parameter
    requires grad (frozen-backbone fine-tuning, saliency, adversarial gradients).  The ctypes kernels return tensors
    without a grad_fn, so such a call takes the module's plain torch graph on the same device — what the reference's
    eval-mode modules do.  Under `torch.no_grad()` / `inference_mode()` (every pipeline of the reference:
    `@torch.no_grad()` at src/eval.py:27, src/train_detect.py:500,820, src/inference.py:128) this is False at the cost of one flag read."""
    if not torch.is_grad_enabled():
        return False
    for t in inputs:
        if isinstance(t, torch.Tensor):
            if t.requires_grad:
                return True
        elif isinstance(t, (list, tuple)):
            if any(isinstance(u, torch.Tensor) and u.requires_grad for u in t):
                return True
    st = _state(module)
    if st["dirty"]:
        state_token(module, torch.device("cpu"))
    return any(p.requires_grad for p in st["params"])
