"""Host-side runtime around the kernels: frame sharding across GPUs and the double-buffered
host->device->host frame pipeline.

Frames are independent (eval-mode BatchNorm has no cross-sample term), so multi-GPU is a
contiguous split of the frame batch with no data-path collective: one process per GPU, weights
replicated once.  torch.distributed is used for the launch barrier, for the max-over-ranks step
time and, if a caller wants one list, for gathering the per-frame detection counts.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [begin, end) of rank `rank`; the first n_frames % world ranks get one more."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_frames, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """MAX-reduce a scalar over the process group (identity without one)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_counts(local_counts: Sequence[int], device: Optional[torch.device] = None) -> List[int]:
    """All ranks' per-frame detection counts in global frame order (shards may differ by one frame)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(local_counts)
    world = dist.get_world_size()
    dev = device if device is not None else "cpu"
    n = torch.tensor([len(local_counts)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    width = int(max(s.item() for s in sizes))
    padded = torch.full((width,), -1, dtype=torch.int64, device=dev)
    padded[: len(local_counts)] = torch.tensor(list(local_counts), dtype=torch.int64, device=dev)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    out: List[int] = []
    for p, s in zip(parts, sizes):
        out.extend(p[: int(s.item())].tolist())
    return out


@dataclass
class _Slot:
    dev: Dict[str, torch.Tensor]
    ready: torch.cuda.Event     # H2D of this slot finished
    free: torch.cuda.Event      # kernels that read this slot finished


class FramePipeline:
    """Streams host-resident frames through a per-chunk step function.

    `host_inputs` maps names to PINNED host tensors whose first dimension is the frame index.  Chunks
    of `chunk` frames are copied on a copy stream into one of two device slots while the previous
    chunk's kernels run on the compute stream; `step(dev_inputs, frame_begin, frame_end)` is called
    on the compute stream and must enqueue its own device->host copies of results.
    """

    def __init__(self, host_inputs: Dict[str, torch.Tensor], chunk: int, device: torch.device):
        for k, t in host_inputs.items():
            if t.is_cuda or not t.is_pinned():
                raise ValueError(f"host input {k!r} must be a pinned CPU tensor")
        self.host = host_inputs
        self.n_frames = next(iter(host_inputs.values())).shape[0]
        self.chunk = max(1, min(int(chunk), self.n_frames))
        self.device = device
        self.copy_stream = torch.cuda.Stream(device)
        self.slots = [
            _Slot({k: torch.empty((self.chunk, *t.shape[1:]), dtype=t.dtype, device=device) for k, t in host_inputs.items()},
                  torch.cuda.Event(), torch.cuda.Event())
            for _ in range(2)
        ]
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in host_inputs.values())

    def run(self, step: Callable[[Dict[str, torch.Tensor], int, int], None]) -> None:
        compute = torch.cuda.current_stream(self.device)
        bounds = [(b, min(b + self.chunk, self.n_frames)) for b in range(0, self.n_frames, self.chunk)]

        def upload(i: int) -> None:
            b, e = bounds[i]
            slot = self.slots[i & 1]
            with torch.cuda.stream(self.copy_stream):
                if i >= 2:
                    self.copy_stream.wait_event(slot.free)
                for k, t in self.host.items():
                    slot.dev[k][: e - b].copy_(t[b:e], non_blocking=True)
                slot.ready.record(self.copy_stream)

        self.copy_stream.wait_stream(compute)
        upload(0)
        for i, (b, e) in enumerate(bounds):
            if i + 1 < len(bounds):
                upload(i + 1)
            slot = self.slots[i & 1]
            compute.wait_event(slot.ready)
            step({k: v[: e - b] for k, v in slot.dev.items()}, b, e)
            slot.free.record(compute)
        compute.wait_stream(self.copy_stream)


class GraphedStep:
    """A fixed-shape device step captured once in a CUDA graph and replayed.

    `fn()` must read its inputs from tensors that stay allocated (copy new data INTO them between replays) and must
    not synchronise with the host: every kernel of this package qualifies (they only enqueue on the current stream; the
    list-of-dicts form of `decode_centernet_predictions` does not — it reads the counts back — so a step ends with
    `ops.centernet_decode`, whose outputs have fixed shapes).  Weight caches (folded / packed parameters) are built by the
    warm-up calls, outside the capture.  The tensors `fn` returned during capture are the step's outputs: a replay
    overwrites them in place.
    """

    def __init__(self, fn: Callable[[], object], device: torch.device, warmup: int = 2):
        self.device = device
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):
                fn()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.outputs = fn()

    def replay(self):
        """Enqueues the whole step on the current stream; returns the (static) output tensors."""
        self.graph.replay()
        return self.outputs
