"""Host-side runtime around the kernels: frame sharding across GPUs and the double-buffered
host->device->host frame pipeline.

Frames are independent (eval-mode BatchNorm has no cross-sample term), so multi-GPU is a
contiguous split of the frame batch with no data-path collective: one process per GPU, weights
replicated once.  torch.distributed is used for the launch barrier, for the max-over-ranks step
time and, if a caller wants one list, for gathering the per-frame detection counts.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from pathlib import Path
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


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int) -> Optional[int]:
    """NUMA node of a CUDA device, from its PCI address in sysfs (None when the platform does not say)."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int((Path("/sys/bus/pci/devices") / bus / "numa_node").read_text().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa(local_rank: int, local_world: int = 1) -> Dict[str, object]:
    """Pins this process (and so its pinned-memory allocations, by first touch) to CPUs next to its GPU.

    One process per GPU feeds its device from pinned host memory; on a multi-socket host a rank whose buffers sit on
    the other socket pays the inter-socket link on every copy, and ranks left on "all CPUs" migrate and share cores.
    The CPUs of the GPU's NUMA node (all CPUs when sysfs gives no node) are divided evenly among the ranks that
    share that node.  Call before allocating pinned buffers.  Returns what was done, for the bench record."""
    info: Dict[str, object] = {"numa_node": None, "cpus": None, "bound": False}
    try:
        node = gpu_numa_node(local_rank)
        info["numa_node"] = node
        all_cpus = sorted(os.sched_getaffinity(0))
        cpus = all_cpus
        sharers, my_pos = local_world, local_rank
        if node is not None:
            node_cpus = set(_parse_cpulist((Path("/sys/devices/system/node") / f"node{node}" / "cpulist").read_text()))
            cpus = [c for c in all_cpus if c in node_cpus] or all_cpus
            same = [r for r in range(local_world) if gpu_numa_node(r) == node]
            sharers, my_pos = max(1, len(same)), (same.index(local_rank) if local_rank in same else 0)
        per = max(1, len(cpus) // sharers)
        mine = cpus[my_pos * per:(my_pos + 1) * per] or cpus
        os.sched_setaffinity(0, mine)
        torch.set_num_threads(max(1, min(len(mine), 4)))
        info.update(cpus=f"{mine[0]}-{mine[-1]} ({len(mine)} of {len(all_cpus)})", bound=True, ranks_on_node=sharers)
    except Exception as e:      # affinity is an optimisation, never a reason to fail
        info["error"] = str(e)[:120]
    return info


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
        self.graphs: List[GraphedStep] = []

    def _bounds(self) -> List[Tuple[int, int]]:
        return [(b, min(b + self.chunk, self.n_frames)) for b in range(0, self.n_frames, self.chunk)]

    def _upload(self, i: int, bounds) -> None:
        b, e = bounds[i]
        slot = self.slots[i & 1]
        with torch.cuda.stream(self.copy_stream):
            if i >= 2:
                self.copy_stream.wait_event(slot.free)
            for k, t in self.host.items():
                slot.dev[k][: e - b].copy_(t[b:e], non_blocking=True)
            slot.ready.record(self.copy_stream)

    def capture(self, step: Callable[[Dict[str, torch.Tensor], int, int], object]) -> None:
        """Captures `step(slot_inputs, 0, chunk)` once per slot in a CUDA graph (the slot tensors keep their addresses, so
        a replay computes on whatever the last upload put there).  Needs n_frames % chunk == 0.  `step` must not touch
        frame-indexed tensors outside the slot (every chunk replays the same graph)."""
        if self.n_frames % self.chunk:
            raise ValueError("capture needs whole chunks")
        self.graphs = [GraphedStep(lambda s=slot: step(s.dev, 0, self.chunk), self.device) for slot in self.slots]

    def run_captured(self, after: Callable[[object, int, int], None]) -> None:
        """One pass over all frames: upload chunk i+1 while the graph of chunk i runs; `after(outputs, begin, end)` is
        called on the compute stream after each replay (enqueue the device->host copies of the results there)."""
        compute = torch.cuda.current_stream(self.device)
        bounds = self._bounds()
        self.copy_stream.wait_stream(compute)
        self._upload(0, bounds)
        for i, (b, e) in enumerate(bounds):
            if i + 1 < len(bounds):
                self._upload(i + 1, bounds)
            slot = self.slots[i & 1]
            compute.wait_event(slot.ready)
            after(self.graphs[i & 1].replay(), b, e)
            slot.free.record(compute)
        compute.wait_stream(self.copy_stream)

    def h2d_only_ms(self, reps: int = 3) -> float:
        """Milliseconds of one pass of the uploads alone (no kernels): what the host side allows this pipeline."""
        bounds = self._bounds()
        best = float("inf")
        for _ in range(max(1, reps)):
            torch.cuda.synchronize(self.device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(self.copy_stream):
                e0.record(self.copy_stream)
                for i, (b, e) in enumerate(bounds):
                    for k, t in self.host.items():
                        self.slots[i & 1].dev[k][: e - b].copy_(t[b:e], non_blocking=True)
                e1.record(self.copy_stream)
            torch.cuda.synchronize(self.device)
            best = min(best, e0.elapsed_time(e1))
        return best

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


class BranchStreams:
    """Side streams for work that is independent of what the caller's stream is doing (the modality branches of the fusion
    module, the radar encoder next to the LiDAR encoder): `fork(i)` returns a context in which kernels are enqueued on side
    stream i, ordered after everything enqueued on the caller's stream so far; `join()` makes the caller's stream wait for
    every side stream used since the last join.  Works eagerly and inside a CUDA-graph capture (the capture then records
    parallel branches).  Tensors allocated inside a fork belong to the side stream's pool: keep their use inside the fork,
    or on the caller's stream AFTER the join (what the fused fusion path does: the branches write slices of a tensor the
    caller allocated).  The streams are created once per device and reused.
    """

    _pool: dict = {}

    def __init__(self, device: torch.device, n: int = 2):
        key = (device.index if device.index is not None else torch.cuda.current_device(), n)
        if key not in BranchStreams._pool:
            BranchStreams._pool[key] = [torch.cuda.Stream(device=device) for _ in range(n)]
        self.streams = BranchStreams._pool[key]
        self.device = device
        self.used: list = []

    def fork(self, i: int):
        side = self.streams[i]
        side.wait_stream(torch.cuda.current_stream(self.device))
        self.used.append(side)
        return torch.cuda.stream(side)

    def join(self):
        cur = torch.cuda.current_stream(self.device)
        for side in self.used:
            cur.wait_stream(side)
        self.used = []
