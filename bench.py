#!/usr/bin/env python
"""Benchmark of the BEV encode + decode hot path (BASELINE.json metric: frames/sec, HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME] [--precision bf16|f32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One STEP = one complete inference pass behind the camera backbone over a batch of synthetic nuScenes-shaped frames: the
reference's own chain (FlexibleMultiModal3DDetector.forward without the ResNet, src/fusion.py:1113-1137, then eval.py's
decode, src/eval.py:58-62) PLUS the two north_star-form stages the reference lacks, all chained on real data:

    S1a bin_sort            (F,N,4) points -> cell / perm / offsets
    S1b pointnet_encode     shared MLP 4-64-128-256-512-1024 -> per-cell scatter-max canvas (north_star) AND the global
                            max (the reference's PointNetLiDAREncoder output, which feeds the fusion module), one pass
    S1c radar_encode        5 x (F,125,7) -> shared MLP 7-32-64-128-256 + max + concat-FC            (MultiRadarEncoder)
    S2  camera_project      (F,6,512,h,w) -> (F,512,H,W), calibrated projection + bilinear gather    (north_star form)
        fusion              FlexibleBEVFusion.forward: camera mean -> camera_proj convs -> bilinear resize; lidar_init ->
                            convs; radar_proj -> convs; concat -> bev_fusion convs                   (reference form)
        head                CenterNetHead.forward (five 3x3 + five 1x1 convs)
    S3  centernet_decode    sigmoid + 3x3 NMS + top-K + gather + boxes, one launch

`value`  : frames/s, inputs resident in HBM, the whole step captured in ONE CUDA graph and replayed (CUDA events, L2
           flushed between steps).  `kernels` gives the rooflines of the seven hot-path kernels measured one by one,
           `eager_stage_ms` the per-stage times of an eager pass.
`e2e`    : the same step from PINNED HOST buffers to boxes on the host — H2D of every input and D2H of the results inside
           the timed region, chunks double-buffered — plus a second variant with the camera features already on the device
           (where the reference's own camera encoder leaves them, src/fusion.py:1110) and the measured bare-H2D ceiling.
`configs`: every BASELINE.json configuration (lidar_only B=1, camera_only B=8, fusion B=32, full B=64 split over the
           ranks, stress 300k points / 100x100), each timed the same way.
`--impl reference` times the torch-CPU port of the reference's implementation of the same step (oracle/torch_port.py)
on the host cores, same frames per step unless that would not finish in a few minutes (then fewer, and it says so).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

N_CLASSES, TOPK = 10, 100                 # src/eval.py:61
IMG_W, IMG_H = 1600.0, 900.0
FEAT_C = 512
MAC_PER_POINT = 4 * 64 + 64 * 128 + 128 * 256 + 256 * 512 + 512 * 1024   # 696,576
L2_FLUSH_BYTES = 256 << 20
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
FP32_FMA_PEAK_DERIVED = 148 * 128 * 2 * 1965e6 / 1e12     # SMs x FP32 lanes x 2 flop x max SM clock; the measured figure is printed beside it

# BASELINE.json configs -> concrete shapes.  frames = per GPU (weak) unless total_frames is set (strong split).
WORKLOADS = {
    "step": dict(baseline="configs[2] 'camera+lidar fusion with CenterNet decode, batch 32 on 1 B200' plus the 5-radar branch of "
                          "configs[3]; per GPU, so N GPUs process N x frames (configs[3] at N=2)",
                 frames=32, points=35000, valid=34720, grid=50, feat=(57, 100), cam=True, lidar=True, radar=True),
    "lidar_only": dict(baseline="configs[0] lidar_only, batch 1, ~34k points, base.yaml BEV grid", frames=1, points=35000, valid=34720,
                       grid=50, feat=(57, 100), cam=False, lidar=True, radar=False),
    "camera_only": dict(baseline="configs[1] camera_only: 6x900x1600 ResNet-18 features projected to BEV, batch 8 on 1 B200", frames=8,
                        points=0, valid=0, grid=50, feat=(57, 100), cam=True, lidar=False, radar=False),
    "fusion": dict(baseline="configs[2] camera+lidar fusion with CenterNet decode, batch 32 on 1 B200", frames=32, points=35000,
                   valid=34720, grid=50, feat=(57, 100), cam=True, lidar=True, radar=False),
    "full_split": dict(baseline="configs[3] camera+lidar+radar full fusion, batch 64 split over the ranks (strong scaling)",
                       total_frames=64, frames=64, points=35000, valid=34720, grid=50, feat=(57, 100), cam=True, lidar=True, radar=True),
    "stress": dict(baseline="configs[4] dense-scale stress: 10-sweep LiDAR (300k points) + radar, 2x BEV resolution (100x100), "
                            "batch 256 over 8 B200 = 32 per GPU", frames=32, points=300000, valid=298000, grid=100, feat=(57, 100),
                   cam=True, lidar=True, radar=True, lidar_start=50),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="step", choices=sorted(WORKLOADS), help="the workload of the headline numbers")
    ap.add_argument("--frames", type=int, default=None, help="frames per GPU per step (default: the workload's)")
    ap.add_argument("--precision", default=os.environ.get("B200BEV_PRECISION", "bf16"), choices=["auto", "f32", "f32_cudnn", "bf16"])
    ap.add_argument("--chunk", type=int, default=4, help="frames per pipeline chunk in the e2e leg")
    ap.add_argument("--cpu-frames", type=int, default=None, help="frames per CPU pass (default: as many as fit the time budget)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the timed pass at the other precision and the per-kernel side measurements")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configurations")
    ap.add_argument("--no-affinity", action="store_true", help="do not bind the process to the GPU's NUMA node")
    return ap.parse_args()


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return {k: float(d[k]) for k in FALLBACK_PEAKS if k in d} | {"source": "measured"}
        except Exception:
            pass
    return dict(FALLBACK_PEAKS) | {"source": "fallback"}


def load_traffic(stage: str, dtype: str, frames: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the stage's kernel, from the committed
    `ncu --set full` capture (profiles/traffic.json); None when no capture matches this workload."""
    try:
        entry = json.loads((ROOT / "profiles" / "traffic.json").read_text())[f"{stage}:{dtype}"]
        return float(entry["bytes_per_launch"]) if int(entry["frames"]) == frames else None
    except Exception:
        return None


# --------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.12)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0: float, t1: float):
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.1 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        num = lambda s: float(s) if s.replace(".", "", 1).isdigit() else None
        sm = [v for v in (num(r[0]) for r in rows) if v is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        pw = [v for v in (num(r[2]) for r in rows) if v is not None]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": num(rows[0][1]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(pw) if pw else None}


def workload_config(name: str, wl: dict, frames: int, world: int):
    stages = []
    if wl["lidar"]:
        stages += ["bin_sort", "pointnet_encode(canvas+global)"]
    if wl["radar"]:
        stages.append("radar_encode")
    if wl["cam"]:
        stages.append("camera_project")
    stages += ["fusion(FlexibleBEVFusion.forward)", "head(CenterNetHead.forward)", "centernet_decode"]
    return {
        "workload": f"{name}: BASELINE {wl['baseline']}",
        "frames_per_gpu": frames, "lidar_points": wl["points"] if wl["lidar"] else 0,
        "radars": "5x125x7" if wl["radar"] else None,
        "camera_features": f"6x{FEAT_C}x{wl['feat'][0]}x{wl['feat'][1]} (900x1600 / stride 16)" if wl["cam"] else None,
        "bev_grid": f"{wl['grid']}x{wl['grid']}", "classes": N_CLASSES, "topk": TOPK, "stages": stages,
        "l2": "256 MiB written between timed steps (L2 flush); with cameras the 2.2 GB of features alone exceed L2",
    }


def frames_for(wl: dict, world: int, rank: int, override=None):
    if override is not None:
        return int(override)
    if "total_frames" in wl:
        from bevfusion_multimodal_3d_object_detection_b200 import runtime
        b, e = runtime.shard_range(wl["total_frames"], rank, world)
        return e - b
    return wl["frames"]


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: torch-CPU port of the reference's implementation of the same step
# --------------------------------------------------------------------------------------------------
class CpuWorkload:
    def __init__(self, name: str, frames: int):
        import torch

        import bevfusion_multimodal_3d_object_detection_b200 as b200bev
        from bevfusion_multimodal_3d_object_detection_b200 import synthetic as syn
        from oracle import bev_oracle as orc
        from oracle import torch_port as tp

        wl = WORKLOADS[name]
        self.tp, self.torch, self.wl = tp, torch, wl
        torch.set_num_threads(os.cpu_count() or 1)
        self.cores = torch.get_num_threads()
        self.frames = frames
        G, (fh, fw) = wl["grid"], wl["feat"]
        self.G = G
        chain = b200bev.BEVDetectorChain(use_camera=wl["cam"], use_lidar=wl["lidar"], use_radar=wl["radar"], bev_h=G, bev_w=G,
                                         lidar_start_size=wl.get("lidar_start"))
        self.sd = {k: torch.from_numpy(v) for k, v in syn.detector_state(42, chain.state_shapes()).items()}
        del chain
        g = torch.Generator().manual_seed(42)
        self.lidar = torch.from_numpy(syn.lidar_batch(42, frames, n_valid=wl["valid"], n_total=wl["points"])) if wl["lidar"] else None
        self.radars = [torch.from_numpy(r) for r in syn.radar_batch(43, frames)] if wl["radar"] else None
        self.feats = torch.relu(torch.randn((frames, 6, FEAT_C, fh, fw), generator=g)) if wl["cam"] else None
        if wl["lidar"]:
            self.lidar_layers = tp.mlp_layers_from_state(self.sd, "lidar_encoder.")
        if wl["cam"]:
            K, E = syn.camera_rig(IMG_W, IMG_H)
            self.table = torch.from_numpy(orc.project_cells(K, E, (IMG_W, IMG_H), (fh, fw), (G, G), syn.PC_RANGE))
        self.pc_range = syn.PC_RANGE

    def step(self):
        tp, sd, G = self.tp, self.sd, self.G
        canvas = proj = lidar_feat = radar_feat = None
        if self.lidar is not None:
            cell, perm = tp.cell_index_and_sort(self.lidar, self.pc_range, G, G)                 # S1a
            canvas, lidar_feat = tp.cell_canvas(self.lidar, self.lidar_layers, cell, G * G, with_global=True)   # S1b
        if self.radars is not None:
            radar_feat = tp.multi_radar(self.radars, tp.mlp_layers_from_state(sd, "radar_encoder.radar_encoder."),
                                        sd["radar_encoder.fusion_fc.weight"], sd["radar_encoder.fusion_fc.bias"])   # S1c
        if self.feats is not None:
            proj = tp.camera_project(self.feats, self.table, (G, G))                              # S2, north_star form
        bev = tp.fusion_forward(sd, self.feats, lidar_feat, radar_feat, (G, G))                   # S2, reference form + glue
        pred = tp.head_forward(sd, bev)
        dets = tp.decode(pred, score_thresh=0.0, max_detections=TOPK, voxel_size=0.512)           # S3
        return canvas, proj, dets


def time_cpu(name: str, frames, steps: int, warmup: int, budget_s: float, max_frames: int):
    """Times `steps` passes of the CPU port.  frames=None: as many frames per pass as the b200 arm steps, unless
    (steps + warmup) passes would exceed `budget_s` — then fewer (frames are independent: the per-frame cost is what is
    measured).  Returns (workload, seconds, passes)."""
    if frames is None:
        probe = CpuWorkload(name, 1)
        probe.step()
        t0 = time.perf_counter()
        probe.step()
        per_frame = time.perf_counter() - t0
        del probe
        frames = int(max(1, min(max_frames, budget_s / max(per_frame * (steps + warmup), 1e-9))))
    wl = CpuWorkload(name, frames)
    for _ in range(warmup):
        wl.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        wl.step()
    return wl, time.perf_counter() - t0, steps


def cpu_sample_text(wl, name, passes, dt, full):
    return (f"{passes} passes over {wl.frames} full-size frames per pass (the b200 arm steps {full} per GPU; frames are "
            f"independent) of workload '{name}': bin/sort, MLP + per-cell scatter-max + global max, radar, calibrated camera "
            f"projection, FlexibleBEVFusion, CenterNetHead, decode — torch-CPU port of the reference ops, {dt:.1f} s of CPU work")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    full = args.frames or WORKLOADS[name]["frames"]
    wl, dt, passes = time_cpu(name, args.cpu_frames, max(args.steps, 1), max(args.warmup, 1), budget_s=170.0, max_frames=full)
    fps = wl.frames * passes / dt
    line = {
        "impl": "reference", "metric": "bev_encode_decode_frames_per_sec", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / passes * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, WORKLOADS[name], wl.frames, 1),
        "same_step_as_b200_arm": True, "frames_per_step_b200_arm": full,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": wl.cores, "kind": "port", "sample": cpu_sample_text(wl, name, passes, dt, full)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def _event(torch):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


class HotPath:
    """One workload on one GPU: modules, device-resident inputs, the step, its CUDA graph."""

    def __init__(self, name: str, frames: int, precision: str, dev, seed: int):
        import numpy as np
        import torch

        import bevfusion_multimodal_3d_object_detection_b200 as b200bev
        from bevfusion_multimodal_3d_object_detection_b200 import _lib, conv_blocks, encoders, ops
        from bevfusion_multimodal_3d_object_detection_b200 import synthetic as syn

        wl = WORKLOADS[name]
        self.name, self.wl, self.F, self.dev, self.precision = name, wl, frames, dev, precision
        self.torch, self.ops, self._lib, self.encoders, self.conv_blocks = torch, ops, _lib, encoders, conv_blocks
        G, (fh, fw) = wl["grid"], wl["feat"]
        self.G = G
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.chain = b200bev.BEVDetectorChain(use_camera=wl["cam"], use_lidar=wl["lidar"], use_radar=wl["radar"], bev_h=G, bev_w=G,
                                              lidar_start_size=wl.get("lidar_start"), precision=precision)
        sd = syn.detector_state(42, self.chain.state_shapes())
        self.chain.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        del sd
        self.chain = self.chain.eval().to(dev)
        if not PARALLEL_BRANCHES:
            self.chain.fusion.b200_parallel_branches = False
        self.inputs = {}
        if wl["lidar"]:
            self.inputs["lidar"] = to(syn.lidar_batch(seed, frames, n_valid=wl["valid"], n_total=wl["points"]))
        if wl["radar"]:
            for i, r in enumerate(syn.radar_batch(seed + 1, frames)):
                self.inputs[f"radar{i}"] = to(r)
        if wl["cam"]:
            g = torch.Generator(device=dev).manual_seed(seed + 2)
            self.inputs["feats"] = torch.relu(torch.randn((frames, 6, FEAT_C, fh, fw), device=dev, generator=g))
            Kc, Ec = syn.camera_rig(IMG_W, IMG_H)
            self.K, self.E = to(Kc), to(Ec)
            _, table = ops.camera_project(self.inputs["feats"][:1, :, :1].contiguous(), self.K, self.E, (IMG_W, IMG_H), (G, G),
                                          return_table=True)
            self.hits = int(table[0, :, :, 2].sum().item())     # (cell, camera) pairs in view, from the kernel's own table
        self.graph = None

    def lidar_params(self):
        enc = self.chain.lidar_encoder
        prec = self.encoders._precision_of(enc)
        blob, dims, tc = self.encoders.packed_params(enc, self.dev, precision=prec)
        return prec, blob, dims, tc

    # the step, on any dict of inputs with this workload's keys (device-resident tensors or a pipeline slot)
    def step(self, inp, marks=None):
        torch, ops, wl, G = self.torch, self.ops, self.wl, self.G
        mark = (lambda n: marks.append((n, _event(torch)))) if marks is not None else (lambda n: None)
        out = {}
        lidar_feat = radar_feat = None
        # Graph capture (no stage marks): the radar encoder and the calibrated projection do not depend on the LiDAR encoder and
        # go to side streams — parallel branches of the captured graph.  The eager, stage-by-stage timing pass stays serial.
        import contextlib
        from bevfusion_multimodal_3d_object_detection_b200 import runtime
        n_frames = int(inp["lidar" if wl["lidar"] else "feats"].shape[0])      # a pipeline chunk is smaller than the workload
        small = n_frames * G * G <= 32 * 50 * 50    # measured: the side streams gain 2.5 % up to here and lose 1-2 % beyond
        fork = runtime.BranchStreams(self.dev) if (marks is None and PARALLEL_BRANCHES and small and
                                                   inp["lidar" if wl["lidar"] else "feats"].is_cuda) else None
        side = (lambda i: fork.fork(i)) if fork else (lambda i: contextlib.nullcontext())
        with torch.no_grad():
            if fork and wl["radar"]:
                with side(1):
                    radar_feat = self.chain.radar_encoder([inp[f"radar{i}"] for i in range(5)])
            if fork and wl["cam"]:
                with side(0):
                    out["proj"] = ops.camera_project(inp["feats"], self.K, self.E, (IMG_W, IMG_H), (G, G))
            if wl["lidar"]:
                prec, blob, dims, tc = self.lidar_params()
                mark("bin_sort")
                _, perm, off = ops.bin_sort(inp["lidar"], G, G)
                mark("pointnet_encode")
                lidar_feat, out["canvas"] = ops.pointnet_encode(inp["lidar"], blob, dims, perm=perm, offsets=off, n_cells=G * G,
                                                                precision=prec, tc_params=tc)
            if wl["radar"] and not fork:
                mark("radar_encode")
                radar_feat = self.chain.radar_encoder([inp[f"radar{i}"] for i in range(5)])
            if wl["cam"] and not fork:
                mark("camera_project")
                out["proj"] = ops.camera_project(inp["feats"], self.K, self.E, (IMG_W, IMG_H), (G, G))
            if fork:
                fork.join()
            mark("fusion")
            bev = self.chain.fusion(camera_features=inp.get("feats"), lidar_features=lidar_feat, radar_features=radar_feat)
            mark("head")
            pred = self.chain.det_head(bev)
            mark("centernet_decode")
            logits = self.conv_blocks.logits_of(pred["heatmap"])
            out["det"] = ops.centernet_decode(pred["heatmap"] if logits is None else logits, pred["offset"], pred["size"], pred["rot"],
                                              pred["vel"], TOPK, 0.512, score_thresh=0.0, heat_is_logit=logits is not None)
            mark("end")
        return out

    def capture(self):
        from bevfusion_multimodal_3d_object_detection_b200 import runtime
        self.graph = runtime.GraphedStep(lambda: self.step(self.inputs), self.dev)
        return self.graph

    def stage_names(self):
        wl = self.wl
        return (["bin_sort", "pointnet_encode"] if wl["lidar"] else []) + (["radar_encode"] if wl["radar"] else []) + \
               (["camera_project"] if wl["cam"] else []) + ["fusion", "head", "centernet_decode"]


PARALLEL_BRANCHES = os.environ.get("B200BEV_BENCH_SERIAL", "0") != "1"   # B200BEV_BENCH_SERIAL=1: one stream (A/B timing)


def timed_steps(hp: HotPath, steps: int, warmup: int, flush, barrier, clocks=None):
    """(graph ms per step, eager per-stage ms, warm-up passes run).  Graph replays and eager passes are timed separately."""
    torch, dev = hp.torch, hp.dev
    graph = hp.capture()
    t_warm, n_warm = time.time(), 0
    while n_warm < max(warmup, 3) or (clocks is not None and not clocks.rows and time.time() - t_warm < 3.0):
        graph.replay()
        torch.cuda.synchronize(dev)
        n_warm += 1
    barrier()
    step_ms = []
    for _ in range(steps):
        flush.zero_()                                   # L2 flush, outside the per-step event pair
        e0 = _event(torch)
        graph.replay()
        e1 = _event(torch)
        torch.cuda.synchronize(dev)
        step_ms.append(e0.elapsed_time(e1))
    barrier()
    # eager pass with an event per stage: where the step's time goes (launch gaps included, so the sum exceeds the graph)
    stage_ms = {n: 0.0 for n in hp.stage_names()}
    reps = max(3, min(steps, 10))
    hp.step(hp.inputs)
    for _ in range(reps):
        flush.zero_()
        marks = []
        hp.step(hp.inputs, marks)
        torch.cuda.synchronize(dev)
        for (n, e), (_, e_next) in zip(marks[:-1], marks[1:]):
            stage_ms[n] += e.elapsed_time(e_next)
    return step_ms, {n: v / reps for n, v in stage_ms.items()}, n_warm


def kernel_rooflines(hp: HotPath, peaks, fp32_peak, flush):
    """The hot-path kernels one by one (median of 5, L2 cold and clean): algorithmic work per launch
    over the measured duration.  Tensor fractions are given against the burst AND the sustained bf16 peak."""
    torch, ops, wl, F, G, dev = hp.torch, hp.ops, hp.wl, hp.F, hp.G, hp.dev
    fh, fw = wl["feat"]
    hw, HW = fh * fw, G * G

    def med_ms(fn, reps=5):
        fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            flush.view(torch.int64).sum()               # leave L2 cold AND clean: the flush's dirty lines are not this kernel's
            e0 = _event(torch)
            fn()
            e1 = _event(torch)
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    out = {}

    def put(name, ms, bound, work, note=None):
        sec = ms * 1e-3
        if bound == "hbm":
            ach = work / sec / 1e9
            row = {"ms": round(ms, 4), "bound": bound, "achieved": round(ach, 2), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                   "frac": round(ach / peaks["hbm_gbs"], 4)}
        elif bound == "fp32_fma":
            ach = work / sec / 1e12
            row = {"ms": round(ms, 4), "bound": bound, "achieved": round(ach, 2), "peak": round(fp32_peak["tflops"], 2), "unit": "TFLOP/s",
                   "frac": round(ach / fp32_peak["tflops"], 4), "peak_source": fp32_peak["source"]}
        else:
            ach = work / sec / 1e12
            row = {"ms": round(ms, 4), "bound": bound, "achieved": round(ach, 2), "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                   "frac": round(ach / peaks["bf16_tflops"], 4), "peak_kind": "burst (the kernel is timed alone)",
                   "frac_of_sustained": round(ach / peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]), 4)}
        if note:
            row["note"] = note
        out[name] = row

    with torch.no_grad():
        if wl["lidar"]:
            prec, blob, dims, tc = hp.lidar_params()
            lidar = hp.inputs["lidar"]
            N = lidar.shape[1]
            _, perm, off = ops.bin_sort(lidar, G, G)
            put("bin_sort", med_ms(lambda: ops.bin_sort(lidar, G, G)), "hbm", F * (24.0 * N + 4.0 * (HW + 1)))
            flops = F * N * 2.0 * MAC_PER_POINT
            ms_cell = med_ms(lambda: ops.pointnet_encode(lidar, blob, dims, perm=perm, offsets=off, n_cells=HW, precision=prec, tc_params=tc))
            ms_glob = med_ms(lambda: ops.pointnet_encode(lidar, blob, dims, precision=prec, tc_params=tc))
            if prec == hp._lib.BF16_TENSOR:
                put("pointnet_encode", ms_cell, "tensor", flops, "cell canvas + global max in one pass, bf16 tcgen05 (includes the canvas memset)")
                put("pointnet_encode_global_only", ms_glob, "tensor", flops, "the reference's PointNetLiDAREncoder.forward alone")
            else:
                note = ("fp32 accuracy (parity 1e-5): on the tensor-core path every fp32 product is three fp16 tcgen05 products, so the "
                        "tensor pipe executes 3x the algorithmic flops; `achieved` counts the ALGORITHMIC flops")
                put("pointnet_encode", ms_cell, "tensor", flops, "cell canvas + global max; " + note)
                put("pointnet_encode_global_only", ms_glob, "tensor", flops, note)
            if N <= 40000:
                # N3, the step in front of the path (not part of the timed step): raw sweeps -> filtered / padded points, alone and
                # fused into the bin-and-sort launch
                import numpy as np
                from bevfusion_multimodal_3d_object_detection_b200 import synthetic as syn
                rows = N + 8000
                raw = torch.from_numpy(np.concatenate([syn.raw_sweep(9000 + i, rows) for i in range(F)], axis=0)).to(dev)
                offs = torch.tensor([rows * i for i in range(F + 1)], dtype=torch.int64, device=dev)
                put("lidar_prepare", med_ms(lambda: ops.lidar_prepare(raw, offs, N, syn.PC_RANGE, max_frame_rows=rows)), "hbm",
                    F * 16.0 * (rows + N), f"{F} x {rows} raw rows -> {N}; the stage in front of the step")
                put("lidar_prepare_bin_sort", med_ms(lambda: ops.lidar_prepare_bin_sort(raw, offs, N, G, G, syn.PC_RANGE, max_frame_rows=rows)),
                    "hbm", F * (16.0 * (rows + N) + 24.0 * N + 4.0 * (HW + 1)), "range filter + compaction + padding + bin-and-sort, one launch")
                del raw
            try:
                # N2: the dense layers behind the encoder (FlexibleBEVFusion.lidar_init; the second layer's weight is the largest
                # single read of the step) as the step runs them: tensor cores at fp32 accuracy from the split-fp16 image
                from bevfusion_multimodal_3d_object_detection_b200.fusion import lidar_init_dense
                fus = hp.chain.fusion
                l0, l2 = fus.lidar_init[0], fus.lidar_init[2]
                gf = torch.rand((F, l0.in_features), device=dev)
                put("lidar_init", med_ms(lambda: lidar_init_dense(fus, gf)), "hbm",
                    4.0 * (l0.weight.numel() + l2.weight.numel() + l0.out_features + 2 * l2.out_features
                           + F * (l0.in_features + 2 * l0.out_features + l2.out_features)),
                    f"Linear({l0.in_features},{l0.out_features}) + ReLU + Linear({l2.in_features},{l2.out_features}), two launches")
                del gf
            except Exception as e:          # a chain without a lidar branch
                out["lidar_init"] = {"error": str(e)[:120]}
        if wl["radar"]:
            radars = [hp.inputs[f"radar{i}"] for i in range(5)]
            put("radar_encode", med_ms(lambda: hp.chain.radar_encoder(radars)), "fp32_fma",
                F * 625 * 2.0 * (7 * 32 + 32 * 64 + 64 * 128 + 128 * 256) + F * 2.0 * 1280 * 256)
        if wl["cam"]:
            feats = hp.inputs["feats"]
            put("camera_mean", med_ms(lambda: ops.camera_mean(feats)), "hbm", F * 4.0 * FEAT_C * hw * 7)
            put("camera_mean_nhwc_bf16", med_ms(lambda: ops.camera_mean_nhwc_bf16(feats)), "hbm", F * FEAT_C * hw * (6 * 4.0 + 2.0),
                "the mean delivered as camera_proj's channels-last bf16 input (what the bf16 step runs)")
            x = torch.rand((F, 256, fh, fw), device=dev)
            put("bilinear_resize", med_ms(lambda: ops.bilinear_resize(x, (G, G))), "hbm", F * 4.0 * 256 * (hw + HW))
            put("camera_project", med_ms(lambda: ops.camera_project(feats, hp.K, hp.E, (IMG_W, IMG_H), (G, G))), "hbm",
                F * 4.0 * FEAT_C * (min(6 * hw, 4 * hp.hits) + HW))
            del x
        maps = {k: torch.rand((F, c, G, G), device=dev) for k, c in (("heatmap", N_CLASSES), ("offset", 2), ("size", 3), ("rot", 2), ("vel", 2))}
        put("centernet_decode", med_ms(lambda: ops.centernet_decode(maps["heatmap"], maps["offset"], maps["size"], maps["rot"], maps["vel"],
                                                                    TOPK, 0.512, heat_is_logit=True)),
            "hbm", F * (4.0 * N_CLASSES * HW + 9 * 4.0 * TOPK + TOPK * (11 * 4.0 + 3 * 8.0)))
    return out


def measure_fp32_fma_peak(torch, dev):
    """The fp32-FMA yardstick, MEASURED: cuBLAS SGEMM (TF32 off) at 8192^3, best of 5 — the highest fp32-FMA rate a tuned
    kernel reaches on this part.  The derived lane peak (148 SM x 128 lanes x 2 x 1.965 GHz) is printed beside it."""
    try:
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        n = 8192
        a = torch.randn((n, n), device=dev)
        b = torch.randn((n, n), device=dev)
        torch.matmul(a, b)
        best = 1e9
        for _ in range(5):
            e0 = _event(torch)
            torch.matmul(a, b)
            e1 = _event(torch)
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        torch.backends.cuda.matmul.allow_tf32 = old
        del a, b
        return {"tflops": 2.0 * n ** 3 / (best * 1e-3) / 1e12, "source": "measured: cuBLAS fp32 SGEMM 8192^3, TF32 off, best of 5",
                "derived_lane_peak_tflops": round(FP32_FMA_PEAK_DERIVED, 2)}
    except Exception as e:
        return {"tflops": FP32_FMA_PEAK_DERIVED, "source": f"derived (148 SM x 128 lanes x 2 x 1.965 GHz); measurement failed: {str(e)[:80]}"}


def run_e2e(hp: HotPath, steps: int, warmup: int, chunk: int, barrier, features_from_host: bool):
    """Host buffers in, boxes out.  Every step: H2D of the step's inputs from pinned memory, the graphed step per chunk,
    D2H of the detections.  features_from_host=False leaves the camera features on the device (only points and radar cross)."""
    import torch

    from bevfusion_multimodal_3d_object_detection_b200 import runtime

    dev, F = hp.dev, hp.F
    keys = [k for k in hp.inputs if features_from_host or k != "feats"]
    if not keys:
        return None
    host = {k: hp.inputs[k].cpu().pin_memory() for k in keys}
    chunk = max(1, min(chunk, F))
    while F % chunk:
        chunk -= 1                                             # whole chunks only: every chunk replays the same graph
    if not features_from_host:
        chunk = F                                              # a few MB of points: one copy, one graph replay
    pipe = runtime.FramePipeline(host, chunk, dev)
    resident = {k: v for k, v in hp.inputs.items() if k not in keys}
    out_host = {
        "boxes": torch.empty((F, TOPK, 7), dtype=torch.float32).pin_memory(),
        "scores": torch.empty((F, TOPK), dtype=torch.float32).pin_memory(),
        "vel": torch.empty((F, TOPK, 2), dtype=torch.float32).pin_memory(),
        "count": torch.empty((F,), dtype=torch.int32).pin_memory(),
    }
    d2h_bytes = sum(t.numel() * t.element_size() for t in out_host.values())

    def slot_step(slot_inputs, b, e):
        inp = dict(slot_inputs)
        for k, v in resident.items():
            inp[k] = v[b:e]
        return hp.step(inp)["det"]

    def after(det, b, e):
        out_host["boxes"][b:e].copy_(det["boxes"], non_blocking=True)
        out_host["scores"][b:e].copy_(det["scores"], non_blocking=True)
        out_host["vel"][b:e].copy_(det["velocities"], non_blocking=True)
        out_host["count"][b:e].copy_(det["count"], non_blocking=True)

    graphed = True
    try:
        pipe.capture(slot_step)
    except Exception:
        graphed = False
    run = (lambda: pipe.run_captured(after)) if graphed else (lambda: pipe.run(lambda d, b, e: after(slot_step(d, b, e), b, e)))
    for _ in range(max(warmup, 3)):
        run()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
        torch.cuda.synchronize(dev)                      # results are on the host here
    barrier()
    dt = runtime.max_over_ranks(time.perf_counter() - t0, dev)
    # bare H2D of the same bytes, all ranks at once: the ceiling the host gives this pipeline
    barrier()
    ceil_ms = runtime.max_over_ranks(pipe.h2d_only_ms(reps=3), dev)
    counts = out_host["count"].tolist()
    res = {"ms_per_step": dt / steps * 1e3, "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "chunk_frames": pipe.chunk,
           "graphed_chunks": graphed, "h2d_gbs": pipe.h2d_bytes / (dt / steps) / 1e9, "h2d_ceiling_ms": ceil_ms,
           "h2d_ceiling_gbs": pipe.h2d_bytes / (ceil_ms * 1e-3) / 1e9 if ceil_ms > 0 else None,
           "of_h2d_ceiling": (ceil_ms / (dt / steps * 1e3)) if ceil_ms > 0 else None, "detections_frame0": counts[0]}
    del pipe, host
    return res


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    from bevfusion_multimodal_3d_object_detection_b200 import _lib, runtime

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    affinity = None if args.no_affinity else runtime.bind_to_gpu_numa(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    # the fp32 path must BE fp32: torch's default lets cuDNN run fp32 convolutions in TF32 (3.5e-3)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    _lib.enable_call_counting()
    peaks = load_peaks()
    state = {"precision": "bf16" if args.precision in ("auto", "bf16") else args.precision}
    seed = 42 + 1000 * rank
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def run_workload(name, steps, warmup, with_e2e, clocks=None, frames_override=None, main=False):
        """value / e2e / per-stage times of one workload at the current precision; returns (result dict, HotPath)."""
        wl = WORKLOADS[name]
        F = frames_for(wl, world, rank, frames_override)
        hp = HotPath(name, F, state["precision"], dev, seed)
        hp.step(hp.inputs)                      # builds the packed-weight caches
        _lib.reset_call_counts()
        hp.step(hp.inputs)
        launches = _lib.kernel_launches()
        step_ms, stage_ms, n_warm = timed_steps(hp, steps, warmup, flush, barrier, clocks)
        total_ms = runtime.max_over_ranks(sum(step_ms), dev)
        frames_all = runtime.sum_over_ranks(F, dev)
        res = {"name": name, "frames_per_gpu": F, "frames_total": int(frames_all), "steps": steps,
               "scaling": "strong" if "total_frames" in wl else "weak",
               "value": frames_all * steps / (total_ms * 1e-3), "unit": "frames/s", "ms_per_step": total_ms / steps,
               "eager_stage_ms": {k: round(v, 4) for k, v in stage_ms.items()}, "eager_ms_per_step": round(sum(stage_ms.values()), 4),
               "launches_per_step": launches, "warmup_run": n_warm, "config": workload_config(name, wl, F, world)}
        if with_e2e:
            e_steps = steps if main else max(3, min(steps, 10))
            host_leg = run_e2e(hp, e_steps, warmup, args.chunk, barrier, features_from_host=True)
            if host_leg is not None:
                host_leg.update(value=frames_all / (host_leg["ms_per_step"] * 1e-3), unit="frames/s")
                res["e2e"] = host_leg
            if wl["cam"] and (wl["lidar"] or wl["radar"]):
                dev_leg = run_e2e(hp, e_steps, warmup, args.chunk, barrier, features_from_host=False)
                if dev_leg is not None:
                    dev_leg.update(value=frames_all / (dev_leg["ms_per_step"] * 1e-3), unit="frames/s")
                    res["e2e_features_on_device"] = dev_leg
        return res, hp

    # ---- headline workload ----
    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    t_wall0 = time.time()
    main, hp = run_workload(args.workload, args.steps, args.warmup, not args.no_e2e, clocks, args.frames, main=True)
    clock_summary = clocks.summary(t_wall0, time.time())
    clocks.__exit__(None, None, None)
    F = hp.F

    kernels = fp32_peak = None
    if not args.no_alt:
        fp32_peak = measure_fp32_fma_peak(torch, dev)
        try:
            kernels = kernel_rooflines(hp, peaks, fp32_peak, flush)
        except Exception as e:      # never let a side measurement break the bench line
            kernels = {"error": str(e)[:300]}

    # dominant hot-path kernel of the step -> roofline.  A kernel timed alone is held against the BURST tensor peak; the
    # fraction of the sustained peak (what a kernel inside a long step can expect) is printed beside it.
    roofline = None
    if isinstance(kernels, dict) and "error" not in kernels:
        cand = {k: v for k, v in kernels.items() if k in ("bin_sort", "pointnet_encode", "radar_encode", "camera_project", "centernet_decode")}
        dom = max(cand, key=lambda k: cand[k]["ms"])
        roofline = {k: v for k, v in cand[dom].items() if k != "ms"}
        roofline.update({"kernel": dom, "kernel_ms": cand[dom]["ms"], "traffic": load_traffic(dom, state["precision"], F),
                         "peak_source": ("measured (MEASURED_PEAKS.json)" if peaks["source"] == "measured" else "fallback (B200_PROFILING.md)")})
    del hp
    torch.cuda.empty_cache()

    # ---- the other precisions, TIMED in the same contract (graph replays, CUDA events, L2 flush) ----
    NOTES = {"f32": "fp32 path (the drop-in's default precision, parity 1e-5): PointNet MLP and every convolution block at fp32 accuracy "
                    "on the tensor cores (three fp16 tcgen05 products per fp32 product), dense layers / radar MLP on fp32 FFMA",
             "f32_cudnn": "as f32, but the convolution blocks on the reference's own cuDNN layers in true fp32 (TF32 switched off)",
             "bf16": "bf16 path: tcgen05 MLP and convolution kernels (parity 1e-2)"}
    others = {}
    if not args.no_alt:
        saved = state["precision"]
        for prec in [p for p in ("f32", "f32_cudnn", "bf16") if p != saved]:
            try:
                state["precision"] = prec
                o_res, o_hp = run_workload(args.workload, max(3, min(args.steps, 10)), args.warmup, False, None, args.frames)
                other = {"dtype": prec, "value": o_res["value"], "unit": "frames/s", "ms_per_step": o_res["ms_per_step"],
                         "timed": True, "steps": o_res["steps"], "eager_stage_ms": o_res["eager_stage_ms"], "note": NOTES[prec]}
                if prec != "f32_cudnn":
                    try:
                        ok = kernel_rooflines(o_hp, peaks, fp32_peak, flush)
                        other["pointnet_encode"] = ok.get("pointnet_encode")
                        other["pointnet_encode_global_only"] = ok.get("pointnet_encode_global_only")
                    except Exception as e:
                        other["kernel_error"] = str(e)[:200]
                del o_hp
            except Exception as e:
                other = {"error": str(e)[:300]}
            others[prec] = other
            torch.cuda.empty_cache()
        state["precision"] = saved

    # ---- every BASELINE configuration ----
    configs = None
    if not args.no_configs:
        configs = []
        for name in ("lidar_only", "camera_only", "fusion", "full_split", "stress"):
            try:
                res, chp = run_workload(name, max(5, min(args.steps, 10)), args.warmup, not args.no_e2e)
                res.pop("config")
                res["baseline"] = WORKLOADS[name]["baseline"]
                del chp
                configs.append(res)
            except Exception as e:
                configs.append({"name": name, "error": str(e)[:300]})
            torch.cuda.empty_cache()

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only): bounded sample of the same step ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))        # the CPU arm gets every core again
        except Exception:
            pass
        try:
            wl_cpu, dt, passes = time_cpu(args.workload, args.cpu_frames, 2, 1, budget_s=20.0, max_frames=F)
            cpu_baseline = {"value": wl_cpu.frames * passes / dt, "unit": "frames/s", "cores": wl_cpu.cores, "kind": "port",
                            "sample": cpu_sample_text(wl_cpu, args.workload, passes, dt, F)}
        except Exception as e:
            cpu_baseline = {"error": str(e)[:300]}

    if rank == 0:
        e2e = main.get("e2e")
        precision = state["precision"]
        line = {
            "metric": "bev_encode_decode_frames_per_sec", "value": main["value"], "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "warmup_run": main["warmup_run"], "ms_per_step": main["ms_per_step"],
            "higher_is_better": True, "scaling": main["scaling"], "vs_baseline": None, "dtype": precision, "data": "synthetic",
            "config": main["config"], "roofline": roofline, "cpu_baseline": cpu_baseline,
            "e2e": None if e2e is None else {**e2e, "api": "BEVDetectorChain modules (the functions patch() installs on the reference's "
                                                          "classes) + ops.bin_sort / pointnet_encode(canvas) / camera_project, through "
                                                          "runtime.FramePipeline: pinned host -> device, double-buffered chunks, one CUDA "
                                                          "graph per chunk slot, boxes copied back to pinned host memory"},
            "e2e_features_on_device": main.get("e2e_features_on_device"),
            "gpu_launches": main["launches_per_step"] * args.steps, "launches_per_step": main["launches_per_step"],
            "clocks": clock_summary, "eager_stage_ms": main["eager_stage_ms"], "eager_ms_per_step": main["eager_ms_per_step"],
            "kernels": kernels, "fp32_fma_peak": fp32_peak,
            "fp32_path": others.get("f32"), "fp32_cudnn_path": others.get("f32_cudnn"), "bf16_path": others.get("bf16"),
            "configs": configs, "affinity": affinity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
