"""N>1 host logic on CPU: two `gloo` ranks exercise the frame sharding, the max-over-ranks step time
and the count gather of runtime.py (frames are independent, no data-path collective)."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_frames: int, out_dir: str):
    sys.path.insert(0, str(ROOT))
    from bevfusion_multimodal_3d_object_detection_b200 import runtime

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        begin, end = runtime.shard_range(n_frames, rank, world)
        # every frame's "detection count" is a function of its global index, so the gather is checkable
        local = [(7 * i) % 11 for i in range(begin, end)]
        res = {
            "range": [begin, end],
            "max": runtime.max_over_ranks(10.0 + rank),
            "sum": runtime.sum_over_ranks(float(end - begin)),
            "counts": runtime.gather_counts(local),
        }
        Path(out_dir, f"rank{rank}.json").write_text(json.dumps(res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [64, 7, 1])
def test_two_gloo_ranks_shard_and_reduce(tmp_path, n_frames):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, n_frames, str(tmp_path)), nprocs=world, join=True)
    res = [json.loads((tmp_path / f"rank{r}.json").read_text()) for r in range(world)]
    # contiguous, disjoint, covering shards; sizes differ by at most one
    assert res[0]["range"][0] == 0 and res[0]["range"][1] == res[1]["range"][0] and res[1]["range"][1] == n_frames
    sizes = [r["range"][1] - r["range"][0] for r in res]
    assert max(sizes) - min(sizes) <= 1
    for r in res:
        assert r["max"] == 11.0                       # the slower rank's time is what gets reported
        assert r["sum"] == float(n_frames)            # units all ranks processed
        assert r["counts"] == [(7 * i) % 11 for i in range(n_frames)]   # global frame order, ragged shards


def test_shard_range_validation_and_single_process_identities():
    from bevfusion_multimodal_3d_object_detection_b200 import runtime

    with pytest.raises(ValueError):
        runtime.shard_range(8, 2, 2)
    for world in (1, 2, 3, 8):
        covered = []
        for r in range(world):
            b, e = runtime.shard_range(13, r, world)
            covered.extend(range(b, e))
        assert covered == list(range(13))
    assert runtime.max_over_ranks(3.5) == 3.5 and runtime.sum_over_ranks(2.0) == 2.0
    assert runtime.gather_counts([1, 2, 3]) == [1, 2, 3]


def test_reference_arm_prints_one_line_from_rank0_only():
    """`bench.py --impl reference` under a 2-rank launch: rank 0 alone measures and prints, rank 1 exits 0."""
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    outs = []
    for rank in (1, 0):
        env.update(RANK=str(rank), LOCAL_RANK=str(rank))
        p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                            "--warmup", "0", "--cpu-frames", "1"], env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(p.stdout.strip())
    assert outs[0] == ""
    line = json.loads(outs[1].splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["unit"] == "frames/s" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["cpu_baseline"]["cores"] >= 1
