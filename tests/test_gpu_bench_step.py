"""The bench's step (bench.HotPath.step: bin/sort -> MLP with canvas + global max -> radar -> calibrated projection -> fusion ->
head -> decode) gives the same outputs with its independent branches on side streams (parallel branches of the captured graph)
and on one stream, eagerly and as a replayed CUDA graph."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu


def _flat(out):
    res = {}
    for k, v in out.items():
        if isinstance(v, dict):
            res.update({f"{k}.{kk}": vv for kk, vv in v.items() if torch.is_tensor(vv)})
        elif torch.is_tensor(v):
            res[k] = v
    return res


def test_step_with_parallel_branches_equals_the_serial_step(cuda):
    import bench

    hp = bench.HotPath("step", 2, "bf16", cuda, 42)
    outs = {}
    for parallel in (False, True):
        bench.PARALLEL_BRANCHES = parallel
        hp.chain.fusion.b200_parallel_branches = parallel
        for _ in range(2):
            out = hp.step(hp.inputs)
        torch.cuda.synchronize()
        outs[parallel] = {k: v.clone() for k, v in _flat(out).items()}
    assert outs[True].keys() == outs[False].keys() and len(outs[True]) >= 4
    for k in outs[True]:
        assert torch.equal(outs[True][k], outs[False][k]), k
    # the captured graph (parallel branches) replays to the same tensors
    bench.PARALLEL_BRANCHES = True
    hp.chain.fusion.b200_parallel_branches = True
    graph = hp.capture()
    for _ in range(3):
        rep = graph.replay()
    torch.cuda.synchronize()
    for k, v in _flat(rep).items():
        assert torch.equal(v, outs[False][k]), k
