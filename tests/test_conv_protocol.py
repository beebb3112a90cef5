"""The mbarrier protocol of the 3x3 tcgen05 convolution kernel, replayed on the CPU (tests/conv_protocol_model.py): no
schedule of the pixel producers, the weight warp, the MMA issuer, the epilogue warps, the copies and the tensor pipe may
deadlock, refill a slot that MMAs still read, issue on data that has not landed, or hand an accumulator half over early."""
import pytest

from tests import conv_protocol_model as model


@pytest.mark.parametrize("n_tiles,ncc", [(1, 1), (1, 4), (2, 2), (3, 2), (5, 1), (4, 12)])
def test_conv_protocol_has_no_deadlock_or_hazard(n_tiles, ncc):
    for seed in range(8):
        assert model.run(n_tiles, ncc, seed) == "ok", f"seed {seed}"
    # adversarial schedules: one kind of agent (or the copies, or the tensor pipe) forty times slower than the rest
    for seed, slow in enumerate(("epilogue", "producer", "weights", "issuer", "copy", "retire")):
        assert model.run(n_tiles, ncc, 100 + seed, slow=(slow,)) == "ok", f"slow {slow}"


@pytest.mark.parametrize("n_tiles,ncc,seg", [(1, 3, 3), (2, 6, 3), (3, 7, 3), (2, 36, 3), (4, 2, 3)])
def test_conv_protocol_with_k_segments(n_tiles, ncc, seg):
    """fp32-accuracy mode: every `seg` chunks the accumulator is handed to the epilogue (which adds it to the running total),
    so accumulator hand-overs outnumber tiles: the first chain of a tile lands in the half that keeps its total, the later
    ones in the other half — same barriers, each half counting its own uses."""
    for seed in range(6):
        assert model.run(n_tiles, ncc, seed, seg=seg) == "ok", f"seed {seed}"
    for seed, slow in enumerate(("epilogue", "producer", "weights", "issuer", "copy", "retire")):
        assert model.run(n_tiles, ncc, 200 + seed, slow=(slow,), seg=seg) == "ok", f"slow {slow}"


def test_model_catches_a_missing_accumulator_handshake():
    """Without the acc_empty wait the third tile's first MMA overwrites a half the epilogue has not read."""
    outcomes = set()
    for seed in range(20):
        bars = model.make_bars(4)
        for b in bars["acc_empty"]:
            b.done = lambda parity: True
        outcomes.add(model.run(4, 1, seed, bars=bars, slow=("epilogue",)))
    assert any(o.startswith("hazard: accumulator") or o.startswith("hazard: epilogue") for o in outcomes), outcomes


def test_model_catches_a_ring_refilled_too_early():
    """If the weight warp did not wait for `empty_a`, a stage would be overwritten under the MMAs that read it."""
    outcomes = set()
    for seed in range(20):
        bars = model.make_bars(4)
        for b in bars["empty_a"]:
            b.done = lambda parity: True
        outcomes.add(model.run(2, 2, seed, bars=bars, slow=("retire",)))
    assert any(o.startswith("hazard") or o.startswith("deadlock") for o in outcomes), outcomes


@pytest.mark.parametrize("n_tiles,ncc,seg", [(1, 1, None), (2, 2, None), (3, 4, None), (5, 1, None), (2, 6, 3), (3, 7, 3)])
def test_pair_protocol_has_no_deadlock_or_hazard(n_tiles, ncc, seg):
    """conv3x3_tc_halo_kernel<2> (and, with the same hand-overs, split_gemm_kernel<FINAL, 2>): leader issues, follower's producers
    and relay report to the leader, commits arrive in both CTAs, both CTAs' epilogue warps release the halves at the leader."""
    for seed in range(6):
        assert model.run_pair(n_tiles, ncc, seed, seg=seg) == "ok", f"seed {seed}"
    for seed, slow in enumerate(("epilogue", "producer", "weights", "relay", "issuer", "copy", "retire", "follower")):
        assert model.run_pair(n_tiles, ncc, 300 + seed, slow=(slow,), seg=seg) == "ok", f"slow {slow}"


def test_pair_model_catches_a_leader_that_ignores_the_follower():
    """Without the waits on peer_a / peer_blk the leader issues MMAs on operands the follower has not received yet."""
    outcomes = set()
    for seed in range(20):
        cta = model.make_pair_bars(4)
        for b in cta[0]["peer_a"] + cta[0]["peer_blk"]:
            b.done = lambda parity: True
        outcomes.add(model.run_pair(2, 2, seed, cta=cta, slow=("follower",)))
    assert any(o.startswith("hazard: MMAs") for o in outcomes), outcomes


def test_pair_model_catches_epilogues_that_release_locally():
    """If only the leader's four epilogue warps counted (acc_empty of 4), a half could be overwritten before the follower read it."""
    outcomes = set()
    for seed in range(30):
        cta = model.make_pair_bars(4)
        cta[0]["acc_empty"] = [model.Bar(model.N_EPILOGUE) for _ in range(2)]
        outcomes.add(model.run_pair(4, 1, seed, cta=cta, slow=("follower",)))
    assert any(o.startswith("hazard") or o.startswith("deadlock") for o in outcomes), outcomes
