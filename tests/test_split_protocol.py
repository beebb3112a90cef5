"""The mbarrier protocol of the fp32-accuracy GEMM kernel (split_gemm_kernel), replayed on the CPU
(tests/split_protocol_model.py): no schedule may deadlock, rewrite an operand slot that MMAs still read, issue on data that has
not landed, or hand an accumulator half over early."""
import pytest

from tests import split_protocol_model as model


@pytest.mark.parametrize("n_tiles,nkb", [(1, 1), (3, 1), (2, 2), (4, 4), (3, 8), (7, 3)])
def test_split_gemm_protocol_has_no_deadlock_or_hazard(n_tiles, nkb):
    for seed in range(6):
        assert model.run(n_tiles, nkb, seed) == "ok", f"seed {seed}"
    for seed, slow in enumerate(("epilogue", "producer", "weights", "issuer", "copy", "retire")):
        assert model.run(n_tiles, nkb, 100 + seed, slow=(slow,)) == "ok", f"slow {slow}"


def test_model_catches_an_x_slot_rewritten_too_early():
    """If the producers did not wait for `empty_x`, a slot would be rewritten under the MMAs that read it."""
    outcomes = set()
    for seed in range(20):
        bars = model.make_bars()
        for b in bars["empty_x"]:
            b.done = lambda parity: True
        outcomes.add(model.run(2, 4, seed, bars=bars, slow=("retire",)))
    assert any(o.startswith("hazard") or o.startswith("deadlock") for o in outcomes), outcomes


@pytest.mark.parametrize("n_tiles,nkb", [(1, 1), (3, 1), (2, 2), (4, 4), (3, 8)])
def test_split_gemm_pair_protocol_has_no_deadlock_or_hazard(n_tiles, nkb):
    """split_gemm_kernel<FINAL, 2>: the leader issues, the follower's producers and relay report to the leader, commits arrive in
    both CTAs, both CTAs' epilogue warps release the accumulator halves at the leader."""
    for seed in range(5):
        assert model.run_pair(n_tiles, nkb, seed) == "ok", f"seed {seed}"
    for seed, slow in enumerate(("epilogue", "producer", "weights", "relay", "issuer", "copy", "retire", "follower")):
        assert model.run_pair(n_tiles, nkb, 200 + seed, slow=(slow,)) == "ok", f"slow {slow}"


def test_pair_model_catches_a_leader_that_ignores_the_followers_x_rows():
    outcomes = set()
    for seed in range(20):
        cta = model.make_pair_bars()
        for b in cta[0]["peer_x"]:
            b.done = lambda parity: True
        outcomes.add(model.run_pair(2, 4, seed, cta=cta, slow=("follower",)))
    assert any(o.startswith("hazard: MMAs") for o in outcomes), outcomes
