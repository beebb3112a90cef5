"""The mbarrier protocol of the tcgen05 MLP kernel, replayed on the CPU (tests/tc_protocol_model.py): no
schedule of the producers, the MMA issuer, the epilogues, the relay, the bulk copies and the tensor pipe may
deadlock or touch a ring slot / accumulator out of turn.  Ring depths are the kernel's: one CTA — global mode 6
pairs, cell mode 4; a cta_group::2 pair — 12 and 8 (each CTA holds half of every pair)."""
import pytest

from tests import tc_protocol_model as model


@pytest.mark.parametrize("ring_pairs,trip,cg", [(4, 1, 1), (6, 1, 1), (6, 2, 1), (4, 2, 1), (8, 1, 2), (12, 1, 2), (12, 2, 2)])
@pytest.mark.parametrize("n_tiles", [1, 2, 3])
def test_protocol_has_no_deadlock_or_hazard(ring_pairs, trip, cg, n_tiles):
    for seed in range(25):
        assert model.run(n_tiles, ring_pairs, seed, trip=trip, cg=cg) == "ok", f"seed {seed}"


def test_model_catches_a_shared_act_ready_barrier():
    """The bug that hung a B200: one act_ready barrier for every K-pair lets the epilogue get two phases ahead."""
    outcomes = set()
    for seed in range(20):
        bars = model.make_bars(6, 1)
        bars["act_ready"] = [bars["act_ready"][0]] * 4
        outcomes.add(model.run(1, 6, seed, bars=bars))
    assert any(o.startswith("deadlock") or o.startswith("hazard") for o in outcomes), outcomes
    assert model.run(1, 6, 0) == "ok"


def test_model_catches_a_pair_without_the_relay():
    """cta_group::2: if the leader did not wait for the follower's half (peer_full), it would issue on stale weights."""
    outcomes = set()
    for seed in range(20):
        bars = model.make_bars(12, 2)
        for b in bars["peer_full"]:
            b.done = lambda parity: True           # the leader never waits
        outcomes.add(model.run(2, 12, seed, cg=2, bars=bars))
    assert any(o.startswith("hazard") for o in outcomes), outcomes


@pytest.mark.parametrize("ring_pairs,cg", [(4, 1), (8, 2)])
@pytest.mark.parametrize("n_tiles", [1, 2, 3])
def test_cell_mode_protocol_has_no_deadlock_or_hazard(ring_pairs, cg, n_tiles):
    """Cell mode: two warp sets per CTA, each draining one layer-5 accumulator buffer and arriving for two."""
    for seed in range(25):
        assert model.run(n_tiles, ring_pairs, seed, cg=cg, cell=True) == "ok", f"seed {seed}"


def test_model_catches_cell_mode_without_the_tile_barrier():
    """The launch failure of round 2: without the block barrier at the end of a tile, the warps that do not drain chunk 7
    reach layer 2 of the next tile while chunk 7 is in flight, and their parity wait on acc_full[1] passes four phases early."""
    outcomes = {model.run(3, 8, seed, cg=2, cell=True, tile_barrier=False) for seed in range(40)}
    assert any(o.startswith("hazard") or o.startswith("deadlock") for o in outcomes), outcomes
