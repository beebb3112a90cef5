"""The mbarrier protocol of the tcgen05 MLP kernel, replayed on the CPU (tests/tc_protocol_model.py): no
schedule of the producer, the MMA issuer, the epilogue, the bulk copies and the tensor pipe may deadlock or
touch a ring slot / accumulator out of turn.  Ring depths are the kernel's (global mode 6 pairs, cell mode 4)."""
import pytest

from tests import tc_protocol_model as model


@pytest.mark.parametrize("ring_pairs,trip", [(4, 1), (6, 2), (4, 2), (6, 1)])   # cell mode, global mode, and the crosses
@pytest.mark.parametrize("n_tiles", [1, 2, 3])
def test_protocol_has_no_deadlock_or_hazard(ring_pairs, trip, n_tiles):
    for seed in range(40):
        assert model.run(n_tiles, ring_pairs, seed, trip=trip) == "ok", f"seed {seed}"


def test_model_catches_a_shared_act_ready_barrier(monkeypatch):
    """The bug that hung a B200: one act_ready barrier for every K-pair lets the epilogue get two phases ahead."""
    real_run = model.run

    class SharedList(list):
        pass

    def run_shared(n_tiles, ring_pairs, seed):
        orig_bar = model.Bar
        made = []

        def factory():
            b = orig_bar()
            made.append(b)
            return b

        monkeypatch.setattr(model, "Bar", factory)
        try:
            # build normally, then alias the four act_ready barriers onto one object
            import random
            from collections import deque
            bars = {"full": [model.Bar() for _ in range(ring_pairs)], "empty": [model.Bar() for _ in range(ring_pairs)],
                    "acc_full": [model.Bar() for _ in range(3)], "acc_empty": [model.Bar() for _ in range(3)]}
            one = model.Bar()
            bars["act_ready"] = [one] * 4
            rnd = random.Random(seed)
            inflight, pipe = [], deque()
            state = {"ring": [None] * ring_pairs, "acc_owner": [None] * 3, "acc_done": [None] * 3}
            agents = {"producer": model.producer(n_tiles, ring_pairs, bars, inflight),
                      "issuer": model.issuer(n_tiles, ring_pairs, bars, pipe, state),
                      "epilogue": model.epilogue(n_tiles, bars, state)}
            pending = {k: next(g) for k, g in agents.items()}
            for _ in range(200000):
                ch = [k for k, r in pending.items() if r[0] == "step" or r[1].done(r[2])]
                if inflight:
                    ch.append("land")
                if pipe:
                    ch.append("retire")
                if not ch:
                    return "ok" if not pending else "deadlock"
                p = rnd.choice(ch)
                if p == "land":
                    pair, j = inflight.pop(rnd.randrange(len(inflight)))
                    state["ring"][pair] = j
                    bars["full"][pair].complete()
                elif p == "retire":
                    _, pair, acc = pipe.popleft()
                    bars["empty"][pair].complete()
                    if acc:
                        state["acc_done"][acc[0]] = acc[1]
                        bars["acc_full"][acc[0]].complete()
                else:
                    try:
                        pending[p] = next(agents[p])
                    except StopIteration:
                        del pending[p]
            return "step limit"
        except AssertionError:
            return "hazard"
        finally:
            monkeypatch.setattr(model, "Bar", orig_bar)

    outcomes = {run_shared(1, 6, seed) for seed in range(20)}
    assert outcomes - {"ok"}, "the model no longer detects the shared-barrier deadlock"
    assert real_run(1, 6, 0) == "ok"
