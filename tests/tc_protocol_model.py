"""Executable model of the synchronisation protocol of csrc/pointnet_mlp_tc.cu (test infrastructure).

Three agents — the weight producer, the MMA issuer and the epilogue — talk through mbarriers whose waiters
only see a PHASE PARITY (mbarrier.try_wait.parity): a waiter that falls two phases behind, or a barrier
that is re-armed before a waiter's second look at it, deadlocks or reads stale data.  Two such bugs were
found on the GPU the expensive way (a hung box); this model replays the protocol under random schedules with
asynchronous completions (bulk copies land in any order, the tensor pipe retires MMAs in issue order at
arbitrary times) and reports deadlocks and data hazards.  It mirrors the loops of the kernel one to one:
change both together.
"""
from __future__ import annotations

import random
from collections import deque

PAIRS_PER_TILE = 1 + 2 + 8 + 32           # kPairsPerTile
K_PAIRS = (1, 1, 2, 4)                    # K-pairs per chunk of network layers 2..5 (layer 2: half a pair)
N_CHUNKS = (1, 2, 4, 8)


def acc_buffer(layer: int, c: int) -> int:
    """acc_buffer() of the kernel."""
    if layer == 3:
        return c & 1
    if layer == 0:
        return 1
    if layer == 1:
        return 1 if c == 0 else 2
    return 2 if c == 1 else 1


class Bar:
    def __init__(self):
        self.bit = 0

    def done(self, parity: int) -> bool:      # mbarrier.try_wait.parity
        return self.bit != parity

    def complete(self):
        self.bit ^= 1


def producer(n_tiles, ring_pairs, bars, inflight):
    pair = phase = 0
    img = 0
    for j in range(n_tiles * PAIRS_PER_TILE):
        yield ("wait", bars["empty"][pair], phase ^ 1)
        inflight.append((pair, j))             # expect_tx + cp.async.bulk
        yield ("step",)
        img = (img + 1) % PAIRS_PER_TILE
        pair += 1
        if pair == ring_pairs:
            pair, phase = 0, phase ^ 1


def issuer(n_tiles, ring_pairs, bars, pipe, state, trip=2):
    pair = phase = 0
    act_phase = [0, 0, 0, 0]
    acc_parity = [0, 0, 0]
    stream = 0
    for t in range(n_tiles):
        for layer in range(4):
            for c in range(N_CHUNKS[layer]):
                buf = acc_buffer(layer, c)
                yield ("wait", bars["acc_empty"][buf], acc_parity[buf] ^ 1)
                acc_parity[buf] ^= 1
                kp = 0
                while kp < K_PAIRS[layer]:
                    two = trip == 2 and kp + 1 < K_PAIRS[layer]
                    pr1, ph1 = pair + 1, phase
                    if pr1 == ring_pairs:
                        pr1, ph1 = 0, ph1 ^ 1
                    if c == 0:
                        yield ("wait", bars["act_ready"][kp], act_phase[kp])
                        act_phase[kp] ^= 1
                        if two:
                            yield ("wait", bars["act_ready"][kp + 1], act_phase[kp + 1])
                            act_phase[kp + 1] ^= 1
                    yield ("wait", bars["full"][pair], phase)
                    if two:
                        yield ("wait", bars["full"][pr1], ph1)
                    # data checks at issue time
                    assert state["ring"][pair] == stream, f"ring pair {pair} holds {state['ring'][pair]}, wanted {stream}"
                    if two:
                        assert state["ring"][pr1] == stream + 1, "second ring pair holds the wrong weights"
                    assert state["acc_owner"][buf] in (None, (t, layer, c)), f"accumulator {buf} overwritten before it was drained"
                    state["acc_owner"][buf] = (t, layer, c)
                    last = kp + trip >= K_PAIRS[layer]
                    pipe.append(("mma", pair, (buf, (t, layer, c)) if (last and not two) else None))
                    if two:
                        pipe.append(("mma", pr1, (buf, (t, layer, c)) if last else None))
                    stream += 2 if two else 1
                    yield ("step",)
                    if two:
                        pair, phase = pr1, ph1
                    pair += 1
                    if pair == ring_pairs:
                        pair, phase = 0, phase ^ 1
                    kp += trip


def epilogue(n_tiles, bars, state):
    full_phase = [0, 0, 0]
    bars["act_ready"][0].complete()            # layer 1 of the first tile
    yield ("step",)
    for t in range(n_tiles):
        more = t + 1 < n_tiles
        for layer in range(3):
            for c in range(N_CHUNKS[layer]):
                buf = acc_buffer(layer, c)
                yield ("wait", bars["acc_full"][buf], full_phase[buf])
                full_phase[buf] ^= 1
                assert state["acc_done"][buf] == (t, layer, c), "epilogue drains an accumulator that holds another chunk"
                state["acc_owner"][buf] = None
                bars["acc_empty"][buf].complete()
                yield ("step",)
                if layer == 2 and c == 2:
                    yield ("wait", bars["acc_empty"][2], 1)
                bars["act_ready"][c].complete()
                yield ("step",)
        for c in range(8):
            buf = c & 1
            yield ("wait", bars["acc_full"][buf], full_phase[buf])
            full_phase[buf] ^= 1
            assert state["acc_done"][buf] == (t, 3, c)
            if c == 7 and more:
                bars["act_ready"][0].complete()   # layer 1 of the next tile
                yield ("step",)
            state["acc_owner"][buf] = None
            bars["acc_empty"][buf].complete()
            yield ("step",)


def run(n_tiles: int, ring_pairs: int, seed: int, max_steps: int = 2_000_000, trip: int = 2):
    """Returns 'ok' or a description of the failure."""
    rnd = random.Random(seed)
    bars = {"full": [Bar() for _ in range(ring_pairs)], "empty": [Bar() for _ in range(ring_pairs)],
            "acc_full": [Bar() for _ in range(3)], "acc_empty": [Bar() for _ in range(3)],
            "act_ready": [Bar() for _ in range(4)]}
    inflight, pipe = [], deque()
    state = {"ring": [None] * ring_pairs, "acc_owner": [None] * 3, "acc_done": [None] * 3}
    agents = {"producer": producer(n_tiles, ring_pairs, bars, inflight),
              "issuer": issuer(n_tiles, ring_pairs, bars, pipe, state, trip),
              "epilogue": epilogue(n_tiles, bars, state)}
    pending = {}
    for name, g in list(agents.items()):
        pending[name] = next(g)
    try:
        for _ in range(max_steps):
            choices = []
            for name, req in pending.items():
                if req[0] == "step" or (req[0] == "wait" and req[1].done(req[2])):
                    choices.append(name)
            if inflight:
                choices.append("land")
            if pipe:
                choices.append("retire")
            if not choices:
                return "ok" if not pending else f"deadlock: waiting {sorted(pending)}"
            pick = rnd.choice(choices)
            if pick == "land":
                pair, j = inflight.pop(rnd.randrange(len(inflight)))
                state["ring"][pair] = j
                bars["full"][pair].complete()
            elif pick == "retire":
                _, pair, acc = pipe.popleft()
                bars["empty"][pair].complete()     # tcgen05.commit -> empty[pair]
                if acc is not None:
                    state["acc_done"][acc[0]] = acc[1]
                    bars["acc_full"][acc[0]].complete()
            else:
                try:
                    pending[pick] = next(agents[pick])
                except StopIteration:
                    del pending[pick]
        return "step limit"
    except AssertionError as e:
        return f"hazard: {e}"
