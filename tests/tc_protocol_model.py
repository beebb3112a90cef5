"""Executable model of the synchronisation protocol of csrc/pointnet_mlp_tc.cu (test infrastructure).

The kernel's agents — weight producer(s), the MMA issuer, the epilogue(s) and, for a cta_group::2 pair, the
follower's relay — talk through mbarriers whose waiters only see a PHASE PARITY (mbarrier.try_wait.parity): a
waiter that falls two phases behind, or a barrier that is re-armed before a waiter's second look at it,
deadlocks or reads stale data.  Two such bugs were found on the GPU the expensive way (a hung box); this model
replays the protocol under random schedules with asynchronous completions (bulk copies land in any order, the
tensor pipe retires MMAs in issue order at arbitrary times) and reports deadlocks and data hazards.  It mirrors
the loops of the kernel one to one: change both together.

cg = 1: one CTA.  cg = 2: a CTA pair — two producers (each fills its own half of every ring pair), the relay
(follower: full -> leader's peer_full), ONE issuer (the leader's) whose commits reach the barriers of both CTAs,
two epilogues whose acc_empty / act_ready arrivals all go to the leader's barriers (count 2 here: one per CTA).

cell = True: the per-cell mode.  The 16 epilogue warps of a CTA are two agents ("sets" of 8 warps): both take part in
layers 2-4; in layer 5 set s drains only the chunks of accumulator buffer s and arrives for two on acc_empty, and set 1
— which drains chunk 7 — stores the next tile's layer 1 for both.  A set does not look at the other buffer's acc_full
during layer 5 (four phases: its parity bookkeeping stays right), which is only safe because of the block-wide barrier
at the end of the tile: without it set 0 reaches its next wait on acc_full[1] (layer 2 of the next tile) while chunk 7
is still in flight, and the parity wait takes chunk 5's phase for layer 2's (found on the GPU as a launch failure).
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
    def __init__(self, count: int = 1):
        self.bit = 0
        self.count = count
        self.pending = 0

    def done(self, parity: int) -> bool:      # mbarrier.try_wait.parity
        return self.bit != parity

    def arrive(self, n: int = 1):
        self.pending += n
        assert self.pending <= self.count, "more arrivals than the barrier expects in one phase"
        if self.pending == self.count:
            self.pending = 0
            self.bit ^= 1

    complete = arrive                          # a barrier with count 1


def producer(n_tiles, ring_pairs, full, empty, inflight, who):
    pair = phase = 0
    for j in range(n_tiles * PAIRS_PER_TILE):
        yield ("wait", empty[pair], phase ^ 1)
        inflight.append((who, pair, j))        # expect_tx + cp.async.bulk
        yield ("step",)
        pair += 1
        if pair == ring_pairs:
            pair, phase = 0, phase ^ 1


def relay(n_tiles, ring_pairs, full, peer_full):
    pair = phase = 0
    for _ in range(n_tiles * PAIRS_PER_TILE):
        yield ("wait", full[pair], phase)
        peer_full[pair].arrive()               # mbarrier.arrive on the leader's barrier
        yield ("step",)
        pair += 1
        if pair == ring_pairs:
            pair, phase = 0, phase ^ 1


def issuer(n_tiles, ring_pairs, bars, pipe, state, trip, cg):
    pair = phase = 0
    act_phase = [0, 0, 0, 0]
    acc_parity = [0, 0, 0]
    stream = 0
    full = bars["full"][0]
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
                    yield ("wait", full[pair], phase)
                    if cg == 2:
                        yield ("wait", bars["peer_full"][pair], phase)
                    if two:
                        yield ("wait", full[pr1], ph1)
                        if cg == 2:
                            yield ("wait", bars["peer_full"][pr1], ph1)
                    # data checks at issue time, for every CTA the instruction touches
                    for r in range(cg):
                        assert state["ring"][r][pair] == stream, f"CTA {r}: ring pair {pair} holds {state['ring'][r][pair]}, wanted {stream}"
                        if two:
                            assert state["ring"][r][pr1] == stream + 1, f"CTA {r}: second ring pair holds the wrong weights"
                        assert state["acc_owner"][r][buf] in (None, (t, layer, c)), f"CTA {r}: accumulator {buf} overwritten before it was drained"
                        state["acc_owner"][r][buf] = (t, layer, c)
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


def epilogue(n_tiles, bars, state, r):
    """Epilogue of CTA r: waits on ITS acc_full / accx, arrives on the LEADER's acc_empty / act_ready."""
    full_phase = [0, 0, 0]
    accx_parity = 0
    acc_full, accx = bars["acc_full"][r], bars["accx"][r]
    bars["act_ready"][0].arrive()              # layer 1 of the first tile
    yield ("step",)
    for t in range(n_tiles):
        more = t + 1 < n_tiles
        for layer in range(3):
            for c in range(N_CHUNKS[layer]):
                buf = acc_buffer(layer, c)
                yield ("wait", acc_full[buf], full_phase[buf])
                full_phase[buf] ^= 1
                assert state["acc_done"][r][buf] == (t, layer, c), "epilogue drains an accumulator that holds another chunk"
                state["acc_owner"][r][buf] = None
                bars["acc_empty"][buf].arrive()
                if layer == 2 and c == 1:
                    accx.arrive()
                yield ("step",)
                if layer == 2 and c == 2:
                    yield ("wait", accx, accx_parity)
                    accx_parity ^= 1
                bars["act_ready"][c].arrive()
                yield ("step",)
        for c in range(8):
            buf = c & 1
            yield ("wait", acc_full[buf], full_phase[buf])
            full_phase[buf] ^= 1
            assert state["acc_done"][r][buf] == (t, 3, c)
            if c == 7 and more:
                bars["act_ready"][0].arrive()   # layer 1 of the next tile
                yield ("step",)
            state["acc_owner"][r][buf] = None
            bars["acc_empty"][buf].arrive()
            yield ("step",)


def epilogue_cell(n_tiles, bars, state, r, s, tile_barrier=True):
    """Cell mode: warp set s (0 or 1) of CTA r, see the module docstring."""
    full_phase = [0, 0, 0]
    accx_parity = tile_parity = 0
    acc_full, accx, tile_bar = bars["acc_full"][r], bars["accx"][r], bars["tile_bar"][r]
    bars["act_ready"][0].arrive()              # layer 1 of the first tile: every warp stores its own share
    yield ("step",)
    for t in range(n_tiles):
        more = t + 1 < n_tiles
        for layer in range(3):
            for c in range(N_CHUNKS[layer]):
                buf = acc_buffer(layer, c)
                yield ("wait", acc_full[buf], full_phase[buf])
                full_phase[buf] ^= 1
                assert state["acc_done"][r][buf] == (t, layer, c), "epilogue drains an accumulator that holds another chunk"
                bars["acc_empty"][buf].arrive()
                if s == 1:
                    state["acc_owner"][r][buf] = None   # (both sets have read their columns once the barrier completes)
                if layer == 2 and c == 1:
                    accx.arrive()
                yield ("step",)
                if layer == 2 and c == 2:
                    yield ("wait", accx, accx_parity)
                    accx_parity ^= 1
                bars["act_ready"][c].arrive()
                yield ("step",)
        for cc in range(4):
            c = 2 * cc + s                       # the chunks of accumulator buffer s
            yield ("wait", acc_full[s], full_phase[s])
            full_phase[s] ^= 1
            assert state["acc_done"][r][s] == (t, 3, c), "epilogue drains an accumulator that holds another chunk"
            if c == 7 and more:
                bars["act_ready"][0].arrive(2)   # layer 1 of the next tile, for both sets
                yield ("step",)
            state["acc_owner"][r][s] = None
            bars["acc_empty"][s].arrive(2)       # the draining set arrives for two
            yield ("step",)                      # ... and walks the runs
        if tile_barrier:                         # bar.sync over the 16 epilogue warps
            tile_bar.arrive()
            yield ("wait", tile_bar, tile_parity)
            tile_parity ^= 1


def make_bars(ring_pairs: int, cg: int, cell: bool = False):
    if cell:
        return {"full": [[Bar() for _ in range(ring_pairs)] for _ in range(cg)],
                "empty": [[Bar() for _ in range(ring_pairs)] for _ in range(cg)],
                "peer_full": [Bar() for _ in range(ring_pairs)],
                "acc_full": [[Bar() for _ in range(3)] for _ in range(cg)],
                "accx": [Bar(2) for _ in range(cg)],
                "tile_bar": [Bar(2) for _ in range(cg)],
                "acc_empty": [Bar(2 * cg) for _ in range(3)],
                "act_ready": [Bar(2 * cg) for _ in range(4)]}
    return {"full": [[Bar() for _ in range(ring_pairs)] for _ in range(cg)],
            "empty": [[Bar() for _ in range(ring_pairs)] for _ in range(cg)],
            "peer_full": [Bar() for _ in range(ring_pairs)],
            "acc_full": [[Bar() for _ in range(3)] for _ in range(cg)],
            "accx": [Bar() for _ in range(cg)],
            "acc_empty": [Bar(cg) for _ in range(3)],
            "act_ready": [Bar(cg) for _ in range(4)]}


def run(n_tiles: int, ring_pairs: int, seed: int, max_steps: int = 4_000_000, trip: int = 1, cg: int = 1, bars=None,
        cell: bool = False, tile_barrier: bool = True):
    """Returns 'ok' or a description of the failure.  n_tiles = tile slots per cluster."""
    rnd = random.Random(seed)
    bars = bars if bars is not None else make_bars(ring_pairs, cg, cell)
    inflight, pipe = [], deque()
    state = {"ring": [[None] * ring_pairs for _ in range(cg)], "acc_owner": [[None] * 3 for _ in range(cg)],
             "acc_done": [[None] * 3 for _ in range(cg)]}
    agents = {"issuer": issuer(n_tiles, ring_pairs, bars, pipe, state, trip, cg)}
    for r in range(cg):
        agents[f"producer{r}"] = producer(n_tiles, ring_pairs, bars["full"][r], bars["empty"][r], inflight, r)
        if cell:
            for st in range(2):
                agents[f"epilogue{r}.{st}"] = epilogue_cell(n_tiles, bars, state, r, st, tile_barrier)
        else:
            agents[f"epilogue{r}"] = epilogue(n_tiles, bars, state, r)
    if cg == 2:
        agents["relay"] = relay(n_tiles, ring_pairs, bars["full"][1], bars["peer_full"])
    pending = {name: next(g) for name, g in agents.items()}
    try:
        for _ in range(max_steps):
            choices = [name for name, req in pending.items() if req[0] == "step" or req[1].done(req[2])]
            if inflight:
                choices.append("land")
            if pipe:
                choices.append("retire")
            if not choices:
                return "ok" if not pending else f"deadlock: waiting {sorted(pending)}"
            pick = rnd.choice(choices)
            if pick == "land":
                who, pair, j = inflight.pop(rnd.randrange(len(inflight)))
                state["ring"][who][pair] = j
                bars["full"][who][pair].arrive()
            elif pick == "retire":
                _, pair, acc = pipe.popleft()
                for r in range(cg):                # tcgen05.commit (multicast for a pair) -> every CTA's barriers
                    bars["empty"][r][pair].arrive()
                    if acc is not None:
                        state["acc_done"][r][acc[0]] = acc[1]
                        bars["acc_full"][r][acc[0]].arrive()
            else:
                try:
                    pending[pick] = next(agents[pick])
                except StopIteration:
                    del pending[pick]
        return "step limit"
    except AssertionError as e:
        return f"hazard: {e}"
