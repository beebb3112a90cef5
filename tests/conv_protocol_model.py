"""Executable model of the synchronisation protocol of conv3x3_tc_halo_kernel (csrc/conv_tc.cu) — test infrastructure.

Agents, as in the kernel: 8 pixel-producer warps (cp.async into one of two pixel blocks, one `full_blk` arrival per
warp), the weight warp (bulk copies into a ring of `ring` stages, `full_a` by transaction bytes), the MMA issuer (waits
`acc_empty` at a tile's first MMA, `full_blk` per 64-channel chunk, `full_a` per tap; its commits release `empty_a`,
`empty_blk`, `acc_full` when the tensor pipe retires the MMAs) and 4 epilogue warps (read one accumulator half, arrive on
`acc_empty`).  Waiters only see a phase PARITY, so the model replays the loops under random schedules with asynchronous
completions (copies land at arbitrary times, the pipe retires MMA groups in issue order at arbitrary times) and reports
deadlocks and hazards: a slot refilled while MMAs that read it are in flight, an MMA issued on a slot whose data has
not landed, an accumulator half overwritten before every epilogue warp has read it or read before its MMAs retired.
It mirrors the kernel's loops one to one: change both together.
"""
from __future__ import annotations

import random
from collections import deque

N_PRODUCERS = 8
N_EPILOGUE = 4
TAPS = 9


class Bar:
    def __init__(self, count: int = 1):
        self.bit, self.count, self.pending = 0, count, 0

    def done(self, parity: int) -> bool:       # mbarrier.try_wait.parity
        return self.bit != parity

    def arrive(self):
        self.pending += 1
        if self.pending == self.count:
            self.pending, self.bit = 0, self.bit ^ 1


def make_bars(ring: int):
    return {"full_a": [Bar() for _ in range(ring)], "empty_a": [Bar() for _ in range(ring)],
            "full_blk": [Bar(N_PRODUCERS) for _ in range(2)], "empty_blk": [Bar() for _ in range(2)],
            "acc_full": [Bar() for _ in range(2)], "acc_empty": [Bar(N_EPILOGUE) for _ in range(2)]}


def pixel_producer(w, n_tiles, ncc, bars, st):
    n_blocks = n_tiles * ncc
    for bi in range(n_blocks + 1):
        if bi >= 1:
            yield ("landed", ("blk", w, bi - 1))                       # cp.async.wait_group 0
            bars["full_blk"][(bi - 1) & 1].arrive()
            yield ("step",)
        if bi < n_blocks:
            buf = bi & 1
            if bi >= 2:
                yield ("wait", bars["empty_blk"][buf], ((bi >> 1) - 1) & 1)
            if st["blk_reading"][buf]:
                st["hazard"] = f"pixel block {buf} refilled (block {bi}) while MMAs still read it"
            st["inflight"].append(("blk", w, bi))                      # cp.async issued
            yield ("step",)


def weight_producer(n_tiles, ncc, ring, bars, st):
    for g in range(n_tiles * ncc * TAPS):
        slot = g % ring
        if g >= ring:
            yield ("wait", bars["empty_a"][slot], ((g // ring) - 1) & 1)
        if st["a_reading"][slot]:
            st["hazard"] = f"weight slot {slot} refilled (stage {g}) while MMAs still read it"
        st["inflight"].append(("a", slot, g))                          # expect_tx + bulk copy
        yield ("step",)


def issuer(n_tiles, ncc, ring, bars, st, seg=None):
    """seg: k chunks per accumulator chain; None = the whole tile in one chain (bf16 mode).  fp32-accuracy mode: tile t keeps
    its total in accumulator half t & 1 — the first chain accumulates straight into it — and every later chain goes to the
    other half, from where the epilogue adds it to the total; a half is reused when the epilogue has released it."""
    seg = seg or ncc
    g = bi = 0
    buf = 0
    uses = [0, 0]
    seq = (0, 0)
    for t in range(n_tiles):
        for cc in range(ncc):
            seg_first, seg_last = cc % seg == 0, (cc % seg == seg - 1 or cc == ncc - 1)
            if seg_first:
                buf = (t & 1) ^ (0 if cc == 0 else 1)
                seq = (t, cc // seg)
                k = uses[buf]
                uses[buf] += 1
                if k >= 1:
                    yield ("wait", bars["acc_empty"][buf], (k - 1) & 1)
            yield ("wait", bars["full_blk"][bi & 1], (bi >> 1) & 1)
            if st["blk_data"][bi & 1] != bi:
                st["hazard"] = f"MMAs of block {bi} issued on pixel data of block {st['blk_data'][bi & 1]}"
            for tap in range(TAPS):
                slot = g % ring
                yield ("wait", bars["full_a"][slot], (g // ring) & 1)
                if st["a_data"][slot] != g:
                    st["hazard"] = f"MMAs of stage {g} issued on weights of stage {st['a_data'][slot]}"
                commits = [("empty_a", slot)]
                if tap == TAPS - 1:
                    commits.append(("empty_blk", bi & 1))
                    if seg_last:
                        commits.append(("acc_full", buf))
                st["a_reading"][slot] += 1
                st["blk_reading"][bi & 1] += 1
                st["pipe"].append({"slot": slot, "blk": bi & 1, "commits": commits, "tile": seq, "buf": buf,
                                   "first": seg_first and tap == 0})
                g += 1
                yield ("step",)
            bi += 1


def epilogue(q, n_tiles, n_seg, bars, st):
    """Per tile: the total's half T = t & 1 completes with the first chain; every later chain lands in the other half P and is
    added to T (P is released after each addition); after the last one the total is stored and T is released."""
    uses = [0, 0]

    def wait_full(buf, what):
        k = uses[buf]
        uses[buf] += 1
        return ("wait", bars["acc_full"][buf], k & 1), what

    for t in range(n_tiles):
        T, P = t & 1, (t & 1) ^ 1
        req, what = wait_full(T, (t, 0))
        yield req
        if st["acc_tile"][T] != what:
            st["hazard"] = f"epilogue warp {q} took accumulator {T} for {what} but it holds {st['acc_tile'][T]}"
        for sg in range(1, n_seg):
            req, what = wait_full(P, (t, sg))
            yield req
            if st["acc_tile"][P] != what:
                st["hazard"] = f"epilogue warp {q} took accumulator {P} for {what} but it holds {st['acc_tile'][P]}"
            yield ("step",)                                            # tcgen05.ld of both halves, add, tcgen05.st of the total
            st["acc_unread"][P] -= 1
            bars["acc_empty"][P].arrive()
        yield ("step",)                                                # tcgen05.ld of the total + stores
        st["acc_unread"][T] -= 1
        bars["acc_empty"][T].arrive()
        yield ("step",)


def run(n_tiles: int, ncc: int, seed: int, ring: int = 4, bars=None, max_steps: int = 400_000, slow=(), slow_factor: int = 40,
        seg=None) -> str:
    """slow: agent kinds ('epilogue', 'producer', 'weights', 'issuer', 'copy', 'retire') scheduled `slow_factor` times less often.
    seg: k chunks per accumulator chain (the kernel's halo_segment_chunks: 3 in fp32-accuracy mode)."""
    n_seg = -(-ncc // (seg or ncc))
    rng = random.Random(seed)
    bars = bars or make_bars(ring)
    st = {"inflight": [], "pipe": deque(), "hazard": None, "a_reading": [0] * ring, "blk_reading": [0, 0],
          "a_data": [None] * ring, "blk_data": [None, None], "blk_landed": {}, "acc_tile": [None, None], "acc_unread": [0, 0],
          "landed": set()}
    agents = [pixel_producer(w, n_tiles, ncc, bars, st) for w in range(N_PRODUCERS)]
    agents += [weight_producer(n_tiles, ncc, ring, bars, st), issuer(n_tiles, ncc, ring, bars, st, seg)]
    agents += [epilogue(q, n_tiles, n_seg, bars, st) for q in range(N_EPILOGUE)]
    kinds = ["producer"] * N_PRODUCERS + ["weights", "issuer"] + ["epilogue"] * N_EPILOGUE
    pending = [None] * len(agents)
    alive = set(range(len(agents)))
    for _ in range(max_steps):
        if st["hazard"]:
            return "hazard: " + st["hazard"]
        if not alive and not st["inflight"] and not st["pipe"]:
            return "ok"
        moves = []
        for i in alive:
            p = pending[i]
            if p is None or p[0] == "step" or (p[0] == "wait" and p[1].done(p[2])) or (p[0] == "landed" and p[1] in st["landed"]):
                moves.append(("agent", i))
        if st["inflight"]:
            moves.append(("copy", None))
        if st["pipe"]:
            moves.append(("retire", None))
        if not moves:
            stuck = {i: pending[i] and (pending[i][0], pending[i][2] if pending[i][0] == "wait" else pending[i][1]) for i in alive}
            return f"deadlock: {stuck}"
        weights = [1 if (kinds[i] if k == "agent" else k) in slow else slow_factor for k, i in moves]
        kind, i = rng.choices(moves, weights=weights)[0]
        if kind == "agent":
            try:
                pending[i] = next(agents[i])
            except StopIteration:
                alive.discard(i)
        elif kind == "copy":                                           # one copy lands (any order)
            c = st["inflight"].pop(rng.randrange(len(st["inflight"])))
            if c[0] == "a":
                _, slot, g = c
                st["a_data"][slot] = g
                bars["full_a"][slot].arrive()
            else:
                _, w, bi = c
                st["landed"].add(c)
                got = st["blk_landed"].setdefault(bi, 0) + 1
                st["blk_landed"][bi] = got
                if got == N_PRODUCERS:
                    st["blk_data"][bi & 1] = bi
        else:                                                          # the pipe retires the oldest MMA group
            m = st["pipe"].popleft()
            if m["first"]:                                             # accumulate = 0: the half is overwritten from here on
                if st["acc_unread"][m["buf"]]:
                    st["hazard"] = f"accumulator {m['buf']} overwritten by tile {m['tile']} before every epilogue warp read it"
                st["acc_tile"][m["buf"]] = None
            st["a_reading"][m["slot"]] -= 1
            st["blk_reading"][m["blk"]] -= 1
            for name, idx in m["commits"]:
                if name == "acc_full":
                    st["acc_tile"][m["buf"]] = m["tile"]
                    st["acc_unread"][m["buf"]] = N_EPILOGUE
                bars[name][idx].arrive()
    return "timeout"


# ---------------------------------------------------------------------------------------------------------------------
# CTA pairs (conv3x3_tc_halo_kernel<2>; the same hand-overs as split_gemm_kernel<FINAL, 2>): two CTAs, each with its own
# barriers, pixel blocks and weight ring.  The leader (CTA 0) issues every MMA; an MMA reads the operands of BOTH CTAs and
# its commits arrive on the barriers of both.  The follower's producer warps arrive on the leader's `peer_blk`, its relay warp
# forwards the completion of its weight stages to the leader's `peer_a`, and the epilogue warps of both CTAs arrive on the
# leader's `acc_empty` (count 8).
# ---------------------------------------------------------------------------------------------------------------------
def make_pair_bars(ring: int):
    cta = [make_bars(ring) for _ in range(2)]
    cta[0]["acc_empty"] = [Bar(2 * N_EPILOGUE) for _ in range(2)]
    cta[0]["peer_a"] = [Bar() for _ in range(ring)]
    cta[0]["peer_blk"] = [Bar(N_PRODUCERS) for _ in range(2)]
    return cta


def pair_pixel_producer(r, w, n_tiles, ncc, cta, st):
    n_blocks = n_tiles * ncc
    for bi in range(n_blocks + 1):
        if bi >= 1:
            yield ("landed", ("blk", r, w, bi - 1))                    # cp.async.wait_group 0
            (cta[0]["full_blk"] if r == 0 else cta[0]["peer_blk"])[(bi - 1) & 1].arrive()
            yield ("step",)
        if bi < n_blocks:
            buf = bi & 1
            if bi >= 2:
                yield ("wait", cta[r]["empty_blk"][buf], ((bi >> 1) - 1) & 1)
            if st["blk_reading"][r][buf]:
                st["hazard"] = f"CTA {r}: pixel block {buf} refilled (block {bi}) while MMAs still read it"
            st["inflight"].append(("blk", r, w, bi))
            yield ("step",)


def pair_weight_producer(r, n_tiles, ncc, ring, cta, st):
    for g in range(n_tiles * ncc * TAPS):
        slot = g % ring
        if g >= ring:
            yield ("wait", cta[r]["empty_a"][slot], ((g // ring) - 1) & 1)
        if st["a_reading"][r][slot]:
            st["hazard"] = f"CTA {r}: weight slot {slot} refilled (stage {g}) while MMAs still read it"
        st["inflight"].append(("a", r, slot, g))
        yield ("step",)


def pair_relay(n_tiles, ncc, ring, cta, st):
    """The follower's MMA warp: forwards every weight stage's completion to the leader."""
    for g in range(n_tiles * ncc * TAPS):
        slot = g % ring
        yield ("wait", cta[1]["full_a"][slot], (g // ring) & 1)
        cta[0]["peer_a"][slot].arrive()
        yield ("step",)


def pair_issuer(n_tiles, ncc, ring, cta, st, seg=None):
    seg = seg or ncc
    lead = cta[0]
    g = bi = 0
    buf = 0
    uses = [0, 0]
    seq = (0, 0)
    for t in range(n_tiles):
        for cc in range(ncc):
            seg_first, seg_last = cc % seg == 0, (cc % seg == seg - 1 or cc == ncc - 1)
            if seg_first:
                buf = (t & 1) ^ (0 if cc == 0 else 1)
                seq = (t, cc // seg)
                k = uses[buf]
                uses[buf] += 1
                if k >= 1:
                    yield ("wait", lead["acc_empty"][buf], (k - 1) & 1)
            yield ("wait", lead["full_blk"][bi & 1], (bi >> 1) & 1)
            yield ("wait", lead["peer_blk"][bi & 1], (bi >> 1) & 1)
            for r in range(2):
                if st["blk_data"][r][bi & 1] != bi:
                    st["hazard"] = f"MMAs of block {bi} issued on CTA {r}'s pixel data of block {st['blk_data'][r][bi & 1]}"
            for tap in range(TAPS):
                slot = g % ring
                yield ("wait", lead["full_a"][slot], (g // ring) & 1)
                yield ("wait", lead["peer_a"][slot], (g // ring) & 1)
                for r in range(2):
                    if st["a_data"][r][slot] != g:
                        st["hazard"] = f"MMAs of stage {g} issued on CTA {r}'s weights of stage {st['a_data'][r][slot]}"
                commits = [("empty_a", slot)]
                if tap == TAPS - 1:
                    commits.append(("empty_blk", bi & 1))
                    if seg_last:
                        commits.append(("acc_full", buf))
                for r in range(2):
                    st["a_reading"][r][slot] += 1
                    st["blk_reading"][r][bi & 1] += 1
                st["pipe"].append({"slot": slot, "blk": bi & 1, "commits": commits, "tile": seq, "buf": buf,
                                   "first": seg_first and tap == 0})
                g += 1
                yield ("step",)
            bi += 1


def pair_epilogue(r, q, n_tiles, n_seg, cta, st):
    uses = [0, 0]

    def wait_full(buf, what):
        k = uses[buf]
        uses[buf] += 1
        return ("wait", cta[r]["acc_full"][buf], k & 1), what

    for t in range(n_tiles):
        T, P = t & 1, (t & 1) ^ 1
        req, what = wait_full(T, (t, 0))
        yield req
        if st["acc_tile"][T] != what:
            st["hazard"] = f"CTA {r} epilogue warp {q} took accumulator {T} for {what} but it holds {st['acc_tile'][T]}"
        for sg in range(1, n_seg):
            req, what = wait_full(P, (t, sg))
            yield req
            if st["acc_tile"][P] != what:
                st["hazard"] = f"CTA {r} epilogue warp {q} took accumulator {P} for {what} but it holds {st['acc_tile'][P]}"
            yield ("step",)
            st["acc_unread"][P] -= 1
            cta[0]["acc_empty"][P].arrive()                            # at the leader, whoever's warp this is
        yield ("step",)
        st["acc_unread"][T] -= 1
        cta[0]["acc_empty"][T].arrive()
        yield ("step",)


def run_pair(n_tiles: int, ncc: int, seed: int, ring: int = 4, cta=None, max_steps: int = 800_000, slow=(), slow_factor: int = 40,
             seg=None) -> str:
    """The pair protocol under a random schedule.  slow: 'epilogue', 'producer', 'weights', 'relay', 'issuer', 'copy', 'retire',
    or 'follower' (every agent of CTA 1)."""
    n_seg = -(-ncc // (seg or ncc))
    rng = random.Random(seed)
    cta = cta or make_pair_bars(ring)
    st = {"inflight": [], "pipe": deque(), "hazard": None, "a_reading": [[0] * ring, [0] * ring], "blk_reading": [[0, 0], [0, 0]],
          "a_data": [[None] * ring, [None] * ring], "blk_data": [[None, None], [None, None]], "blk_landed": {},
          "acc_tile": [None, None], "acc_unread": [0, 0], "landed": set()}
    agents, kinds, owner = [], [], []
    for r in range(2):
        for w in range(N_PRODUCERS):
            agents.append(pair_pixel_producer(r, w, n_tiles, ncc, cta, st)); kinds.append("producer"); owner.append(r)
        agents.append(pair_weight_producer(r, n_tiles, ncc, ring, cta, st)); kinds.append("weights"); owner.append(r)
        for q in range(N_EPILOGUE):
            agents.append(pair_epilogue(r, q, n_tiles, n_seg, cta, st)); kinds.append("epilogue"); owner.append(r)
    agents.append(pair_issuer(n_tiles, ncc, ring, cta, st, seg)); kinds.append("issuer"); owner.append(0)
    agents.append(pair_relay(n_tiles, ncc, ring, cta, st)); kinds.append("relay"); owner.append(1)
    pending = [None] * len(agents)
    alive = set(range(len(agents)))
    for _ in range(max_steps):
        if st["hazard"]:
            return "hazard: " + st["hazard"]
        if not alive and not st["inflight"] and not st["pipe"]:
            return "ok"
        moves = []
        for i in alive:
            p = pending[i]
            if p is None or p[0] == "step" or (p[0] == "wait" and p[1].done(p[2])) or (p[0] == "landed" and p[1] in st["landed"]):
                moves.append(("agent", i))
        if st["inflight"]:
            moves.append(("copy", None))
        if st["pipe"]:
            moves.append(("retire", None))
        if not moves:
            stuck = {i: pending[i] and pending[i][0] for i in alive}
            return f"deadlock: {stuck}"

        def is_slow(k, i):
            if k != "agent":
                return k in slow
            return kinds[i] in slow or ("follower" in slow and owner[i] == 1)

        weights = [1 if is_slow(k, i) else slow_factor for k, i in moves]
        kind, i = rng.choices(moves, weights=weights)[0]
        if kind == "agent":
            try:
                pending[i] = next(agents[i])
            except StopIteration:
                alive.discard(i)
        elif kind == "copy":
            c = st["inflight"].pop(rng.randrange(len(st["inflight"])))
            if c[0] == "a":
                _, r, slot, g = c
                st["a_data"][r][slot] = g
                cta[r]["full_a"][slot].arrive()                        # a bulk copy signals a barrier of the CTA it writes to
            else:
                _, r, w, bi = c
                st["landed"].add(c)
                got = st["blk_landed"].setdefault((r, bi), 0) + 1
                st["blk_landed"][(r, bi)] = got
                if got == N_PRODUCERS:
                    st["blk_data"][r][bi & 1] = bi
        else:
            m = st["pipe"].popleft()
            if m["first"]:
                if st["acc_unread"][m["buf"]]:
                    st["hazard"] = f"accumulator {m['buf']} overwritten by tile {m['tile']} before every epilogue warp read it"
                st["acc_tile"][m["buf"]] = None
            for r in range(2):
                st["a_reading"][r][m["slot"]] -= 1
                st["blk_reading"][r][m["blk"]] -= 1
            for name, idx in m["commits"]:
                if name == "acc_full":
                    st["acc_tile"][m["buf"]] = m["tile"]
                    st["acc_unread"][m["buf"]] = 2 * N_EPILOGUE
                for r in range(2):                                     # tcgen05.commit ... multicast::cluster: both CTAs
                    cta[r][name][idx].arrive()
    return "timeout"
