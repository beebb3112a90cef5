"""Executable model of the synchronisation protocol of split_gemm_kernel (csrc/pointnet_mlp_split.cu) — test infrastructure.

Agents, as in the kernel: 16 producer warps (fp32 rows -> registers -> hi/lo stores into one of two x slots, one `full_x`
arrival per warp; the generic-proxy stores are complete when the warp arrives), the weight warp (one bulk copy per k block
into a ring of three, `full_w` by transaction bytes), the MMA issuer (waits `acc_empty` at a tile's first k block, `full_x`
and `full_w` per k block; its commits release `empty_x`, `empty_w` and publish `acc_full` when the tensor pipe retires the
twelve MMAs) and 4 epilogue warps.  Same scheduler and hazard checks as tests/conv_protocol_model.py: random and adversarial
schedules, asynchronous copy completion, in-order retirement.  It mirrors the kernel's loops one to one: change both together.
"""
from __future__ import annotations

import random
from collections import deque

from tests.conv_protocol_model import Bar

N_PRODUCERS, N_EPILOGUE, X_RING, W_RING = 16, 4, 2, 3


def make_bars():
    return {"full_x": [Bar(N_PRODUCERS) for _ in range(X_RING)], "empty_x": [Bar() for _ in range(X_RING)],
            "full_w": [Bar() for _ in range(W_RING)], "empty_w": [Bar() for _ in range(W_RING)],
            "acc_full": [Bar() for _ in range(2)], "acc_empty": [Bar(N_EPILOGUE) for _ in range(2)]}


def producer(w, total, bars, st):
    for it in range(total):
        yield ("step",)                                               # convert the rows loaded for this stage, prefetch the next
        slot = it % X_RING
        if it >= X_RING:
            yield ("wait", bars["empty_x"][slot], ((it // X_RING) - 1) & 1)
        if st["x_reading"][slot]:
            st["hazard"] = f"x slot {slot} rewritten (stage {it}) while MMAs still read it"
        st["x_written"][slot][w] = it                                  # st.shared + fence.proxy.async
        yield ("step",)
        bars["full_x"][slot].arrive()


def weight_warp(total, bars, st):
    for g in range(total):
        slot = g % W_RING
        if g >= W_RING:
            yield ("wait", bars["empty_w"][slot], ((g // W_RING) - 1) & 1)
        if st["w_reading"][slot]:
            st["hazard"] = f"weight slot {slot} refilled (stage {g}) while MMAs still read it"
        st["inflight"].append((slot, g))
        yield ("step",)


def issuer(n_tiles, nkb, bars, st):
    c = 0
    for t in range(n_tiles):
        buf = t & 1
        for kb in range(nkb):
            xs, ws = c % X_RING, c % W_RING
            if kb == 0 and t >= 2:
                yield ("wait", bars["acc_empty"][buf], ((t >> 1) - 1) & 1)
            yield ("wait", bars["full_x"][xs], (c // X_RING) & 1)
            yield ("wait", bars["full_w"][ws], (c // W_RING) & 1)
            if any(v != c for v in st["x_written"][xs]):
                st["hazard"] = f"MMAs of stage {c} issued on x data of stages {sorted(set(st['x_written'][xs]))}"
            if st["w_data"][ws] != c:
                st["hazard"] = f"MMAs of stage {c} issued on weights of stage {st['w_data'][ws]}"
            commits = [("empty_x", xs), ("empty_w", ws)] + ([("acc_full", buf)] if kb == nkb - 1 else [])
            st["x_reading"][xs] += 1
            st["w_reading"][ws] += 1
            st["pipe"].append({"xs": xs, "ws": ws, "commits": commits, "tile": t, "buf": buf, "first": kb == 0})
            c += 1
            yield ("step",)


def epilogue(q, n_tiles, bars, st):
    for t in range(n_tiles):
        buf = t & 1
        yield ("wait", bars["acc_full"][buf], (t >> 1) & 1)
        if st["acc_tile"][buf] != t:
            st["hazard"] = f"epilogue warp {q} read accumulator {buf} for tile {t} but it holds {st['acc_tile'][buf]}"
        yield ("step",)
        st["acc_unread"][buf] -= 1
        bars["acc_empty"][buf].arrive()
        yield ("step",)


def run(n_tiles: int, nkb: int, seed: int, bars=None, max_steps: int = 600_000, slow=(), slow_factor: int = 40) -> str:
    rng = random.Random(seed)
    bars = bars or make_bars()
    total = n_tiles * nkb
    st = {"inflight": [], "pipe": deque(), "hazard": None, "x_reading": [0] * X_RING, "w_reading": [0] * W_RING,
          "x_written": [[None] * N_PRODUCERS for _ in range(X_RING)], "w_data": [None] * W_RING, "acc_tile": [None, None],
          "acc_unread": [0, 0]}
    agents = [producer(w, total, bars, st) for w in range(N_PRODUCERS)]
    agents += [weight_warp(total, bars, st), issuer(n_tiles, nkb, bars, st)] + [epilogue(q, n_tiles, bars, st) for q in range(N_EPILOGUE)]
    kinds = ["producer"] * N_PRODUCERS + ["weights", "issuer"] + ["epilogue"] * N_EPILOGUE
    pending = [None] * len(agents)
    alive = set(range(len(agents)))
    for _ in range(max_steps):
        if st["hazard"]:
            return "hazard: " + st["hazard"]
        if not alive and not st["inflight"] and not st["pipe"]:
            return "ok"
        moves = [("agent", i) for i in alive
                 if pending[i] is None or pending[i][0] == "step" or (pending[i][0] == "wait" and pending[i][1].done(pending[i][2]))]
        if st["inflight"]:
            moves.append(("copy", None))
        if st["pipe"]:
            moves.append(("retire", None))
        if not moves:
            return f"deadlock: { {i: pending[i][2] for i in alive if pending[i] and pending[i][0] == 'wait'} }"
        weights = [1 if (kinds[i] if k == "agent" else k) in slow else slow_factor for k, i in moves]
        kind, i = rng.choices(moves, weights=weights)[0]
        if kind == "agent":
            try:
                pending[i] = next(agents[i])
            except StopIteration:
                alive.discard(i)
        elif kind == "copy":
            slot, g = st["inflight"].pop(rng.randrange(len(st["inflight"])))
            st["w_data"][slot] = g
            bars["full_w"][slot].arrive()
        else:
            m = st["pipe"].popleft()
            if m["first"]:
                if st["acc_unread"][m["buf"]]:
                    st["hazard"] = f"accumulator {m['buf']} overwritten by tile {m['tile']} before every epilogue warp read it"
                st["acc_tile"][m["buf"]] = None
            st["x_reading"][m["xs"]] -= 1
            st["w_reading"][m["ws"]] -= 1
            for name, idx in m["commits"]:
                if name == "acc_full":
                    st["acc_tile"][m["buf"]] = m["tile"]
                    st["acc_unread"][m["buf"]] = N_EPILOGUE
                bars[name][idx].arrive()
    return "timeout"


# ---------------------------------------------------------------------------------------------------------------------
# CTA pairs (split_gemm_kernel<FINAL, 2>): two CTAs, each with its own barriers, x slots and weight ring.  The leader (CTA 0)
# issues every MMA; an MMA reads the operands of BOTH CTAs and its commits arrive on the barriers of both.  The follower's
# sixteen producer warps arrive on the leader's `peer_x`, its relay warp forwards every weight stage's completion to the
# leader's `peer_w`, and the epilogue warps of both CTAs arrive on the leader's `acc_empty` (count 8).
# ---------------------------------------------------------------------------------------------------------------------
def make_pair_bars():
    cta = [make_bars() for _ in range(2)]
    cta[0]["acc_empty"] = [Bar(2 * N_EPILOGUE) for _ in range(2)]
    cta[0]["peer_x"] = [Bar(N_PRODUCERS) for _ in range(X_RING)]
    cta[0]["peer_w"] = [Bar() for _ in range(W_RING)]
    return cta


def pair_producer(r, w, total, cta, st):
    for it in range(total):
        yield ("step",)
        slot = it % X_RING
        if it >= X_RING:
            yield ("wait", cta[r]["empty_x"][slot], ((it // X_RING) - 1) & 1)
        if st["x_reading"][r][slot]:
            st["hazard"] = f"CTA {r}: x slot {slot} rewritten (stage {it}) while MMAs still read it"
        st["x_written"][r][slot][w] = it
        yield ("step",)
        (cta[0]["full_x"] if r == 0 else cta[0]["peer_x"])[slot].arrive()


def pair_weight_warp(r, total, cta, st):
    for g in range(total):
        slot = g % W_RING
        if g >= W_RING:
            yield ("wait", cta[r]["empty_w"][slot], ((g // W_RING) - 1) & 1)
        if st["w_reading"][r][slot]:
            st["hazard"] = f"CTA {r}: weight slot {slot} refilled (stage {g}) while MMAs still read it"
        st["inflight"].append((r, slot, g))
        yield ("step",)


def pair_relay(total, cta, st):
    for c in range(total):
        ws = c % W_RING
        yield ("wait", cta[1]["full_w"][ws], (c // W_RING) & 1)
        cta[0]["peer_w"][ws].arrive()
        yield ("step",)


def pair_issuer(n_tiles, nkb, cta, st):
    lead = cta[0]
    c = 0
    for t in range(n_tiles):
        buf = t & 1
        for kb in range(nkb):
            xs, ws = c % X_RING, c % W_RING
            if kb == 0 and t >= 2:
                yield ("wait", lead["acc_empty"][buf], ((t >> 1) - 1) & 1)
            yield ("wait", lead["full_x"][xs], (c // X_RING) & 1)
            yield ("wait", lead["full_w"][ws], (c // W_RING) & 1)
            yield ("wait", lead["peer_x"][xs], (c // X_RING) & 1)
            yield ("wait", lead["peer_w"][ws], (c // W_RING) & 1)
            for r in range(2):
                if any(v != c for v in st["x_written"][r][xs]):
                    st["hazard"] = f"MMAs of stage {c} issued on CTA {r}'s x data of stages {sorted(set(map(str, st['x_written'][r][xs])))}"
                if st["w_data"][r][ws] != c:
                    st["hazard"] = f"MMAs of stage {c} issued on CTA {r}'s weights of stage {st['w_data'][r][ws]}"
                st["x_reading"][r][xs] += 1
                st["w_reading"][r][ws] += 1
            commits = [("empty_x", xs), ("empty_w", ws)] + ([("acc_full", buf)] if kb == nkb - 1 else [])
            st["pipe"].append({"xs": xs, "ws": ws, "commits": commits, "tile": t, "buf": buf, "first": kb == 0})
            c += 1
            yield ("step",)


def pair_epilogue(r, q, n_tiles, cta, st):
    for t in range(n_tiles):
        buf = t & 1
        yield ("wait", cta[r]["acc_full"][buf], (t >> 1) & 1)
        if st["acc_tile"][buf] != t:
            st["hazard"] = f"CTA {r} epilogue warp {q} read accumulator {buf} for tile {t} but it holds {st['acc_tile'][buf]}"
        yield ("step",)
        st["acc_unread"][buf] -= 1
        cta[0]["acc_empty"][buf].arrive()
        yield ("step",)


def run_pair(n_tiles: int, nkb: int, seed: int, cta=None, max_steps: int = 1_200_000, slow=(), slow_factor: int = 40) -> str:
    """slow: 'epilogue', 'producer', 'weights', 'relay', 'issuer', 'copy', 'retire', or 'follower' (every agent of CTA 1)."""
    rng = random.Random(seed)
    cta = cta or make_pair_bars()
    total = n_tiles * nkb
    st = {"inflight": [], "pipe": deque(), "hazard": None, "x_reading": [[0] * X_RING for _ in range(2)],
          "w_reading": [[0] * W_RING for _ in range(2)],
          "x_written": [[[None] * N_PRODUCERS for _ in range(X_RING)] for _ in range(2)],
          "w_data": [[None] * W_RING for _ in range(2)], "acc_tile": [None, None], "acc_unread": [0, 0]}
    agents, kinds, owner = [], [], []
    for r in range(2):
        for w in range(N_PRODUCERS):
            agents.append(pair_producer(r, w, total, cta, st)); kinds.append("producer"); owner.append(r)
        agents.append(pair_weight_warp(r, total, cta, st)); kinds.append("weights"); owner.append(r)
        for q in range(N_EPILOGUE):
            agents.append(pair_epilogue(r, q, n_tiles, cta, st)); kinds.append("epilogue"); owner.append(r)
    agents.append(pair_issuer(n_tiles, nkb, cta, st)); kinds.append("issuer"); owner.append(0)
    agents.append(pair_relay(total, cta, st)); kinds.append("relay"); owner.append(1)
    pending = [None] * len(agents)
    alive = set(range(len(agents)))
    for _ in range(max_steps):
        if st["hazard"]:
            return "hazard: " + st["hazard"]
        if not alive and not st["inflight"] and not st["pipe"]:
            return "ok"
        moves = [("agent", i) for i in alive
                 if pending[i] is None or pending[i][0] == "step" or (pending[i][0] == "wait" and pending[i][1].done(pending[i][2]))]
        if st["inflight"]:
            moves.append(("copy", None))
        if st["pipe"]:
            moves.append(("retire", None))
        if not moves:
            return f"deadlock: { {i: (kinds[i], owner[i]) for i in alive} }"

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
            r, slot, g = st["inflight"].pop(rng.randrange(len(st["inflight"])))
            st["w_data"][r][slot] = g
            cta[r]["full_w"][slot].arrive()                            # a bulk copy signals a barrier of the CTA it writes to
        else:
            m = st["pipe"].popleft()
            if m["first"]:
                if st["acc_unread"][m["buf"]]:
                    st["hazard"] = f"accumulator {m['buf']} overwritten by tile {m['tile']} before every epilogue warp read it"
                st["acc_tile"][m["buf"]] = None
            for r in range(2):
                st["x_reading"][r][m["xs"]] -= 1
                st["w_reading"][r][m["ws"]] -= 1
            for name, idx in m["commits"]:
                if name == "acc_full":
                    st["acc_tile"][m["buf"]] = m["tile"]
                    st["acc_unread"][m["buf"]] = 2 * N_EPILOGUE
                for r in range(2):                                     # tcgen05.commit ... multicast::cluster
                    cta[r][name][idx].arrive()
    return "timeout"
