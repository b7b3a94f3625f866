"""Explicit-state model of the bootstrapping-key ring of fastw::k_phase1_w (csrc/kernels_fast_w.cuh): one producer thread issuing
cp.async.bulk copies into shared-memory slots guarded by full / empty mbarriers, consumed by one warp per unit and kind (A: even
tiles, B: odd tiles).  Every interleaving of the producer, the consumer warps and the COMPLETIONS of the outstanding copies (bulk
copies complete in any order) is explored breadth first.

What is modelled exactly as in the kernel:
  * an mbarrier wait is by phase PARITY: wait(P) succeeds iff the barrier's completed-phase count is odd for P = 0, even (and > 0 or
    not) for P = 1 -- i.e. iff (completed & 1) != P.  A waiter cannot tell "my phase" from "two phases earlier";
  * full[slot]: one arrival (the producer's expect_tx) + the copy's bytes: the phase completes when the copy completes;
  * empty[slot]: one arrival per unit (the elected lane of the consuming warp);
  * the producer waits for empty[slot] only from its second lap on, with the parity of the previous lap;
  * a consumer walks its tiles in order: wait full[slot] -> read the tile -> arrive on empty[slot].
Coupling of the two warps of a unit: `coupled=True` applies the token order of the sweeps (a warp may start its i-th tile only after
the other warp of its unit has finished i tiles); `coupled=False` is what dead units of a small batch and skipped steps (a~ = 0) do:
they only keep the ring moving, nothing holds a warp back.

Checked: a consumer never reads a slot that does not hold ITS tile (stale or half-written data), a copy is never issued into a slot
that is being read or already has a copy in flight, no barrier gets more arrivals than its count, and every run ends with all tiles
consumed (no deadlock).

Ring layouts:  ("one", D)      one ring of D slots shared by both kinds, tile n in slot n mod D  (round-2 build before the fix: D = 5)
               ("two", RA, RB) one ring per kind (shipped: 3 + 2)
               ("all", D)      one ring of D slots, EVERY consumer reads EVERY tile (fast::k_phase1_tma, fast32::k_rgsw_tma, the
                               tiled key switch): every waiter sees every phase, so any depth works

    python tools/models/key_ring_model.py          # prints the verdict for the layouts discussed in profiles/README_r2.md
"""
from __future__ import annotations

from collections import deque


def _place(layout, n):
    """tile n -> (slot, lap of that slot's ring, ring depth).  Kind = n & 1."""
    if layout[0] in ("one", "all"):
        d = layout[1]
        return n % d, n // d, d
    ra, rb = layout[1], layout[2]
    i = n >> 1                               # index of the tile within its kind
    if n & 1:
        return ra + i % rb, i // rb, rb
    return i % ra, i // ra, ra


def _nslots(layout):
    return layout[1] if layout[0] in ("one", "all") else layout[1] + layout[2]


def _first_lap(layout, n):
    """The producer's `if (n >= ...)` test: no wait during the first lap of the ring the tile belongs to."""
    if layout[0] in ("one", "all"):
        return n < layout[1]
    return (n >> 1) < (layout[2] if n & 1 else layout[1])


def check(layout, units=2, ntiles=24, coupled=False, max_states=4_000_000):
    """Returns ("ok", states) or (reason, trace_length)."""
    ns = _nslots(layout)
    every = layout[0] == "all"                 # one kind of consumer that reads every tile
    nk = 1 if every else 2
    tile = (lambda k, i: i) if every else (lambda k, i: 2 * i + k)
    # state = (prod_n, cons, slots); cons[kind * units + u] = (own index i, reading 0/1)
    # slots[s] = (full_done, empty_done, empty_cnt, content, inflight)   content: tile id, -1 nothing, -2 being written
    init = (0, tuple((0, 0) for _ in range(nk * units)), tuple((0, 0, 0, -1, -1) for _ in range(ns)))
    per_kind = [ntiles] if every else [(ntiles + 1) // 2, ntiles // 2]
    seen = {init}
    todo = deque([init])
    while todo:
        st = todo.popleft()
        prod_n, cons, slots = st
        succ = []
        # ---- producer
        if prod_n < ntiles:
            s, lap, _ = _place(layout, prod_n)
            fd, ed, ec, content, infl = slots[s]
            ok = True
            if not _first_lap(layout, prod_n):
                par = (lap & 1) ^ 1                               # initial 1, flipped at every wrap
                ok = (ed & 1) != par
            if ok:
                if infl != -1:
                    return "two copies in flight to one slot", len(seen)
                for k in range(nk):
                    for u in range(units):
                        i, rd = cons[k * units + u]
                        if rd and _place(layout, tile(k, i))[0] == s:
                            return "copy issued into a slot that is being read", len(seen)
                ns_ = list(slots)
                ns_[s] = (fd, ed, ec, -2, prod_n)
                succ.append((prod_n + 1, cons, tuple(ns_)))
        # ---- completions, in any order
        for s in range(ns):
            fd, ed, ec, content, infl = slots[s]
            if infl != -1:
                ns_ = list(slots)
                ns_[s] = (fd + 1, ed, ec, infl, -1)
                succ.append((prod_n, cons, tuple(ns_)))
        # ---- consumers
        for k in range(nk):
            for u in range(units):
                i, rd = cons[k * units + u]
                if i >= per_kind[k]:
                    continue
                n = tile(k, i)
                s, lap, _ = _place(layout, n)
                fd, ed, ec, content, infl = slots[s]
                if not rd:
                    if coupled and not every and cons[(1 - k) * units + u][0] < min(i, per_kind[1 - k]):
                        continue                                  # token: the other warp of my unit has not finished i tiles yet
                    if (fd & 1) != (lap & 1):                      # consumer parity starts at 0 and flips every lap
                        if content != n:
                            return f"consumer of kind {'AB'[k]} took tile {content} in slot {s} for its tile {n}", len(seen)
                        nc = list(cons)
                        nc[k * units + u] = (i, 1)
                        succ.append((prod_n, tuple(nc), slots))
                else:
                    if content != n:
                        return f"slot {s} overwritten while kind {'AB'[k]} read tile {n}", len(seen)
                    ec2, ed2 = ec + 1, ed
                    if ec2 > units:
                        return "too many arrivals on an empty barrier", len(seen)
                    if ec2 == units:
                        ec2, ed2 = 0, ed + 1
                    ns_ = list(slots)
                    ns_[s] = (fd, ed2, ec2, content, infl)
                    nc = list(cons)
                    nc[k * units + u] = (i + 1, 0)
                    succ.append((prod_n, tuple(nc), tuple(ns_)))
        if not succ:
            done = prod_n == ntiles and all(cons[k * units + u][0] >= per_kind[k] for k in range(nk) for u in range(units))
            if not done:
                return "deadlock", len(seen)
        for nx in succ:
            if nx not in seen:
                seen.add(nx)
                todo.append(nx)
        if len(seen) > max_states:
            return "state limit", len(seen)
    return "ok", len(seen)


if __name__ == "__main__":
    for layout in (("one", 5), ("one", 4), ("two", 3, 2), ("two", 2, 2), ("two", 3, 1), ("all", 5), ("all", 10)):
        for coupled in (True, False):
            res, n = check(layout, units=2, ntiles=24, coupled=coupled)
            print(f"{layout!s:18} {'token-coupled' if coupled else 'free-running '}  {res}  ({n} states)")
