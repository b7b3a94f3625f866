"""Explicit-state model of the four-token hand-off between the two warps of a unit in fastw::k_phase1_w and fastw32::k_cggi_w
(csrc/kernels_fast_w.cuh, "own sum first, the other warp's sum second").

Warp A owns the .b half (sum 0), warp B the .a half (sum 1).  Per gadget digit j a warp adds its spectrum into its OWN sum first and
into the OTHER warp's sum second; at the end of a step it waits for the other warp's last addition into its own sum, reads and
clears that sum (inverse transform, accumulator update) and starts the next step.  Tokens are mbarriers with ONE arrival per phase;
the waiter keeps a parity bit (wait(P) succeeds iff (completed & 1) != P):

    token 0  b: A -> B      A's first-pass addition of digit j into .b is done     (B's second pass of digit j waits)
    token 1  b: B -> A      B's second-pass addition of digit j into .b is done    (A's first pass of digit j + 1, A's step end wait)
    token 2  a: B -> A      B's first-pass addition of digit j into .a is done     (A's second pass of digit j waits)
    token 3  a: A -> B      A's second-pass addition of digit j into .a is done    (B's first pass of digit j + 1, B's step end wait)

Every interleaving of the two warps is explored.  Checked:
  * the read-modify-write sweeps of the two warps on one sum and the owner's read-and-clear never overlap;
  * the additions arrive in the fixed order .b: A0, B0, A1, B1, ...  and  .a: B0, A0, B1, A1, ...  (results are deterministic);
  * the owner's read at the end of a step sees exactly the 2l additions of that step;
  * no token is passed a second time before its previous phase was consumed (a lapped parity wait would be ambiguous);
  * no deadlock.  Steps with a~ = 0 are skipped by both warps without touching the tokens (`skip` marks them).

    python tools/models/token_model.py
"""
from __future__ import annotations

from collections import deque

WAIT, PASS, RMW_BEGIN, RMW_END, READ_BEGIN, READ_END = range(6)


def program(w, l, skip):
    """Instruction list of warp w (0 = A, 1 = B) over the steps in `skip` (True = a~ is zero, nothing happens)."""
    own, oth = w, 1 - w
    wait_own, wait_oth = (1, 2) if w == 0 else (3, 0)
    pass_own, pass_oth = (0, 3) if w == 0 else (2, 1)
    ops = []
    for sk in skip:
        if sk:
            continue
        for j in range(l):
            if j > 0:
                ops.append((WAIT, wait_own))
            ops += [(RMW_BEGIN, own, j), (RMW_END, own, j), (PASS, pass_own)]
            ops += [(WAIT, wait_oth), (RMW_BEGIN, oth, j), (RMW_END, oth, j), (PASS, pass_oth)]
        ops += [(WAIT, wait_own), (READ_BEGIN, own), (READ_END, own)]
    return ops


def check(l=3, skip=(False, False, False), drop_wait=None):
    """drop_wait = (warp, ordinal): removes that warp's n-th WAIT -- a seeded defect the checker must find."""
    progs = [program(0, l, skip), program(1, l, skip)]
    if drop_wait is not None:
        w, nth = drop_wait
        idx = [i for i, op in enumerate(progs[w]) if op[0] == WAIT][nth]
        progs[w] = progs[w][:idx] + progs[w][idx + 1:]
    # parity of a warp's k-th wait on a token: k & 1 (starts at 0, flips after each wait)
    nwaits = [[[0] * (len(p) + 1) for _ in range(4)] for p in progs]
    for w, p in enumerate(progs):
        cnt = [0, 0, 0, 0]
        for i, op in enumerate(p):
            for t in range(4):
                nwaits[w][t][i] = cnt[t]
            if op[0] == WAIT:
                cnt[op[1]] += 1
        for t in range(4):
            nwaits[w][t][len(p)] = cnt[t]
    waiter = {0: 1, 1: 0, 2: 0, 3: 1}                      # who waits on each token
    # state: (pcA, pcB, tokens done x4, busy x2 (who holds the sum, -1 none), adds since clear x2)
    init = (0, 0, (0, 0, 0, 0), (-1, -1), (0, 0))
    seen = {init}
    todo = deque([init])
    while todo:
        st = todo.popleft()
        pcs = [st[0], st[1]]
        tok, busy, adds = st[2], st[3], st[4]
        succ = []
        for w in range(2):
            pc = pcs[w]
            if pc >= len(progs[w]):
                continue
            op = progs[w][pc]
            ntok, nbusy, nadds = list(tok), list(busy), list(adds)
            if op[0] == WAIT:
                t = op[1]
                par = nwaits[w][t][pc] & 1
                if (tok[t] & 1) == par:
                    continue                                           # blocked
            elif op[0] == PASS:
                t = op[1]
                consumed = nwaits[waiter[t]][t][pcs[waiter[t]]]
                if tok[t] - consumed >= 1:
                    return f"token {t} passed again before its previous phase was consumed"
                ntok[t] += 1
            elif op[0] == RMW_BEGIN:
                s, j = op[1], op[2]
                if busy[s] != -1:
                    return f"warp {'AB'[w]} enters sum {'ba'[s]} while warp {'AB'[busy[s]]} is in it"
                first = s                                              # the owner adds first: A into .b, B into .a
                expect = 2 * j + (0 if w == first else 1)
                if adds[s] != expect:
                    return f"sum {'ba'[s]}: addition {adds[s]} comes from warp {'AB'[w]} digit {j}, expected position {expect}"
                nbusy[s] = w
            elif op[0] == RMW_END:
                s = op[1]
                nbusy[s] = -1
                nadds[s] += 1
            elif op[0] == READ_BEGIN:
                s = op[1]
                if busy[s] != -1:
                    return f"owner reads sum {'ba'[s]} while warp {'AB'[busy[s]]} is adding"
                if adds[s] != 2 * l:
                    return f"owner reads sum {'ba'[s]} after {adds[s]} of {2 * l} additions"
                nbusy[s] = w
            elif op[0] == READ_END:
                s = op[1]
                nbusy[s] = -1
                nadds[s] = 0
            npc = list(pcs)
            npc[w] += 1
            succ.append((npc[0], npc[1], tuple(ntok), tuple(nbusy), tuple(nadds)))
        if not succ and not (pcs[0] == len(progs[0]) and pcs[1] == len(progs[1])):
            return f"deadlock at A:{pcs[0]} B:{pcs[1]}"
        for nx in succ:
            if nx not in seen:
                seen.add(nx)
                todo.append(nx)
    return "ok"


if __name__ == "__main__":
    for l in (1, 2, 3, 6):
        print("l =", l, check(l, (False, False, False)), "| with a skipped step:", check(l, (False, True, False, False)))
    print("seeded defects:", check(3, (False, False), drop_wait=(0, 1)), "|", check(3, (False, False), drop_wait=(1, 3)))
