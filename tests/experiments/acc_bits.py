"""Round-2 groundwork (not a test; run by hand):  python tests/experiments/acc_bits.py [gates]

Question: how many low bits of the Torus64 RLWE accumulator does KMS phase 1 need?  The planned six-unit phase-1 kernel
 stores the accumulator in 48 or 40 bits to fit more units per SM.  This drives the CPU oracle's
phase-1 step (`orc_cmux_step`, bootstrapping.jl:413-438) row by row from Python, clears the low `drop` bits of both
polynomials after every step, finishes the bootstrap with the oracle's phase 2 and key switch, and reports decryptions
and the output phase-error standard deviation for each `drop`.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from conftest import fresh_inputs, keyset, make_oracle  # noqa: E402
from oracle import oracle as O  # noqa: E402


def bootstrap_with_truncated_acc(orc, ks, lin, drop):
    p = ks.params
    tilde = orc.modswitch(lin)
    mask = np.uint64((~((1 << drop) - 1)) & 0xFFFFFFFFFFFFFFFF)
    lev = np.zeros((p.k, p.l_lev, 2, p.H, 2), dtype=np.float64)      # orc_phase2 layout: party 0 uses row 0 only
    for party in range(p.k):
        rows = 1 if party == 0 else p.l_lev
        ta = tilde[1 + party * p.n: 1 + (party + 1) * p.n]
        for r in range(rows):
            acc = np.zeros((2, p.N), dtype=np.uint64)
            acc[0, 0] = np.uint64(1) << np.uint64(64 - (r + 1) * p.logB_lev)
            for idx in range(p.n):
                if ta[idx] == 0:
                    continue
                acc = orc.cmux_step(party, idx, int(ta[idx]), acc)
                if drop:
                    acc &= mask
            lev[party, r, 0] = O.fft(acc[0])
            lev[party, r, 1] = O.fft(acc[1])
    acc = orc.phase2(lev, int(tilde[0]))
    return orc.keyswitch(acc)


DROPS = tuple(int(x) for x in os.environ.get("ACC_DROPS", "0,16,24,27,32,36").split(","))


def main():
    gates = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    ks = keyset("KMS2party")
    orc = make_oracle(ks)
    b1, c1 = fresh_inputs(ks, gates, seed=501)
    b2, c2 = fresh_inputs(ks, gates, seed=502)
    want = np.array([not (x and y) for x, y in zip(b1, b2)])
    lins = [orc.gate_linear(0, c1[g], c2[g]) for g in range(gates)]
    ref = np.stack([orc.bootstrap(l.copy()) for l in lins])
    for drop in DROPS:
        outs = np.stack([bootstrap_with_truncated_acc(orc, ks, lins[g], drop) for g in range(gates)])
        if drop == 0:
            assert np.array_equal(outs, ref), "the Python-driven phase 1 must reproduce orc_bootstrap bit for bit"
        dec = ks.decrypt_batch(outs)
        err = []
        for g in range(gates):
            mu = (1 << 29) if want[g] else (7 << 29)
            d = (ks.phase(outs[g]) - mu) & 0xFFFFFFFF
            err.append(d - (1 << 32) if d >= (1 << 31) else d)
        err = np.array(err, dtype=np.float64)
        print(f"drop {drop:2d} low bits: {int((dec == want).sum())}/{gates} correct, phase-error std 2^{np.log2(err.std()):.2f}, "
              f"max 2^{np.log2(np.abs(err).max()):.2f}", flush=True)


if __name__ == "__main__":
    main()
