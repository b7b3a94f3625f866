"""The production key switch (k_keyswitch_tiled, csrc/keyswitch.cuh) restated in numpy and compared with the oracle's key switch
(the reference's loops, bootstrapping.jl:564-594 for KMS and :664-695 for KMS_block) on random accumulators -- bit for bit, since the
stage is integer arithmetic.  What the tiled kernel does differently from the reference's loop order and what this pins on the CPU:
sample extraction folded into the digit staging, the f * logD = 16-bit rounding kept unreduced in a uint16 (a carry out of the top
wraps, as in the reference's digit loop), unbalanced digits read as plain 2-bit fields (+ksk row d-1), balanced digits of the block
schemes as two's-complement 2-bit fields with the carry chain (1: +row 0, 3 = -1: -row 0, 2 = -2: -row 1), the first n coefficients
of a block scheme copied instead of switched, and per-party partial sums added into one output.  The GPU leg of the same comparison
is tests/test_gpu_strict.py::test_tiled_keyswitch_bit_exact.  CPU only."""
import numpy as np
import pytest

from conftest import keyset, make_oracle


def _divbits(a, bit):
    a = a.astype(np.uint32)
    carry = (a << np.uint32(32 - bit)) >> np.uint32(31)
    return ((a >> np.uint32(bit)) + carry).astype(np.uint32)


def _tiled_keyswitch(p, ksk, acc):
    N, n, k, f = p.N, p.n, p.k, p.f
    block = p.scheme in (1, 4)
    bits64 = acc.dtype == np.uint64
    A = (acc >> np.uint64(32)).astype(np.uint32) if bits64 else acc.astype(np.uint32)       # [(k+1)][N]
    out = np.zeros(1 + n * k, dtype=np.uint32)
    out[0] = A[0, 0]                                                         # res.b starts from acc.b[0]
    lv = np.arange(f)
    for party in range(k):
        a = A[1 + party]
        ext = np.empty(N, dtype=np.uint32)                                   # sample extraction: a_0, -a_{N-1}, ..., -a_1
        ext[0] = a[0]
        ext[1:] = (np.uint32(0) - a[:0:-1]).astype(np.uint32)
        c0 = n if block else 0
        ai = _divbits(ext[c0:], 32 - f * p.logD)
        if not block:
            packed = ai.astype(np.uint16)
        else:
            acc_bits = np.zeros_like(ai)
            for l_ in range(f - 1, 0, -1):
                d = ai & np.uint32(3)
                ai = (ai >> np.uint32(2)) + (d >> np.uint32(1))
                acc_bits |= d << np.uint32(2 * (f - 1 - l_))
            acc_bits |= (ai & np.uint32(3)) << np.uint32(2 * (f - 1))
            packed = acc_bits.astype(np.uint16)
        d = (packed[:, None].astype(np.uint32) >> (2 * (f - 1 - lv))[None, :].astype(np.uint32)) & np.uint32(3)      # [c][level]
        K = ksk[party]                                                        # [N][Dk][f][1 + n]
        total = np.zeros(n + 1, dtype=np.uint32)
        cs, ls = np.nonzero(d)
        dv = d[cs, ls]
        if not block:
            rows = K[cs + c0, dv - 1, ls]
            total = np.add.reduce(rows, axis=0, dtype=np.uint32)
        else:
            plus = dv == 1
            rows_p = K[cs[plus] + c0, 0, ls[plus]]
            rows_m = K[cs[~plus] + c0, np.where(dv[~plus] == 2, 1, 0), ls[~plus]]
            total = (np.add.reduce(rows_p, axis=0, dtype=np.uint32) - np.add.reduce(rows_m, axis=0, dtype=np.uint32)).astype(np.uint32)
        out[0] = np.uint32((int(out[0]) + int(total[0])) & 0xFFFFFFFF)
        seg = total[1:].copy()
        if block:
            seg += ext[:n]                                                    # the first n coefficients are copied, not switched
        out[1 + party * n: 1 + (party + 1) * n] = seg
    return out


@pytest.mark.parametrize("name", ["KMS2party", "KMS2partyblock", "CGGIparam", "Blockparam", "CCS2party"])
def test_tiled_key_switch_algorithm_equals_the_reference_loops(name):
    ks = keyset(name)
    p = ks.params
    orc = make_oracle(ks)
    rng = np.random.default_rng(len(name))
    dt = np.uint64 if p.N == 2048 else np.uint32
    nparties = p.k if p.is_mk else 1
    for trial in range(3):
        acc = rng.integers(0, 2 ** (64 if dt == np.uint64 else 32), size=(nparties + 1, p.N), dtype=dt)
        if trial == 1:                                                        # rounding carries out of the top field, zeros, extremes
            acc[1, :8] = np.array([0, 1, 2 ** 31, 2 ** 31 - 1, 2 ** 32 - 1, 2 ** 32 - 2 ** 15, 2 ** 32 - 2 ** 15 - 1, 2 ** 15], dtype=np.uint64).astype(dt) \
                << (dt(32) if dt == np.uint64 else dt(0))
        want = np.asarray(orc.keyswitch(acc)).reshape(-1)
        got = _tiled_keyswitch(p, ks.ksk, acc)
        assert np.array_equal(got, want), (name, trial, int(np.sum(got != want)))
