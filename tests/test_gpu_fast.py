"""GPU parity, FAST mode (the production path).  Integer stages are shared with STRICT mode and stay bit-exact
(tests/test_gpu_strict.py).  The floating-point stages use a different operation schedule (fused multiply-add,
twist-free transform), so accumulator coefficients are compared against the CPU oracle
  * per blind-rotation step on identical inputs, within a stated Torus64 tolerance, and
  * through decryptions and output noise statistics over whole bootstraps
(two different FFT roundings diverge completely after the first digit flip, SURVEY.md 8(c), so whole-bootstrap
coefficient comparison is only meaningful in STRICT mode)."""
import numpy as np
import pytest

from conftest import fresh_inputs, keyset, make_oracle
from mktfhe_b200.gate import PLAIN
from mktfhe_b200.scheme import MODE_FAST, MODE_STRICT

pytestmark = pytest.mark.gpu

# Torus64 max-norm tolerance for ONE external-product step on identical inputs.  Float64 carries 53 bits; a step
# sums 2*l*N products of magnitude <= 2^(logB-1) * 2^63, so rounding noise sits near 2^(63 + logB - 1 + 6 - 53):
# measured 2^30.5 between two Float64 schedules at KMS2party.
STEP_TOL = 2.0 ** 33
# Torus32 schemes: digits < 2^9 times 32-bit keys stay far below 2^53, so the transforms are exact up to the last unit; the
# only freedom is `native`'s truncation of a value that lands within rounding of an integer.
STEP_TOL32 = 4.0


def _signed_diff(a, b):
    if a.dtype == np.uint32:
        return (a - b).astype(np.int32).astype(np.float64)
    return (a.astype(np.uint64) - b.astype(np.uint64)).astype(np.int64).astype(np.float64)


@pytest.mark.parametrize("name", ["KMS2party", "CGGIparam"])
def test_cmux_step_within_tolerance(gpu_schemes, name):
    ks = keyset(name)
    orc = make_oracle(ks)
    s = gpu_schemes(name)
    s.set_mode(MODE_FAST)
    p = ks.params
    rng = np.random.default_rng(3)
    dt = s.torus_dtype
    tol = STEP_TOL if dt == np.uint64 else STEP_TOL32
    at = np.array([1, 2, 77, p.N - 1, p.N, p.N + 1, 2 * p.N - 1, 2 * p.N, 1234, 2 * p.N - 95], dtype=np.uint32)
    rows = rng.integers(0, np.iinfo(dt).max, size=(len(at), 2, p.N), dtype=dt)
    rows[0] = 0
    rows[0, 0, 0] = 1 << (p.torus_bits - 7)      # the trivial RLEV row phase 1 starts from
    worst = 0.0
    last = p.k - 1 if p.is_mk else 0
    for party, idx in ((0, 0), (last, 7), (last, p.n - 1)):
        out = s.cmux_step(party, idx, at, rows)
        for g in range(len(at)):
            ref = orc.cmux_step(party, idx, at[g], rows[g])
            d = np.abs(_signed_diff(out[g], ref)).max()
            worst = max(worst, d)
            assert d < tol, (party, idx, g, np.log2(d + 1))
        # a~ = 2N multiplies by X^2N - 1 = 0: the row must come back unchanged, exactly
        assert np.array_equal(out[7], rows[7])
    print(f"{name}: worst per-step |delta| = 2^{np.log2(worst + 1):.2f}")
    s.set_mode(MODE_STRICT)
    out = s.cmux_step(last, 7, at, rows)
    for g in range(len(at)):
        assert np.array_equal(out[g], orc.cmux_step(last, 7, at[g], rows[g]))
    s.set_mode(MODE_FAST)


@pytest.mark.parametrize("name", ["KMS2partyblock", "Blockparam"])
def test_block_step_within_tolerance(gpu_schemes, name):
    """KMS_block / LMSS: one block iteration (3 key bits folded into one accumulator pair in FAST mode) against the
    oracle's reference-order block step; STRICT is bit-exact."""
    ks = keyset(name)
    orc = make_oracle(ks)
    s = gpu_schemes(name)
    p = ks.params
    rng = np.random.default_rng(4)
    dt = s.torus_dtype
    tol = STEP_TOL if dt == np.uint64 else STEP_TOL32
    at = np.array([[1, 2, 3], [0, 77, 0], [p.N, 0, 2 * p.N], [2 * p.N - 1, 2 * p.N - 95, 17], [0, 0, 5], [2 * p.N, 0, 0]], dtype=np.uint32)
    rows = rng.integers(0, np.iinfo(dt).max, size=(len(at), 2, p.N), dtype=dt)
    last = p.k - 1 if p.is_mk else 0
    for mode in (MODE_FAST, MODE_STRICT):
        s.set_mode(mode)
        worst = 0.0
        for party, blk in ((0, 0), (last, 5), (last, p.d - 1)):
            out = s.block_step(party, blk, at, rows)
            for g in range(len(at)):
                ref = orc.block_step(party, blk, at[g], rows[g])
                if mode == MODE_STRICT:
                    assert np.array_equal(out[g], ref), (party, blk, g)
                else:
                    d = np.abs(_signed_diff(out[g], ref)).max()
                    worst = max(worst, d)
                    assert d < tol, (party, blk, g, np.log2(d + 1))
            assert np.array_equal(out[5], rows[5])            # only rotation is 2N: exact no-op
        if mode == MODE_FAST:
            print(f"{name}: worst per-block |delta| = 2^{np.log2(worst + 1):.2f}")
    s.set_mode(MODE_FAST)


@pytest.mark.parametrize("name", ["KMS2party", "KMS2partyblock", "CGGIparam", "Blockparam", "CCS2party"])
def test_fast_gates_decrypt_and_noise(gpu_schemes, name):
    """All six gates over a batch decrypt to the plaintext truth table in FAST mode, identically to STRICT and
    the oracle, and the output phase-error standard deviation matches STRICT mode's."""
    ks = keyset(name)
    s = gpu_schemes(name)
    B = 96
    b1, c1 = fresh_inputs(ks, B, seed=31)
    b2, c2 = fresh_inputs(ks, B, seed=32)
    errs = {}
    for mode in (MODE_FAST, MODE_STRICT):
        s.set_mode(mode)
        e = []
        for op in range(6):
            out = s.gate(op, c1, c2)
            want = np.array([PLAIN[op](bool(x), bool(y)) for x, y in zip(b1, b2)])
            assert np.array_equal(ks.decrypt_batch(out), want), (mode, op)
            for g in range(B):
                mu = (1 << 29) if want[g] else (7 << 29)
                d = (ks.phase(out[g]) - mu) & 0xFFFFFFFF
                e.append(d - (1 << 32) if d >= (1 << 31) else d)
        errs[mode] = np.array(e, dtype=np.float64)
    s.set_mode(MODE_FAST)
    sd_f, sd_s = errs[MODE_FAST].std(), errs[MODE_STRICT].std()
    print(f"phase-error std: FAST 2^{np.log2(sd_f):.2f}  STRICT 2^{np.log2(sd_s):.2f}  (decision margin 2^29)")
    assert 0.8 < sd_f / sd_s < 1.25
    assert np.abs(errs[MODE_FAST]).max() < 2 ** 29


@pytest.mark.parametrize("name", ["KMS2party"])
def test_fast_phase1_first_step_matches(gpu_schemes, name):
    """Phase 1 output in FAST mode is in the reference slot order and agrees with the oracle as long as the two
    trajectories have not diverged: use a ciphertext whose party blocks have exactly one non-zero a~."""
    ks = keyset(name)
    orc = make_oracle(ks)
    s = gpu_schemes(name)
    s.set_mode(MODE_FAST)
    p = ks.params
    ct = np.zeros(p.lwe_words, dtype=np.uint32)
    ct[0] = 0x12345678
    ct[1 + 5] = 0x9ABCDEF0            # party 0, index 5
    ct[1 + p.n + 9] = 0x0FEDCBA9      # party 1, index 9
    lev = s.phase1(ct[None])[0]
    tilde = orc.modswitch(ct)
    r0 = 0
    for party in range(p.k):
        ref = orc.phase1(party, tilde[1 + party * p.n: 1 + (party + 1) * p.n])
        got = lev[r0:r0 + ref.shape[0]]
        scale = np.abs(ref).max()
        assert np.abs(got - ref).max() / scale < 1e-9, party
        r0 += ref.shape[0]


@pytest.mark.parametrize("name", ["CGGIparam", "Blockparam", "CCS2party", "KMS2party", "KMS2partyblock"])
def test_tiled_keyswitch_bit_exact(gpu_schemes, name):
    """The production (gate-tiled) key switch is integer work: bit-exact against the oracle on random accumulators,
    including a ragged last tile (batch not a multiple of the tile) and batch = 1."""
    ks = keyset(name)
    orc = make_oracle(ks)
    s = gpu_schemes(name)
    s.set_mode(MODE_FAST)
    p = ks.params
    rng = np.random.default_rng(41)
    dt = s.torus_dtype
    for B in (1, 19):
        acc = rng.integers(0, np.iinfo(dt).max, size=(B, p.k + 1, p.N), dtype=dt)
        acc[0, 1, :8] = np.iinfo(dt).max                   # rounds up across the top digit
        acc[0, 1, 8:16] = 0
        out = s.keyswitch(acc)
        for g in range(B):
            assert np.array_equal(out[g], orc.keyswitch(acc[g])), (name, B, g)
