"""Edge cases and error behaviour of the C-ABI on the device: empty and ragged batches, batch sizes that straddle
the kernels' internal tile boundaries, batch invariance, and the state / argument errors a caller can provoke.

The reference evaluates one gate at a time (gate.jl:1-52), so "batch of B" must mean exactly "B independent
reference calls": every gate's result may depend on nothing but its own operands.
"""
import numpy as np
import pytest

from conftest import fresh_inputs, keyset, make_oracle
from mktfhe_b200 import params as P
from mktfhe_b200.gate import PLAIN
from mktfhe_b200.scheme import MODE_FAST, MODE_STRICT, MktfheError, Scheme

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["KMS2party", "CGGIparam", "CCS2party"])
def test_empty_batch_is_a_no_op(gpu_schemes, name):
    ks = keyset(name)
    s = gpu_schemes(name)
    w = 1 + ks.params.n * ks.params.k
    empty = np.zeros((0, w), dtype=np.uint32)
    for mode in (MODE_FAST, MODE_STRICT):
        s.set_mode(mode)
        assert s.gate(0, empty, empty).shape == (0, w)
        assert s.bootstrapping(empty).shape == (0, w)
    s.set_mode(MODE_FAST)


# Internal tiles: phase 1 groups 4 (gate, party, row) units per CTA by party; the key switch tiles 16 gates;
# fast32 runs 8 (CGGI / LMSS) or 4 (CCS) gates per CTA.  Sizes below sit on, just under and just over those.
@pytest.mark.parametrize("name", ["KMS2party", "KMS2partyblock", "CGGIparam", "Blockparam", "CCS2party"])
def test_ragged_batches_are_batch_invariant(gpu_schemes, name):
    """STRICT: every ragged batch is a bit-identical prefix of the big one.  FAST: the same holds (each gate's
    arithmetic is a function of its own operands only), and everything decrypts."""
    ks = keyset(name)
    s = gpu_schemes(name)
    B = 37
    b1, c1 = fresh_inputs(ks, B, seed=71)
    b2, c2 = fresh_inputs(ks, B, seed=72)
    want = np.array([PLAIN[0](bool(x), bool(y)) for x, y in zip(b1, b2)])
    for mode in (MODE_STRICT, MODE_FAST):
        s.set_mode(mode)
        full = s.gate(0, c1, c2)
        assert np.array_equal(ks.decrypt_batch(full), want)
        for n in (1, 2, 3, 5, 7, 8, 9, 15, 16, 17, 33):
            part = s.gate(0, c1[:n], c2[:n])
            assert np.array_equal(part, full[:n]), (mode, n)
        # a single ciphertext (1-D input) is the batch-of-one case
        one = s.gate(0, c1[4], c2[4])
        assert one.shape == c1[4].shape and np.array_equal(one, full[4])
    s.set_mode(MODE_FAST)


@pytest.mark.parametrize("name", ["KMS2party", "CGGIparam"])
def test_strict_single_gate_equals_oracle_for_every_opcode(gpu_schemes, name):
    """One gate at a time, like the reference's API, for each of the six opcodes plus the bare bootstrap."""
    ks = keyset(name)
    s = gpu_schemes(name)
    orc = make_oracle(ks)
    _, c1 = fresh_inputs(ks, 1, seed=81)
    _, c2 = fresh_inputs(ks, 1, seed=82)
    s.set_mode(MODE_STRICT)
    for op in range(6):
        got = s.gate(op, c1[0], c2[0])
        ref = orc.gate_batch(op, c1, c2)[0]
        assert np.array_equal(got, ref), op
    assert np.array_equal(s.bootstrapping(c1[0]), orc.bootstrap(c1[0]))
    s.set_mode(MODE_FAST)


def test_extreme_ciphertext_values(gpu_schemes):
    """All-zero and all-ones masks, and bodies at the rounding boundaries of the modulus switch: STRICT stays
    bit-identical to the oracle (these hit a~ = 0 and a~ = 2N, the reference's unreduced divbits corner,
    bootstrapping.jl:8-9), and FAST produces well-formed ciphertexts for them."""
    name = "KMS2party"
    ks = keyset(name)
    s = gpu_schemes(name)
    orc = make_oracle(ks)
    p = ks.params
    w = 1 + p.n * p.k
    cts = np.zeros((6, w), dtype=np.uint32)
    cts[1, :] = 0xFFFFFFFF
    cts[2, 1:] = 0xFFFFFFFF                      # every a rounds up to 2N
    cts[3, 0] = 0x80000000
    step = 1 << (32 - 12)                        # 2N = 4096 buckets
    cts[4, :] = step // 2                        # exactly on a rounding boundary
    cts[5, :] = step // 2 - 1
    s.set_mode(MODE_STRICT)
    got = s.bootstrapping(cts)
    for g in range(cts.shape[0]):
        assert np.array_equal(got[g], orc.bootstrap(cts[g])), g
    s.set_mode(MODE_FAST)
    fast = s.bootstrapping(cts)
    # Every output is a well-formed encryption of +-1/8, whatever the input was ...
    for out in (got, fast):
        for g in range(cts.shape[0]):
            ph = int(ks.phase(out[g])) & 0xFFFFFFFF
            dist = min((ph - mu) & 0xFFFFFFFF if ((ph - mu) & 0xFFFFFFFF) < (1 << 31) else (mu - ph) & 0xFFFFFFFF
                       for mu in (1 << 29, 7 << 29))
            assert dist < (1 << 28), (g, hex(ph))
    # ... and FAST decides like STRICT wherever the input phase is away from the decision boundary (rows 4, 5:
    # phase ~ -2^28; rows 0-3 sit on the boundary itself, where the sign is noise).
    for g in (4, 5):
        assert ks.lwe_decrypt(fast[g]) == ks.lwe_decrypt(got[g]), g


def test_state_and_argument_errors(gpu_schemes):
    ks = keyset("KMS2party")
    p = ks.params
    w = 1 + p.n * p.k
    c = np.zeros((2, w), dtype=np.uint32)

    fresh = Scheme(p, device=0)
    try:
        with pytest.raises(MktfheError, match="not finalized"):
            fresh.gate(0, c, c)
        with pytest.raises(MktfheError, match="missing"):
            fresh.finalize()
        with pytest.raises(MktfheError, match="party"):
            fresh.upload_party(p.k, ks.brk[0], ks.ksk[0], ks.rlk[0], ks.pubb[0])
        with pytest.raises(MktfheError):
            fresh.upload_party(0, ks.brk[0], ks.ksk[0])          # KMS needs rlk and pubb
    finally:
        fresh.close()

    s = gpu_schemes("KMS2party")
    with pytest.raises(MktfheError, match="opcode"):
        s.gate(6, c, c)
    with pytest.raises(MktfheError):
        s.set_mode(7)
    with pytest.raises(ValueError):
        s.gate(0, c, c[:1])
    with pytest.raises(MktfheError):
        s.cmux_step(p.k, 0, np.zeros(1, np.uint32), np.zeros((1, 2, p.N), np.uint64))
    with pytest.raises(MktfheError):
        s.block_step(0, 0, np.zeros((1, p.ell), np.uint32), np.zeros((1, 2, p.N), np.uint64))   # not a block scheme
    cg = gpu_schemes("CGGIparam")
    with pytest.raises(MktfheError, match="KMS"):
        cg.phase1(np.zeros((1, 1 + P.ALL["CGGIparam"].n), dtype=np.uint32))
    # the context is still usable after every rejected call
    b1, c1 = fresh_inputs(ks, 4, seed=91)
    b2, c2 = fresh_inputs(ks, 4, seed=92)
    want = np.array([PLAIN[0](bool(x), bool(y)) for x, y in zip(b1, b2)])
    assert np.array_equal(ks.decrypt_batch(s.gate(0, c1, c2)), want)


def test_chunked_batches_match_unchunked(monkeypatch):
    """A batch larger than the workspace budget runs in chunks inside the call; results must not depend on it."""
    from mktfhe_b200.scheme import setup
    ks = keyset("KMS2party")
    B = 150
    b1, c1 = fresh_inputs(ks, B, seed=101)
    b2, c2 = fresh_inputs(ks, B, seed=102)
    want = np.array([PLAIN[3](bool(x), bool(y)) for x, y in zip(b1, b2)])
    monkeypatch.setenv("MKTFHE_WORKSPACE_MB", "8")             # ~0.23 MB of scratch per gate -> chunks of ~35 gates
    small = setup(ks, device=0)
    monkeypatch.delenv("MKTFHE_WORKSPACE_MB")
    big = setup(ks, device=0)
    try:
        for mode in (MODE_FAST, MODE_STRICT):
            small.set_mode(mode)
            big.set_mode(mode)
            a = small.gate(3, c1, c2)
            b = big.gate(3, c1, c2)
            assert np.array_equal(a, b), mode
            assert np.array_equal(ks.decrypt_batch(a), want)
    finally:
        small.close()
        big.close()
