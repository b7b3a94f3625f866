"""GPU parity, STRICT mode: every stage of the CUDA path through the C-ABI equals the CPU oracle bit for bit
(integer stages AND floating-point stages: same butterflies, same order, no FMA)."""
import numpy as np
import pytest

from conftest import REFERENCE_TEST_SETS, fresh_inputs, keyset, make_oracle
from mktfhe_b200 import params as P
from mktfhe_b200.gate import PLAIN
from mktfhe_b200.scheme import MODE_STRICT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["CGGIparam", "KMS2party"])
def test_fft_ifft_decomp_bit_exact(gpu_schemes, name):
    from oracle import oracle as O
    s = gpu_schemes(name)
    p = s.params
    rng = np.random.default_rng(5)
    for dt in (np.uint32, np.uint64):
        bits = np.dtype(dt).itemsize * 8
        polys = rng.integers(0, 2 ** bits, size=(5, p.N), dtype=dt)
        polys[0, :] = 0
        polys[1, :] = np.iinfo(dt).max
        polys[2, ::2] = dt(1) << dt(bits - 1)             # signed minimum: the negation wraps
        spec = s.fft(polys)
        ref = np.stack([O.fft(q) for q in polys])
        assert np.array_equal(spec.view(np.uint64), ref.view(np.uint64))
        # inverse on spectra of products-like magnitude
        big = spec * 4097.0
        back = s.ifft(big, bits)
        refb = np.stack([O.ifft(q, bits) for q in big])
        assert np.array_equal(back, refb)
        for (l, logB) in ((3, 9), (2, 7), (4, 8)) if bits == 32 else ((3, 12), (2, 7), (16, 2), (6, 7)):
            d = s.decomp(polys, l, logB)
            refd = np.stack([O.decomp(q, l, logB) for q in polys])
            assert np.array_equal(d, refd), (bits, l, logB)


@pytest.mark.parametrize("name", REFERENCE_TEST_SETS)
def test_stages_bit_exact(gpu_schemes, name):
    ks = keyset(name)
    orc = make_oracle(ks)
    s = gpu_schemes(name)
    s.set_mode(MODE_STRICT) if s.mode != MODE_STRICT else None
    p = ks.params
    nin = 4
    bits, cts = fresh_inputs(ks, nin, seed=11)
    # edge ciphertext: a = 0xFFFFFFFF rounds to a~ = 2N (monomial table entry 2N = 0), a = 0 is skipped
    edge = cts[0].copy()
    edge[1:9] = 0xFFFFFFFF
    edge[9:17] = 0
    cts = np.concatenate([cts, edge[None]])
    # gate linear part, every opcode
    for op in range(6):
        lin = s.gate_linear(op, cts[:-1], cts[1:])
        for g in range(lin.shape[0]):
            assert np.array_equal(lin[g], orc.gate_linear(op, cts[g], cts[g + 1])), op
    lin = s.gate_linear(0, cts[:-1], cts[1:])
    lin = np.concatenate([lin, edge[None]])
    # modulus switch
    tilde = s.modswitch(lin)
    for g in range(lin.shape[0]):
        assert np.array_equal(tilde[g], orc.modswitch(lin[g]))
    assert (tilde[-1][1:9] == 2 * p.N).all()
    # blind rotation: all accumulator coefficients
    acc = s.blindrotate(lin)
    ref_acc = np.stack([orc.blindrotate(c) for c in lin])
    assert np.array_equal(acc, ref_acc), f"{name}: {np.count_nonzero(acc != ref_acc)} coefficients differ"
    # key switch on identical accumulators
    out = s.keyswitch(ref_acc)
    ref_out = np.stack([orc.keyswitch(a) for a in ref_acc])
    assert np.array_equal(out, ref_out)
    # whole bootstrap through the public entry point
    boot = s.bootstrapping(lin)
    assert np.array_equal(boot, ref_out)
    want = [PLAIN[0](bool(bits[g]), bool(bits[g + 1])) for g in range(nin - 1)]
    got = ks.decrypt_batch(boot[:nin - 1])
    assert list(got) == want


@pytest.mark.parametrize("name", ["KMS2party", "KMS2partyblock"])
def test_phase1_bit_exact(gpu_schemes, name):
    ks = keyset(name)
    orc = make_oracle(ks)
    s = gpu_schemes(name)
    s.set_mode(MODE_STRICT) if s.mode != MODE_STRICT else None
    p = ks.params
    _, cts = fresh_inputs(ks, 2, seed=21)
    lev = s.phase1(cts)
    for g in range(2):
        tilde = orc.modswitch(cts[g])
        r0 = 0
        for party in range(p.k):
            ref = orc.phase1(party, tilde[1 + party * p.n: 1 + (party + 1) * p.n])
            got = lev[g, r0:r0 + ref.shape[0]]
            assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), (g, party)
            r0 += ref.shape[0]


@pytest.mark.parametrize("name", ["CGGIparam", "KMS2party"])
def test_cmux_step_bit_exact(gpu_schemes, name):
    ks = keyset(name)
    orc = make_oracle(ks)
    s = gpu_schemes(name)
    s.set_mode(MODE_STRICT) if s.mode != MODE_STRICT else None
    p = ks.params
    rng = np.random.default_rng(3)
    dt = s.torus_dtype
    rows = rng.integers(0, np.iinfo(dt).max, size=(6, 2, p.N), dtype=dt)
    at = np.array([1, p.N - 1, p.N, p.N + 1, 2 * p.N - 1, 2 * p.N], dtype=np.uint32)
    party = p.k - 1 if p.is_mk else 0
    out = s.cmux_step(party, 7, at, rows)
    for g in range(6):
        assert np.array_equal(out[g], orc.cmux_step(party, 7, at[g], rows[g])), g


@pytest.mark.parametrize("name", REFERENCE_TEST_SETS)
def test_random_gate_chains(gpu_schemes, name):
    """The reference's own acceptance test (test/*.jl): random chains of all six gates decrypt to the plaintext
    circuit; here the chains of 8 trials run as one batch per level."""
    ks = keyset(name)
    s = gpu_schemes(name)
    p = ks.params
    rng = np.random.default_rng(17)
    trials, nin = 8, (p.k if p.is_mk else 4)
    m = rng.integers(0, 2, size=(trials, nin)).astype(bool)
    cts = [[ks.lwe_ith_encrypt(int(m[t, i]), i, 7000 + t * 10 + i) if p.is_mk else ks.lwe_encrypt(int(m[t, i]), 7000 + t * 10 + i)
            for i in range(nin)] for t in range(trials)]
    res = np.stack([c[0] for c in cts])
    mres = m[:, 0].copy()
    for i in range(1, nin):
        ops = rng.integers(0, 6, size=trials)
        nxt = np.stack([c[i] for c in cts])
        out = np.empty_like(res)
        for op in range(6):
            sel = np.nonzero(ops == op)[0]
            if len(sel):
                out[sel] = s.gate(op, res[sel], nxt[sel])
        res = out
        mres = np.array([PLAIN[int(ops[t])](bool(mres[t]), bool(m[t, i])) for t in range(trials)])
    res = s.bootstrapping(res)               # the extra `@time bootstrapping!` of test/KMS.jl:36
    assert list(ks.decrypt_batch(res)) == list(mres)
