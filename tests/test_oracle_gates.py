"""The reference's own acceptance criterion (test/CGGI.jl:34, LMSS.jl:34, CCS.jl:37, KMS.jl:37, KMSblock.jl:37):
random chains of all six gates, each followed by a bootstrap, decrypt to the plaintext circuit -- run through
the CPU oracle with host-generated keys at the five parameter sets those scripts use."""
import numpy as np
import pytest

from conftest import REFERENCE_TEST_SETS, keyset, make_oracle
from mktfhe_b200.gate import PLAIN


@pytest.mark.parametrize("name", REFERENCE_TEST_SETS)
def test_random_gate_chain_decrypts(name):
    ks = keyset(name)
    orc = make_oracle(ks)
    p = ks.params
    rng = np.random.default_rng(hash(name) % 1000)
    trials = 2
    nin = p.k if p.is_mk else 4
    for t in range(trials):
        m = rng.integers(0, 2, nin).astype(bool)
        cts = [ks.lwe_ith_encrypt(int(m[i]), i, 50 * t + i) if p.is_mk else ks.lwe_encrypt(int(m[i]), 50 * t + i)
               for i in range(nin)]
        for i in range(nin):
            assert ks.lwe_decrypt(cts[i]) == m[i]
        res, mres = cts[0], bool(m[0])
        for i in range(1, nin):
            op = int(rng.integers(0, 6))
            res = orc.bootstrap(orc.gate_linear(op, res, cts[i]))
            mres = PLAIN[op](mres, bool(m[i]))
            assert ks.lwe_decrypt(res) == mres, (name, t, i, op)
        res = orc.bootstrap(res)                      # the extra bootstrapping! of test/KMS.jl:36
        assert ks.lwe_decrypt(res) == mres


@pytest.mark.parametrize("name", ["KMS2party", "CCS2party"])
def test_all_gates_on_full_support_inputs(name):
    """Every opcode on ciphertexts supported on all party blocks (the shape bootstrapped outputs have)."""
    ks = keyset(name)
    orc = make_oracle(ks)
    rng = np.random.default_rng(9)
    for op in range(6):
        a, b = (int(x) for x in rng.integers(0, 2, 2))
        ca, cb = ks.lwe_encrypt_full(a, 900 + op), ks.lwe_encrypt_full(b, 950 + op)
        out = orc.bootstrap(orc.gate_linear(op, ca, cb))
        assert ks.lwe_decrypt(out) == PLAIN[op](bool(a), bool(b)), op
        err = (ks.phase(out) - ((1 << 29) if PLAIN[op](bool(a), bool(b)) else (7 << 29))) & 0xFFFFFFFF
        err = err - (1 << 32) if err >= (1 << 31) else err
        assert abs(err) < (1 << 29)


def test_batch_entry_matches_single_calls():
    ks = keyset("CGGIparam")
    orc = make_oracle(ks)
    c1 = ks.encrypt_batch([0, 1, 1], 10)
    c2 = ks.encrypt_batch([1, 1, 0], 20)
    batch = orc.gate_batch(0, c1, c2, nthreads=2)
    for g in range(3):
        assert np.array_equal(batch[g], orc.bootstrap(orc.gate_linear(0, c1[g], c2[g])))
