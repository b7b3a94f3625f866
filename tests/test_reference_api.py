"""mktfhe_b200/reference_api.py: the reference's names and argument order (src/MKTFHE.jl:21-35).  The host half
(CRS, party_keygen, lwe_encrypt, lwe_ith_encrypt, lwe_decrypt, NOT!) is checked here the way test/KMS.jl uses it, with the
CPU oracle standing in for the evaluator; `setup` + gates on the GPU are `examples/kms_flow.py`."""
import numpy as np
import pytest

from mktfhe_b200 import params as P
from mktfhe_b200.gate import PLAIN
from mktfhe_b200.reference_api import (CRS, LWEkey, NOT_, lwe_decrypt, lwe_encrypt, lwe_ith_encrypt, party_keygen, setup)


def test_kms_flow_like_test_KMS_jl():
    from oracle import oracle as O
    params = P.KMS2party
    a = CRS(params, seed=0x4D4B5446)
    keys = [party_keygen(a, params) for _ in range(params.k)]
    lwekeys, btk = [q[0] for q in keys], [q[-1] for q in keys]
    assert [k.party for k in btk] == [0, 1] and btk[0].brk.shape[0] == params.n
    orc = O.Oracle(params, [k.brk for k in btk], [k.ksk for k in btk], [k.rlk for k in btk], [k.b for k in btk], a.fft)
    rng = np.random.default_rng(12)
    for trial in range(2):
        m = rng.integers(0, 2, params.k).astype(bool)
        ctxts = [lwe_ith_encrypt(m[i - 1], i, lwekeys[i - 1], params) for i in range(1, params.k + 1)]
        for i in range(params.k):
            assert lwe_decrypt(ctxts[i], lwekeys, params) == m[i]
            blocks = ctxts[i][1:].reshape(params.k, params.n)
            assert np.count_nonzero(blocks[i]) > 0 and all(np.count_nonzero(blocks[j]) == 0 for j in range(params.k) if j != i)
        res, mres = ctxts[0], bool(m[0])
        for i in range(2, params.k + 1):
            op = int(rng.integers(0, 6))
            res = orc.bootstrap(orc.gate_linear(op, res, ctxts[i - 1]))
            mres = PLAIN[op](mres, bool(m[i - 1]))
        res = orc.bootstrap(res)
        assert mres == lwe_decrypt(res, lwekeys, params)
        assert lwe_decrypt(NOT_(res.copy()), lwekeys, params) == (not mres)


def test_keys_match_the_keyset_generator():
    """Same seed, same arrays as KeySet (the generator the golden vectors use)."""
    from conftest import keyset
    params = P.KMS2party
    ks = keyset("KMS2party")
    a = CRS(params, seed=ks.seed)
    lwekey, ringkey, btk = party_keygen(a, params, seed=ks.seed)
    assert np.array_equal(lwekey.key, ks.parties[0]["lwekey"]) and np.array_equal(btk.ksk, ks.parties[0]["ksk"])
    assert np.array_equal(btk.brk, ks.parties[0]["brk"]) and np.array_equal(a.fft, ks.crs_fft)


def test_fresh_randomness_by_default():
    params = P.CGGIparam
    key = LWEkey(np.random.default_rng(1).integers(0, 2, params.n).astype(np.uint32))
    c1, c2 = lwe_encrypt(1, key, params), lwe_encrypt(1, key, params)
    assert not np.array_equal(c1, c2) and lwe_decrypt(c1, key, params) and lwe_decrypt(c2, key, params)
    assert np.array_equal(lwe_encrypt(0, key, params, seed=9), lwe_encrypt(0, key, params, seed=9))
    assert lwe_decrypt(lwe_encrypt(0, key, params), key, params) is False


def test_party_secrets_are_independent_of_the_crs_and_of_each_other():
    """a party's secret keys must not be derivable from the public CRS.  By default every party_keygen call
    draws its own 256-bit key from the OS: two key generations over the SAME CRS give unrelated secrets, the CRS object carries
    no seed, and two parties never share a secret."""
    params = P.CCS2party
    a1, a2 = CRS(params, seed=7), CRS(params, seed=7)
    assert np.array_equal(a1.coeff, a2.coeff) and not hasattr(a1, "seed")
    k1 = [party_keygen(a1, params)[0].key for _ in range(params.k)]
    k2 = [party_keygen(a2, params)[0].key for _ in range(params.k)]
    assert not np.array_equal(k1[0], k2[0]) and not np.array_equal(k1[1], k2[1])      # same public CRS, fresh secrets
    assert not np.array_equal(k1[0], k1[1])
    assert not np.array_equal(CRS(params).coeff, CRS(params).coeff)


def test_keyset_default_seed_is_random():
    from mktfhe_b200.keys import KeySet
    p = P.CGGIparam
    k1, k2 = KeySet(p, want_ksk=False, secret_only=True), KeySet(p, want_ksk=False, secret_only=True)
    assert k1.seed is None and not np.array_equal(k1.lwekeys, k2.lwekeys)
    c1, c2 = k1.lwe_encrypt(1), k1.lwe_encrypt(1)                # fresh randomness per encryption
    assert not np.array_equal(c1, c2) and k1.lwe_decrypt(c1) and k1.lwe_decrypt(c2)
    assert np.array_equal(k1.lwe_encrypt(1, 5), k1.lwe_encrypt(1, 5))
    batch = k1.encrypt_batch([1, 0, 1])
    assert not np.array_equal(batch[0, 1:], batch[2, 1:]) and list(k1.decrypt_batch(batch)) == [True, False, True]


def test_argument_errors():
    kms, cggi = P.KMS2party, P.CGGIparam
    with pytest.raises(TypeError):
        CRS(cggi)
    a = CRS(kms, seed=1)
    with pytest.raises(ValueError):
        party_keygen(a, P.KMS4party)
    key = LWEkey(np.zeros(kms.n, dtype=np.uint32))
    for bad in (0, kms.k + 1):
        with pytest.raises(IndexError):
            lwe_ith_encrypt(1, bad, key, kms)
    with pytest.raises(TypeError):
        lwe_encrypt(1, key, kms)
    with pytest.raises(TypeError):
        lwe_ith_encrypt(1, 1, key, cggi)
    with pytest.raises(ValueError):
        lwe_decrypt(np.zeros(kms.lwe_words, np.uint32), [key], kms)          # one key for two parties
    with pytest.raises(ValueError):
        lwe_decrypt(np.zeros(3, np.uint32), [key, key], kms)
    with pytest.raises(TypeError):
        setup(kms)                                                             # multi-key needs (a, btk, params)
    with pytest.raises(TypeError):
        setup(a, [], cggi)
