"""Host side of the drop-in surface: parameter sets, seeded key generation, encryption and decryption
(mirrors /root/reference/src/tfhe/params.jl, scheme.jl:352-410, keygen.jl)."""
import numpy as np
import pytest

from conftest import keyset
from mktfhe_b200 import params as P
from mktfhe_b200.keys import KeySet
from oracle import oracle as O


def test_named_sets_match_params_jl():
    assert len(P.ALL) == 16
    k = P.KMS32party
    assert (k.n, k.N, k.k, k.l_gsw, k.logB_gsw, k.l_lev, k.logB_lev, k.l_uni, k.logB_uni) == (560, 2048, 32, 6, 7, 3, 7, 16, 2)
    b = P.KMS8partyblock
    assert (b.d, b.ell, b.n, b.ksk_rows, b.torus_bits) == (203, 3, 609, 2, 64)
    assert P.CCS16party.brk_polys == 36 and P.CGGIparam.ksk_rows == 3 and P.Blockparam.n == 687
    assert P.KMS2party.rows(0) == 1 and P.KMS2party.rows(1) == 2
    # brk per party 105 MiB at KMS2party, ksk 110.3 MB
    assert P.KMS2party.brk_doubles * 8 == 560 * 12 * 1024 * 16
    assert abs(P.KMS2party.ksk_words * 4 / 1e6 - 110.3) < 0.1


def test_keygen_is_deterministic_in_the_seed():
    p = P.small(P.CGGIparam, n=16)
    a = KeySet(p, seed=7, nthreads=1)
    b = KeySet(p, seed=7, nthreads=4)
    c = KeySet(p, seed=8)
    assert np.array_equal(a.brk[0], b.brk[0]) and np.array_equal(a.ksk[0], b.ksk[0])
    assert not np.array_equal(a.brk[0], c.brk[0])


@pytest.mark.parametrize("name", ["CGGIparam", "CCS2party", "KMS2party", "KMS2partyblock"])
def test_fresh_ciphertexts_decrypt_and_noise_is_alpha(name):
    ks = keyset(name)
    p = ks.params
    errs = []
    for i in range(64):
        m = i & 1
        ct = ks.lwe_encrypt_full(m, 3000 + i) if p.is_mk else ks.lwe_encrypt(m, 3000 + i)
        assert ks.lwe_decrypt(ct) == bool(m)
        e = (ks.phase(ct) - ((1 << 29) if m else (7 << 29))) & 0xFFFFFFFF
        errs.append(e - (1 << 32) if e >= (1 << 31) else e)
    assert 0.6 * p.alpha < np.std(errs) < 1.5 * p.alpha
    if p.is_mk:
        ct = ks.lwe_ith_encrypt(1, p.k - 1, 5)
        assert np.all(ct[1:1 + (p.k - 1) * p.n] == 0) and ks.lwe_decrypt(ct)


def test_block_binary_keys_have_at_most_one_bit_per_block():
    ks = keyset("KMS2partyblock")
    p = ks.params
    for key in ks.lwekeys:
        blocks = key.reshape(p.d, p.ell)
        assert blocks.max() <= 1 and blocks.sum(axis=1).max() <= 1
    # unikey's first n coefficients equal the LWE key (key.jl:71-88)
    assert np.array_equal(ks.parties[0]["ringkey"][:p.n].astype(np.uint32), ks.lwekeys[0])


def test_ksk_rows_encrypt_digit_times_key_coefficient():
    """keygen.jl:110-114: ksk[digit, c].stack[level] is an LWE encryption of digit * key[c] * 2^(32 - 2(level+1))."""
    ks = keyset("KMS2party")
    p = ks.params
    ksk, s, rk = ks.ksk[0], ks.lwekeys[0].astype(np.uint64), ks.parties[0]["ringkey"]
    for c in (0, 5, p.N - 1):
        for dg in range(1, 4):
            for lv in (0, 3, 7):
                row = ksk[c, dg - 1, lv].astype(np.uint64)
                phase = int((row[0] + np.sum(row[1:] * s)) % (1 << 32))
                want = (dg * int(rk[c]) << (32 - 2 * (lv + 1))) % (1 << 32)
                err = (phase - want + (1 << 31)) % (1 << 32) - (1 << 31)
                assert abs(err) < 8 * p.alpha


def test_brk_rows_are_rlwe_encryptions_of_key_bit_times_gadget():
    """keygen.jl:106-108 -> gsw.jl:174-178: basketb.stack[j] = RLWE(bit * g_j) under gswkey; phase = b + a*z."""
    ks = keyset("CGGIparam")
    p = ks.params
    z = ks.parties[0]["ringkey"].astype(np.int64)                 # CGGI: brk is under the ring key
    fz = O.fft(z.astype(np.uint32))
    cz = fz[:, 0] + 1j * fz[:, 1]
    for idx in (0, 3, p.n - 1):
        bit = int(ks.lwekeys[0][idx])
        for j in range(p.l_gsw):
            b = ks.brk[0][idx, (0 * p.l_gsw + j) * 2 + 0]
            a = ks.brk[0][idx, (0 * p.l_gsw + j) * 2 + 1]
            ph = (b[:, 0] + 1j * b[:, 1]) + (a[:, 0] + 1j * a[:, 1]) * cz
            coeffs = O.ifft(np.stack([ph.real, ph.imag], axis=1), 32).astype(np.int64)
            coeffs[coeffs >= 2 ** 31] -= 2 ** 32
            want0 = bit << (32 - (j + 1) * p.logB_gsw)
            assert abs(coeffs[0] - want0) < 16 * p.beta + 2 ** 12      # key noise + Float64 transform rounding
            assert np.abs(coeffs[1:]).max() < 16 * p.beta + 2 ** 12
