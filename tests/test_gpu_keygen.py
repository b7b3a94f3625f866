"""Device key generation (csrc/keygen.cuh) against the host library: same seeded ChaCha20 streams, exact
integer ring products, the reference's Float64 transform -> every flat key array must be BYTE-IDENTICAL to
mktfhe_host_party_keygen's (/root/reference/src/tfhe/keygen.jl:3-155, src/ciphertext/gsw.jl:174-184, unienc.jl:36-90,
lev.jl:31-45 are what both restate), and gates evaluated with device-generated keys equal gates evaluated with uploaded keys."""
import hashlib
import time

import numpy as np
import pytest

from conftest import REFERENCE_TEST_SETS, fresh_inputs, keyset
from mktfhe_b200 import params as P
from mktfhe_b200.scheme import MODE_STRICT, Scheme, setup_generated

pytestmark = pytest.mark.gpu
SEED = 0x4D4B5446


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name", REFERENCE_TEST_SETS + ["KMS4party", "CCS4party"])
def test_device_keys_are_byte_identical_to_host_keys(name):
    ks = keyset(name)                      # host keys, seed 0x4D4B5446
    p = ks.params
    s = Scheme(p, 0)
    try:
        if p.is_mk:
            s.keygen_common(SEED)
        for i in range(p.k if p.is_mk else 1):
            s.keygen_party(i, SEED)
        for i, q in enumerate(ks.parties):
            got = s.download_party_key(i)
            for f in ("ksk", "pubb", "rlk", "brk"):
                if q.get(f) is None:
                    continue
                if not np.array_equal(got[f], q[f]):
                    bad = np.argwhere(got[f] != q[f])
                    raise AssertionError(f"{name} party {i} {f}: {len(bad)} of {q[f].size} elements differ, first at {bad[0]}")
                assert _sha(got[f]) == _sha(q[f])
            if p.is_mk:
                assert np.array_equal(got["crs_fft"], ks.crs_fft)
    finally:
        s.close()


@pytest.mark.parametrize("name", ["KMS2party", "CGGIparam"])
def test_gates_with_device_generated_keys(gpu_schemes, name):
    uploaded = gpu_schemes(name)
    s, secrets = setup_generated(P.ALL[name], SEED, mode=MODE_STRICT)
    try:
        ks = keyset(name)
        assert np.array_equal(secrets.lwekeys, ks.lwekeys)
        b1, c1 = fresh_inputs(ks, 5, seed=131)
        b2, c2 = fresh_inputs(ks, 5, seed=132)
        uploaded.set_mode(MODE_STRICT)
        assert np.array_equal(s.gate(0, c1, c2), uploaded.gate(0, c1, c2))
        assert list(secrets.decrypt_batch(s.gate(3, c1, c2))) == [bool(x) != bool(y) for x, y in zip(b1, b2)]
    finally:
        uploaded.set_mode(1)
        s.close()


def test_kms32_keys_on_device_quickly():
    """North-star config 5: the 10.6 GB key set of 32 parties, generated where it is used."""
    p = P.KMS32party
    t = time.perf_counter()
    s, secrets = setup_generated(p, SEED)
    dt = time.perf_counter() - t
    try:
        print(f"KMS32party: device key generation + finalize {dt:.2f} s (host generation + upload: ~18 s)")
        b1, c1 = fresh_inputs(secrets, 8, seed=141)
        b2, c2 = fresh_inputs(secrets, 8, seed=142)
        out = s.gate(0, c1, c2)
        assert int(np.sum(secrets.decrypt_batch(out) == ~(b1 & b2))) >= 7      # the set fails ~1.6 % of gates by itself
        assert dt < 20.0
    finally:
        s.close()
