import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


from mktfhe_b200 import params as P  # noqa: E402
from mktfhe_b200.keys import KeySet  # noqa: E402

# The five parameter sets the reference's own test scripts run (test/CGGI.jl:5, LMSS.jl:5, CCS.jl:5, KMS.jl:5,
# KMSblock.jl:5).
REFERENCE_TEST_SETS = ["CGGIparam", "Blockparam", "CCS2party", "KMS2party", "KMS2partyblock"]

_keysets = {}


def keyset(name: str) -> KeySet:
    if name not in _keysets:
        _keysets[name] = KeySet(P.ALL[name], seed=0x4D4B5446)
    return _keysets[name]


def make_oracle(ks: KeySet):
    from oracle import oracle as O
    p = ks.params
    return O.Oracle(p, ks.brk, ks.ksk, ks.rlk if p.scheme in (P.KMS, P.KMS_BLOCK) else None,
                    ks.pubb if p.is_mk else None, ks.crs_fft)


def fresh_inputs(ks: KeySet, count: int, seed: int, full: bool = True):
    """`count` fresh ciphertexts of random bits (full support for MK sets unless full=False)."""
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, count).astype(bool)
    p = ks.params
    cts = []
    for i, b in enumerate(bits):
        if not p.is_mk:
            cts.append(ks.lwe_encrypt(int(b), seed * 1000 + i))
        elif full:
            cts.append(ks.lwe_encrypt_full(int(b), seed * 1000 + i))
        else:
            cts.append(ks.lwe_ith_encrypt(int(b), i % p.k, seed * 1000 + i))
    return bits, np.stack(cts)


@pytest.fixture(scope="session")
def gpu_schemes():
    """Lazily created device contexts, one per parameter set, shared by the GPU tests."""
    from mktfhe_b200.scheme import setup
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = setup(keyset(name), device=0)
        return cache[name]
    yield get
    for s in cache.values():
        s.close()
