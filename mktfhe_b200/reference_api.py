"""The reference's own function names and argument order (/root/reference/src/MKTFHE.jl:21-35), for code and tests that
should read like test/KMS.jl or test/CGGI.jl:

    a = CRS(params)                                          # scheme.jl:409-410
    keys = [party_keygen(a, params) for _ in range(params.k)]   # (lwekey, ringkey, btk) per party, scheme.jl:227-242 ...
    lwekeys, btk = [q[0] for q in keys], [q[-1] for q in keys]
    scheme = setup(a, btk, params)                           # scheme.jl:244,292,343  (+ upload to the GPU)
    c = lwe_ith_encrypt(m, i, lwekeys[i - 1], params)        # i is 1-based like the reference, scheme.jl:379-386
    res = NAND(c1, c2, scheme); bootstrapping_(res, scheme)  # gate.jl, bootstrapping.jl:4-27 (`bootstrapping!`)
    assert lwe_decrypt(res, lwekeys, params) == expected     # scheme.jl:388-407

    lwekey, ringkey, scheme = setup(params)                  # single-key schemes, scheme.jl:151,190

Differences that cannot be avoided: Julia's `!` is spelled `_` (`bootstrapping_`, `NOT_`), ciphertexts are uint32 arrays
`[b, a...]` instead of `LWE` structs.  Randomness: by default every call (CRS, each party's keygen, each encryption) draws
its own 256-bit ChaCha20 key from the OS CSPRNG, like the reference's unseeded `ChaCha20Stream()` calls -- a party's secrets
are unrelated to the public CRS and to every other party.  `seed=<int>` makes a call reproducible for tests; such keys and
ciphertexts are not secret (64-bit seed) and a seed must never be reused for two encryptions.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

from . import _host
from .gate import AND, NAND, NOR, OR, XNOR, XOR  # noqa: F401  (re-exported with the reference's names)
from .params import Params
from .scheme import Scheme


def _fresh_seed() -> bytes:
    return _host.fresh_key()


@dataclass
class CommonReferenceString:
    """`a = CRS(params)`: l_uni uniform polynomials (coefficient form for keygen, FFT form for the evaluator).  Public data:
    it carries no seed, and nothing secret is ever derived from it."""
    params: Params
    coeff: np.ndarray
    fft: np.ndarray
    _next_party: int = 0


@dataclass
class LWEkey:
    key: np.ndarray                    # uint32 [n], binary (or block-binary) secret, key.jl


@dataclass
class BootKey:
    """Evaluation key of one party in the flat upload layouts (include/mktfhe_b200.h): BootKey_KMS / _CCS / _bin / _block."""
    party: int
    brk: np.ndarray
    ksk: np.ndarray
    rlk: np.ndarray | None = None
    b: np.ndarray | None = None        # public key part (`pubb`), keygen.jl:85-93


def CRS(params: Params, seed: int | None = None) -> CommonReferenceString:
    if not params.is_mk:
        raise TypeError("CRS is defined for the multi-key parameter sets (CCS*, KMS*) only")
    coeff, fft = _host.crs(params, _fresh_seed() if seed is None else int(seed))
    return CommonReferenceString(params, coeff, fft)


def party_keygen(a: CommonReferenceString, params: Params, nthreads: int = 0, seed: int | None = None):
    """-> (lwekey, ringkey, btk) for the next party, so that `first.(keys)` / `last.(keys)` of test/KMS.jl:10-12 carry over.
    The party's secrets come from its own fresh 256-bit key (seed=<int>: reproducible, for tests only)."""
    if a.params != params:
        raise ValueError("the CRS was made for a different parameter set")
    if a._next_party >= params.k:
        raise ValueError(f"all {params.k} parties of this CRS already have keys")
    party = a._next_party
    a._next_party += 1
    q = _host.party_keygen(params, _fresh_seed() if seed is None else int(seed), party, a.coeff, nthreads)
    return LWEkey(q["lwekey"]), q["ringkey"], BootKey(party, q["brk"], q["ksk"], q["rlk"], q["pubb"])


def setup(*args, device: int = 0, seed: int | None = None, mode: int | None = None):
    """`setup(params)` -> (lwekey, ringkey, scheme) for CGGI / LMSS; `setup(a, btk, params)` -> scheme for CCS / KMS."""
    if len(args) == 1:
        params, = args
        if params.is_mk:
            raise TypeError("multi-key parameter sets need setup(a, btk, params)")
        q = _host.party_keygen(params, _fresh_seed() if seed is None else int(seed), 0, None)
        s = Scheme(params, device)
        s.upload_party(0, q["brk"], q["ksk"], None, None)
        s.finalize()
        if mode is not None:
            s.set_mode(mode)
        return LWEkey(q["lwekey"]), q["ringkey"], s
    if len(args) != 3:
        raise TypeError("setup(params) or setup(a, btk, params)")
    a, btk, params = args
    if not params.is_mk or a.params != params:
        raise TypeError("setup(a, btk, params) needs a multi-key parameter set and its own CRS")
    if len(btk) != params.k or sorted(k.party for k in btk) != list(range(params.k)):
        raise ValueError(f"need the boot keys of all {params.k} parties of this CRS")
    s = Scheme(params, device)
    for k in btk:
        s.upload_party(k.party, k.brk, k.ksk, k.rlk, k.b)
    s.upload_common(a.fft)
    s.finalize()
    if mode is not None:
        s.set_mode(mode)
    return s


def lwe_encrypt(m, key: LWEkey, params: Params, seed: int | None = None) -> np.ndarray:
    """Single-key `lwe_encrypt(m, key, params)`: scheme.jl:352-368."""
    if params.is_mk:
        raise TypeError("use lwe_ith_encrypt for multi-key parameter sets")
    return _host.encrypt_batch(params, _fresh_seed() if seed is None else int(seed), _host.ENC_SINGLE, 0, [bool(m)],
                               np.ascontiguousarray(key.key), 1)[0]


def lwe_ith_encrypt(m, i: int, key: LWEkey, params: Params, seed: int | None = None) -> np.ndarray:
    """`lwe_ith_encrypt(m, i, key, params)` with the reference's 1-based party index: scheme.jl:370-386."""
    if not params.is_mk:
        raise TypeError("use lwe_encrypt for single-key parameter sets")
    if not 1 <= int(i) <= params.k:
        raise IndexError(f"party index {i} outside 1..{params.k}")
    return _host.encrypt_batch(params, _fresh_seed() if seed is None else int(seed), _host.ENC_ITH, int(i) - 1, [bool(m)],
                               np.ascontiguousarray(key.key), 1)[0]


def lwe_decrypt(lwe, key, params: Params | None = None) -> bool:
    """`lwe_decrypt(lwe, key)` (single key) or `lwe_decrypt(lwe, keys, params)` (all parties' keys): scheme.jl:388-407."""
    keys = [key] if isinstance(key, LWEkey) else list(key)
    if params is None:
        raise TypeError("pass the parameter set (the reference infers it from the key type)")
    expect = params.k if params.is_mk else 1
    if len(keys) != expect:
        raise ValueError(f"need {expect} LWE key(s), got {len(keys)}")
    flat = np.ascontiguousarray(np.stack([k.key for k in keys]))
    ct = np.ascontiguousarray(lwe, dtype=np.uint32)
    if ct.shape != (params.lwe_words,):
        raise ValueError(f"ciphertext has shape {ct.shape}, expected ({params.lwe_words},)")
    cp = params.c_struct()
    return bool(_host.lib().mktfhe_host_lwe_decrypt(ctypes.byref(cp), _host.ptr(flat), _host.ptr(ct)))


def bootstrapping_(ctxt: np.ndarray, scheme: Scheme) -> np.ndarray:
    """`bootstrapping!(ctxt, scheme)`: refreshes the ciphertext in place (bootstrapping.jl:4-27) and returns it."""
    ctxt[...] = scheme.bootstrapping(ctxt)
    return ctxt


def NOT_(ctxt: np.ndarray) -> np.ndarray:
    """`NOT!(ctxt)`: in-place negation, no bootstrap (gate.jl:55-58)."""
    np.negative(ctxt, out=ctxt)
    return ctxt
