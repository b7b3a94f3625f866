"""Host-side key material in the flat upload layouts of include/mktfhe_b200.h.

Mirrors the reference's host flow:  a = CRS(params); keys = [party_keygen(a, params) ...]
(/root/reference/test/KMS.jl:6-12, src/tfhe/scheme.jl:227-242,273-287,324-338,409-410) and
`setup(params)` for the single-key schemes (scheme.jl:151-166,190-205).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _host
from .params import Params


class KeySet:
    """All parties' secret and evaluation keys for one parameter set, held in ONE process: the shape tests, benchmarks and
    single-owner deployments need.  (In a real multi-party run every party calls `reference_api.party_keygen` on its own
    machine and only the evaluation keys travel.)

    seed=None (default): the CRS and every party draw independent 256-bit ChaCha20 keys from the OS CSPRNG, like the
    reference's unseeded streams.  seed=<int>: everything derives from that integer -- reproducible and therefore NOT secret
    (64 bits, and the CRS shares it); tests, benchmarks and golden vectors only."""

    def __init__(self, params: Params, seed: int | None = None, nthreads: int = 0, want_ksk: bool = True,
                 secret_only: bool = False):
        self.params = p = params
        self.seed = seed
        self.crs_coeff = self.crs_fft = None
        nparties = p.k if p.is_mk else 1
        if seed is None:
            crs_seed, party_seeds = _host.fresh_key(), [_host.fresh_key() for _ in range(nparties)]
        else:
            crs_seed, party_seeds = int(seed), [int(seed)] * nparties
        if p.is_mk:
            self.crs_coeff, self.crs_fft = _host.crs(p, crs_seed)
        self.parties = [_host.party_keygen(p, party_seeds[i], i, self.crs_coeff, nthreads, want_ksk, not secret_only)
                        for i in range(nparties)]
        self.lwekeys = np.ascontiguousarray(np.stack([q["lwekey"] for q in self.parties]))   # [k][n]

    # flat views -------------------------------------------------------------------------
    @property
    def brk(self):
        return [q["brk"] for q in self.parties]

    @property
    def ksk(self):
        return [q["ksk"] for q in self.parties]

    @property
    def rlk(self):
        return [q["rlk"] for q in self.parties]

    @property
    def pubb(self):
        return [q["pubb"] for q in self.parties]

    # encrypt / decrypt (scheme.jl:352-407) -------------------------------------------------
    def _cp(self):
        return self.params.c_struct()

    # Encryption randomness: seed=None (default) draws a fresh 256-bit ChaCha20 key from the OS for every call.  An int seed
    # makes mask and noise a pure function of it: REUSING A SEED FOR TWO CIPHERTEXTS REPEATS BOTH (c1 - c0 reveals m1 - m0) and a
    # known seed reveals the noise, i.e. one exact linear equation in the secret per ciphertext.  Tests and benchmarks only.
    def lwe_encrypt(self, m: int, seed: int | None = None) -> np.ndarray:
        """Single-key `lwe_encrypt(m, key, params)`."""
        self._cp()
        return _host.encrypt_batch(self.params, _host.fresh_key() if seed is None else seed, _host.ENC_SINGLE, 0, [m], self.lwekeys[0], 1)[0]

    def lwe_ith_encrypt(self, m: int, i: int, seed: int | None = None) -> np.ndarray:
        """`lwe_ith_encrypt(m, i, lwekeys[i], params)` with 0-based party index."""
        self._cp()
        if not 0 <= int(i) < self.params.k:
            raise ValueError("bad party index")
        return _host.encrypt_batch(self.params, _host.fresh_key() if seed is None else seed, _host.ENC_ITH, int(i), [m], self.lwekeys[i], 1)[0]

    def lwe_encrypt_full(self, m: int, seed: int | None = None) -> np.ndarray:
        """Fresh ciphertext supported on all k blocks (bench/test input; no reference counterpart)."""
        self._cp()
        return _host.encrypt_batch(self.params, _host.fresh_key() if seed is None else seed, _host.ENC_FULL, 0, [m], self.lwekeys, 1)[0]

    def encrypt_batch(self, bits, seed0: int | None = None, party: int | None = None, nthreads: int = 0) -> np.ndarray:
        """One native call for the whole batch (OpenMP over ciphertexts).  MK sets: full-support ciphertexts, or party i's
        `lwe_ith_encrypt` when `party` is given.  With an int seed0, ciphertext g equals the single call with seed0 + g."""
        p = self.params
        self._cp()
        seed = _host.fresh_key() if seed0 is None else seed0
        if not p.is_mk:
            return _host.encrypt_batch(p, seed, _host.ENC_SINGLE, 0, bits, self.lwekeys[0], nthreads)
        if party is not None:
            return _host.encrypt_batch(p, seed, _host.ENC_ITH, int(party), bits, self.lwekeys[int(party)], nthreads)
        return _host.encrypt_batch(p, seed, _host.ENC_FULL, 0, bits, self.lwekeys, nthreads)

    def phase(self, ct) -> int:
        cp = self._cp()
        return int(_host.lib().mktfhe_host_lwe_phase(ctypes.byref(cp), _host.ptr(self.lwekeys), _host.ptr(np.ascontiguousarray(ct, dtype=np.uint32))))

    def lwe_decrypt(self, ct) -> bool:
        cp = self._cp()
        return bool(_host.lib().mktfhe_host_lwe_decrypt(ctypes.byref(cp), _host.ptr(self.lwekeys), _host.ptr(np.ascontiguousarray(ct, dtype=np.uint32))))

    def decrypt_batch(self, cts, nthreads: int = 0) -> np.ndarray:
        """lwe_decrypt over a batch in one native call (OpenMP over ciphertexts)."""
        self._cp()
        return _host.decrypt_batch(self.params, self.lwekeys, cts, nthreads)

    def phase_batch(self, cts, nthreads: int = 0) -> np.ndarray:
        self._cp()
        return _host.phase_batch(self.params, self.lwekeys, cts, nthreads)
