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
    """All parties' secret and evaluation keys for one parameter set, generated from a seed."""

    def __init__(self, params: Params, seed: int = 0x4D4B5446, nthreads: int = 0, want_ksk: bool = True,
                 secret_only: bool = False):
        self.params = p = params
        self.seed = seed
        self.crs_coeff = self.crs_fft = None
        if p.is_mk:
            self.crs_coeff, self.crs_fft = _host.crs(p, seed)
        nparties = p.k if p.is_mk else 1
        self.parties = [_host.party_keygen(p, seed, i, self.crs_coeff, nthreads, want_ksk, not secret_only)
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

    def lwe_encrypt(self, m: int, seed: int) -> np.ndarray:
        """Single-key `lwe_encrypt(m, key, params)`."""
        p = self.params
        out = np.empty(p.lwe_words, dtype=np.uint32)
        cp = self._cp()
        _host.lib().mktfhe_host_lwe_encrypt(ctypes.byref(cp), seed, int(m), _host.ptr(self.lwekeys[0]), _host.ptr(out))
        return out

    def lwe_ith_encrypt(self, m: int, i: int, seed: int) -> np.ndarray:
        """`lwe_ith_encrypt(m, i, lwekeys[i], params)` with 0-based party index."""
        p = self.params
        out = np.empty(p.lwe_words, dtype=np.uint32)
        cp = self._cp()
        rc = _host.lib().mktfhe_host_lwe_ith_encrypt(ctypes.byref(cp), seed, int(m), i, _host.ptr(self.lwekeys[i]), _host.ptr(out))
        if rc != 0:
            raise ValueError("bad party index")
        return out

    def lwe_encrypt_full(self, m: int, seed: int) -> np.ndarray:
        """Fresh ciphertext supported on all k blocks (bench/test input; no reference counterpart)."""
        p = self.params
        out = np.empty(p.lwe_words, dtype=np.uint32)
        cp = self._cp()
        _host.lib().mktfhe_host_lwe_encrypt_full(ctypes.byref(cp), seed, int(m), _host.ptr(self.lwekeys), _host.ptr(out))
        return out

    def encrypt_batch(self, bits, seed0: int) -> np.ndarray:
        p = self.params
        enc = self.lwe_encrypt_full if p.is_mk else self.lwe_encrypt
        return np.stack([enc(int(b), seed0 + i) for i, b in enumerate(bits)])

    def phase(self, ct) -> int:
        cp = self._cp()
        return int(_host.lib().mktfhe_host_lwe_phase(ctypes.byref(cp), _host.ptr(self.lwekeys), _host.ptr(np.ascontiguousarray(ct, dtype=np.uint32))))

    def lwe_decrypt(self, ct) -> bool:
        cp = self._cp()
        return bool(_host.lib().mktfhe_host_lwe_decrypt(ctypes.byref(cp), _host.ptr(self.lwekeys), _host.ptr(np.ascontiguousarray(ct, dtype=np.uint32))))

    def decrypt_batch(self, cts) -> np.ndarray:
        return np.array([self.lwe_decrypt(c) for c in cts], dtype=bool)
