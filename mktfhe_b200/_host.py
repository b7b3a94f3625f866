"""ctypes binding of libmktfhe_host.so (include/mktfhe_host.h): host key generation, encryption, decryption."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build
from .params import CParams, Params

_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        path = build.HOST_LIB
        if not os.path.exists(path):
            build.build_host()
        _lib = ctypes.CDLL(path)
        P = ctypes.POINTER(CParams)
        vp, u64, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int
        _lib.mktfhe_host_crs.argtypes = [P, u64, vp, vp]
        _lib.mktfhe_host_party_keygen.argtypes = [P, u64, i32, vp, vp, vp, vp, vp, vp, vp, i32]
        _lib.mktfhe_host_lwe_encrypt.argtypes = [P, u64, i32, vp, vp]
        _lib.mktfhe_host_lwe_ith_encrypt.argtypes = [P, u64, i32, i32, vp, vp]
        _lib.mktfhe_host_lwe_encrypt_full.argtypes = [P, u64, i32, vp, vp]
        _lib.mktfhe_host_lwe_phase.argtypes = [P, vp, vp]
        _lib.mktfhe_host_lwe_phase.restype = ctypes.c_uint32
        _lib.mktfhe_host_lwe_decrypt.argtypes = [P, vp, vp]
        _lib.mktfhe_host_crs_key.argtypes = [P, vp, vp, vp]
        _lib.mktfhe_host_party_keygen_key.argtypes = [P, vp, i32, vp, vp, vp, vp, vp, vp, vp, i32]
        _lib.mktfhe_host_lwe_encrypt_key.argtypes = [P, vp, i32, vp, vp]
        _lib.mktfhe_host_lwe_ith_encrypt_key.argtypes = [P, vp, i32, i32, vp, vp]
        _lib.mktfhe_host_lwe_encrypt_full_key.argtypes = [P, vp, i32, vp, vp]
        _lib.mktfhe_host_encrypt_batch.argtypes = [P, u64, i32, i32, vp, ctypes.c_size_t, vp, vp, i32]
        _lib.mktfhe_host_encrypt_batch_key.argtypes = [P, vp, i32, i32, vp, ctypes.c_size_t, vp, vp, i32]
        _lib.mktfhe_host_decrypt_batch.argtypes = [P, vp, vp, ctypes.c_size_t, vp, i32]
        _lib.mktfhe_host_phase_batch.argtypes = [P, vp, vp, ctypes.c_size_t, vp, i32]
        _lib.mktfhe_host_fft_tables.argtypes = [i32, vp, vp, vp, vp]
        _lib.mktfhe_host_fft_tables.restype = None
    return _lib


def ptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def torus_dtype(p: Params):
    return np.uint64 if p.torus_bits == 64 else np.uint32


def fresh_key() -> bytes:
    """A 256-bit ChaCha20 key from the OS CSPRNG: the production source of every secret and of encryption randomness."""
    return os.urandom(32)


def is_key(seed) -> bool:
    """Randomness arguments are either a 32-byte key (production) or an int seed (reproducible: tests, benchmarks)."""
    if isinstance(seed, (bytes, bytearray)):
        if len(seed) != 32:
            raise ValueError("a ChaCha20 key has 32 bytes")
        return True
    if isinstance(seed, (int, np.integer)):
        return False
    raise TypeError("seed must be an int (reproducible) or 32 bytes (a ChaCha20 key)")


def _kbuf(key: bytes):
    return ctypes.cast(ctypes.create_string_buffer(bytes(key), 32), ctypes.c_void_p)


def crs(p: Params, seed):
    cp = p.c_struct()
    coeff = np.empty((p.l_uni, p.N), dtype=torus_dtype(p))
    fft = np.empty((p.l_uni, p.H, 2), dtype=np.float64)
    if is_key(seed):
        rc = lib().mktfhe_host_crs_key(ctypes.byref(cp), _kbuf(seed), ptr(coeff), ptr(fft))
    else:
        rc = lib().mktfhe_host_crs(ctypes.byref(cp), int(seed), ptr(coeff), ptr(fft))
    if rc != 0:
        raise RuntimeError(f"mktfhe_host_crs failed: {rc}")
    return coeff, fft


def party_keygen(p: Params, seed, party: int, crs_coeff, nthreads: int = 0, want_ksk: bool = True,
                 want_eval: bool = True):
    cp = p.c_struct()
    out = {
        "lwekey": np.empty(p.n, dtype=np.uint32),
        "ringkey": np.empty(p.N, dtype=torus_dtype(p)),
        "brk": np.empty((p.n, p.brk_polys, p.H, 2), dtype=np.float64) if want_eval else None,
        "rlk": np.empty((p.l_uni, 3, p.H, 2), dtype=np.float64) if (p.scheme in (3, 4) and want_eval) else None,
        "pubb": np.empty((p.l_uni, p.H, 2), dtype=np.float64) if (p.is_mk and want_eval) else None,
        "ksk": np.empty((p.N, p.ksk_rows, p.f, p.n + 1), dtype=np.uint32) if (want_ksk and want_eval) else None,
    }
    keyed = is_key(seed)
    fn = lib().mktfhe_host_party_keygen_key if keyed else lib().mktfhe_host_party_keygen
    rc = fn(ctypes.byref(cp), _kbuf(seed) if keyed else int(seed), party, ptr(crs_coeff), ptr(out["lwekey"]),
            ptr(out["ringkey"]), ptr(out["brk"]), ptr(out["rlk"]), ptr(out["pubb"]), ptr(out["ksk"]), nthreads)
    if rc != 0:
        raise RuntimeError(f"mktfhe_host_party_keygen failed: {rc}")
    return out


ENC_SINGLE, ENC_ITH, ENC_FULL = 0, 1, 2


def encrypt_batch(p: Params, seed, kind: int, party: int, bits, lwekeys, nthreads: int = 0) -> np.ndarray:
    """`count` ciphertexts in one native call (OpenMP over ciphertexts): scheme.jl:352-386 over a batch."""
    bits = np.ascontiguousarray(np.asarray(bits).astype(bool), dtype=np.uint8)
    out = np.empty((bits.shape[0], p.lwe_words), dtype=np.uint32)
    cp = p.c_struct()
    keys = np.ascontiguousarray(lwekeys, dtype=np.uint32)
    if is_key(seed):
        rc = lib().mktfhe_host_encrypt_batch_key(ctypes.byref(cp), _kbuf(seed), kind, party, ptr(bits), bits.shape[0], ptr(keys), ptr(out), nthreads)
    else:
        rc = lib().mktfhe_host_encrypt_batch(ctypes.byref(cp), int(seed), kind, party, ptr(bits), bits.shape[0], ptr(keys), ptr(out), nthreads)
    if rc != 0:
        raise ValueError(f"mktfhe_host_encrypt_batch failed: {rc}")
    return out


def decrypt_batch(p: Params, lwekeys, cts, nthreads: int = 0) -> np.ndarray:
    cts = np.ascontiguousarray(cts, dtype=np.uint32).reshape(-1, p.lwe_words)
    out = np.empty(cts.shape[0], dtype=np.uint8)
    cp = p.c_struct()
    rc = lib().mktfhe_host_decrypt_batch(ctypes.byref(cp), ptr(np.ascontiguousarray(lwekeys, dtype=np.uint32)), ptr(cts), cts.shape[0], ptr(out), nthreads)
    if rc != 0:
        raise ValueError(f"mktfhe_host_decrypt_batch failed: {rc}")
    return out.astype(bool)


def phase_batch(p: Params, lwekeys, cts, nthreads: int = 0) -> np.ndarray:
    cts = np.ascontiguousarray(cts, dtype=np.uint32).reshape(-1, p.lwe_words)
    out = np.empty(cts.shape[0], dtype=np.uint32)
    cp = p.c_struct()
    rc = lib().mktfhe_host_phase_batch(ctypes.byref(cp), ptr(np.ascontiguousarray(lwekeys, dtype=np.uint32)), ptr(cts), cts.shape[0], ptr(out), nthreads)
    if rc != 0:
        raise ValueError(f"mktfhe_host_phase_batch failed: {rc}")
    return out


def fft_tables(N: int):
    H = N // 2
    t = [np.empty((H, 2), dtype=np.float64) for _ in range(4)]
    lib().mktfhe_host_fft_tables(N, *[ptr(x) for x in t])
    return dict(zip(("psi", "psiinv", "roots", "rootsinv"), t))
