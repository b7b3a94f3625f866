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


def crs(p: Params, seed: int):
    cp = p.c_struct()
    coeff = np.empty((p.l_uni, p.N), dtype=torus_dtype(p))
    fft = np.empty((p.l_uni, p.H, 2), dtype=np.float64)
    rc = lib().mktfhe_host_crs(ctypes.byref(cp), seed, ptr(coeff), ptr(fft))
    if rc != 0:
        raise RuntimeError(f"mktfhe_host_crs failed: {rc}")
    return coeff, fft


def party_keygen(p: Params, seed: int, party: int, crs_coeff, nthreads: int = 0, want_ksk: bool = True,
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
    rc = lib().mktfhe_host_party_keygen(ctypes.byref(cp), seed, party, ptr(crs_coeff), ptr(out["lwekey"]),
                                        ptr(out["ringkey"]), ptr(out["brk"]), ptr(out["rlk"]), ptr(out["pubb"]),
                                        ptr(out["ksk"]), nthreads)
    if rc != 0:
        raise RuntimeError(f"mktfhe_host_party_keygen failed: {rc}")
    return out


def fft_tables(N: int):
    H = N // 2
    t = [np.empty((H, 2), dtype=np.float64) for _ in range(4)]
    lib().mktfhe_host_fft_tables(N, *[ptr(x) for x in t])
    return dict(zip(("psi", "psiinv", "roots", "rootsinv"), t))
