"""ctypes wrapper of the CPU oracle (oracle/mktfhe_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libmktfhe_oracle.so")


class OrcParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("scheme", "n", "d", "ell", "f", "logD", "N", "k",
                 "l_gsw", "logB_gsw", "l_lev", "logB_lev", "l_uni", "logB_uni")] + \
               [("alpha", ctypes.c_double), ("beta", ctypes.c_double)]


class OrcKeys(ctypes.Structure):
    _fields_ = [("brk", ctypes.POINTER(ctypes.c_void_p)), ("rlk", ctypes.POINTER(ctypes.c_void_p)),
                ("pubb", ctypes.POINTER(ctypes.c_void_p)), ("ksk", ctypes.POINTER(ctypes.c_void_p)),
                ("crs", ctypes.c_void_p)]


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in ("mktfhe_oracle.c", "orc_ring.inc", "mktfhe_oracle.h", "Makefile")]
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        vp, i32, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32
        _lib.orc_create.restype = vp
        _lib.orc_create.argtypes = [ctypes.POINTER(OrcParams), ctypes.POINTER(OrcKeys)]
        _lib.orc_destroy.argtypes = [vp]
        _lib.orc_fft_tables.argtypes = [i32, vp, vp, vp, vp]
        _lib.orc_monomials.argtypes = [i32, vp]
        _lib.orc_fft.argtypes = [i32, i32, vp, vp]
        _lib.orc_ifft.argtypes = [i32, i32, vp, vp]
        _lib.orc_decomp.argtypes = [i32, i32, i32, i32, vp, vp]
        _lib.orc_modswitch.argtypes = [vp, vp, vp]
        _lib.orc_gate_linear.argtypes = [vp, i32, vp, vp, vp]
        _lib.orc_cmux_step.argtypes = [vp, i32, i32, u32, vp]
        _lib.orc_block_step.argtypes = [vp, i32, i32, vp, vp]
        _lib.orc_block_step.restype = None
        _lib.orc_phase1.argtypes = [vp, i32, vp, vp]
        _lib.orc_blindrotate.argtypes = [vp, vp, vp]
        _lib.orc_phase2.argtypes = [vp, vp, u32, vp]
        _lib.orc_keyswitch.argtypes = [vp, vp, vp]
        _lib.orc_bootstrap.argtypes = [vp, vp]
        _lib.orc_gate_batch.argtypes = [vp, i32, vp, vp, vp, ctypes.c_size_t, i32]
        _lib.orc_max_threads.restype = i32
        for f in ("orc_destroy", "orc_fft_tables", "orc_monomials", "orc_fft", "orc_ifft", "orc_decomp", "orc_modswitch",
                  "orc_gate_linear", "orc_cmux_step", "orc_phase1", "orc_blindrotate", "orc_phase2", "orc_keyswitch",
                  "orc_bootstrap", "orc_gate_batch"):
            getattr(_lib, f).restype = None
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be C-contiguous"
    return a.ctypes.data_as(ctypes.c_void_p)


def fft_tables(N):
    H = N // 2
    t = [np.empty((H, 2), dtype=np.float64) for _ in range(4)]
    lib().orc_fft_tables(N, *[_p(x) for x in t])
    return dict(zip(("psi", "psiinv", "roots", "rootsinv"), t))


def monomials(N):
    out = np.empty((2 * N, N // 2, 2), dtype=np.float64)
    lib().orc_monomials(N, _p(out))
    return out


def fft(poly):
    poly = np.ascontiguousarray(poly)
    bits = poly.dtype.itemsize * 8
    out = np.empty((poly.shape[0] // 2, 2), dtype=np.float64)
    lib().orc_fft(poly.shape[0], bits, _p(poly), _p(out))
    return out


def ifft(t, bits):
    t = np.array(t, dtype=np.float64, order="C", copy=True)
    N = t.shape[0] * 2
    out = np.empty(N, dtype=np.uint64 if bits == 64 else np.uint32)
    lib().orc_ifft(N, bits, _p(t), _p(out))
    return out


def decomp(poly, l, logB):
    poly = np.ascontiguousarray(poly)
    bits = poly.dtype.itemsize * 8
    out = np.empty((l, poly.shape[0]), dtype=poly.dtype)
    lib().orc_decomp(poly.shape[0], bits, l, logB, _p(poly), _p(out))
    return out


class Oracle:
    """Oracle context over one key set (arrays as produced by mktfhe_b200.scheme.KeySet)."""

    def __init__(self, params, brk, ksk, rlk=None, pubb=None, crs=None):
        self.p = params
        k = len(brk)
        self._keep = (brk, ksk, rlk, pubb, crs)
        cp = OrcParams(*[getattr(params, n) for n in
                         ("scheme", "n", "d", "ell", "f", "logD", "N", "k", "l_gsw", "logB_gsw", "l_lev", "logB_lev",
                          "l_uni", "logB_uni")], float(params.alpha), float(params.beta))

        def arr(lst):
            if lst is None or lst[0] is None:
                return None
            a = (ctypes.c_void_p * k)(*[x.ctypes.data for x in lst])
            return a
        self._arrs = [arr(brk), arr(rlk), arr(pubb), arr(ksk)]
        keys = OrcKeys(*[ctypes.cast(a, ctypes.POINTER(ctypes.c_void_p)) if a is not None else None for a in self._arrs],
                       crs.ctypes.data if crs is not None else None)
        self.h = lib().orc_create(ctypes.byref(cp), ctypes.byref(keys))
        self.lwe_words = 1 + params.n * params.k
        self.tdtype = np.uint64 if params.scheme in (3, 4) else np.uint32

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def modswitch(self, lwe):
        out = np.empty(self.lwe_words, dtype=np.uint32)
        lib().orc_modswitch(self.h, _p(np.ascontiguousarray(lwe)), _p(out))
        return out

    def gate_linear(self, op, a, b):
        out = np.empty(self.lwe_words, dtype=np.uint32)
        lib().orc_gate_linear(self.h, op, _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out))
        return out

    def cmux_step(self, party, idx, atilde, acc_row):
        acc = np.array(acc_row, dtype=self.tdtype, order="C", copy=True)
        lib().orc_cmux_step(self.h, party, idx, int(atilde), _p(acc))
        return acc

    def block_step(self, party, blk, at, acc_row):
        acc = np.array(acc_row, dtype=self.tdtype, order="C", copy=True)
        lib().orc_block_step(self.h, party, blk, _p(np.ascontiguousarray(at, dtype=np.uint32)), _p(acc))
        return acc

    def phase1(self, party, tildea_party):
        rows = 1 if party == 0 else self.p.l_lev
        out = np.empty((rows, 2, self.p.N // 2, 2), dtype=np.float64)
        lib().orc_phase1(self.h, party, _p(np.ascontiguousarray(tildea_party, dtype=np.uint32)), _p(out))
        return out

    def phase2(self, levkeys, btilde):
        acc = np.empty((self.p.k + 1, self.p.N), dtype=np.uint64)
        lib().orc_phase2(self.h, _p(np.ascontiguousarray(levkeys)), int(btilde), _p(acc))
        return acc

    def blindrotate(self, lwe):
        acc = np.empty((self.p.k + 1, self.p.N), dtype=self.tdtype)
        lib().orc_blindrotate(self.h, _p(np.ascontiguousarray(lwe, dtype=np.uint32)), _p(acc))
        return acc

    def keyswitch(self, acc):
        out = np.empty(self.lwe_words, dtype=np.uint32)
        lib().orc_keyswitch(self.h, _p(np.ascontiguousarray(acc, dtype=self.tdtype)), _p(out))
        return out

    def bootstrap(self, lwe):
        out = np.array(lwe, dtype=np.uint32, order="C", copy=True)
        lib().orc_bootstrap(self.h, _p(out))
        return out

    def gate_batch(self, op, in1, in2, nthreads=0):
        """op >= 0: gate + bootstrap; op < 0: bootstrap of in1 only."""
        in1 = np.ascontiguousarray(in1, dtype=np.uint32)
        in2 = np.ascontiguousarray(in2, dtype=np.uint32) if in2 is not None else in1
        out = np.empty_like(in1)
        lib().orc_gate_batch(self.h, op, _p(in1), _p(in2), _p(out), in1.shape[0], nthreads)
        return out


def max_threads():
    return lib().orc_max_threads()
