"""Named parameter sets, mirroring /root/reference/src/tfhe/params.jl:1-125 field for field.

`Params` is the Python face of `mktfhe_params` (include/mktfhe_params.h), which flattens the
reference's five parameter structs (src/tfhe/scheme.jl:6-101).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

CGGI, LMSS, CCS, KMS, KMS_BLOCK = range(5)
SCHEME_NAMES = {CGGI: "CGGI", LMSS: "LMSS", CCS: "CCS", KMS: "KMS", KMS_BLOCK: "KMS_block"}


class CParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("scheme", "n", "d", "ell", "f", "logD", "N", "k",
                 "l_gsw", "logB_gsw", "l_lev", "logB_lev", "l_uni", "logB_uni")] + \
               [("alpha", ctypes.c_double), ("beta", ctypes.c_double)]


@dataclass(frozen=True)
class Params:
    name: str
    scheme: int
    n: int
    N: int
    k: int
    alpha: float
    beta: float
    f: int = 8
    logD: int = 2
    d: int = 0
    ell: int = 0
    l_gsw: int = 0
    logB_gsw: int = 0
    l_lev: int = 0
    logB_lev: int = 0
    l_uni: int = 0
    logB_uni: int = 0

    # ---- derived -------------------------------------------------------------------------
    @property
    def H(self) -> int:
        return self.N // 2

    @property
    def torus_bits(self) -> int:
        return 64 if self.scheme in (KMS, KMS_BLOCK) else 32

    @property
    def is_mk(self) -> bool:
        return self.scheme in (CCS, KMS, KMS_BLOCK)

    @property
    def is_block(self) -> bool:
        return self.scheme in (LMSS, KMS_BLOCK)

    @property
    def ksk_rows(self) -> int:
        D = 1 << self.logD
        return D // 2 if self.is_block else D - 1

    @property
    def lwe_words(self) -> int:
        return 1 + self.n * self.k

    @property
    def brk_polys(self) -> int:
        return 3 * self.l_uni if self.scheme == CCS else 4 * self.l_gsw

    @property
    def brk_doubles(self) -> int:
        return self.n * self.brk_polys * self.N

    @property
    def rlk_doubles(self) -> int:
        return 3 * self.l_uni * self.N

    @property
    def pubb_doubles(self) -> int:
        return self.l_uni * self.N

    @property
    def ksk_words(self) -> int:
        return self.N * self.ksk_rows * self.f * (self.n + 1)

    def rows(self, party: int) -> int:
        """RLEV rows phase 1 rotates for `party` (0-based): bootstrapping.jl:400."""
        return 1 if party == 0 else self.l_lev

    def c_struct(self) -> CParams:
        return CParams(self.scheme, self.n, self.d, self.ell, self.f, self.logD, self.N, self.k,
                       self.l_gsw, self.logB_gsw, self.l_lev, self.logB_lev, self.l_uni, self.logB_uni,
                       float(self.alpha), float(self.beta))


def _bin(name, n, alpha, f, logD, N, k, beta, l, logB):
    return Params(name, CGGI, n=n, N=N, k=k, alpha=alpha, beta=beta, f=f, logD=logD, l_gsw=l, logB_gsw=logB)


def _block(name, d, ell, alpha, f, logD, N, k, beta, l, logB):
    return Params(name, LMSS, n=d * ell, d=d, ell=ell, N=N, k=k, alpha=alpha, beta=beta, f=f, logD=logD,
                  l_gsw=l, logB_gsw=logB)


def _ccs(name, n, alpha, f, logD, N, beta, l, logB, k):
    return Params(name, CCS, n=n, N=N, k=k, alpha=alpha, beta=beta, f=f, logD=logD, l_uni=l, logB_uni=logB)


def _kms(name, n, alpha, f, logD, N, beta, lg, bg, ll, bl, lu, bu, k):
    return Params(name, KMS, n=n, N=N, k=k, alpha=alpha, beta=beta, f=f, logD=logD,
                  l_gsw=lg, logB_gsw=bg, l_lev=ll, logB_lev=bl, l_uni=lu, logB_uni=bu)


def _kmsb(name, d, ell, alpha, f, logD, N, beta, lg, bg, ll, bl, lu, bu, k):
    return Params(name, KMS_BLOCK, n=d * ell, d=d, ell=ell, N=N, k=k, alpha=alpha, beta=beta, f=f, logD=logD,
                  l_gsw=lg, logB_gsw=bg, l_lev=ll, logB_lev=bl, l_uni=lu, logB_uni=bu)


A = float(1 << 17)
B64 = 85.4084

CGGIparam = _bin("CGGIparam", 630, A, 8, 2, 1 << 10, 1, float(1 << 7), 3, 9)                       # params.jl:1-6
Blockparam = _block("Blockparam", 229, 3, A, 8, 2, 1 << 10, 1, float(1 << 7), 3, 9)                # :8-13
CCS2party = _ccs("CCS2party", 560, A, 8, 2, 1 << 10, float(1 << 4), 3, 8, 2)                       # :15-21
CCS4party = _ccs("CCS4party", 560, A, 8, 2, 1 << 10, float(1 << 4), 4, 8, 4)                       # :23-29
CCS8party = _ccs("CCS8party", 560, A, 8, 2, 1 << 10, float(1 << 4), 5, 6, 8)                       # :31-37
CCS16party = _ccs("CCS16party", 560, A, 8, 2, 1 << 10, float(1 << 4), 12, 2, 16)                   # :39-45
KMS2party = _kms("KMS2party", 560, A, 8, 2, 1 << 11, B64, 3, 12, 2, 7, 3, 10, 2)                   # :47-53
KMS4party = _kms("KMS4party", 560, A, 8, 2, 1 << 11, B64, 5, 8, 2, 8, 7, 6, 4)                     # :55-61
KMS8party = _kms("KMS8party", 560, A, 8, 2, 1 << 11, B64, 4, 9, 3, 6, 8, 4, 8)                     # :63-69
KMS16party = _kms("KMS16party", 560, A, 8, 2, 1 << 11, B64, 5, 8, 3, 6, 9, 4, 16)                  # :71-77
KMS32party = _kms("KMS32party", 560, A, 8, 2, 1 << 11, B64, 6, 7, 3, 7, 16, 2, 32)                 # :79-85
KMS2partyblock = _kmsb("KMS2partyblock", 203, 3, A, 8, 2, 1 << 11, B64, 3, 12, 2, 7, 3, 10, 2)     # :87-93
KMS4partyblock = _kmsb("KMS4partyblock", 203, 3, A, 8, 2, 1 << 11, B64, 5, 8, 2, 8, 7, 6, 4)       # :95-101
KMS8partyblock = _kmsb("KMS8partyblock", 203, 3, A, 8, 2, 1 << 11, B64, 4, 9, 3, 6, 8, 4, 8)       # :103-109
KMS16partyblock = _kmsb("KMS16partyblock", 203, 3, A, 8, 2, 1 << 11, B64, 5, 8, 3, 6, 9, 4, 16)    # :111-117
KMS32partyblock = _kmsb("KMS32partyblock", 203, 3, A, 8, 2, 1 << 11, B64, 6, 7, 3, 7, 16, 2, 32)   # :119-125

ALL = {p.name: p for p in (
    CGGIparam, Blockparam, CCS2party, CCS4party, CCS8party, CCS16party,
    KMS2party, KMS4party, KMS8party, KMS16party, KMS32party,
    KMS2partyblock, KMS4partyblock, KMS8partyblock, KMS16partyblock, KMS32partyblock)}


def small(base: Params, **kw) -> Params:
    """A reduced copy of a named set for fast tests (e.g. fewer LWE coefficients); never used by bench.py."""
    from dataclasses import replace
    return replace(base, name=base.name + "-small", **kw)
