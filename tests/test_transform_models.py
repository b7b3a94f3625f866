"""The index math of the FAST transforms (which thread holds which point / slot in which pass, which twiddle it needs) is modelled
in numpy under tools/models/ and checked against the oracle's transform (oracle.fft = the reference's fftto!, fft.jl:57-63):
slot n of every mapping must hold the reference's slot n, the inverse must undo the forward transform, and the per-slot closed form
of the monomial X^a - 1 must equal the reference's table (scheme.jl:121-146).  CPU only."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", "models", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name,bits", [("fft32_model", 64), ("fft16x32_model", 32)])
def test_forward_slots_are_the_reference_slots_and_inverse_undoes_them(name, bits):
    from oracle import oracle as O
    m = _load(name)
    rng = np.random.default_rng(3)
    p = rng.integers(0, 1 << 20, m.N).astype(np.uint64 if bits == 64 else np.uint32)      # small coefficients: exact in double
    ref = O.fft(p)
    ref = ref[:, 0] + 1j * ref[:, 1]
    sp = p.astype(np.int64).astype(float)
    c = sp[:m.H] - 1j * sp[m.H:]
    got = m.fwd(c)
    assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.abs(m.inv(got) / m.H - c).max() < 1e-6


def test_decimation_in_time_inverse_undoes_the_forward_transform():
    f = _load("fft32_model")
    d = _load("ifft_dit_model")
    rng = np.random.default_rng(4)
    p = rng.integers(-(1 << 20), 1 << 20, f.N).astype(float)
    c = p[:f.H] - 1j * p[f.H:]
    assert np.abs(d.inv_dit(f.fwd(c)) / f.H - c).max() < 1e-6
    assert d.counts() == 31872


@pytest.mark.parametrize("name,tbits", [("fft32_model", 5), ("fft16x32_model", 4)])
def test_monomial_closed_form_per_thread_and_register_slot(name, tbits):
    from oracle import oracle as O
    m = _load(name)
    table = O.monomials(m.N)
    for a in (1, 2, 777, m.N - 1, m.N, m.N + 1, 2 * m.N - 1):
        mono = table[a - 1]
        mono = mono[:, 0] + 1j * mono[:, 1]
        worst = 0.0
        for t in range(1 << tbits):
            m1 = np.exp(-1j * np.pi * (((4 * m.brv(t, tbits) + 1) * a) % (2 * m.N)) / m.N)
            for e in range(32):
                z = m1 * np.exp(-1j * np.pi * ((a * m.brv(e, 5)) % 32) / 16) - 1
                worst = max(worst, abs(z - mono[32 * t + e]))
        assert worst < 1e-12, (name, a, worst)


def _gadgets():
    """Every (word width, l, logB) a FAST kernel decomposes with: the three gadgets of each named parameter set (params.jl)."""
    from mktfhe_b200 import params as P
    seen = set()
    for p in P.ALL.values():
        w = 64 if p.N == 2048 else 32
        for l, b in ((p.l_gsw, p.logB_gsw), (p.l_lev, p.logB_lev), (p.l_uni, p.logB_uni)):
            if l:
                seen.add((w, l, b))
    return sorted(seen)


@pytest.mark.parametrize("w,l,logB", _gadgets())
def test_one_add_field_extraction_equals_the_reference_digits(w, l, logB):
    """FAST decomposition (kernels_fast*.cuh): the accumulator is kept with ONE constant added -- half an ulp of the kept precision
    (divbits rounding, arithmetic.jl:23-27) plus B/2 at every digit position -- after which digit j is the plain bit field
    ((x + cadd) >> shift_j) & (B - 1), minus B/2.  Must equal the reference's rounding + carry chain (gsw.jl:86-96) digit for digit,
    including at the wrap-around values, for every gadget of params.jl."""
    from oracle import oracle as O
    dt = np.uint64 if w == 64 else np.uint32
    rng = np.random.default_rng(w * 1000 + l * 37 + logB)
    x = rng.integers(0, 2 ** w, size=4096, dtype=dt)
    bit = w - l * logB
    edges = [0, 1, 2 ** w - 1, 2 ** (w - 1), 2 ** (w - 1) - 1, 2 ** (w - 1) + 1]
    for j in range(l + 1):                       # values sitting exactly on and next to every rounding / carry boundary
        pos = bit + j * logB - 1
        if pos >= 0:
            edges += [(1 << pos) % 2 ** w, ((1 << pos) - 1) % 2 ** w, (2 ** w - (1 << pos)) % 2 ** w, (2 ** w - (1 << pos) - 1) % 2 ** w]
    x[:len(edges)] = np.array(edges, dtype=object).astype(dt)
    ref = O.decomp(x, l, logB)                   # [l][n], two's complement in the word type, digit 0 most significant
    cadd = (1 << (bit - 1)) if bit > 0 else 0
    for j in range(l):
        cadd += 1 << (bit + j * logB + logB - 1)
    mask, half = (1 << logB) - 1, 1 << (logB - 1)
    for j in range(l):
        sh = bit + (l - 1 - j) * logB
        got = [(((int(v) + cadd) % 2 ** w) >> sh & mask) - half for v in x]
        want = ref[j].astype(np.int64) if w == 64 else ref[j].astype(np.int32).astype(np.int64)
        assert got == [int(v) for v in want], (w, l, logB, j)
