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
