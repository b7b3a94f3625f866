"""Pins the CPU oracle (oracle/) on mathematical ground truth.  The reference ships no golden vectors and Julia
is not installed, so these identities -- together with the reference's own acceptance criterion in
test_oracle_gates.py -- are what the oracle is pinned against (oracle/mktfhe_oracle.h)."""
import mpmath as mp
import numpy as np
import pytest

from oracle import oracle as O
from mktfhe_b200 import _host


def brv(x, bits):
    return int(format(x, f"0{bits}b")[::-1], 2)


@pytest.mark.parametrize("N", [1024, 2048])
def test_tables_are_correctly_rounded_and_shared(N):
    """fft.jl:26-44: tables are BigFloat exp() rounded to Float64.  Oracle, host library and mpmath agree bit for bit."""
    H = N // 2
    t = O.fft_tables(N)
    h = _host.fft_tables(N)
    for k in t:
        assert np.array_equal(t[k].view(np.uint64), h[k].view(np.uint64)), k
    mp.mp.prec = 200
    logH = H.bit_length() - 1
    for j in list(range(0, H, 37)) + [1, H - 1]:
        th = mp.pi * j / H
        r = brv(j, logH)          # table position after bit_reverse!
        assert t["psi"][r, 0] == float(mp.cos(th)) and t["psi"][r, 1] == float(-mp.sin(th))
        assert t["psiinv"][r, 1] == float(mp.sin(th))
        ph = mp.pi * j / N
        assert t["roots"][j, 0] == float(mp.cos(ph)) and t["roots"][j, 1] == float(mp.sin(ph))
        assert t["rootsinv"][j, 0] == float(mp.cos(ph) / H) and t["rootsinv"][j, 1] == float(-mp.sin(ph) / H)


@pytest.mark.parametrize("N,dt", [(1024, np.uint32), (2048, np.uint64)])
def test_fft_slot_j_is_evaluation_at_Zj(N, dt):
    """slot j holds p(Z_j), Z_j = exp(-i*pi*(4*brv(j)+1)/N)."""
    H = N // 2
    rng = np.random.default_rng(1)
    p = rng.integers(-2 ** 20, 2 ** 20, size=N).astype(np.int64)
    spec = O.fft(p.astype(dt))
    spec = spec[:, 0] + 1j * spec[:, 1]
    logH = H.bit_length() - 1
    n = np.arange(N)
    for j in (0, 1, 5, H // 2, H - 1):
        z = np.exp(-1j * np.pi * (4 * brv(j, logH) + 1) / N)
        want = np.sum(p * z ** n)
        assert abs(spec[j] - want) < 1e-6 * np.abs(p).sum()


def negacyclic(a, b, mod_bits):
    N = len(a)
    res = [0] * N
    for i in range(N):
        if b[i] == 0:
            continue
        for j in range(N):
            k = i + j
            v = int(a[j]) * int(b[i])
            if k >= N:
                res[k - N] -= v
            else:
                res[k] += v
    return np.array([r % (1 << mod_bits) for r in res], dtype=object)


def test_fft_product_is_negacyclic_convolution_u32():
    """ifft(fft(a) .* fft(b)) equals the product in Z_{2^32}[X]/(X^N+1) when the true result needs < 53 bits --
    up to one unit, because `native` (arithmetic.jl:1-4) truncates toward -inf instead of rounding to nearest."""
    N = 1024
    rng = np.random.default_rng(2)
    a = rng.integers(0, 2 ** 32, size=N, dtype=np.uint32)
    b = np.zeros(N, dtype=np.int64)
    b[rng.choice(N, 24, replace=False)] = rng.integers(-256, 256, 24)
    fa, fb = O.fft(a), O.fft(b.astype(np.uint32))
    ca, cb = fa[:, 0] + 1j * fa[:, 1], fb[:, 0] + 1j * fb[:, 1]
    prod = ca * cb
    got = O.ifft(np.stack([prod.real, prod.imag], axis=1), 32)
    sa = a.astype(np.int64)
    sa[sa >= 2 ** 31] -= 2 ** 32            # the transform reads coefficients as signed
    want = negacyclic(sa, b, 32)
    diff = [((int(g) - int(w) + (1 << 31)) % (1 << 32)) - (1 << 31) for g, w in zip(got, want)]
    assert set(diff) <= {-1, 0} and diff.count(0) > 0


def test_fft_product_u64_within_float64_rounding():
    N = 2048
    rng = np.random.default_rng(3)
    a = rng.integers(0, 2 ** 64, size=N, dtype=np.uint64)
    b = np.zeros(N, dtype=np.int64)
    b[rng.choice(N, 16, replace=False)] = rng.integers(-2048, 2048, 16)
    fa, fb = O.fft(a), O.fft(b.astype(np.uint64))
    prod = (fa[:, 0] + 1j * fa[:, 1]) * (fb[:, 0] + 1j * fb[:, 1])
    got = O.ifft(np.stack([prod.real, prod.imag], axis=1), 64)
    sa = a.astype(object)
    sa = np.array([int(x) - (1 << 64) if int(x) >= (1 << 63) else int(x) for x in sa], dtype=object)
    want = negacyclic(sa, b, 64)
    diff = np.array([((int(g) - int(w) + (1 << 63)) % (1 << 64)) - (1 << 63) for g, w in zip(got, want)], dtype=np.float64)
    assert np.abs(diff).max() < 2.0 ** 34       # 64-bit inputs carry only 53 bits through the transform


@pytest.mark.parametrize("N,dt,bits", [(1024, np.uint32, 32)])
def test_fft_roundtrip_for_small_polys(N, dt, bits):
    """Round trip returns p or p - 1 per coefficient: `native` floors, so a value that comes back as 41.99999 is 41."""
    rng = np.random.default_rng(4)
    p = rng.integers(1, 2 ** 12, size=N).astype(dt)
    back = O.ifft(O.fft(p), bits).astype(np.int64)
    d = back - p.astype(np.int64)
    assert set(np.unique(d)) <= {-1, 0} and (d == 0).sum() > N // 4


@pytest.mark.parametrize("dt,l,logB", [(np.uint32, 3, 9), (np.uint32, 4, 8), (np.uint32, 12, 2), (np.uint64, 3, 12),
                                      (np.uint64, 2, 7), (np.uint64, 16, 2), (np.uint64, 6, 7), (np.uint64, 8, 4)])
def test_decomposition_digits_balanced_and_recompose(dt, l, logB):
    """gsw.jl:86-96: digits in [-B/2, B/2), sum d_j * 2^(w - j*logB) = a up to 2^(w - l*logB - 1) (mod 2^w)."""
    w = np.dtype(dt).itemsize * 8
    rng = np.random.default_rng(5)
    a = rng.integers(0, 2 ** w, size=1024, dtype=dt)
    a[:4] = [0, np.iinfo(dt).max, 1 << (w - 1), (1 << (w - 1)) - 1]
    d = O.decomp(a, l, logB)
    sd = d.astype(np.int64) if w == 64 else d.astype(np.int32).astype(np.int64)
    B = 1 << logB
    assert sd.min() >= -B // 2 and sd.max() < B // 2
    for c in range(len(a)):
        rec = sum(int(sd[j, c]) << (w - (j + 1) * logB) for j in range(l)) % (1 << w)
        err = (rec - int(a[c]) + (1 << (w - 1))) % (1 << w) - (1 << (w - 1))
        bound = (1 << (w - l * logB - 1)) if w > l * logB else 0
        assert abs(err) <= bound, (c, err, bound)


def test_monomial_table_matches_closed_form():
    """scheme.jl:121-146: entry a-1 = FFT(X^a - 1) = Z_j^a - 1; entry 2N-1 = 0."""
    N = 1024
    H = N // 2
    mono = O.monomials(N)
    mono = mono[..., 0] + 1j * mono[..., 1]
    assert np.all(mono[2 * N - 1] == 0)
    j = np.arange(H)
    r = np.array([brv(x, 9) for x in j])
    for a in (1, 2, N - 1, N, N + 1, 2 * N - 1):
        Za = np.exp(-1j * np.pi * (((4 * r + 1) * a) % (2 * N)) / N)        # Z_j^a with exact angle reduction
        assert np.abs(mono[a - 1] - (Za - 1)).max() < 1e-12


def test_modswitch_can_return_2N_and_gate_linear_wraps():
    """App. A.2: divbits is not reduced; a = 0xFFFFFFFF rounds to 2N."""
    from conftest import keyset, make_oracle
    ks = keyset("CGGIparam")
    orc = make_oracle(ks)
    p = ks.params
    ct = np.zeros(p.lwe_words, dtype=np.uint32)
    ct[0] = 0xFFFFFFFF
    ct[1] = 0xFFE00000
    ct[2] = 0x000FFFFF
    t = orc.modswitch(ct)
    assert t[0] == 2 * p.N and t[1] == 2 * p.N - 1 and t[2] == 0
    one = np.full(p.lwe_words, 1, dtype=np.uint32)
    nand = orc.gate_linear(0, one, one)
    assert nand[0] == (1 << 29) - 2 and nand[1] == 0xFFFFFFFE
    xnor = orc.gate_linear(4, one, one)
    assert xnor[0] == ((3 << 30) - 4) & 0xFFFFFFFF and xnor[1] == (-4) & 0xFFFFFFFF
