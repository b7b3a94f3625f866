"""numpy model of the 32-threads x 32-points transform of csrc/kernels_fast_w.cuh (index math only): checks that
slot n of the forward transform holds the reference's slot n (oracle.fft) and that the inverse undoes it."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O
from mpmath import mp, mpf, cos, sin, pi
mp.prec = 113
H, N = 1024, 2048

def tables():
    theta = [mpf(1) / 2]
    tw = np.zeros(1024, dtype=complex)
    for s in range(10):
        nxt = []
        for i, th in enumerate(theta):
            ang = pi * th / 2
            tw[(1 << s) + i] = complex(float(cos(ang)), float(-sin(ang)))
            nxt += [th / 2, th / 2 + 1]
        theta = nxt
    return tw

TW = tables()
T2 = np.zeros((16, 32), dtype=complex)
for t in range(32):
    T2[0, t] = TW[32 + t]
    T2[1, t] = TW[64 + 2 * t]
    for g in range(2): T2[2 + g, t] = TW[128 + 4 * t + 2 * g]
    for g in range(4): T2[4 + g, t] = TW[256 + 8 * t + 2 * g]
    for g in range(8): T2[8 + g, t] = TW[512 + 16 * t + 2 * g]

def bf(a, b, w): return a + w * b, a - w * b
def bf_mi(a, b, w): return bf(a, b, -1j * w)
def bi(a, b, w): return a + b, (a - b) * np.conj(w)
def bi_mi(a, b, w): return bi(a, b, -1j * w)

def fwd(c):            # c: 1024 complex in natural order; returns slots
    x = np.array([[c[t + 32 * m] for m in range(32)] for t in range(32)])      # x[t][m]
    for s in range(5):
        half = 16 >> s
        for m in range(32):
            if m & half: continue
            node = m >> (5 - s)
            w = TW[(1 << s) + (node & ~1)]
            f = bf_mi if node & 1 else bf
            x[:, m], x[:, m + half] = f(x[:, m], x[:, m + half], w)
    xb = np.zeros(1056, dtype=complex)
    for t in range(32):
        for m in range(32): xb[t + 33 * m] = x[t, m]
    y = np.array([[xb[33 * t + e] for e in range(32)] for t in range(32)])      # y[t][e] = slot 32t+e
    for e in range(16):
        y[:, e], y[:, e + 16] = bf(y[:, e], y[:, e + 16], T2[0])
    for s in range(6, 10):
        half = 16 >> (s - 5)
        for e in range(32):
            if e & half: continue
            sub = e >> (10 - s)
            row = (1 << (s - 6)) + (sub >> 1)
            f = bf_mi if sub & 1 else bf
            y[:, e], y[:, e + half] = f(y[:, e], y[:, e + half], T2[row])
    return y.reshape(-1)

def inv(slots):        # unscaled inverse
    y = slots.reshape(32, 32).copy()
    for s in range(9, 5, -1):
        half = 16 >> (s - 5)
        for e in range(32):
            if e & half: continue
            sub = e >> (10 - s)
            row = (1 << (s - 6)) + (sub >> 1)
            f = bi_mi if sub & 1 else bi
            y[:, e], y[:, e + half] = f(y[:, e], y[:, e + half], T2[row])
    for e in range(16):
        y[:, e], y[:, e + 16] = bi(y[:, e], y[:, e + 16], T2[0])
    xb = np.zeros(1056, dtype=complex)
    for t in range(32):
        for e in range(32): xb[33 * t + e] = y[t, e]
    x = np.array([[xb[t + 33 * m] for m in range(32)] for t in range(32)])
    for s in range(4, -1, -1):
        half = 16 >> s
        for m in range(32):
            if m & half: continue
            node = m >> (5 - s)
            w = TW[(1 << s) + (node & ~1)]
            f = bi_mi if node & 1 else bi
            x[:, m], x[:, m + half] = f(x[:, m], x[:, m + half], w)
    c = np.zeros(1024, dtype=complex)
    for t in range(32):
        for m in range(32): c[t + 32 * m] = x[t, m]
    return c

def brv(v, bits): return int(format(v, f"0{bits}b")[::-1], 2)

if __name__ == "__main__":
    rng = np.random.default_rng(1)
    p = rng.integers(0, 1 << 20, N).astype(np.uint64)          # small coefficients: exact in double
    ref = O.fft(p); ref = ref[:, 0] + 1j * ref[:, 1]
    sp = p.astype(np.int64).astype(float)
    c = sp[:H] - 1j * sp[H:]
    got = fwd(c)
    print("forward max rel err vs oracle slots:", np.abs(got - ref).max() / np.abs(ref).max())
    back = inv(got) / H
    print("inverse round trip:", np.abs(back - c).max())
    # monomial: slot n = 32t+e evaluates at exp(-i pi (4 brv10(n)+1)/N); brv10(n) = 32 brv5(e) + brv5(t)
    a = 777
    mono = O.monomials(N)[a - 1]; mono = mono[:, 0] + 1j * mono[:, 1]          # table[a] = FFT(X^a - 1), 1-based a
    worst = 0
    for t in range(32):
        m1 = np.exp(-1j * np.pi * (((4 * brv(t, 5) + 1) * a) % 4096) / 2048)
        for e in range(32):
            z = m1 * np.exp(-1j * np.pi * ((a * brv(e, 5)) % 32) / 16) - 1
            worst = max(worst, abs(z - mono[32 * t + e]))
    print("monomial closed form max err:", worst)
