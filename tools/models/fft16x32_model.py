"""numpy model of the 16-threads x 32-points transform of csrc/kernels_fast32_w.cuh (H = 512, N = 1024; index math only):
slot n of the forward transform must hold the reference's slot n (oracle.fft of a UInt32 polynomial), the inverse must undo it,
and the monomial closed form per (thread, register slot) must equal the reference's table.
Passes: 5 stages on elements t + 16m (m < 32), one transposition inside the half-warp, 4 stages on the thread's 32 contiguous
slots 32t + e (two 16-point blocks)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle as O
from mpmath import mp, mpf, cos, sin, pi
mp.prec = 113
H, N, T = 512, 1024, 16

def tables():
    theta = [mpf(1) / 2]
    tw = np.zeros(H, dtype=complex)
    for s in range(9):
        nxt = []
        for i, th in enumerate(theta):
            ang = pi * th / 2
            tw[(1 << s) + i] = complex(float(cos(ang)), float(-sin(ang)))
            nxt += [th / 2, th / 2 + 1]
        theta = nxt
    return tw

TW = tables()
# pass-2 table [15][16]: stage 6 row 0: TW[32 + 2t]; stage 7 rows 1, 2: TW[64 + 4t + 2g]; stage 8 rows 3..6: TW[128 + 8t + 2g];
# stage 9 rows 7..14: TW[256 + 16t + 2g]
T2 = np.zeros((15, T), dtype=complex)
for t in range(T):
    T2[0, t] = TW[32 + 2 * t]
    for g in range(2): T2[1 + g, t] = TW[64 + 4 * t + 2 * g]
    for g in range(4): T2[3 + g, t] = TW[128 + 8 * t + 2 * g]
    for g in range(8): T2[7 + g, t] = TW[256 + 16 * t + 2 * g]

def bf(a, b, w): return a + w * b, a - w * b
def bf_mi(a, b, w): return bf(a, b, -1j * w)
def bi(a, b, w): return a + b, (a - b) * np.conj(w)
def bi_mi(a, b, w): return bi(a, b, -1j * w)

def pass2(y, fwd_dir):
    stages = range(6, 10) if fwd_dir else range(9, 5, -1)
    for s in stages:
        half = 1 << (9 - s)
        for e in range(32):
            if e & half: continue
            node = e >> (10 - s)
            row = ((1 << (s - 6)) - 1) + (node >> 1)
            if fwd_dir: f = bf_mi if node & 1 else bf
            else: f = bi_mi if node & 1 else bi
            y[:, e], y[:, e + half] = f(y[:, e], y[:, e + half], T2[row])
    return y

def pass1(x, fwd_dir):
    stages = range(5) if fwd_dir else range(4, -1, -1)
    for s in stages:
        half = 16 >> s
        for m in range(32):
            if m & half: continue
            node = m >> (5 - s)
            w = TW[(1 << s) + (node & ~1)]
            if fwd_dir: f = bf_mi if node & 1 else bf
            else: f = bi_mi if node & 1 else bi
            x[:, m], x[:, m + half] = f(x[:, m], x[:, m + half], w)
    return x

def fwd(c):            # c: 512 complex in natural order; returns slots
    x = np.array([[c[t + T * m] for m in range(32)] for t in range(T)])         # x[t][m] = point t + 16m
    x = pass1(x, True)
    pos = np.zeros(H, dtype=complex)
    for t in range(T):
        for m in range(32): pos[t + T * m] = x[t, m]
    y = pos.reshape(T, 32).copy()                                               # y[t][e] = position 32t + e
    return pass2(y, True).reshape(-1)

def inv(slots):        # unscaled inverse
    y = pass2(slots.reshape(T, 32).copy(), False)
    pos = y.reshape(-1)
    x = np.array([[pos[t + T * m] for m in range(32)] for t in range(T)])
    x = pass1(x, False)
    c = np.zeros(H, dtype=complex)
    for t in range(T):
        for m in range(32): c[t + T * m] = x[t, m]
    return c

def brv(v, bits): return int(format(v, f"0{bits}b")[::-1], 2)

if __name__ == "__main__":
    rng = np.random.default_rng(1)
    p = rng.integers(0, 1 << 20, N).astype(np.uint32)
    ref = O.fft(p); ref = ref[:, 0] + 1j * ref[:, 1]
    sp = p.astype(np.int64).astype(float)
    c = sp[:H] - 1j * sp[H:]
    got = fwd(c)
    print("forward max rel err vs oracle slots:", np.abs(got - ref).max() / np.abs(ref).max())
    print("inverse round trip:", np.abs(inv(got) / H - c).max())
    a = 777
    mono = O.monomials(N)[a - 1]; mono = mono[:, 0] + 1j * mono[:, 1]
    worst = 0
    for t in range(T):
        m1 = np.exp(-1j * np.pi * (((4 * brv(t, 4) + 1) * a) % 2048) / 1024)
        for e in range(32):
            z = m1 * np.exp(-1j * np.pi * ((a * brv(e, 5)) % 32) / 16) - 1
            worst = max(worst, abs(z - mono[32 * t + e]))
    print("monomial closed form max err:", worst)
