"""numpy model of the inverse transform used by csrc/kernels_fast_w.cuh since round 2: a plain cyclic decimation-in-time
network on the slot array followed by an untwist, instead of walking the forward product tree backwards.

Slot n of the forward transform holds c(exp(-i pi (4 brv10(n) + 1) / N)) = sum_k (c_k zeta^k) exp(-2 pi i k f / H) with
zeta = exp(-i pi / N), f = brv10(n): the slot array IS the bit-reversed spectrum of the twisted sequence c_k zeta^k.  A DIT
network takes bit-reversed input to natural-order output with multiply-then-add butterflies (6 FMA-pipe instructions, the
product tree's add-then-multiply inverse butterfly needs 8), its first two stages have twiddles 1 and i only (4 additions), and
the untwist zeta^{-k} costs one complex multiplication per point: 31.9k FP64 instructions per transform instead of 41.0k.

Thread mapping (as in the kernel): stages 0..4 on the thread's 32 contiguous slots 32t+e (twiddles are compile-time constants),
one warp-local transposition, stages 5..9 on elements t+32m (per-thread twiddles exp(2 pi i (t + 32 m') / 2^(s+1)), siblings
m' + dm/2 differ by the factor i), untwist with T[k] = exp(i pi k / 2048), k = t + 32m; T[k] for k > 512 is T[1024-k] with real
and imaginary parts swapped."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fft32_model import fwd, H, N

def bf(a, b, w): return a + w * b, a - w * b

def counts():
    n = 0
    for s in range(5):
        half = 1 << s
        for e in range(32):
            if e & half: continue
            j = e & (half - 1)
            trivial = (4 * j) % (2 * half) == 0 if half > 1 else True      # w in {1, i}
            n += 32 * (4 if trivial else 6)
    n += 5 * 512 * 6 + 4 * 1024
    return n

def inv_dit(slots):
    y = slots.reshape(32, 32).copy()                       # y[t][e] = slot 32t + e
    for s in range(5):
        half = 1 << s
        for e in range(32):
            if e & half: continue
            w = np.exp(2j * np.pi * (e & (half - 1)) / (2 * half))
            y[:, e], y[:, e + half] = bf(y[:, e], y[:, e + half], w)
    x = y.reshape(-1).reshape(32, 32).T.copy()             # x[t][m] = element t + 32m
    t = np.arange(32)
    for s in range(5, 10):
        dm = 1 << (s - 5)
        for m in range(32):
            if m & dm: continue
            mp = m & (dm - 1)
            if s >= 6 and mp >= dm // 2:                   # sibling: i * w(m' - dm/2)
                w = 1j * np.exp(2j * np.pi * (t + 32 * (mp - dm // 2)) / (1 << (s + 1)))
            else:
                w = np.exp(2j * np.pi * (t + 32 * mp) / (1 << (s + 1)))
            x[:, m], x[:, m + dm] = bf(x[:, m], x[:, m + dm], w)
    T = np.exp(1j * np.pi * np.arange(513) / 2048)
    c = np.zeros(1024, dtype=complex)
    for tt in range(32):
        for m in range(32):
            k = tt + 32 * m
            tw = T[k] if k <= 512 else complex(T[1024 - k].imag, T[1024 - k].real)
            c[k] = x[tt, m] * tw
    return c

if __name__ == "__main__":
    rng = np.random.default_rng(2)
    p = rng.integers(-(1 << 20), 1 << 20, N).astype(float)
    c = p[:H] - 1j * p[H:]
    back = inv_dit(fwd(c)) / H
    print("DIT inverse of the forward transform, max abs error:", np.abs(back - c).max())
    print("FP64 instructions per inverse:", counts(), "(product-tree inverse: 40960, forward: 30720)")
