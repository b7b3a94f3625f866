"""FAST KMS phase 2 (fast::k_phase2) chains gadget products and re-decomposes their outputs, so its coefficients cannot be compared
with another rounding over the whole phase (one digit flip changes everything downstream).  The products themselves
can: feed the ORACLE's intermediate polynomial (the accumulator component, y = ifft(ty), v) into ONE product built from the same
device functions as k_phase2 and compare with the reference-order product within a stated Torus64 tolerance.
Reference: /root/reference/src/tfhe/bootstrapping.jl:483-499 (LEV product with levkey[idx]), :520-535 (u, v from y), :538-550 (w from v)."""
import numpy as np
import pytest

from conftest import fresh_inputs, keyset, make_oracle
from mktfhe_b200.scheme import MODE_FAST

pytestmark = pytest.mark.gpu

PRODUCT_TOL = 2.0 ** 33          # same bound as one blind-rotation step (tests/test_gpu_fast.py); measured values are printed


def oracle_product(poly, keys, l, logB):
    """native(ifft(Sum_j fft(D_j(poly)) * keys[j][c])) with the oracle's decomposition and transforms (reference order)."""
    from oracle import oracle as O
    digits = O.decomp(poly, l, logB)
    ncomp, H = keys.shape[1], keys.shape[2]
    acc = np.zeros((ncomp, H), dtype=np.complex128)
    for j in range(l):
        x = O.fft(digits[j])
        x = x[:, 0] + 1j * x[:, 1]
        for c in range(ncomp):
            acc[c] += x * (keys[j, c, :, 0] + 1j * keys[j, c, :, 1])
    return np.stack([O.ifft(np.stack([acc[c].real, acc[c].imag], axis=1), 64) for c in range(ncomp)])


def _diff(a, b):
    return np.abs((a.astype(np.uint64) - b.astype(np.uint64)).astype(np.int64).astype(np.float64)).max()


@pytest.mark.parametrize("name", ["KMS2party", "KMS8party", "KMS32party"])
def test_phase2_products_within_tolerance(gpu_schemes, name):
    ks = keyset(name)
    orc = make_oracle(ks)
    s = gpu_schemes(name)
    s.set_mode(MODE_FAST)
    p = ks.params
    rng = np.random.default_rng(7)
    _, ct = fresh_inputs(ks, 1, seed=151)
    tilde = orc.modswitch(ct[0])
    party = p.k - 1
    lev = orc.phase1(party, tilde[1 + party * p.n: 1 + (party + 1) * p.n])            # [l_lev][2][H][2]: real levkey rows
    B = 6
    acc = rng.integers(0, np.iinfo(np.uint64).max, size=(B, p.N), dtype=np.uint64)   # an accumulator component
    worst = {}
    # (1) LEV product of an accumulator component with levkey[idx]: tx (kept in FFT form by the kernel) and y = ifft(ty)
    got = s.gadget_product(acc, lev, p.l_lev, p.logB_lev)
    ys = []
    for g in range(B):
        ref = oracle_product(acc[g], lev, p.l_lev, p.logB_lev)
        worst["lev"] = max(worst.get("lev", 0.0), _diff(got[g], ref))
        ys.append(ref[1])                                                            # the oracle's y feeds the next product
    # (2) the oracle's y through the uni-encryption gadget: u = y [.] rlk.d, and the public-key / CRS products (v)
    y = np.stack(ys)
    rlk = ks.rlk[party]                                                              # [l_uni][3][H][2]
    got = s.gadget_product(y, rlk, p.l_uni, p.logB_uni)
    vs = []
    for g in range(B):
        ref = oracle_product(y[g], rlk, p.l_uni, p.logB_uni)
        worst["uni"] = max(worst.get("uni", 0.0), _diff(got[g], ref))
        vs.append(ref[0])
    pub = np.stack([ks.pubb[0], ks.crs_fft], axis=1)                                 # [l_uni][2][H][2]
    got = s.gadget_product(y, pub, p.l_uni, p.logB_uni)
    for g in range(B):
        worst["pub"] = max(worst.get("pub", 0.0), _diff(got[g], oracle_product(y[g], pub, p.l_uni, p.logB_uni)))
    # (3) a computed polynomial (the oracle's, standing for v) through rlk.f: w
    v = np.stack(vs)
    got = s.gadget_product(v, rlk[:, 1:], p.l_uni, p.logB_uni)
    for g in range(B):
        worst["w"] = max(worst.get("w", 0.0), _diff(got[g], oracle_product(v[g], rlk[:, 1:], p.l_uni, p.logB_uni)))
    print(f"{name}: worst |delta| per product: " + ", ".join(f"{k} 2^{np.log2(v_ + 1):.2f}" for k, v_ in worst.items()) + " (tolerance 2^33)")
    assert max(worst.values()) < PRODUCT_TOL
