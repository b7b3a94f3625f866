"""FAST CCS (fast32::k_ccs_fast) re-decomposes a computed polynomial inside every hybrid product (the `v` of
/root/reference/src/tfhe/bootstrapping.jl:286-320), so whole-step coefficients cannot be compared with another rounding: one digit
flip at a rounding boundary changes everything downstream.  The gadget products the hybrid product is made of can:
feed the ORACLE's polynomial (an accumulator component, then the oracle's own v) into ONE product built from the same device
functions as k_ccs_fast (mktfhe_gadget_product32_batch) and compare with the reference-order product within a stated Torus32 tolerance.
Reference: bootstrapping.jl:277-284 (u_c against d[j]), :286-294 (v_c against crs[j] / b[j]), :313-320 (w against f[j])."""
import numpy as np
import pytest

from conftest import keyset
from mktfhe_b200.scheme import MODE_FAST

pytestmark = pytest.mark.gpu

PRODUCT_TOL = 2          # Torus32 units: the two roundings may fall on different sides of a boundary; measured values are printed


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
    return np.stack([O.ifft(np.stack([acc[c].real, acc[c].imag], axis=1), 32) for c in range(ncomp)])


def _diff(a, b):
    return int(np.abs((a.astype(np.uint32) - b.astype(np.uint32)).astype(np.int32).astype(np.int64)).max())


@pytest.mark.parametrize("name", ["CCS2party", "CCS4party", "CCS16party"])
def test_hybrid_product_parts_within_tolerance(gpu_schemes, name):
    ks = keyset(name)
    s = gpu_schemes(name)
    s.set_mode(MODE_FAST)
    p = ks.params
    l, logB, H = p.l_uni, p.logB_uni, p.H
    rng = np.random.default_rng(11)
    party, idx = p.k - 1, 5
    uni = ks.brk[party][idx].reshape(l, 3, H, 2)                                    # [j][d, f.b, f.a][H][2]
    B = 6
    acc = rng.integers(0, 1 << 32, size=(B, p.N), dtype=np.uint64).astype(np.uint32)   # an accumulator component
    worst = {}
    # (1) u = acc [.] d and v = acc [.] crs (component 0) / acc [.] b_party (the others)
    vs = []
    for label, second in (("crs", ks.crs_fft), ("pub", ks.pubb[party])):
        keys = np.stack([uni[:, 0], second], axis=1)                                 # [l][2][H][2]
        got = s.gadget_product(acc, keys, l, logB)
        for g in range(B):
            ref = oracle_product(acc[g], keys, l, logB)
            worst["u"] = max(worst.get("u", 0), _diff(got[g, 0], ref[0]))
            worst[label] = max(worst.get(label, 0), _diff(got[g, 1], ref[1]))
            if label == "crs":
                vs.append((0 - ref[1].astype(np.uint32)).astype(np.uint32))          # v_0 = - sum (mulsubto!, :288-290)
    # (2) the oracle's v, a computed polynomial, through f: w_b, w_a
    v = np.stack(vs)
    got = s.gadget_product(v, uni[:, 1:], l, logB)
    for g in range(B):
        ref = oracle_product(v[g], uni[:, 1:], l, logB)
        worst["w"] = max(worst.get("w", 0), _diff(got[g], ref))
    print(f"{name}: worst |delta| per product (Torus32 units): " + ", ".join(f"{k} {v_}" for k, v_ in worst.items()) + f" (tolerance {PRODUCT_TOL})")
    assert max(worst.values()) <= PRODUCT_TOL


def test_hook_rejects_the_wrong_torus(gpu_schemes):
    from mktfhe_b200 import _lib
    s = gpu_schemes("KMS2party")
    z = np.zeros((1, 2048), np.uint32)
    k = np.zeros((1, 1, 1024, 2), np.float64)
    o = np.zeros((1, 1, 2048), np.uint32)
    rc = _lib.lib().mktfhe_gadget_product32_batch(s._h, 1, 8, z.ctypes.data, k.ctypes.data, 1, o.ctypes.data, 1)
    assert rc != 0
