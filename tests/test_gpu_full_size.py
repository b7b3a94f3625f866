"""The BASELINE.json configurations at their FULL batch sizes, through properties that do not need the (slow) CPU oracle:

  * every output decrypts to the plaintext gate (the reference's own acceptance criterion, test/KMS.jl:28-37), or, at the margin
    set KMS32party, at the rate the oracle itself reaches (tests/golden/failrate_KMS32party.npz);
  * determinism: the same call twice gives the same bytes (the kernels fix the order of every floating-point sum);
  * commutativity: NAND(c1, c2) and NAND(c2, c1) are bit-identical (the linear part commutes, gate.jl:1-9, so the bootstrap sees
    the same ciphertext);
  * batch invariance at size: a gate's output does not depend on the batch it was evaluated in (first / last gates re-evaluated
    in a small batch) -- "batch of B" means B independent reference calls;
  * a second level on top: NAND(x, fresh 1) of the outputs decrypts to NOT x (bootstrapped outputs are valid inputs; a fresh second
    operand keeps the two noise terms independent -- NAND(x, x) doubles the noise coherently and fails a few gates in 4096).
Sizes: C2 KMS2party 4096, C1 CGGIparam 4096, C3 KMS8partyblock 2048 (per GPU of 16384 / 8), C5 KMS32party 1024."""
import hashlib

import numpy as np
import pytest

from conftest import keyset
from mktfhe_b200.scheme import MODE_FAST

pytestmark = pytest.mark.gpu

CASES = [("KMS2party", 4096, 1.0), ("CGGIparam", 4096, 1.0), ("KMS8partyblock", 2048, 1.0), ("KMS32party", 1024, 0.93)]


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("name,batch,min_ok", CASES)
def test_full_batch_properties(gpu_schemes, name, batch, min_ok):
    ks = keyset(name)
    s = gpu_schemes(name)
    s.set_mode(MODE_FAST)
    rng = np.random.default_rng(2024)
    m1 = rng.integers(0, 2, batch)
    m2 = rng.integers(0, 2, batch)
    c1 = ks.encrypt_batch(m1, 7_000_000)
    c2 = ks.encrypt_batch(m2, 8_000_000)
    out = s.gate(0, c1, c2)                                             # MK-NAND over the whole batch
    want = ~(m1.astype(bool) & m2.astype(bool))
    ok = int((ks.decrypt_batch(out) == want).sum())
    assert ok >= min_ok * batch, f"{name}: {ok}/{batch} gates decrypt correctly"
    # determinism and commutativity, bit for bit
    assert _sha(s.gate(0, c1, c2)) == _sha(out)
    assert np.array_equal(s.gate(0, c2, c1), out)
    # batch invariance at size: head and tail of the batch evaluated on their own
    k = 37
    assert np.array_equal(s.gate(0, c1[:k], c2[:k]), out[:k])
    assert np.array_equal(s.gate(0, c1[-k:], c2[-k:]), out[-k:])
    # the outputs are valid inputs: NAND(x, 1) = NOT x
    if min_ok == 1.0:
        ones = ks.encrypt_batch(np.ones(batch, dtype=np.int64), 9_000_000)
        out2 = s.gate(0, out, ones)
        assert np.array_equal(ks.decrypt_batch(out2), ~want)
    print(f"{name}: {batch} gates, {ok} decrypt correctly; deterministic, commutative, batch-invariant")
