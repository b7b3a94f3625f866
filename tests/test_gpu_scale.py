"""Larger party counts (parameter sets nobody runs in the reference's own tests):
KMS 8-party, KMS 8-party block (BASELINE config 3) and CCS 4-party -- STRICT bit-exact against the oracle on one
gate, FAST/production gates decrypt correctly over a batch, noise stays inside the decision margin."""
import numpy as np
import pytest

from conftest import fresh_inputs, keyset, make_oracle
from mktfhe_b200.gate import PLAIN
from mktfhe_b200.scheme import MODE_FAST, MODE_STRICT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["KMS8party", "KMS8partyblock", "CCS4party"])
def test_larger_party_counts(gpu_schemes, name):
    ks = keyset(name)
    orc = make_oracle(ks)
    s = gpu_schemes(name)
    p = ks.params
    B = 12
    b1, c1 = fresh_inputs(ks, B, seed=51)
    b2, c2 = fresh_inputs(ks, B, seed=52)
    # STRICT: one NAND, every output word equals the oracle's
    s.set_mode(MODE_STRICT)
    lin = orc.gate_linear(0, c1[0], c2[0])
    assert np.array_equal(s.gate(0, c1[:1], c2[:1])[0], orc.bootstrap(lin))
    acc = s.blindrotate(lin[None])[0]
    assert np.array_equal(acc, orc.blindrotate(lin))
    if name.startswith("CCS"):
        # CCS4party sits at the edge of its noise budget in the reference algorithm itself: the oracle-identical STRICT
        # path decrypts ~96 % of fresh NANDs (output error std 2^28.07 against a margin of 2^29).  Production mode
        # must show the same statistics, not exact decryptions.
        stats = {}
        for mode in (MODE_STRICT, MODE_FAST):
            s.set_mode(mode)
            errs, ok = [], 0
            for op in (0, 3, 5):
                out = s.gate(op, c1, c2)
                want = np.array([PLAIN[op](bool(x), bool(y)) for x, y in zip(b1, b2)])
                ok += int(np.sum(ks.decrypt_batch(out) == want))
                for g in range(B):
                    e = (ks.phase(out[g]) - ((1 << 29) if want[g] else (7 << 29))) & 0xFFFFFFFF
                    errs.append(e - (1 << 32) if e >= (1 << 31) else e)
            stats[mode] = (ok, float(np.std(errs)))
        print(f"{name}: STRICT ok {stats[MODE_STRICT][0]}/{3 * B} std 2^{np.log2(stats[MODE_STRICT][1]):.2f}; "
              f"FAST ok {stats[MODE_FAST][0]}/{3 * B} std 2^{np.log2(stats[MODE_FAST][1]):.2f}")
        assert 0.75 < stats[MODE_FAST][1] / stats[MODE_STRICT][1] < 1.33
        assert stats[MODE_FAST][0] >= 0.85 * 3 * B and stats[MODE_STRICT][0] >= 0.85 * 3 * B
        s.set_mode(MODE_FAST)
        return
    # production mode: all gates over the batch
    s.set_mode(MODE_FAST)
    worst = 0
    for op in (0, 3, 5):
        out = s.gate(op, c1, c2)
        want = np.array([PLAIN[op](bool(x), bool(y)) for x, y in zip(b1, b2)])
        assert np.array_equal(ks.decrypt_batch(out), want), (name, op)
        for g in range(B):
            e = (ks.phase(out[g]) - ((1 << 29) if want[g] else (7 << 29))) & 0xFFFFFFFF
            worst = max(worst, abs(e - (1 << 32) if e >= (1 << 31) else e))
    print(f"{name}: worst |phase error| = 2^{np.log2(worst + 1):.2f} (margin 2^29)")
    assert worst < (1 << 29)
    # chained gates: outputs of bootstraps feed the next level (full-support ciphertexts)
    lvl = s.gate(0, c1, c2)
    lvl2 = s.gate(2, lvl, c1)
    want = np.array([(not (x and y)) or bool(x) for x, y in zip(b1, b2)])
    assert np.array_equal(ks.decrypt_batch(lvl2), want)
