"""Committed regression vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU oracle).

They are oracle outputs, not reference outputs (the reference has no golden vectors and cannot run here), so they
do not change the "parity unpinned" status; they pin key generation, the oracle and the CUDA STRICT path to the same
bytes on every machine and across commits.
"""
import hashlib
import os

import numpy as np
import pytest

from conftest import REFERENCE_TEST_SETS, keyset, make_oracle
from mktfhe_b200.gate import PLAIN

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def key_array(ks, label):
    if label == "crs_fft":
        return ks.crs_fft
    party, field = label.split(".")
    return ks.parties[int(party[1:])][field]


@pytest.mark.parametrize("name", REFERENCE_TEST_SETS)
def test_keygen_reproduces_the_committed_digests(name):
    g = load(name)
    ks = keyset(name)
    assert int(g["key_seed"]) == ks.seed
    for label, want in zip(g["digest_names"], g["digest_values"]):
        got = hashlib.sha256(np.ascontiguousarray(key_array(ks, str(label))).tobytes()).hexdigest()
        assert got == str(want), (name, str(label))


@pytest.mark.parametrize("name", REFERENCE_TEST_SETS)
def test_oracle_reproduces_the_committed_vectors(name):
    g = load(name)
    ks = keyset(name)
    orc = make_oracle(ks)
    c1, c2 = g["in1"], g["in2"]
    for gi in range(c1.shape[0]):
        lin = orc.gate_linear(0, c1[gi], c2[gi])
        assert np.array_equal(lin, g["nand_linear"][gi])
        assert np.array_equal(orc.modswitch(lin), g["nand_tilde"][gi])
        assert np.array_equal(orc.blindrotate(lin), g["nand_acc"][gi])
    for op in (0, 3):                                  # NAND and XOR in full; the GPU test covers all six
        assert np.array_equal(orc.gate_batch(op, c1, c2), g["out"][op]), (name, op)
    # and the vectors themselves decrypt to the truth table
    for op in range(6):
        want = [PLAIN[op](bool(x), bool(y)) for x, y in zip(g["bits1"], g["bits2"])]
        assert list(ks.decrypt_batch(g["out"][op])) == want, (name, op)


@pytest.mark.gpu
@pytest.mark.parametrize("name", REFERENCE_TEST_SETS)
def test_gpu_strict_reproduces_the_committed_vectors(gpu_schemes, name):
    from mktfhe_b200.scheme import MODE_FAST, MODE_STRICT
    g = load(name)
    ks = keyset(name)
    s = gpu_schemes(name)
    c1, c2 = g["in1"], g["in2"]
    s.set_mode(MODE_STRICT)
    lin = s.gate_linear(0, c1, c2)
    assert np.array_equal(lin, g["nand_linear"])
    assert np.array_equal(s.modswitch(lin), g["nand_tilde"])
    assert np.array_equal(s.blindrotate(lin), g["nand_acc"])
    assert np.array_equal(s.keyswitch(g["nand_acc"]), g["out"][0])
    for op in range(6):
        assert np.array_equal(s.gate(op, c1, c2), g["out"][op]), (name, op)
    s.set_mode(MODE_FAST)
    for op in range(6):
        want = [PLAIN[op](bool(x), bool(y)) for x, y in zip(g["bits1"], g["bits2"])]
        assert list(ks.decrypt_batch(s.gate(op, c1, c2))) == want, (name, op)
