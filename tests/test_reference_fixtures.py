"""Reference-pinned parity: fixtures dumped from a RUN OF THE REFERENCE (julia/dump_fixtures.jl, which follows
/root/reference/test/KMS.jl:5-37) hold real reference keys in the flat upload layouts, input ciphertexts, the outputs of all six
gates and the intermediate stages of NAND.  The CPU oracle and the GPU STRICT path must reproduce every array bit for bit from
those keys (/root/reference/src/tfhe/keygen.jl:85-118 builds them, src/tfhe/bootstrapping.jl:4-27 consumes them).

The build image has no Julia, so `tests/golden_ref/` is absent here and the two reference tests skip; the consumer itself is
exercised on mock fixtures written in the same container format with the oracle standing in for the reference."""
import os

import numpy as np
import pytest

from conftest import REFERENCE_TEST_SETS, fresh_inputs, keyset, make_oracle
from mktfhe_b200 import blob
from mktfhe_b200 import params as P

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.environ.get("MKTFHE_GOLDEN_REF", os.path.join(HERE, "golden_ref"))


class OracleEvaluator:
    def __init__(self, lk):
        from oracle import oracle as O
        p = lk.params
        self.orc = O.Oracle(p, [np.ascontiguousarray(a) for a in lk.brk], [np.ascontiguousarray(a) for a in lk.ksk],
                            [np.ascontiguousarray(a) for a in lk.rlk] if p.scheme in (P.KMS, P.KMS_BLOCK) else None,
                            [np.ascontiguousarray(a) for a in lk.pubb] if p.is_mk else None,
                            np.ascontiguousarray(lk.crs_fft) if p.is_mk else None)

    def gate(self, op, c1, c2):
        return self.orc.gate_batch(op, c1, c2)

    def gate_linear(self, op, c1, c2):
        return np.stack([self.orc.gate_linear(op, a, b) for a, b in zip(c1, c2)])

    def modswitch(self, c):
        return np.stack([self.orc.modswitch(x) for x in c])

    def blindrotate(self, c):
        return np.stack([self.orc.blindrotate(x) for x in c])

    def close(self):
        pass


class GpuStrictEvaluator:
    def __init__(self, lk):
        from mktfhe_b200.scheme import MODE_STRICT, setup
        self.s = setup(lk, device=0, mode=MODE_STRICT)

    def gate(self, op, c1, c2):
        return self.s.gate(op, c1, c2)

    def gate_linear(self, op, c1, c2):
        return self.s.gate_linear(op, c1, c2)

    def modswitch(self, c):
        return self.s.modswitch(c)

    def blindrotate(self, c):
        return self.s.blindrotate(c)

    def close(self):
        self.s.close()


def check_fixture(directory, name, make_evaluator):
    """Every array of <name>.fixture.blob reproduced bit for bit from the keys of <name>.keys.blob."""
    lk = blob.load_keys(os.path.join(directory, f"{name}.keys.blob"), verify=True)
    p, a = blob.load_fixture(os.path.join(directory, f"{name}.fixture.blob"))
    assert p == P.ALL[name] and lk.params == p, f"{name}: fixture parameters differ from params.jl"
    in1, in2 = np.ascontiguousarray(a["in1"]), np.ascontiguousarray(a["in2"])
    b1, b2 = a["bits1"].astype(bool), a["bits2"].astype(bool)
    ev = make_evaluator(lk)
    try:
        lin = ev.gate_linear(0, in1, in2)
        assert np.array_equal(lin, a["nand_linear"]), f"{name}: gate linear part differs from the reference"
        assert np.array_equal(ev.modswitch(lin), a["nand_tilde"]), f"{name}: modulus switch differs from the reference"
        acc = ev.blindrotate(lin)
        assert acc.dtype == a["nand_acc"].dtype and np.array_equal(acc, a["nand_acc"]), f"{name}: blind-rotation accumulator differs from the reference"
        from mktfhe_b200.gate import PLAIN
        for op, g in enumerate(blob.FIXTURE_GATES):
            out = ev.gate(op, in1, in2)
            assert np.array_equal(out, a[f"out_{g}"]), f"{name}: {g} output differs from the reference"
            if lk.lwekeys is not None:
                want = np.array([PLAIN[op](bool(x), bool(y)) for x, y in zip(b1, b2)])
                assert np.array_equal(lk.decrypt_batch(out), want), f"{name}: {g} decrypts wrongly"
    finally:
        ev.close()
    return a["in1"].shape[0]


def write_mock_fixture(directory, name, pairs=3):
    """Same files as julia/dump_fixtures.jl writes, with the oracle standing in for the reference (NOT reference data)."""
    ks = keyset(name)
    p = ks.params
    orc = make_oracle(ks)
    blob.save_keys(os.path.join(directory, f"{name}.keys.blob"), ks, include_secret=True)
    b1, c1 = fresh_inputs(ks, pairs, seed=301, full=False)
    b2, c2 = fresh_inputs(ks, pairs, seed=302, full=False)
    c1 = np.concatenate([c1, orc.gate_batch(0, c1[:1], c2[:1])])          # one pair of bootstrapped (full-support) operands
    c2 = np.concatenate([c2, orc.gate_batch(2, c1[1:2], c2[1:2])])
    b1 = np.append(b1, not (b1[0] and b2[0])); b2 = np.append(b2, b1[1] or b2[1])
    arrays = {"in1": c1, "in2": c2, "bits1": b1.astype(np.uint8), "bits2": b2.astype(np.uint8)}
    lin = np.stack([orc.gate_linear(0, x, y) for x, y in zip(c1, c2)])
    arrays["nand_linear"] = lin
    arrays["nand_tilde"] = np.stack([orc.modswitch(x) for x in lin])
    arrays["nand_acc"] = np.stack([orc.blindrotate(x) for x in lin])
    for op, g in enumerate(blob.FIXTURE_GATES):
        arrays[f"out_{g}"] = orc.gate_batch(op, c1, c2)
    blob.save_fixture(os.path.join(directory, f"{name}.fixture.blob"), p, arrays)


@pytest.mark.parametrize("name", ["CGGIparam", "CCS2party"])
def test_fixture_consumer_on_mock_fixtures(tmp_path, name):
    write_mock_fixture(str(tmp_path), name)
    assert check_fixture(str(tmp_path), name, OracleEvaluator) == 4
    # a corrupted output must be caught
    p, a = blob.load_fixture(str(tmp_path / f"{name}.fixture.blob"))
    bad = {k: np.array(v) for k, v in a.items()}
    bad["out_XOR"][1, 5] ^= 1
    blob.save_fixture(str(tmp_path / f"{name}.fixture.blob"), p, bad)
    with pytest.raises(AssertionError, match="XOR output differs"):
        check_fixture(str(tmp_path), name, OracleEvaluator)
    del bad["nand_acc"]
    blob.save_fixture(str(tmp_path / f"{name}.fixture.blob"), p, bad)
    with pytest.raises(blob.BlobError, match="lacks"):
        blob.load_fixture(str(tmp_path / f"{name}.fixture.blob"))


def _have(name):
    return os.path.exists(os.path.join(REF_DIR, f"{name}.keys.blob")) and os.path.exists(os.path.join(REF_DIR, f"{name}.fixture.blob"))


@pytest.mark.parametrize("name", REFERENCE_TEST_SETS)
def test_oracle_reproduces_the_reference_fixtures(name):
    if not _have(name):
        pytest.skip(f"no reference fixtures for {name} under {REF_DIR} (make them with julia/dump_fixtures.jl)")
    check_fixture(REF_DIR, name, OracleEvaluator)


@pytest.mark.gpu
@pytest.mark.parametrize("name", REFERENCE_TEST_SETS)
def test_gpu_strict_reproduces_the_reference_fixtures(name):
    if not _have(name):
        pytest.skip(f"no reference fixtures for {name} under {REF_DIR} (make them with julia/dump_fixtures.jl)")
    check_fixture(REF_DIR, name, GpuStrictEvaluator)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["KMS2party", "KMS2partyblock", "Blockparam"])
def test_gpu_strict_on_mock_fixtures(tmp_path, name):
    """The GPU leg of the consumer, through blob.load_keys -> setup, on oracle-made fixtures."""
    write_mock_fixture(str(tmp_path), name)
    check_fixture(str(tmp_path), name, GpuStrictEvaluator)
