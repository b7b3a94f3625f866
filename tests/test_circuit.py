"""Gate-circuit scheduler (mktfhe_b200/circuit.py): netlist -> levels of independent gates -> one batched
`mktfhe_gate_level` call per level over a device-resident wire table.  The reference's own tests are chains of
per-gate calls (test/KMS.jl:28-36); the GPU tests evaluate the same kind of chain gate by gate through the oracle and
require the circuit evaluator to reproduce every output bit-exactly in STRICT mode."""
import numpy as np
import pytest

from conftest import keyset, make_oracle
from mktfhe_b200.circuit import Circuit, equality, greater_than, ripple_adder
from mktfhe_b200.scheme import BOOTSTRAP_OP, NOT_OP


def bits_of(v, n):
    return [(v >> i) & 1 for i in range(n)]


def test_plain_evaluation_of_the_standard_circuits():
    add, eq, gt = ripple_adder(5), equality(5), greater_than(5)
    rng = np.random.default_rng(3)
    for a, b in rng.integers(0, 32, (200, 2)):
        x = bits_of(int(a), 5) + bits_of(int(b), 5)
        o = add.evaluate_plain(x)
        assert sum(int(o[i]) << i for i in range(6)) == a + b
        assert bool(eq.evaluate_plain(x)[0]) == (a == b)
        assert bool(gt.evaluate_plain(x)[0]) == (a > b)
    batch = rng.integers(0, 2, (7, 10)).astype(bool)
    assert add.evaluate_plain(batch).shape == (7, 6)


@pytest.mark.parametrize("make", [lambda: ripple_adder(8), lambda: equality(7), lambda: greater_than(6)])
def test_schedule_is_a_valid_levelisation(make):
    c = make()
    steps = c.schedule()
    defined = set(range(c.n_inputs))
    seen = 0
    for ops, s1, s2, dst in steps:
        assert len(set(dst.tolist())) == len(dst)                       # SSA: one definition per wire
        assert len(set(ops.tolist()) & {NOT_OP}) in (0, 1) and (NOT_OP not in ops or set(ops.tolist()) == {NOT_OP})
        for op, a, b, d in zip(ops, s1, s2, dst):
            assert int(a) in defined and int(b) in defined              # reads only earlier steps
            assert int(d) not in defined
        defined |= set(dst.tolist())
        seen += len(dst)
    assert seen == len(c.gates) and defined == set(range(c.n_wires))
    assert c.depth() == sum(1 for s in steps if s[0][0] != NOT_OP)
    # ASAP: the adder's carry chain costs two levels per bit
    if c.n_inputs == 16 and len(c.outputs) == 9:
        assert c.depth() == 2 * 8 - 1


def test_construction_errors():
    c = Circuit(2)
    with pytest.raises(ValueError):
        c.gate(0, 0, 5)                       # undefined wire
    with pytest.raises(ValueError):
        c.gate(NOT_OP, 0, 1)                  # not a two-input opcode
    w = c.NAND(0, 1)
    c.set_outputs([w])
    with pytest.raises(ValueError):
        c.evaluate_plain([True])              # wrong input count


def oracle_eval(c, orc, cts):
    """The reference's way: one gate call at a time."""
    wires = [np.asarray(x, dtype=np.uint32) for x in cts]
    for op, a, b in c.gates:
        if op == NOT_OP:
            wires.append((np.uint32(0) - wires[a]).astype(np.uint32))
        elif op == BOOTSTRAP_OP:
            wires.append(orc.bootstrap(wires[a]))
        else:
            wires.append(orc.bootstrap(orc.gate_linear(op, wires[a], wires[b])))
    return np.stack([wires[w] for w in c.outputs])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["KMS2party", "CGGIparam"])
def test_circuit_matches_gate_by_gate_oracle_bit_exactly(gpu_schemes, name):
    from mktfhe_b200.scheme import MODE_FAST, MODE_STRICT
    ks = keyset(name)
    s = gpu_schemes(name)
    p = ks.params
    orc = make_oracle(ks)
    c = Circuit(3)
    t = c.XOR(0, 1)
    u = c.MUX(2, t, c.NOT(0))                 # exercises NOT between levels
    v = c.bootstrap(c.NOR(u, c.NOT(c.NOT(1))))
    c.set_outputs([t, u, v])
    bits = np.array([1, 0, 1], dtype=bool)
    enc = (lambda m, i: ks.lwe_ith_encrypt(int(m), i % p.k, 700 + i)) if p.is_mk else (lambda m, i: ks.lwe_encrypt(int(m), 700 + i))
    cts = np.stack([enc(m, i) for i, m in enumerate(bits)])
    s.set_mode(MODE_STRICT)
    got = c.evaluate(s, cts)
    want = oracle_eval(c, orc, cts)
    assert np.array_equal(got, want)
    assert list(ks.decrypt_batch(got)) == list(c.evaluate_plain(bits))
    s.set_mode(MODE_FAST)
    assert list(ks.decrypt_batch(c.evaluate(s, cts))) == list(c.evaluate_plain(bits))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["KMS2party", "KMS2partyblock", "CGGIparam", "CCS2party"])
def test_encrypted_adder_and_comparators(gpu_schemes, name):
    """`instances` independent 4-bit additions / comparisons in one evaluation; for the multi-key sets operand a is
    encrypted by party 0 and operand b by party 1 (test/KMS.jl:17-22: lwe_ith_encrypt)."""
    ks = keyset(name)
    s = gpu_schemes(name)
    p = ks.params
    nb, inst = 4, 6
    rng = np.random.default_rng(11)
    vals = rng.integers(0, 1 << nb, (inst, 2))
    vals[0] = (9, 9)
    plain = np.array([bits_of(int(a), nb) + bits_of(int(b), nb) for a, b in vals], dtype=bool)

    def enc(m, inst_i, j):
        seed = 5000 + inst_i * 64 + j
        if not p.is_mk:
            return ks.lwe_encrypt(int(m), seed)
        return ks.lwe_ith_encrypt(int(m), 0 if j < nb else 1, seed)
    cts = np.stack([np.stack([enc(plain[i, j], i, j) for j in range(2 * nb)]) for i in range(inst)])
    for c in (ripple_adder(nb), equality(nb), greater_than(nb)):
        out = c.evaluate(s, cts)
        assert out.shape == (inst, len(c.outputs), p.lwe_words)
        dec = np.array([ks.decrypt_batch(out[i]) for i in range(inst)])
        assert np.array_equal(dec, c.evaluate_plain(plain)), name
    sums = ripple_adder(nb).evaluate_plain(plain)
    assert [sum(int(r[i]) << i for i in range(nb + 1)) for r in sums] == [int(a + b) for a, b in vals]


@pytest.mark.gpu
def test_gate_level_argument_errors(gpu_schemes):
    from mktfhe_b200.scheme import MktfheError
    s = gpu_schemes("CGGIparam")
    z = np.zeros(1, dtype=np.int32)
    s.wires_resize(0)
    with pytest.raises(MktfheError, match="wire table"):
        s.gate_level(z, z, z, z)
    s.wires_resize(4)
    with pytest.raises(MktfheError, match="opcode"):
        s.gate_level(z + 9, z, z, z)
    with pytest.raises(MktfheError, match="index"):
        s.gate_level(z, z + 4, z, z)
    with pytest.raises(MktfheError, match="range"):
        s.wires_read(3, 2)
    s.gate_level(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32))
    s.wires_resize(0)


def test_random_netlists_schedule_and_evaluate_consistently():
    """Random SSA netlists (all opcodes, NOT chains, bare bootstraps): executing the schedule step by step on plaintext bits
    gives the same wire values as executing the gates in creation order, and every step reads only earlier steps."""
    rng = np.random.default_rng(2024)
    plain2 = {0: lambda x, y: not (x and y), 1: lambda x, y: x and y, 2: lambda x, y: x or y,
              3: lambda x, y: x != y, 4: lambda x, y: x == y, 5: lambda x, y: not (x or y)}
    for trial in range(40):
        nin = int(rng.integers(1, 6))
        c = Circuit(nin)
        for _ in range(int(rng.integers(1, 40))):
            kind = rng.random()
            a, b = int(rng.integers(0, c.n_wires)), int(rng.integers(0, c.n_wires))
            if kind < 0.15:
                c.NOT(a)
            elif kind < 0.22:
                c.bootstrap(a)
            else:
                c.gate(int(rng.integers(0, 6)), a, b)
        c.set_outputs(list(range(c.n_wires)))
        bits = rng.integers(0, 2, nin).astype(bool)
        want = c.evaluate_plain(bits)
        wires = {i: bool(bits[i]) for i in range(nin)}
        for ops, s1, s2, dst in c.schedule():
            new = {}
            for op, a, b, d in zip(ops.tolist(), s1.tolist(), s2.tolist(), dst.tolist()):
                assert a in wires and b in wires and d not in wires and d not in new
                new[d] = (not wires[a]) if op == NOT_OP else wires[a] if op == BOOTSTRAP_OP else bool(plain2[op](wires[a], wires[b]))
            wires.update(new)
        assert [wires[w] for w in range(c.n_wires)] == [bool(v) for v in want], trial
