"""Gate circuits on top of the gates of /root/reference/src/tfhe/gate.jl.

The reference evaluates a circuit as a chain of per-gate calls (test/KMS.jl:28-36).  A GPU wants batches, so a
`Circuit` records the gates as a netlist, sorts them into levels of mutually independent gates (ASAP schedule: a
gate's level is one more than the deepest of its operands) and `evaluate` runs each level as one
`mktfhe_gate_level` call over a device-resident wire table.  `instances` independent copies of the circuit (the
same netlist on different encrypted inputs) widen every level, which is what fills the device.

NOT is free (no bootstrap, gate.jl:55-58): it is executed between levels and does not deepen the circuit.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .scheme import (AND_OP, BOOTSTRAP_OP, NAND_OP, NOR_OP, NOT_OP, OR_OP, XNOR_OP, XOR_OP, Scheme)

_TWO_INPUT = (NAND_OP, AND_OP, OR_OP, XOR_OP, XNOR_OP, NOR_OP)
_PLAIN = {
    NAND_OP: lambda x, y: ~(x & y), AND_OP: lambda x, y: x & y, OR_OP: lambda x, y: x | y,
    XOR_OP: lambda x, y: x ^ y, XNOR_OP: lambda x, y: ~(x ^ y), NOR_OP: lambda x, y: ~(x | y),
}


@dataclass
class Circuit:
    """A netlist in SSA form: wire ids 0 .. n_inputs-1 are the inputs, every gate defines one new wire."""
    n_inputs: int
    gates: list = field(default_factory=list)            # (op, a, b) defining wire n_inputs + index
    outputs: list = field(default_factory=list)

    # -- construction -----------------------------------------------------------------------------
    @property
    def n_wires(self) -> int:
        return self.n_inputs + len(self.gates)

    def _check(self, w):
        if not 0 <= int(w) < self.n_wires:
            raise ValueError(f"wire {w} is not defined yet")
        return int(w)

    def gate(self, op: int, a: int, b: int) -> int:
        if op not in _TWO_INPUT:
            raise ValueError(f"not a two-input gate opcode: {op}")
        self.gates.append((op, self._check(a), self._check(b)))
        return self.n_wires - 1

    def NAND(self, a, b): return self.gate(NAND_OP, a, b)
    def AND(self, a, b): return self.gate(AND_OP, a, b)
    def OR(self, a, b): return self.gate(OR_OP, a, b)
    def XOR(self, a, b): return self.gate(XOR_OP, a, b)
    def XNOR(self, a, b): return self.gate(XNOR_OP, a, b)
    def NOR(self, a, b): return self.gate(NOR_OP, a, b)

    def NOT(self, a) -> int:
        self.gates.append((NOT_OP, self._check(a), self._check(a)))
        return self.n_wires - 1

    def bootstrap(self, a) -> int:
        """A bare bootstrapping! of wire a (noise refresh), bootstrapping.jl:4-27."""
        self.gates.append((BOOTSTRAP_OP, self._check(a), self._check(a)))
        return self.n_wires - 1

    def MUX(self, sel, a, b) -> int:
        """sel ? a : b from the reference's gate set: OR(AND(sel, a), AND(NOT sel, b))."""
        return self.OR(self.AND(sel, a), self.AND(self.NOT(sel), b))

    def set_outputs(self, wires):
        self.outputs = [self._check(w) for w in wires]
        return self

    # -- scheduling ---------------------------------------------------------------------------------
    def schedule(self):
        """-> list of steps; a step is (ops, src1, src2, dst) int32 arrays of mutually independent gates.
        Bootstrapped gates of level L form one step; the NOTs that hang off level L's wires form the steps after it
        (one step per NOT-chain depth), so every step only reads wires written by earlier steps."""
        depth = np.zeros(self.n_wires, dtype=np.int64)          # bootstrap depth
        sub = np.zeros(self.n_wires, dtype=np.int64)            # NOT-chain depth below that level
        keyed = {}
        for i, (op, a, b) in enumerate(self.gates):
            w = self.n_inputs + i
            if op == NOT_OP:
                depth[w], sub[w] = depth[a], sub[a] + 1
            else:
                depth[w], sub[w] = 1 + max(depth[a], depth[b]), 0
            keyed.setdefault((int(depth[w]), int(sub[w])), []).append((op, a, b, w))
        steps = []
        for key in sorted(keyed):
            g = np.array(keyed[key], dtype=np.int32)
            steps.append((g[:, 0].copy(), g[:, 1].copy(), g[:, 2].copy(), g[:, 3].copy()))
        return steps

    def depth(self) -> int:
        """Number of bootstrapped levels (the circuit's latency in units of one batched gate call)."""
        return sum(1 for (ops, _s1, _s2, _dst) in self.schedule() if ops[0] != NOT_OP)

    def bootstrapped_gates(self) -> int:
        return sum(1 for (op, _a, _b) in self.gates if op != NOT_OP)

    # -- plaintext and encrypted evaluation ---------------------------------------------------------------
    def evaluate_plain(self, inputs) -> np.ndarray:
        """inputs: bool [n_inputs] or [instances, n_inputs] -> bool outputs, same leading shape."""
        x = np.asarray(inputs, dtype=bool)
        single = x.ndim == 1
        if single:
            x = x[None]
        if x.shape[1] != self.n_inputs:
            raise ValueError("wrong number of inputs")
        w = np.zeros((x.shape[0], self.n_wires), dtype=bool)
        w[:, :self.n_inputs] = x
        for i, (op, a, b) in enumerate(self.gates):
            if op == NOT_OP:
                w[:, self.n_inputs + i] = ~w[:, a]
            elif op == BOOTSTRAP_OP:
                w[:, self.n_inputs + i] = w[:, a]
            else:
                w[:, self.n_inputs + i] = _PLAIN[op](w[:, a], w[:, b])
        out = w[:, self.outputs]
        return out[0] if single else out

    def evaluate(self, scheme: Scheme, inputs) -> np.ndarray:
        """inputs: uint32 [n_inputs, words] or [instances, n_inputs, words] ciphertexts -> output ciphertexts
        [len(outputs), words] or [instances, len(outputs), words].

        Wire w of instance i lives in row w * instances + i of the device table, so one level of the netlist over all
        instances is a single batched call."""
        cts = np.ascontiguousarray(inputs, dtype=np.uint32)
        single = cts.ndim == 2
        if single:
            cts = cts[None]
        inst, nin, words = cts.shape
        if nin != self.n_inputs or words != scheme.params.lwe_words:
            raise ValueError("inputs do not match the circuit / parameter set")
        if not self.outputs:
            raise ValueError("circuit has no outputs: call set_outputs")
        scheme.wires_resize(self.n_wires * inst)
        try:
            scheme.wires_write(0, np.ascontiguousarray(cts.transpose(1, 0, 2)).reshape(nin * inst, words))
            lane = np.arange(inst, dtype=np.int64)
            for ops, s1, s2, dst in self.schedule():
                def rows(w):
                    return (w.astype(np.int64)[:, None] * inst + lane[None, :]).reshape(-1)
                scheme.gate_level(np.repeat(ops, inst), rows(s1), rows(s2), rows(dst))
            out = np.stack([scheme.wires_read(w * inst, inst) for w in self.outputs], axis=1)   # [inst, nout, words]
        finally:
            scheme.wires_resize(0)
        return out[0] if single else out


# ---- a few standard circuits --------------------------------------------------------------------------

def ripple_adder(bits: int) -> Circuit:
    """a + b -> bits + 1 outputs (little endian).  Inputs: a_0 .. a_{bits-1}, b_0 .. b_{bits-1}.
    Full adder = 2 XOR + 2 AND + 1 OR; depth 2 per bit after the first."""
    c = Circuit(2 * bits)
    a, b = list(range(bits)), list(range(bits, 2 * bits))
    outs = [c.XOR(a[0], b[0])]
    carry = c.AND(a[0], b[0])
    for i in range(1, bits):
        t = c.XOR(a[i], b[i])
        outs.append(c.XOR(t, carry))
        carry = c.OR(c.AND(a[i], b[i]), c.AND(t, carry))
    outs.append(carry)
    return c.set_outputs(outs)


def equality(bits: int) -> Circuit:
    """a == b -> 1 output: XNOR per bit, AND tree (depth 1 + ceil(log2 bits))."""
    c = Circuit(2 * bits)
    layer = [c.XNOR(i, bits + i) for i in range(bits)]
    while len(layer) > 1:
        nxt = [c.AND(layer[i], layer[i + 1]) for i in range(0, len(layer) - 1, 2)]
        if len(layer) % 2:
            nxt.append(layer[-1])
        layer = nxt
    return c.set_outputs(layer)


def greater_than(bits: int) -> Circuit:
    """a > b (unsigned) -> 1 output, scanning from the least significant bit:
    gt = MUX(a_i XOR b_i, a_i, gt)."""
    c = Circuit(2 * bits)
    gt = c.AND(0, c.NOT(bits))                              # a_0 AND NOT b_0
    for i in range(1, bits):
        gt = c.MUX(c.XOR(i, bits + i), i, gt)
    return c.set_outputs([gt])
