"""Gates with the reference's names and argument order: /root/reference/src/tfhe/gate.jl:1-58.
Each takes one ciphertext or a batch and returns bootstrapped ciphertext(s)."""
from __future__ import annotations

import numpy as np

from .scheme import AND_OP, NAND_OP, NOR_OP, OR_OP, XNOR_OP, XOR_OP, Scheme


def NAND(c1, c2, scheme: Scheme):
    return scheme.gate(NAND_OP, c1, c2)


def AND(c1, c2, scheme: Scheme):
    return scheme.gate(AND_OP, c1, c2)


def OR(c1, c2, scheme: Scheme):
    return scheme.gate(OR_OP, c1, c2)


def XOR(c1, c2, scheme: Scheme):
    return scheme.gate(XOR_OP, c1, c2)


def XNOR(c1, c2, scheme: Scheme):
    return scheme.gate(XNOR_OP, c1, c2)


def NOR(c1, c2, scheme: Scheme):
    return scheme.gate(NOR_OP, c1, c2)


def NOT(c):
    """NOT! (gate.jl:55-58): negation only, no bootstrap; returns a new array."""
    return (np.uint32(0) - np.asarray(c, dtype=np.uint32)).astype(np.uint32)


def bootstrapping(c, scheme: Scheme):
    return scheme.bootstrapping(c)


PLAIN = {
    NAND_OP: lambda x, y: not (x and y), AND_OP: lambda x, y: x and y, OR_OP: lambda x, y: x or y,
    XOR_OP: lambda x, y: x != y, XNOR_OP: lambda x, y: x == y, NOR_OP: lambda x, y: not (x or y),
}
