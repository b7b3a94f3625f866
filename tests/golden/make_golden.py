"""Generate the committed regression vectors under tests/golden/.

    python tests/golden/make_golden.py

WHAT THESE ARE: outputs of this repository's CPU oracle (oracle/, reference operation order, no FMA) on seeded keys
and seeded ciphertexts, at the five parameter sets the reference's own test scripts use (test/CGGI.jl:5, LMSS.jl:5,
CCS.jl:5, KMS.jl:5, KMSblock.jl:5).  The reference holds no golden vectors and Julia is not installed, so these are
NOT reference outputs and do not pin ciphertext-level parity with the reference (see oracle/mktfhe_oracle.h).  They
pin everything downstream of that: key generation, the oracle and the CUDA STRICT path must all reproduce these bytes
on every machine, and the FAST path must decrypt to the same bits.

Per set `<name>.npz` holds: the key / ciphertext seeds, the SHA-256 of every flat key array (keygen determinism), two
pairs of full-support input ciphertexts with their plaintext bits, and for each of the six gates the bootstrapped
output ciphertexts plus the intermediate stages of gate 0 / NAND (linear part, modulus switch, blind-rotation
accumulator) so a failure can be localised.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import REFERENCE_TEST_SETS, fresh_inputs, keyset, make_oracle  # noqa: E402

KEY_SEED = 0x4D4B5446
PAIRS = 2
SEED_A, SEED_B = 9001, 9002


def key_digests(ks):
    p = ks.params
    out = {}
    for i, q in enumerate(ks.parties):
        for field in ("lwekey", "brk", "ksk", "rlk", "pubb"):
            a = q.get(field)
            if a is not None:
                out[f"p{i}.{field}"] = hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    if p.is_mk:
        out["crs_fft"] = hashlib.sha256(np.ascontiguousarray(ks.crs_fft).tobytes()).hexdigest()
    return out


def generate(name):
    ks = keyset(name)
    assert ks.seed == KEY_SEED
    orc = make_oracle(ks)
    b1, c1 = fresh_inputs(ks, PAIRS, seed=SEED_A)
    b2, c2 = fresh_inputs(ks, PAIRS, seed=SEED_B)
    outs = np.stack([orc.gate_batch(op, c1, c2) for op in range(6)])
    lin = np.stack([orc.gate_linear(0, c1[g], c2[g]) for g in range(PAIRS)])
    tilde = np.stack([orc.modswitch(lin[g]) for g in range(PAIRS)])
    acc = np.stack([orc.blindrotate(lin[g]) for g in range(PAIRS)])
    dig = key_digests(ks)
    return dict(key_seed=np.uint64(KEY_SEED), seeds=np.array([SEED_A, SEED_B]), bits1=b1, bits2=b2, in1=c1, in2=c2,
                out=outs, nand_linear=lin, nand_tilde=tilde, nand_acc=acc,
                digest_names=np.array(sorted(dig)), digest_values=np.array([dig[k] for k in sorted(dig)]))


if __name__ == "__main__":
    for name in REFERENCE_TEST_SETS:
        d = generate(name)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **d)
        print(name, os.path.getsize(path), "bytes")
