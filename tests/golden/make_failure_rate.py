"""Failure-rate fixtures for the parameter sets whose noise sits at the decision margin in the reference algorithm itself
(CCS16party, KMS32party: /root/reference/src/tfhe/params.jl:39-45,79-85; the reference's own tests never run them).

    python tests/golden/make_failure_rate.py KMS32party 512 [threads]

Runs the CPU oracle (oracle/, the C restatement of the reference) on `count` fresh MK-NAND gates with seeded keys and
inputs and stores, per gate: the first 8 bytes of SHA-256 over the output ciphertext, the decrypted bit, the plaintext
truth and the output phase error on Torus32.  The GPU tests (tests/test_gpu_big_sets.py) regenerate the same keys and
inputs from the seeds, require the STRICT path to reproduce every digest, and compare the FAST path's failure count and
noise with these oracle statistics.  The vectors are oracle outputs, not reference outputs (Julia is absent)."""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

KEY_SEED = 0x4D4B5446
SEED1, SEED2 = 71, 72


def inputs(ks, count):
    from conftest import fresh_inputs
    b1, c1 = fresh_inputs(ks, count, seed=SEED1)
    b2, c2 = fresh_inputs(ks, count, seed=SEED2)
    return b1, b2, c1, c2


def phase_errors(ks, out, want):
    errs = np.empty(len(out), dtype=np.int64)
    for g in range(len(out)):
        e = (ks.phase(out[g]) - ((1 << 29) if want[g] else (7 << 29))) & 0xFFFFFFFF
        errs[g] = e - (1 << 32) if e >= (1 << 31) else e
    return errs


def digest8(ct) -> np.uint64:
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(ct, dtype=np.uint32).tobytes()).digest()[:8], dtype=np.uint64)[0]


def main():
    name, count = sys.argv[1], int(sys.argv[2])
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    from conftest import keyset, make_oracle
    ks = keyset(name)
    orc = make_oracle(ks)
    b1, b2, c1, c2 = inputs(ks, count)
    t = time.time()
    out = orc.gate_batch(0, c1, c2, threads)
    dt = time.time() - t
    want = ~(b1 & b2)
    dec = ks.decrypt_batch(out)
    errs = phase_errors(ks, out, want)
    dig = np.array([digest8(o) for o in out], dtype=np.uint64)
    path = os.path.join(HERE, f"failrate_{name}.npz")
    np.savez_compressed(path, name=name, key_seed=KEY_SEED, seed1=SEED1, seed2=SEED2, count=count, want=want, dec=dec,
                        phase_err=errs, digest8=dig, first_output=out[0])
    fails = int(np.sum(dec != want))
    print(f"{name}: {count} MK-NAND gates by the oracle in {dt:.0f} s; failures {fails} ({100.0 * fails / count:.2f} %), "
          f"phase-error std 2^{np.log2(errs.std()):.2f} -> {path}")


if __name__ == "__main__":
    main()
