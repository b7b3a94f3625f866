"""Two parties add their encrypted 8-bit numbers on the GPU (needs a B200):  python examples/adder.py [instances]

Party 1 encrypts a, party 2 encrypts b (lwe_ith_encrypt, test/KMS.jl:17-22); the evaluator holds only evaluation keys and runs a
ripple-carry adder level by level (`mktfhe_gate_level`), `instances` independent additions per call; decryption needs both keys."""
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from mktfhe_b200 import params as P  # noqa: E402
from mktfhe_b200.circuit import ripple_adder  # noqa: E402
from mktfhe_b200.reference_api import CRS, lwe_decrypt, lwe_ith_encrypt, party_keygen, setup  # noqa: E402


def main():
    inst = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    bits, params = 8, P.KMS2party
    a = CRS(params)
    keys = [party_keygen(a, params) for _ in range(params.k)]
    lwekeys, btk = [q[0] for q in keys], [q[-1] for q in keys]
    scheme = setup(a, btk, params)
    rng = np.random.default_rng()
    x, y = rng.integers(0, 1 << bits, inst), rng.integers(0, 1 << bits, inst)
    cts = np.stack([np.stack([lwe_ith_encrypt((v >> i) & 1, party, lwekeys[party - 1], params)
                              for v, party in ((x[j], 1), (y[j], 2)) for i in range(bits)]) for j in range(inst)])
    adder = ripple_adder(bits)
    t0 = time.perf_counter()
    out = adder.evaluate(scheme, cts)
    dt = time.perf_counter() - t0
    got = np.array([sum(int(lwe_decrypt(out[j, i], lwekeys, params)) << i for i in range(bits + 1)) for j in range(inst)])
    assert np.array_equal(got, x + y), "wrong sums"
    gates = adder.bootstrapped_gates() * inst
    print(f"{inst} encrypted {bits}-bit additions: {gates} gate bootstraps in {adder.depth()} levels, {dt:.2f} s "
          f"({gates / dt:.0f} gates/s), all sums correct")


if __name__ == "__main__":
    main()
