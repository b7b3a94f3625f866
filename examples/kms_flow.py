"""test/KMS.jl of the reference, line by line, on the GPU library (needs a B200):  python examples/kms_flow.py [KMS2party]"""
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from mktfhe_b200 import params as P  # noqa: E402
from mktfhe_b200.reference_api import (AND, CRS, NAND, NOR, OR, XNOR, XOR, bootstrapping_, lwe_decrypt, lwe_ith_encrypt,  # noqa: E402
                                       party_keygen, setup)


def main():
    params = P.ALL[sys.argv[1] if len(sys.argv) > 1 else "KMS2party"]
    a = CRS(params)
    print("KEY GENERATION ...")
    keys = [party_keygen(a, params) for _ in range(params.k)]
    lwekeys, btk = [q[0] for q in keys], [q[-1] for q in keys]
    print(f"BRK SIZE : {btk[0].brk.nbytes >> 20} MiB, RLK SIZE : {btk[0].rlk.nbytes >> 10} KiB, KSK SIZE : {btk[0].ksk.nbytes >> 20} MiB\n")
    scheme = setup(a, btk, params)
    gates = [(NAND, lambda x, y: not (x and y), "NAND "), (AND, lambda x, y: x and y, "& "), (OR, lambda x, y: x or y, "|| "),
             (XOR, lambda x, y: x != y, "XOR "), (XNOR, lambda x, y: x == y, "XNOR "), (NOR, lambda x, y: not (x or y), "NOR ")]
    rng = np.random.default_rng()
    for idx in range(1, 6):
        m = rng.integers(0, 2, params.k).astype(bool)
        ctxts = [lwe_ith_encrypt(m[i - 1], i, lwekeys[i - 1], params) for i in range(1, params.k + 1)]
        res, mres, circuit = ctxts[0], bool(m[0]), "m1 "
        for i in range(2, params.k + 1):
            gate, plain, name = gates[rng.integers(0, 6)]
            res = gate(res, ctxts[i - 1], scheme)
            mres = plain(mres, bool(m[i - 1]))
            circuit += name + f"m{i} "
        t0 = time.perf_counter()
        bootstrapping_(res, scheme)
        dt = time.perf_counter() - t0
        assert mres == lwe_decrypt(res, lwekeys, params)
        print(f"Trial {idx} : {circuit}= {mres}   (bootstrapping! {dt * 1e3:.2f} ms)")


if __name__ == "__main__":
    main()
