"""Latency of small `mktfhe_gate_batch` calls (the reference API is per gate: bootstrapping.jl:4 handles ONE ciphertext).

    python bench/latency_probe.py PARAMS REPS [BATCH ...]      # default batches: 1 3

Each call goes through the C-ABI with host buffers (`Scheme.gate`), is repeated REPS times and must return the same
words every time -- this is also the repeat test that exposed the key-ring race of the one-warp phase-1 kernel
(profiles/README_r2.md).  Prints one JSON line per batch size."""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np

import bench as B
from mktfhe_b200 import params as P
from mktfhe_b200.scheme import MODE_FAST, setup_generated


def main():
    name, reps = sys.argv[1], int(sys.argv[2])
    batches = [int(b) for b in sys.argv[3:]] or [1, 3]
    p = P.ALL[name]
    s, ks = setup_generated(p, B.KEY_SEED, mode=MODE_FAST)
    for batch in batches:
        m1, m2, c1, c2 = B.make_inputs(ks, batch, 0)
        first = s.gate(0, c1, c2)
        ok = int(np.sum(np.asarray(ks.decrypt_batch(first)).astype(bool) == ~(m1 & m2)))
        t = time.perf_counter()
        for _ in range(reps):
            assert np.array_equal(s.gate(0, c1, c2), first), "a repeated call returned different words"
        ms = 1e3 * (time.perf_counter() - t) / reps
        print(json.dumps({"params": name, "batch": batch, "calls": reps, "identical": True, "ms_per_call": round(ms, 3),
                          "decrypt_ok": ok, "of": batch}), flush=True)


if __name__ == "__main__":
    main()
