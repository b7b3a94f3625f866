#!/usr/bin/env python
"""One process, ONE call, N GPUs: end-to-end MK-NAND throughput through a front context (mktfhe_ctx_create_multi).

    python bench/front_context_bench.py [--workload kms2] [--batch-per-gpu 4096] [--steps 3] [--warmup 2] [--devices 0,1,...]

This is the path a single-process caller (the Julia binding: `upload(scheme, params; devices = 0:7)`) uses instead of one
process per GPU: keys are generated / uploaded once and replicated device to device, every `mktfhe_gate_batch` call on HOST
buffers is sharded in contiguous slices, one worker thread and stream per device, H2D and D2H copies inside the call.
Prints one JSON line: gates/s over all devices (wall clock around the synchronous calls: the call returns when the results
are in the host buffer), the same for a single-device context of the same process, and whether the two results are
bit-identical (gates are independent, so they must be).  bench.py (one rank per GPU, torchrun) is the driver's contract; this
script measures the C-ABI route of SURVEY 8(b)/(e)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench as B
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="kms2", choices=sorted(B.WORKLOADS))
    ap.add_argument("--batch-per-gpu", type=int, default=None)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--devices", default=None, help="comma-separated device indices (default: all visible)")
    args = ap.parse_args()

    import torch
    from mktfhe_b200 import _lib, params as P
    from mktfhe_b200.scheme import MODE_FAST, setup_generated
    pname, default_batch = B.WORKLOADS[args.workload]
    p = P.ALL[pname]
    devs = [int(d) for d in args.devices.split(",")] if args.devices else list(range(torch.cuda.device_count()))
    per = args.batch_per_gpu or default_batch
    batch = per * len(devs)

    t0 = time.perf_counter()
    front, ks = setup_generated(p, B.KEY_SEED, devices=devs, mode=MODE_FAST)
    setup_s = time.perf_counter() - t0
    m1, m2, c1, c2 = B.make_inputs(ks, batch, 0)
    h1 = torch.from_numpy(c1.view(np.int32)).pin_memory()
    h2 = torch.from_numpy(c2.view(np.int32)).pin_memory()
    hout = torch.empty_like(h1).pin_memory()
    L = _lib.lib()

    def run(scheme, n, out):
        rc = L.mktfhe_gate_batch(scheme._h, 0, h1.data_ptr(), h2.data_ptr(), out.data_ptr(), n)
        if rc != 0:
            raise RuntimeError(L.mktfhe_last_error(scheme._h).decode())

    def timed(scheme, n, out):
        for _ in range(args.warmup):
            run(scheme, n, out)
        t = time.perf_counter()
        for _ in range(args.steps):
            run(scheme, n, out)
        return (time.perf_counter() - t) / args.steps

    s_front = timed(front, batch, hout)
    res_front = hout.numpy().view(np.uint32).copy()
    dec = ks.decrypt_batch(res_front[: min(batch, 512)])
    ok = int((dec == ~(m1 & m2)[: len(dec)]).sum())
    front.close()

    single, _ = setup_generated(p, B.KEY_SEED, device=devs[0], mode=MODE_FAST)
    hout1 = torch.empty_like(h1).pin_memory()
    s_single = timed(single, per, hout1)            # the first device's share of the same inputs
    run(single, batch, hout1)                       # and the whole batch, for the bit-for-bit comparison
    identical = bool(np.array_equal(hout1.numpy().view(np.uint32), res_front))
    single.close()

    line = {"metric": B.METRIC, "unit": B.UNIT, "route": "one process, one front context (mktfhe_ctx_create_multi), host buffers",
            "n_gpus": len(devs), "devices": devs, "steps": args.steps, "warmup": args.warmup,
            "config": {"workload": f"{pname} MK-NAND, {per} gates per GPU per call", "params": pname, "batch_total": batch},
            "e2e": {"value": batch / s_front, "ms_per_call": 1e3 * s_front,
                    "h2d_bytes_per_step": 2 * batch * p.lwe_words * 4, "d2h_bytes_per_step": batch * p.lwe_words * 4},
            "single_device_e2e": {"value": per / s_single, "ms_per_call": 1e3 * s_single},
            "speedup_over_single_device": (batch / s_front) / (per / s_single),
            "bit_identical_to_single_device": identical,
            "decrypt_check": {"checked": len(dec), "ok": ok},
            "key_setup_s": setup_s, "timing": "host wall clock around synchronous mktfhe_gate_batch calls"}
    print(json.dumps(line))
    return 0 if identical and ok >= 0.97 * len(dec) else 3


if __name__ == "__main__":
    sys.exit(main())
