"""Multi-GPU host logic on CPU: world_size-2 gloo processes exercise the key broadcast and the result gather
(mktfhe_b200/dist.py).  The hot path itself has no collective (gates are independent)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mktfhe_b200 import params as P
from mktfhe_b200.dist import broadcast_keys, gather_results, key_arrays, shard_range


def test_shard_range_partitions_contiguously():
    for batch in (0, 1, 7, 16, 4096, 16385):
        for world in (1, 2, 3, 8):
            spans = [shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(16384, 3, 8) == (6144, 8192)          # BASELINE config 3: 2048 gates per GPU


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mktfhe_b200.keys import KeySet
        p = P.small(P.KMS2party, n=6)
        ks = KeySet(p, seed=11) if rank == 0 else None
        parties, crs = broadcast_keys(ks, p, torch.device("cpu"))
        ref = KeySet(p, seed=11)                               # every rank can regenerate to check what it received
        ok = True
        for i, d in enumerate(parties):
            for name, shape, dt in key_arrays(p):
                got = d[name].numpy()
                want = ref.parties[i][name]
                ok &= got.shape == tuple(shape) and np.array_equal(got.view(np.uint8), want.view(np.uint8))
        ok &= np.array_equal(crs.numpy(), ref.crs_fft)
        # result gather of uneven shards
        batch, words = 7, p.lwe_words
        full = np.arange(batch * words, dtype=np.uint32).reshape(batch, words)
        lo, hi = shard_range(batch, rank, world)
        out = gather_results(full[lo:hi].copy(), batch, dst=0)
        if rank == 0:
            ok &= np.array_equal(out, full)
        else:
            ok &= out is None
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_key_broadcast_and_result_gather_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
    assert res == [(0, True), (1, True)]
