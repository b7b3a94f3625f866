"""Multi-GPU: one process per GPU, replicated keys, gates sharded by contiguous batch slices.

Bootstrapped gates share nothing but read-only keys, so the only collectives are a one-time key
broadcast from rank 0 (NCCL over NVLink; `gloo` in the CPU tests) and an optional gather of result slices.
Nothing here runs inside the timed hot path.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .keys import KeySet
from .params import Params


def shard_range(batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of a batch owned by `rank`; sizes differ by at most one gate."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def key_arrays(p: Params):
    """(name, shape, numpy dtype) of every per-party evaluation-key array, in upload order."""
    out = [("brk", (p.n, p.brk_polys, p.H, 2), np.float64), ("ksk", (p.N, p.ksk_rows, p.f, p.n + 1), np.uint32)]
    if p.scheme in (3, 4):
        out.append(("rlk", (p.l_uni, 3, p.H, 2), np.float64))
    if p.is_mk:
        out.append(("pubb", (p.l_uni, p.H, 2), np.float64))
    return out


_TORCH = {np.float64: torch.float64, np.uint32: torch.int32}


def broadcast_keys(ks: KeySet | None, p: Params, device, src: int = 0):
    """Rank `src` holds `ks`; every rank returns [{name: tensor on device}] per party plus the CRS tensor."""
    nparties = p.k if p.is_mk else 1
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    parties = []
    for i in range(nparties):
        d = {}
        for name, shape, dt in key_arrays(p):
            if rank == src:
                arr = ks.parties[i][name]
                t = torch.from_numpy(arr.view(np.int32) if dt == np.uint32 else arr).to(device)
            else:
                t = torch.empty(shape, dtype=_TORCH[dt], device=device)
            if world > 1:
                dist.broadcast(t, src=src)
            d[name] = t
        parties.append(d)
    crs = None
    if p.is_mk:
        crs = torch.from_numpy(ks.crs_fft).to(device) if rank == src else torch.empty((p.l_uni, p.H, 2), dtype=torch.float64, device=device)
        if world > 1:
            dist.broadcast(crs, src=src)
    return parties, crs


def setup_replicated(p: Params, seed: int, device_index: int, rank: int, world: int, timings: bool = False, keygen: str = "host"):
    """keygen="host": key generation on rank 0 (host library), NCCL broadcast, upload from device memory on every rank.
    keygen="device": every rank generates the (identical, seed-determined) key set on its own GPU -- nothing crosses PCIe or NVLink.
    Returns (Scheme, KeySet); ranks other than 0 hold secret keys only (for encrypting / checking their shard).
    timings=True appends {"keygen_s", "broadcast_s", "upload_finalize_s", "key_bytes"} (this rank's wall clock)."""
    from .scheme import Scheme, setup_generated
    import os
    import time
    t0 = time.perf_counter()
    if keygen == "device":
        s, ks = setup_generated(p, seed, device=device_index)
        if timings:
            nparties = p.k if p.is_mk else 1
            nbytes = nparties * sum(int(np.prod(shape)) * np.dtype(dt).itemsize for _, shape, dt in key_arrays(p))
            return s, ks, {"mode": "device", "keygen_s": time.perf_counter() - t0, "broadcast_s": 0.0, "upload_finalize_s": 0.0, "key_bytes": nbytes,
                           "note": "evaluation keys generated on each GPU from the seed (csrc/keygen.cuh), byte-identical to the host library's; includes finalize"}
        return s, ks
    ks = KeySet(p, seed=seed, secret_only=(rank != 0), nthreads=max(1, len(os.sched_getaffinity(0)) // max(1, min(world, 8))) if rank else len(os.sched_getaffinity(0)))
    t1 = t2 = time.perf_counter()
    s = Scheme(p, device_index)
    if world == 1:
        for i, q in enumerate(ks.parties):
            s.upload_party(i, q["brk"], q["ksk"], q["rlk"], q["pubb"])
        if p.is_mk:
            s.upload_common(ks.crs_fft)
    else:
        dev = torch.device("cuda", device_index)
        dist.barrier()                                   # the other ranks wait here for rank 0's key generation
        t1 = time.perf_counter()
        parties, crs = broadcast_keys(ks if rank == 0 else None, p, dev)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        for i, d in enumerate(parties):
            s.upload_party_ptr(i, d["brk"].data_ptr(), d["ksk"].data_ptr(),
                               d["rlk"].data_ptr() if "rlk" in d else None, d["pubb"].data_ptr() if "pubb" in d else None)
        if p.is_mk:
            s.upload_common(crs.data_ptr())
        del parties, crs
    s.finalize()
    if timings:
        nparties = p.k if p.is_mk else 1
        nbytes = nparties * sum(int(np.prod(shape)) * np.dtype(dt).itemsize for _, shape, dt in key_arrays(p))
        return s, ks, {"mode": "host", "keygen_s": t1 - t0, "broadcast_s": (t2 - t1) if world > 1 else 0.0, "upload_finalize_s": time.perf_counter() - t2,
                       "key_bytes": nbytes, "note": "rank 0 generates (host, OpenMP); NCCL broadcast GPU->GPU; upload = device-to-device copy + FAST layouts"}
    return s, ks


def gather_results(local: np.ndarray, batch: int, dst: int = 0):
    """Result gather: every rank contributes its slice; rank `dst` returns the [batch, words] array."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return local
    words = local.shape[1]
    sizes = [shard_range(batch, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = np.zeros((mx, words), dtype=np.int32)
    pad[: local.shape[0]] = local.view(np.int32)
    t = torch.from_numpy(pad)
    bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, bufs, dst=dst)
    if rank != dst:
        return None
    return np.concatenate([bufs[r].numpy()[: hi - lo] for r, (lo, hi) in enumerate(sizes)]).view(np.uint32)
