"""Flat on-disk format for key material and ciphertexts.

The reference never serialises keys (every test and benchmark regenerates them: test/KMS.jl:6-12).  A blob keeps the
flat upload layouts of include/mktfhe_b200.h byte for byte, so loading is `mmap` + `mktfhe_upload_party_key` with no
conversion, and a blob written on one machine is the upload input on another.

    offset 0      magic  b"MKTFHEB2"
           8      uint32 version (1), uint32 header bytes (little endian)
           16     JSON header, UTF-8: {"kind": "keys" | "ciphertexts", "params": {...}, "seed": .. (null unless include_secret),
                  "arrays": [{"name", "dtype", "shape", "offset", "nbytes", "sha256"}, ...]}
           ...    each array at a 4096-byte aligned offset, C order, little endian

Array names: `crs_fft`, `crs_coeff`, and per party i `p{i}.brk`, `p{i}.ksk`, `p{i}.rlk`, `p{i}.pubb` (evaluation keys)
and, only with include_secret=True, `p{i}.lwekey`, `p{i}.ringkey`.
"""
from __future__ import annotations

import dataclasses
import hashlib
import json
import struct

import numpy as np

from .keys import KeySet
from .params import Params

MAGIC = b"MKTFHEB2"
VERSION = 1
ALIGN = 4096
_EVAL = ("brk", "ksk", "rlk", "pubb")
_SECRET = ("lwekey", "ringkey")


class BlobError(ValueError):
    pass


def _write(path, kind, params: Params, seed, arrays):
    directory, off = [], 0
    for name, a in arrays:
        a = np.ascontiguousarray(a)
        directory.append({"name": name, "dtype": a.dtype.str, "shape": list(a.shape), "offset": off, "nbytes": a.nbytes,
                          "sha256": hashlib.sha256(memoryview(a).cast("B")).hexdigest()})
        off += (a.nbytes + ALIGN - 1) // ALIGN * ALIGN
    header = {"kind": kind, "params": dataclasses.asdict(params), "seed": None if seed is None else int(seed), "arrays": directory}
    hj = json.dumps(header).encode()
    data0 = (16 + len(hj) + ALIGN - 1) // ALIGN * ALIGN
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<II", VERSION, len(hj)) + hj)
        for (name, a), d in zip(arrays, directory):
            f.seek(data0 + d["offset"])
            f.write(memoryview(np.ascontiguousarray(a)).cast("B"))
        f.truncate(data0 + off)
    return data0 + off


def _read(path, want_kind, mmap=True, verify=False):
    with open(path, "rb") as f:
        head = f.read(16)
        if len(head) < 16 or head[:8] != MAGIC:
            raise BlobError(f"{path}: not a mktfhe-b200 blob")
        version, hlen = struct.unpack("<II", head[8:])
        if version != VERSION:
            raise BlobError(f"{path}: blob version {version}, this build reads {VERSION}")
        header = json.loads(f.read(hlen).decode())
    if header["kind"] != want_kind:
        raise BlobError(f"{path}: holds {header['kind']}, expected {want_kind}")
    data0 = (16 + hlen + ALIGN - 1) // ALIGN * ALIGN
    raw = np.memmap(path, dtype=np.uint8, mode="r") if mmap else np.fromfile(path, dtype=np.uint8)
    out = {}
    for d in header["arrays"]:
        lo = data0 + d["offset"]
        if lo + d["nbytes"] > raw.shape[0]:
            raise BlobError(f"{path}: truncated (array {d['name']})")
        a = raw[lo:lo + d["nbytes"]].view(np.dtype(d["dtype"])).reshape(d["shape"])
        if verify and hashlib.sha256(memoryview(np.ascontiguousarray(a)).cast("B")).hexdigest() != d["sha256"]:
            raise BlobError(f"{path}: checksum mismatch in array {d['name']}")
        out[d["name"]] = a
    params = Params(**header["params"])
    return header, params, out


# ---- keys -----------------------------------------------------------------------------------------------

def save_keys(path: str, keys: KeySet, include_secret: bool = False) -> int:
    """Write every party's evaluation keys (and the CRS) of a KeySet; returns the file size."""
    p = keys.params
    arrays = []
    if p.is_mk:
        arrays += [("crs_fft", keys.crs_fft), ("crs_coeff", keys.crs_coeff)]
    for i, q in enumerate(keys.parties):
        for f in _EVAL + (_SECRET if include_secret else ()):
            if q.get(f) is not None:
                arrays.append((f"p{i}.{f}", q[f]))
    # The seed regenerates every secret (an int seed is a test convenience): it goes into the file only together with the secrets.
    seed = keys.seed if (include_secret and isinstance(keys.seed, int)) else None
    return _write(path, "keys", p, seed, arrays)


class LoadedKeys(KeySet):
    """A KeySet backed by a blob: same attributes, arrays are read-only views of the mapped file.
    Encrypt / decrypt helpers work only if the blob was written with include_secret=True."""

    def __init__(self, params, seed, arrays):
        self.params, self.seed = params, seed
        self.crs_fft, self.crs_coeff = arrays.get("crs_fft"), arrays.get("crs_coeff")
        n = params.k if params.is_mk else 1
        self.parties = [{f: arrays.get(f"p{i}.{f}") for f in _EVAL + _SECRET} for i in range(n)]
        for i, q in enumerate(self.parties):
            if q["brk"] is None or q["ksk"] is None:
                raise BlobError(f"blob lacks the evaluation keys of party {i}")
        have = [q["lwekey"] for q in self.parties]
        self.lwekeys = np.ascontiguousarray(np.stack(have)) if all(k is not None for k in have) else None

    def _cp(self):
        if self.lwekeys is None:
            raise BlobError("this blob holds evaluation keys only (written without include_secret)")
        return super()._cp()


def load_keys(path: str, mmap: bool = True, verify: bool = False) -> LoadedKeys:
    header, params, arrays = _read(path, "keys", mmap, verify)
    return LoadedKeys(params, header["seed"], arrays)


# ---- ciphertexts ----------------------------------------------------------------------------------------

def save_ciphertexts(path: str, params: Params, cts) -> int:
    cts = np.ascontiguousarray(cts, dtype=np.uint32)
    if cts.ndim != 2 or cts.shape[1] != params.lwe_words:
        raise ValueError(f"expected [count, {params.lwe_words}] uint32 LWE records")
    return _write(path, "ciphertexts", params, None, [("lwe", cts)])


def load_ciphertexts(path: str, params: Params | None = None, mmap: bool = False) -> np.ndarray:
    _h, p, arrays = _read(path, "ciphertexts", mmap, verify=True)
    if params is not None and dataclasses.asdict(p) != dataclasses.asdict(params):
        raise BlobError(f"{path}: ciphertexts of parameter set {p.name}, expected {params.name}")
    return arrays["lwe"]


# ---- reference fixtures (julia/dump_fixtures.jl) -----------------------------------------------------------

FIXTURE_GATES = ("NAND", "AND", "OR", "XOR", "XNOR", "NOR")


def save_fixture(path: str, params: Params, arrays: dict) -> int:
    """Write a "fixture" blob: input pairs, the outputs of all six gates and the intermediate stages of NAND
    (the arrays julia/dump_fixtures.jl writes from a run of the reference)."""
    return _write(path, "fixture", params, None, list(arrays.items()))


def load_fixture(path: str, verify: bool = True):
    """-> (params, {name: array}) of a fixture blob; checks that every expected array is present and consistent."""
    _h, p, a = _read(path, "fixture", mmap=False, verify=verify)
    need = ["in1", "in2", "bits1", "bits2", "nand_linear", "nand_tilde", "nand_acc"] + [f"out_{g}" for g in FIXTURE_GATES]
    missing = [n for n in need if n not in a]
    if missing:
        raise BlobError(f"{path}: fixture lacks {missing}")
    B = a["in1"].shape[0]
    for n in need:
        if a[n].shape[0] != B:
            raise BlobError(f"{path}: array {n} has {a[n].shape[0]} rows, expected {B}")
    if a["in1"].shape[1] != p.lwe_words or a["nand_acc"].shape[1:] != (p.k + 1, p.N):
        raise BlobError(f"{path}: array shapes do not match parameter set {p.name}")
    return p, a
