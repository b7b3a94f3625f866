"""Scheme objects: the Python mirror of the reference's `CGGI / LMSS / CCS / KMS / KMS_block` structs
(/root/reference/src/tfhe/scheme.jl:107-116,168-179,209-219,256-265,301-312) whose `btk` now lives in HBM
behind an opaque C-ABI context.  `setup(...)` mirrors scheme.jl:151,190,244,292,343.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .keys import KeySet
from .params import Params

MODE_STRICT, MODE_FAST = 0, 1
NAND_OP, AND_OP, OR_OP, XOR_OP, XNOR_OP, NOR_OP = range(6)
NOT_OP, BOOTSTRAP_OP = 6, -1                     # circuit levels only (include/mktfhe_params.h)
STAGES = ("prep", "phase1", "phase2", "keyswitch")


class MktfheError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class Scheme:
    """One device context holding the uploaded evaluation keys of every party."""

    def __init__(self, params: Params, device: int = 0, devices=None):
        """device: one GPU.  devices: a list of GPUs (or "all") behind ONE front context (mktfhe_ctx_create_multi): keys are
        replicated device to device at finalize() and every gate / bootstrapping batch is sharded in contiguous slices."""
        self.params = params
        self._h = ctypes.c_void_p()
        L = _lib.lib()
        cp = params.c_struct()
        if devices is None:
            self.device, self.devices = device, [device]
            rc = L.mktfhe_ctx_create(ctypes.byref(cp), device, ctypes.byref(self._h))
        else:
            if isinstance(devices, str):
                arr, n = None, 0
            else:
                devs = [int(d) for d in devices]
                arr, n = (ctypes.c_int * len(devs))(*devs), len(devs)
            rc = L.mktfhe_ctx_create_multi(ctypes.byref(cp), n, arr, ctypes.byref(self._h))
            if rc == 0:
                out = (ctypes.c_int * 64)()
                cnt = L.mktfhe_ctx_devices(self._h, out, 64)
                self.devices = [out[i] for i in range(cnt)]
                self.device = self.devices[0]
        if rc != 0:
            raise MktfheError(f"mktfhe_ctx_create: {rc}: {L.mktfhe_last_error(None).decode()}")

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.lib().mktfhe_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise MktfheError(f"{what}: {rc}: {_lib.lib().mktfhe_last_error(self._h).decode()}")

    # -- key upload -------------------------------------------------------------------------------
    def upload_party(self, party: int, brk, ksk, rlk=None, pubb=None):
        for a in (brk, ksk, rlk, pubb):
            assert a is None or a.flags["C_CONTIGUOUS"]
        self._ck(_lib.lib().mktfhe_upload_party_key(self._h, party, _ptr(brk), _ptr(rlk), _ptr(pubb), _ptr(ksk)),
                 "mktfhe_upload_party_key")

    def upload_party_ptr(self, party: int, brk: int, ksk: int, rlk: int | None = None, pubb: int | None = None):
        """Raw-pointer upload (host or device addresses), e.g. of tensors received by NCCL broadcast."""
        self._ck(_lib.lib().mktfhe_upload_party_key(self._h, party, brk, rlk, pubb, ksk), "mktfhe_upload_party_key")

    def upload_common(self, crs_fft):
        ptr = crs_fft if isinstance(crs_fft, int) or crs_fft is None else _ptr(crs_fft)
        self._ck(_lib.lib().mktfhe_upload_common(self._h, ptr), "mktfhe_upload_common")

    def finalize(self):
        self._ck(_lib.lib().mktfhe_finalize_keys(self._h), "mktfhe_finalize_keys")

    # -- key generation on the device (include/mktfhe_b200.h, "key generation on the device") --------------
    @staticmethod
    def _seed_args(seed):
        from . import _host
        if _host.is_key(seed):
            buf = ctypes.create_string_buffer(bytes(seed), 32)
            return 0, ctypes.cast(buf, ctypes.c_void_p), buf
        return int(seed), None, None

    def keygen_common(self, seed):
        """CRS + its FFT form generated on the device (scheme.jl:409-410).  seed: int (reproducible) or 32-byte key."""
        s, k, keep = self._seed_args(seed)
        self._ck(_lib.lib().mktfhe_keygen_common(self._h, s, k), "mktfhe_keygen_common")

    def keygen_party(self, party: int, seed):
        """Evaluation keys of `party` generated on the device from the same streams as the host library (keygen.jl:3-155)."""
        s, k, keep = self._seed_args(seed)
        self._ck(_lib.lib().mktfhe_keygen_party(self._h, party, s, k), "mktfhe_keygen_party")

    def download_party_key(self, party: int) -> dict:
        """Parity hook: a party's keys as they sit in device memory, in the flat upload layouts."""
        p = self.params
        out = {"brk": np.empty((p.n, p.brk_polys, p.H, 2), dtype=np.float64),
               "ksk": np.empty((p.N, p.ksk_rows, p.f, p.n + 1), dtype=np.uint32),
               "rlk": np.empty((p.l_uni, 3, p.H, 2), dtype=np.float64) if p.scheme in (3, 4) else None,
               "pubb": np.empty((p.l_uni, p.H, 2), dtype=np.float64) if p.is_mk else None,
               "crs_fft": np.empty((p.l_uni, p.H, 2), dtype=np.float64) if p.is_mk else None}
        self._ck(_lib.lib().mktfhe_download_party_key(self._h, party, _ptr(out["brk"]), _ptr(out["rlk"]), _ptr(out["pubb"]),
                                                      _ptr(out["ksk"]), _ptr(out["crs_fft"])), "mktfhe_download_party_key")
        return out

    def set_mode(self, mode: int):
        self._ck(_lib.lib().mktfhe_set_mode(self._h, mode), "mktfhe_set_mode")

    @property
    def mode(self) -> int:
        return _lib.lib().mktfhe_get_mode(self._h)

    # -- hot path ---------------------------------------------------------------------------------
    def _batchify(self, c):
        c = np.ascontiguousarray(c, dtype=np.uint32)
        single = c.ndim == 1
        if single:
            c = c[None, :]
        if c.shape[1] != self.params.lwe_words:
            raise ValueError(f"ciphertext has {c.shape[1]} words, expected {self.params.lwe_words}")
        return c, single

    def gate(self, op: int, c1, c2):
        """out = bootstrap(linear_op(c1, c2)): gate.jl:1-52.  Accepts one ciphertext or a batch [B, 1+n*k]."""
        c1, single = self._batchify(c1)
        c2, _ = self._batchify(c2)
        if c1.shape != c2.shape:
            raise ValueError("operand shapes differ")
        out = np.empty_like(c1)
        self._ck(_lib.lib().mktfhe_gate_batch(self._h, op, _ptr(c1), _ptr(c2), _ptr(out), c1.shape[0]), "mktfhe_gate_batch")
        return out[0] if single else out

    def bootstrapping(self, c):
        """bootstrapping!(ctxt, scheme): bootstrapping.jl:4-27 (returns the result instead of mutating)."""
        c, single = self._batchify(c)
        out = np.empty_like(c)
        self._ck(_lib.lib().mktfhe_bootstrap_batch(self._h, _ptr(c), _ptr(out), c.shape[0]), "mktfhe_bootstrap_batch")
        return out[0] if single else out

    def gate_dev(self, op: int, in1_ptr: int, in2_ptr: int, out_ptr: int, batch: int):
        """Device-pointer variant (no copies, asynchronous on the context stream)."""
        self._ck(_lib.lib().mktfhe_gate_batch_dev(self._h, op, in1_ptr, in2_ptr, out_ptr, batch), "mktfhe_gate_batch_dev")

    def sync(self):
        self._ck(_lib.lib().mktfhe_sync(self._h), "mktfhe_sync")

    @property
    def stream(self) -> int:
        return _lib.lib().mktfhe_stream(self._h) or 0

    # -- circuits: device-resident wire table (include/mktfhe_b200.h, "gate circuits") -------------------
    def wires_resize(self, nwires: int):
        self._ck(_lib.lib().mktfhe_wires_resize(self._h, int(nwires)), "mktfhe_wires_resize")

    def wires_write(self, first: int, cts):
        cts, _ = self._batchify(cts)
        self._ck(_lib.lib().mktfhe_wires_write(self._h, int(first), cts.shape[0], _ptr(cts)), "mktfhe_wires_write")

    def wires_read(self, first: int, count: int):
        out = np.empty((int(count), self.params.lwe_words), dtype=np.uint32)
        self._ck(_lib.lib().mktfhe_wires_read(self._h, int(first), int(count), _ptr(out)), "mktfhe_wires_read")
        return out

    def gate_level(self, ops, src1, src2, dst):
        """One level of independent gates over the wire table: wires[dst] = bootstrap(op(wires[src1], wires[src2]))."""
        arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (ops, src1, src2, dst)]
        n = arrs[0].shape[0]
        if any(a.shape != (n,) for a in arrs):
            raise ValueError("ops, src1, src2 and dst must be 1-D arrays of one length")
        self._ck(_lib.lib().mktfhe_gate_level(self._h, *[_ptr(a) for a in arrs], n), "mktfhe_gate_level")

    # -- parity hooks -----------------------------------------------------------------------------
    @property
    def torus_dtype(self):
        return np.uint64 if self.params.torus_bits == 64 else np.uint32

    def gate_linear(self, op, c1, c2):
        c1, single = self._batchify(c1)
        c2, _ = self._batchify(c2)
        out = np.empty_like(c1)
        self._ck(_lib.lib().mktfhe_gate_linear_batch(self._h, op, _ptr(c1), _ptr(c2), _ptr(out), c1.shape[0]), "gate_linear")
        return out[0] if single else out

    def modswitch(self, c):
        c, single = self._batchify(c)
        out = np.empty_like(c)
        self._ck(_lib.lib().mktfhe_modswitch_batch(self._h, _ptr(c), _ptr(out), c.shape[0]), "modswitch")
        return out[0] if single else out

    def blindrotate(self, c):
        c, single = self._batchify(c)
        p = self.params
        acc = np.empty((c.shape[0], p.k + 1, p.N), dtype=self.torus_dtype)
        self._ck(_lib.lib().mktfhe_blindrotate_batch(self._h, _ptr(c), _ptr(acc), c.shape[0]), "blindrotate")
        return acc[0] if single else acc

    def phase1(self, c):
        c, single = self._batchify(c)
        p = self.params
        R = 1 + (p.k - 1) * p.l_lev
        lev = np.empty((c.shape[0], R, 2, p.H, 2), dtype=np.float64)
        self._ck(_lib.lib().mktfhe_phase1_batch(self._h, _ptr(c), _ptr(lev), c.shape[0]), "phase1")
        return lev[0] if single else lev

    def keyswitch(self, acc):
        p = self.params
        acc = np.ascontiguousarray(acc, dtype=self.torus_dtype)
        single = acc.ndim == 2
        if single:
            acc = acc[None]
        out = np.empty((acc.shape[0], p.lwe_words), dtype=np.uint32)
        self._ck(_lib.lib().mktfhe_keyswitch_batch(self._h, _ptr(acc), _ptr(out), acc.shape[0]), "keyswitch")
        return out[0] if single else out

    def cmux_step(self, party, idx, atilde, acc_rows):
        rows = np.array(acc_rows, dtype=self.torus_dtype, order="C", copy=True)
        at = np.ascontiguousarray(atilde, dtype=np.uint32)
        assert rows.ndim == 3 and rows.shape[0] == at.shape[0]
        self._ck(_lib.lib().mktfhe_cmux_step_batch(self._h, party, idx, _ptr(at), _ptr(rows), rows.shape[0]), "cmux_step")
        return rows

    def block_step(self, party, blk, atilde, acc_rows):
        rows = np.array(acc_rows, dtype=self.torus_dtype, order="C", copy=True)
        at = np.ascontiguousarray(atilde, dtype=np.uint32)
        assert rows.ndim == 3 and at.shape == (rows.shape[0], self.params.ell)
        self._ck(_lib.lib().mktfhe_block_step_batch(self._h, party, blk, _ptr(at), _ptr(rows), rows.shape[0]), "block_step")
        return rows

    def gadget_product(self, polys, keys, l, logB):
        """out[g][c] = native(ifft(Sum_j fft(D_j(polys[g])) * keys[j][c])): one product of FAST phase 2 (KMS*, Torus64) or of the FAST
        CCS hybrid product (N = 1024, Torus32)."""
        dt = self.torus_dtype
        polys = np.ascontiguousarray(polys, dtype=dt)
        keys = np.ascontiguousarray(keys, dtype=np.float64)            # [l][ncomp][H][2]
        assert keys.shape[0] == l and keys.shape[2:] == (self.params.H, 2)
        out = np.empty((polys.shape[0], keys.shape[1], self.params.N), dtype=dt)
        fn = _lib.lib().mktfhe_gadget_product_batch if dt == np.uint64 else _lib.lib().mktfhe_gadget_product32_batch
        self._ck(fn(self._h, l, logB, _ptr(polys), _ptr(keys), keys.shape[1], _ptr(out), polys.shape[0]), "gadget_product")
        return out

    def fft(self, polys):
        polys = np.ascontiguousarray(polys)
        bits = polys.dtype.itemsize * 8
        out = np.empty((polys.shape[0], self.params.H, 2), dtype=np.float64)
        self._ck(_lib.lib().mktfhe_fft_batch(self._h, bits, _ptr(polys), _ptr(out), polys.shape[0]), "fft")
        return out

    def ifft(self, spectra, bits):
        spectra = np.ascontiguousarray(spectra, dtype=np.float64)
        out = np.empty((spectra.shape[0], self.params.N), dtype=np.uint64 if bits == 64 else np.uint32)
        self._ck(_lib.lib().mktfhe_ifft_batch(self._h, bits, _ptr(spectra), _ptr(out), spectra.shape[0]), "ifft")
        return out

    def decomp(self, polys, l, logB):
        polys = np.ascontiguousarray(polys)
        bits = polys.dtype.itemsize * 8
        out = np.empty((polys.shape[0], l, self.params.N), dtype=polys.dtype)
        self._ck(_lib.lib().mktfhe_decomp_batch(self._h, bits, l, logB, _ptr(polys), _ptr(out), polys.shape[0]), "decomp")
        return out

    # -- measurement ------------------------------------------------------------------------------
    def last_stage_ms(self):
        ms = (ctypes.c_float * 4)()
        n = ctypes.c_int()
        self._ck(_lib.lib().mktfhe_last_stage_ms(self._h, ms, ctypes.byref(n)), "last_stage_ms")
        return dict(zip(STAGES, [float(x) for x in ms])), n.value

    def dfma_peak_tflops(self) -> float:
        v = ctypes.c_double()
        self._ck(_lib.lib().mktfhe_measure_dfma_peak(self._h, ctypes.byref(v)), "dfma_peak")
        return v.value


def setup_generated(params: Params, seed, device: int = 0, mode: int | None = None, devices=None):
    """Like `setup`, with every evaluation key generated ON THE DEVICE (no host key generation, no upload): returns
    (Scheme, KeySet) where the KeySet holds the secret keys only (same seed, host) for encrypting and decrypting.
    seed: int.  For a given seed the device keys are byte-identical to KeySet(params, seed)'s."""
    s = Scheme(params, device, devices)
    if params.is_mk:
        s.keygen_common(seed)
    for i in range(params.k if params.is_mk else 1):
        s.keygen_party(i, seed)
    s.finalize()
    if mode is not None:
        s.set_mode(mode)
    return s, KeySet(params, seed=seed, secret_only=True)


def setup(keys: KeySet, device: int = 0, mode: int | None = None, devices=None) -> Scheme:
    """`scheme = setup(a, btk, params)` / `setup(params)`: create the device context and upload every party's keys.
    devices=[0, 1, ...] (or "all"): one front context over several GPUs of this process."""
    p = keys.params
    s = Scheme(p, device, devices)
    for i, q in enumerate(keys.parties):
        s.upload_party(i, q["brk"], q["ksk"], q["rlk"], q["pubb"])
    if p.is_mk:
        s.upload_common(keys.crs_fft)
    s.finalize()
    if mode is not None:
        s.set_mode(mode)
    return s
