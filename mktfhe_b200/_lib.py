"""ctypes binding of libmktfhe_b200.so (include/mktfhe_b200.h).  Fails loudly when the CUDA library is
missing or no device is present: there is no CPU fallback."""
from __future__ import annotations

import ctypes
import os

from . import build
from .params import CParams

_lib = None

EXPORTS = [
    "mktfhe_ctx_create", "mktfhe_ctx_destroy", "mktfhe_last_error", "mktfhe_set_mode", "mktfhe_get_mode",
    "mktfhe_upload_party_key", "mktfhe_upload_common", "mktfhe_finalize_keys",
    "mktfhe_gate_batch", "mktfhe_bootstrap_batch", "mktfhe_gate_batch_dev", "mktfhe_sync", "mktfhe_stream",
    "mktfhe_gate_linear_batch", "mktfhe_modswitch_batch", "mktfhe_blindrotate_batch", "mktfhe_phase1_batch",
    "mktfhe_keyswitch_batch", "mktfhe_cmux_step_batch", "mktfhe_block_step_batch", "mktfhe_fft_batch", "mktfhe_ifft_batch",
    "mktfhe_decomp_batch", "mktfhe_last_stage_ms", "mktfhe_measure_dfma_peak",
    "mktfhe_wires_resize", "mktfhe_wires_write", "mktfhe_wires_read", "mktfhe_gate_level",
    "mktfhe_ctx_create_multi", "mktfhe_ctx_devices",
    "mktfhe_gadget_product_batch", "mktfhe_gadget_product32_batch", "mktfhe_keygen_common", "mktfhe_keygen_party", "mktfhe_download_party_key",
]


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = build.CUDA_LIB
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -m mktfhe_b200.build` (needs nvcc). "
                           "The hot path has no CPU fallback.")
    L = ctypes.CDLL(path)
    vp, i32, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    L.mktfhe_ctx_create.argtypes = [ctypes.POINTER(CParams), i32, ctypes.POINTER(vp)]
    L.mktfhe_ctx_create_multi.argtypes = [ctypes.POINTER(CParams), i32, vp, ctypes.POINTER(vp)]
    L.mktfhe_ctx_devices.argtypes = [vp, vp, i32]
    L.mktfhe_keygen_common.argtypes = [vp, ctypes.c_uint64, vp]
    L.mktfhe_keygen_party.argtypes = [vp, i32, ctypes.c_uint64, vp]
    L.mktfhe_download_party_key.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.mktfhe_gadget_product_batch.argtypes = [vp, i32, i32, vp, vp, i32, vp, sz]
    L.mktfhe_gadget_product32_batch.argtypes = [vp, i32, i32, vp, vp, i32, vp, sz]
    L.mktfhe_ctx_destroy.argtypes = [vp]
    L.mktfhe_ctx_destroy.restype = None
    L.mktfhe_last_error.argtypes = [vp]
    L.mktfhe_last_error.restype = ctypes.c_char_p
    L.mktfhe_set_mode.argtypes = [vp, i32]
    L.mktfhe_get_mode.argtypes = [vp]
    L.mktfhe_upload_party_key.argtypes = [vp, i32, vp, vp, vp, vp]
    L.mktfhe_upload_common.argtypes = [vp, vp]
    L.mktfhe_finalize_keys.argtypes = [vp]
    L.mktfhe_gate_batch.argtypes = [vp, i32, vp, vp, vp, sz]
    L.mktfhe_bootstrap_batch.argtypes = [vp, vp, vp, sz]
    L.mktfhe_gate_batch_dev.argtypes = [vp, i32, vp, vp, vp, sz]
    L.mktfhe_sync.argtypes = [vp]
    L.mktfhe_stream.argtypes = [vp]
    L.mktfhe_stream.restype = vp
    L.mktfhe_wires_resize.argtypes = [vp, sz]
    L.mktfhe_wires_write.argtypes = [vp, sz, sz, vp]
    L.mktfhe_wires_read.argtypes = [vp, sz, sz, vp]
    L.mktfhe_gate_level.argtypes = [vp, vp, vp, vp, vp, sz]
    L.mktfhe_gate_linear_batch.argtypes = [vp, i32, vp, vp, vp, sz]
    L.mktfhe_modswitch_batch.argtypes = [vp, vp, vp, sz]
    L.mktfhe_blindrotate_batch.argtypes = [vp, vp, vp, sz]
    L.mktfhe_phase1_batch.argtypes = [vp, vp, vp, sz]
    L.mktfhe_keyswitch_batch.argtypes = [vp, vp, vp, sz]
    L.mktfhe_cmux_step_batch.argtypes = [vp, i32, i32, vp, vp, sz]
    L.mktfhe_block_step_batch.argtypes = [vp, i32, i32, vp, vp, sz]
    L.mktfhe_fft_batch.argtypes = [vp, i32, vp, vp, sz]
    L.mktfhe_ifft_batch.argtypes = [vp, i32, vp, vp, sz]
    L.mktfhe_decomp_batch.argtypes = [vp, i32, i32, i32, vp, vp, sz]
    L.mktfhe_last_stage_ms.argtypes = [vp, vp, vp]
    L.mktfhe_measure_dfma_peak.argtypes = [vp, vp]
    _lib = L
    return L
