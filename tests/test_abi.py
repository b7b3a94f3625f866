"""The drop-in boundary: the C-ABI library loads without a GPU, exports every symbol its header declares, fails
loudly instead of falling back to the CPU, and the product never reaches into oracle/."""
import ctypes
import os
import re
import subprocess

import pytest

from mktfhe_b200 import _lib, build, params as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mktfhe_[a-z0-9_]+)\s*\(", src)) - {"mktfhe_torus_bits", "mktfhe_ksk_rows",
                  "mktfhe_brk_doubles", "mktfhe_rlk_doubles", "mktfhe_pubb_doubles", "mktfhe_crs_doubles",
                  "mktfhe_ksk_words", "mktfhe_lwe_words"})


def test_cuda_library_exports_every_declared_symbol():
    build.build_cuda()
    lib = ctypes.CDLL(build.CUDA_LIB)
    names = declared("mktfhe_b200.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == names


def test_host_library_exports_every_declared_symbol():
    build.build_host()
    lib = ctypes.CDLL(build.HOST_LIB)
    for n in declared("mktfhe_host.h"):
        assert hasattr(lib, n), n


def test_no_gpu_means_loud_failure_not_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib.lib()
    cp = P.KMS2party.c_struct()
    h = ctypes.c_void_p()
    rc = L.mktfhe_ctx_create(ctypes.byref(cp), 0, ctypes.byref(h))
    assert rc == -2 and not h.value
    assert b"no CUDA device" in L.mktfhe_last_error(None)
    from mktfhe_b200.scheme import MktfheError, Scheme
    with pytest.raises(MktfheError):
        Scheme(P.KMS2party)


def test_bad_parameters_are_rejected_before_touching_the_device():
    from dataclasses import replace
    L = _lib.lib()
    h = ctypes.c_void_p()
    for bad in (replace(P.KMS2party, N=1024), replace(P.CGGIparam, k=2), replace(P.Blockparam, ell=2, n=458),
                replace(P.KMS2party, l_gsw=17), replace(P.CGGIparam, n=900)):
        cp = bad.c_struct()
        assert L.mktfhe_ctx_create(ctypes.byref(cp), 0, ctypes.byref(h)) == -1
    assert L.mktfhe_ctx_create(None, 0, ctypes.byref(h)) == -4
    # null context on every entry point that takes one
    assert L.mktfhe_finalize_keys(None) == -4 and L.mktfhe_sync(None) == -4
    assert L.mktfhe_gate_batch(None, 0, None, None, None, 1) == -4


def test_product_never_uses_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may touch oracle/."""
    hits = subprocess.run(["grep", "-rlE", r"(import +oracle|from +oracle|#include.*oracle|mktfhe_oracle|orc_[a-z]+\(|_build/lib)", os.path.join(ROOT, "mktfhe_b200"), os.path.join(ROOT, "include"),
                           "--include=*.py", "--include=*.h", "--include=*.cu", "--include=*.cuh", "--include=*.cpp"],
                          capture_output=True, text=True).stdout.split()
    assert hits == [], hits
    bench = open(os.path.join(ROOT, "bench.py")).read()
    assert bench.count("from oracle import") == 1 and "def cpu_baseline" in bench
