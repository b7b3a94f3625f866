"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

  libmktfhe_host.so   host key generation / encrypt / decrypt   (g++, no CUDA)
  libmktfhe_b200.so   CUDA kernels + the C-ABI of include/mktfhe_b200.h   (nvcc, sm_100a)

`python -m mktfhe_b200.build` builds both; `--force` rebuilds.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIBDIR = os.path.join(PKG, "lib")

HOST_LIB = os.path.join(LIBDIR, "libmktfhe_host.so")
# MKTFHE_CUDA_LIB: load / build another copy of the CUDA library (A/B and knock-out timing variants built with MKTFHE_NVCC_EXTRA)
CUDA_LIB = os.environ.get("MKTFHE_CUDA_LIB") or os.path.join(LIBDIR, "libmktfhe_b200.so")

HOST_SRCS = ["host_keygen.cpp"]
CUDA_SRCS = ["capi.cu"]
CUDA_DEPS = ["common.cuh", "fft_strict.cuh", "kernels_strict.cuh", "kernels_fast.cuh", "kernels_fast_w.cuh", "kernels_fast32.cuh", "kernels_fast32_w.cuh", "keyswitch.cuh", "keygen.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",            # contraction is written explicitly (fma()) where the fast path wants it
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _run(cmd: list[str], log: str | None = None) -> None:
    res = subprocess.run(cmd, capture_output=True, text=True)
    if log:
        with open(log, "w") as fh:
            fh.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))


def build_host(force: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in HOST_SRCS]
    deps = srcs + [os.path.join(INCLUDE, h) for h in ("mktfhe_host.h", "mktfhe_params.h")]
    if force or _stale(HOST_LIB, deps):
        tmp = HOST_LIB + ".tmp"
        # -ffp-contract=off: the uploaded FFT form of the keys is the reference's Float64 transform (fft.jl:57-63,105-155), and
        # Julia never contracts a*b + c into an fma; with contraction the spectra differ in the last bit from the oracle's and
        # from the device key generation (csrc/keygen.cuh), which both follow the reference operation for operation.
        _run(["g++", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-std=gnu++17", "-fopenmp", "-fPIC", "-shared", "-Wall",
              "-o", tmp] + srcs + ["-lquadmath"])
        os.replace(tmp, HOST_LIB)            # atomic: a repo snapshot never sees a half-written library
    return HOST_LIB


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def build_cuda(force: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(CSRC, s) for s in CUDA_SRCS]
    deps = srcs + [os.path.join(CSRC, d) for d in CUDA_DEPS] + \
        [os.path.join(INCLUDE, h) for h in ("mktfhe_b200.h", "mktfhe_params.h")]
    if force or _stale(CUDA_LIB, deps):
        tmp = CUDA_LIB + ".tmp"
        # MKTFHE_DEBUG_SPIN=1: every mbarrier wait is bounded and traps instead of hanging (csrc/kernels_fast.cuh)
        dbg = ["-DMKTFHE_DEBUG_SPIN"] if os.environ.get("MKTFHE_DEBUG_SPIN") == "1" else []
        dbg += os.environ.get("MKTFHE_NVCC_EXTRA", "").split()
        _run([nvcc_path()] + NVCC_FLAGS + dbg + ["-I", INCLUDE, "-o", tmp] + srcs + ["-lquadmath"],
             log=CUDA_LIB + ".nvcc.log" if os.environ.get("MKTFHE_CUDA_LIB") else os.path.join(LIBDIR, "nvcc_build.log"))
        os.replace(tmp, CUDA_LIB)
    return CUDA_LIB


def build_all(force: bool = False) -> None:
    build_host(force)
    build_cuda(force)


if __name__ == "__main__":
    build_all("--force" in sys.argv)
    print("built:", HOST_LIB, CUDA_LIB)
