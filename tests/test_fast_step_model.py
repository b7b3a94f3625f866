"""One blind-rotation step of the FAST path, assembled on the CPU from the numpy models of its parts, against the oracle's step
(bootstrapping.jl:413-438, `orc_cmux_step`) on the same accumulator row, key and rotation:

    field-extraction digits (csrc/kernels_fast_w.cuh)  ->  product-tree transform in the kernel's thread / register order
    (tools/models/fft32_model.py)  ->  spectrum x key into both sums  ->  x (X^a - 1) from the reference's monomial table
    ->  inverse in the kernel's order  ->  floor onto the torus (fast::d2torus)  ->  accumulator +=

The model uses numpy complex arithmetic instead of the kernel's fused multiply-adds, so it is not bit-equal to the GPU; what it pins
on the CPU is everything else a FAST step consists of -- which digit meets which key polynomial, the slot order, the monomial
convention, the sign of the folded halves -- and that this schedule of Float64 operations stays within the tolerance the GPU test
states (2^33 on Torus64, tests/test_gpu_fast.py) of the reference's schedule.  CPU only."""
import importlib.util
import math
import os

import numpy as np
import pytest

from conftest import keyset, make_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_LOG2 = 33


def _model(name="fft32_model"):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", "models", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _digits(x, l, logB, w=64):
    """FAST decomposition: bit fields of x + cadd (tests/test_transform_models.py checks them against gsw.jl:86-96)."""
    bit = w - l * logB
    cadd = ((1 << (bit - 1)) if bit > 0 else 0) + sum(1 << (bit + j * logB + logB - 1) for j in range(l))
    mask, half = (1 << logB) - 1, 1 << (logB - 1)
    v = [(int(c) + cadd) % (1 << w) for c in x]
    return [np.array([((c >> (bit + (l - 1 - j) * logB)) & mask) - half for c in v], dtype=np.float64) for j in range(l)]


@pytest.mark.parametrize("name,party,idx,at", [("KMS2party", 0, 0, 1), ("KMS2party", 1, 17, 2047), ("KMS2party", 1, 559, 2048),
                                               ("KMS2party", 0, 300, 4095), ("KMS2partyblock", 1, 5, 1234)])
def test_fast_step_model_within_tolerance_of_the_reference_step(name, party, idx, at):
    from oracle import oracle as O
    m = _model()
    ks = keyset(name)
    p = ks.params
    orc = make_oracle(ks)
    N, H, l, logB = p.N, p.N // 2, p.l_gsw, p.logB_gsw
    rng = np.random.default_rng(1000 * idx + at)
    acc = rng.integers(0, 2 ** 64, size=(2, N), dtype=np.uint64)             # one RLWE row: b, a
    want = orc.cmux_step(party, idx, at, acc)

    brk = ks.brk[party][idx].reshape(2, l, 2, H, 2)                           # [basket b | a][digit][component b | a][slot](re, im)
    key = brk[..., 0] + 1j * brk[..., 1]
    sums = np.zeros((2, H), dtype=complex)
    for basket in range(2):                                                   # digits of acc.b meet basket 0, of acc.a basket 1
        for j, d in enumerate(_digits(acc[basket], l, logB)):
            spec = m.fwd(d[:H] - 1j * d[H:])                                  # the kernel's folding: c_k = p_k - i p_{k+H}
            for comp in range(2):
                sums[comp] += spec * key[basket, j, comp]
    mono = O.monomials(N)[at - 1]
    mono = mono[:, 0] + 1j * mono[:, 1]                                       # FFT(X^at - 1), scheme.jl:121-146
    got = acc.copy()
    for comp in range(2):
        y = m.inv(sums[comp] * mono) / H
        add = [math.floor(v) for v in y.real] + [math.floor(-v) for v in y.imag]
        got[comp] = np.array([(int(a) + b) % 2 ** 64 for a, b in zip(acc[comp], add)], dtype=np.uint64)

    diff = (got.astype(np.int64) - want.astype(np.int64))                     # wraps mod 2^64 = centred difference
    worst = int(np.abs(diff).max())
    assert worst < 2 ** TOL_LOG2, math.log2(max(worst, 1))
    assert worst > 0 or at == 4096                                            # two Float64 schedules: equal words would mean the model IS the oracle


@pytest.mark.parametrize("idx,at", [(0, 1), (629, 1023), (77, 1024), (400, 2047)])
def test_fast_step_model_torus32(idx, at):
    """The same for the Torus32 single-key step (bootstrapping.jl:47-74) with the half-warp transform order of fastw32::k_cggi_w
    (tools/models/fft16x32_model.py): within 1 unit of Torus32 (the GPU test allows 4)."""
    from oracle import oracle as O
    m = _model("fft16x32_model")
    ks = keyset("CGGIparam")
    p = ks.params
    orc = make_oracle(ks)
    N, H, l, logB = p.N, p.N // 2, p.l_gsw, p.logB_gsw
    rng = np.random.default_rng(7000 + 10 * idx + at)
    acc = rng.integers(0, 2 ** 32, size=(2, N), dtype=np.uint32)
    want = orc.cmux_step(0, idx, at, acc)
    brk = ks.brk[0][idx].reshape(2, l, 2, H, 2)
    key = brk[..., 0] + 1j * brk[..., 1]
    sums = np.zeros((2, H), dtype=complex)
    for basket in range(2):
        for j, d in enumerate(_digits(acc[basket], l, logB, w=32)):
            spec = m.fwd(d[:H] - 1j * d[H:])
            for comp in range(2):
                sums[comp] += spec * key[basket, j, comp]
    mono = O.monomials(N)[at - 1]
    mono = mono[:, 0] + 1j * mono[:, 1]
    got = acc.copy()
    for comp in range(2):
        y = m.inv(sums[comp] * mono) / H
        add = [math.floor(v) for v in y.real] + [math.floor(-v) for v in y.imag]
        got[comp] = np.array([(int(a) + b) % 2 ** 32 for a, b in zip(acc[comp], add)], dtype=np.uint32)
    diff = (got.astype(np.int64) - want.astype(np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31
    assert int(np.abs(diff).max()) <= 1, int(np.abs(diff).max())


@pytest.mark.parametrize("party,blk,ats", [(0, 0, (1, 2, 3)), (1, 40, (4095, 0, 2048)), (1, 100, (0, 0, 0)), (0, 186, (777, 777, 1))])
def test_block_step_with_monomials_folded_into_the_keys(party, blk, ats):
    """KMS_block step (bootstrapping.jl:624-655; LMSS :124-163): the reference forms one external product per key bit of a block and
    adds mono_bit x product.  fast::k_phase1_tma<3> keeps ONE pair of sums per block by folding the monomials into the keys,
    Sum_bit mono_bit (Sum_dg D_dg K_bit,dg) = Sum_dg D_dg (Sum_bit mono_bit K_bit,dg), and skips bits with a~ = 0 like the reference.
    The folded form, assembled from the numpy models, must land within 2^33 of the oracle's block step (the GPU measures 2^31.9)."""
    from oracle import oracle as O
    m = _model()
    ks = keyset("KMS2partyblock")
    p = ks.params
    orc = make_oracle(ks)
    N, H, l, logB, ell = p.N, p.N // 2, p.l_gsw, p.logB_gsw, p.ell
    rng = np.random.default_rng(blk * 7 + sum(ats))
    acc = rng.integers(0, 2 ** 64, size=(2, N), dtype=np.uint64)
    want = orc.block_step(party, blk, np.array(ats, dtype=np.uint32), acc)
    table = O.monomials(N)
    folded = np.zeros((2, l, 2, H), dtype=complex)                            # [basket][digit][component][slot]
    for bit, at in enumerate(ats):
        if at == 0:
            continue                                                          # :637 `if tildea[idx] > 0`
        brk = ks.brk[party][blk * ell + bit].reshape(2, l, 2, H, 2)
        mono = table[at - 1][:, 0] + 1j * table[at - 1][:, 1]
        folded += (brk[..., 0] + 1j * brk[..., 1]) * mono
    sums = np.zeros((2, H), dtype=complex)
    for basket in range(2):
        for j, d in enumerate(_digits(acc[basket], l, logB)):
            spec = m.fwd(d[:H] - 1j * d[H:])
            for comp in range(2):
                sums[comp] += spec * folded[basket, j, comp]
    got = acc.copy()
    for comp in range(2):
        y = m.inv(sums[comp]) / H
        add = [math.floor(v) for v in y.real] + [math.floor(-v) for v in y.imag]
        got[comp] = np.array([(int(a) + b) % 2 ** 64 for a, b in zip(acc[comp], add)], dtype=np.uint64)
    worst = int(np.abs(got.astype(np.int64) - want.astype(np.int64)).max())
    assert worst < 2 ** TOL_LOG2, math.log2(max(worst, 1))
    if not any(ats):
        assert worst <= 1                                                     # nothing to add: floor(0) against native(0 +- rounding)
