"""The key ring of the one-warp phase-1 kernel (csrc/kernels_fast_w.cuh) as an explicit-state model: every interleaving of the
producer, the consumer warps and the out-of-order completions of the bulk copies (tools/models/key_ring_model.py).  The model
reproduces the race of the build before the fix (one ring of odd depth: a warp that runs ahead takes an older phase of a barrier
for its own) and shows the shipped layouts free of it.  CPU only; the GPU regression test is
tests/test_gpu_big_sets.py::test_single_gate_calls_at_kms32_do_not_deadlock."""
import importlib.util
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model():
    spec = importlib.util.spec_from_file_location("key_ring_model", os.path.join(ROOT, "tools", "models", "key_ring_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _define(path, name):
    """Default value of `#define NAME <int>` in a kernel source."""
    src = open(os.path.join(ROOT, "mktfhe_b200", "csrc", path)).read()
    m = re.search(r"#define\s+" + name + r"\s+\(?\s*(?:W_INV_DIT\s*\?\s*\d+\s*:\s*)?(\d+)", src)
    assert m, f"{name} not found in {path}"
    return int(m.group(1))


def test_odd_shared_ring_races_when_a_warp_runs_ahead():
    m = _model()
    res, _ = m.check(("one", 5), units=2, ntiles=24, coupled=False)
    assert "took tile" in res or "overwritten" in res or res == "deadlock", res
    # with the token order of live units nothing can run ahead: this is why only small batches and skipped steps hit it
    assert m.check(("one", 5), units=2, ntiles=24, coupled=True)[0] == "ok"


@pytest.mark.parametrize("coupled", [True, False])
def test_shipped_rings_are_safe(coupled):
    m = _model()
    ra, rb = _define("kernels_fast_w.cuh", "W_RING_A"), _define("kernels_fast_w.cuh", "W_RING_B")
    assert (ra, rb) == (3, 2)
    assert m.check(("two", ra, rb), units=2, ntiles=26, coupled=coupled)[0] == "ok"
    if coupled:                                            # three units: 175k states coupled, 840k free-running (18 s): coupled only
        assert m.check(("two", ra, rb), units=3, ntiles=14, coupled=True)[0] == "ok"
    # the half-warp CGGI kernel shares ONE ring between the kinds; its depth is even (static_assert in the source)
    d32 = _define("kernels_fast32_w.cuh", "W32_RING")
    assert d32 % 2 == 0
    assert m.check(("one", d32), units=2, ntiles=28, coupled=coupled)[0] == "ok"


def test_even_shared_ring_and_other_splits_are_safe():
    m = _model()
    for layout in (("one", 4), ("two", 2, 2), ("two", 3, 1), ("two", 1, 1)):
        for coupled in (True, False):
            assert m.check(layout, units=2, ntiles=20, coupled=coupled)[0] == "ok", (layout, coupled)
    # every odd shared depth has the defect, not just five
    for d in (3, 7):
        assert m.check(("one", d), units=2, ntiles=4 * d, coupled=False)[0] != "ok", d


def _token_model():
    spec = importlib.util.spec_from_file_location("token_model", os.path.join(ROOT, "tools", "models", "token_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("l", [1, 2, 3, 5, 6])
def test_token_order_of_the_two_warps(l):
    """Four tokens per unit: exclusive sweeps, fixed order of the additions, complete sums at the step end, no lapped token, no
    deadlock -- for every gadget length of params.jl's KMS sets (l_gsw = 3, 5, 6) and with skipped steps (a~ = 0)."""
    m = _token_model()
    assert m.check(l, (False, False, False)) == "ok"
    assert m.check(l, (False, True, False, True, False)) == "ok"


def test_token_model_finds_every_missing_wait():
    m = _token_model()
    for w in range(2):
        nwait = sum(1 for op in m.program(w, 3, (False, False)) if op[0] == m.WAIT)
        assert nwait == 12
        for k in range(nwait):
            assert m.check(3, (False, False), drop_wait=(w, k)) != "ok", (w, k)


@pytest.mark.parametrize("depth,ntiles", [(5, 13), (10, 23), (8, 19)])
def test_rings_read_by_every_consumer_are_safe_at_any_depth(depth, ntiles):
    """fast::k_phase1_tma (5 whole / 10 half tiles), fast32::k_rgsw_tma (8 whole tiles): every consumer thread waits on every tile,
    so every waiter sees every phase of its barriers and an odd depth is harmless.  More than two laps each."""
    m = _model()
    assert m.check(("all", depth), units=2, ntiles=ntiles, coupled=False)[0] == "ok"
