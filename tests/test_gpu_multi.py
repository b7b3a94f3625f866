"""Multi-GPU behind the C-ABI (mktfhe_ctx_create_multi): one process, one front context, keys uploaded once and
replicated device to device, a host batch sharded in contiguous slices.  Results must be bit-identical to the single-device
call (gates are independent); the only parallelism the reference itself has is Threads.@threads over parties
(/root/reference/src/tfhe/bootstrapping.jl:376-378,573)."""
import numpy as np
import pytest

from conftest import fresh_inputs, keyset
from mktfhe_b200.scheme import MODE_FAST, MODE_STRICT, MktfheError, setup

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("name", ["KMS2party", "CGGIparam"])
def test_front_context_matches_single_device(gpu_schemes, name):
    ks = keyset(name)
    single = gpu_schemes(name)
    ndev = min(_ndev(), 8)
    front = setup(ks, devices=list(range(ndev)))
    try:
        assert front.devices == list(range(ndev))
        B = 37                                               # ragged: slices of different sizes, some possibly empty on 8 GPUs
        b1, c1 = fresh_inputs(ks, B, seed=111)
        b2, c2 = fresh_inputs(ks, B, seed=112)
        for mode in (MODE_STRICT, MODE_FAST):
            single.set_mode(mode); front.set_mode(mode)
            assert front.mode == mode
            for op in (0, 3):
                assert np.array_equal(front.gate(op, c1, c2), single.gate(op, c1, c2)), (name, mode, op, ndev)
            assert np.array_equal(front.bootstrapping(c1[:3]), single.bootstrapping(c1[:3]))
            assert np.array_equal(front.gate(0, c1[:1], c2[:1]), single.gate(0, c1[:1], c2[:1]))      # batch 1: one device works
        ms, launches = front.last_stage_ms()
        assert launches >= 3 and ms["phase1"] > 0
        with pytest.raises(MktfheError, match="single-device"):
            front.blindrotate(c1[:1])
    finally:
        single.set_mode(MODE_FAST)
        front.close()


def test_front_context_argument_errors():
    from mktfhe_b200 import params as P
    from mktfhe_b200.scheme import Scheme
    with pytest.raises(MktfheError):
        Scheme(P.CGGIparam, devices=[0, 0])
    with pytest.raises(MktfheError):
        Scheme(P.CGGIparam, devices=[_ndev()])
    s = Scheme(P.CGGIparam, devices=[0])
    with pytest.raises(MktfheError, match="not finalized"):
        s.gate(0, np.zeros((1, P.CGGIparam.lwe_words), np.uint32), np.zeros((1, P.CGGIparam.lwe_words), np.uint32))
    s.close()
