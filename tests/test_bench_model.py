"""The roofline numerator bench.py reports (`algorithmic_gflop_per_gate`) must be SURVEY.md section 8(d)'s figure for every
BASELINE configuration, and the bench line must carry every key of the measurement contract (checked on the source: the
bench itself needs a GPU)."""
import importlib.util
import os
import re

import pytest

from mktfhe_b200 import params as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bench_module():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# SURVEY.md 8(d), "Totals": GFLOP per gate; key-switch algorithmic MB per gate
SURVEY_TOTALS = {"KMS2party": 0.959, "KMS4party": 3.40, "KMS8party": 8.87, "KMS8partyblock": 4.58, "KMS16party": 22.4,
                 "KMS32party": 54.1, "KMS32partyblock": 28.4, "CGGIparam": 0.166, "CCS16party": 76.8, "CCS2party": 0.75,
                 "CCS8party": 10.3}
SURVEY_KS_MB = {"KMS2party": 55, "KMS8partyblock": 169, "KMS32party": 882, "CGGIparam": 15.5}


@pytest.mark.parametrize("name", sorted(SURVEY_TOTALS))
def test_algorithmic_flops_match_the_survey(name):
    alg = bench_module().algorithmic_gflop_per_gate(P.ALL[name])
    assert alg["total"] == pytest.approx(SURVEY_TOTALS[name], rel=0.01), (name, alg)
    assert alg["phase1"] + alg["phase2"] == pytest.approx(alg["total"])


@pytest.mark.parametrize("name", sorted(SURVEY_KS_MB))
def test_keyswitch_bytes_match_the_survey(name):
    alg = bench_module().algorithmic_gflop_per_gate(P.ALL[name])
    assert alg["ks_bytes"] / 1e6 == pytest.approx(SURVEY_KS_MB[name], rel=0.02), (name, alg["ks_bytes"])


def test_bench_line_carries_the_contract_keys():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "h2d_bytes_per_step", "d2h_bytes_per_step", "gpu_launches", "roofline",
                "bound", "achieved", "peak", "frac", "traffic", "cpu_baseline", "cores", "kind", "sample", "impl"):
        assert re.search(r'["\']%s["\']' % key, src), key
    mod = bench_module()
    assert mod.WORKLOADS["kms2"] == ("KMS2party", 4096)                 # BASELINE configs[1]
    assert mod.WORKLOADS["kms32"][1] == 1024 and mod.WORKLOADS["kms8block"][1] == 2048
