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


def test_reference_arm_runs_on_cpu_and_prints_one_json_line():
    """`bench.py --impl reference` needs no GPU: it times the CPU port and prints exactly one JSON line on stdout."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, MKTFHE_REF_BUDGET_S="0.5")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--batch", "16"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "MK-NAND gate bootstraps/sec" and d["unit"] == "gates/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["params"] == "KMS2party"
