#!/usr/bin/env python
"""bench.py -- MK-NAND gate bootstraps/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--batch B]

One "step" = one pass of the hot path (gate linear part -> modswitch -> blind rotation -> key switch) over one
batch of synthetic MK-NAND inputs.  N = 1 runs BASELINE.json configs[1]: KMS 2-party MK-NAND, batch 4096.
N > 1 (torchrun, one rank per GPU): every rank holds a replicated key set and its own batch of the same size
(weak scaling); there is no collective inside the timed region, only the barriers around it.

Output: ONE JSON line on rank 0 (see README / DESIGN.md for the keys).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MK-NAND gate bootstraps/sec"
UNIT = "gates/s"
KEY_SEED = 0x4D4B5446
GATE_SEED = 0x47415445

WORKLOADS = {  # name -> (parameter set, default per-GPU batch)
    "kms2": ("KMS2party", 4096),
    "kms8block": ("KMS8partyblock", 2048),
    "kms8": ("KMS8party", 1024),
    "kms32": ("KMS32party", 1024),
    "kms32block": ("KMS32partyblock", 1024),
    "cggi": ("CGGIparam", 4096),
    "lmss": ("Blockparam", 4096),
    "ccs2": ("CCS2party", 1024),
    "ccs16": ("CCS16party", 592),
}


def algorithmic_gflop_per_gate(p) -> dict:
    """SURVEY.md 8(d) conventions: FFT/iFFT of H points = 5*H*log2(H), twist = 6H, complex MAC = 8H, mul = 6H."""
    import math
    H = p.H
    fft = 5 * H * math.log2(H) + 6 * H
    out = {}
    if p.scheme in (3, 4):
        R = 1 + (p.k - 1) * p.l_lev
        if p.scheme == 3:
            step_row = (2 * p.l_gsw + 2) * fft + 4 * p.l_gsw * 8 * H + 2 * 6 * H
            ph1 = p.n * R * step_row
        else:
            ph1 = p.d * R * ((2 * p.l_gsw + 2) * fft + p.ell * (4 * p.l_gsw + 2) * 8 * H)
        ph1 += 2 * R * fft
        ph2 = 0.0
        for idx in range(1, p.k + 1):
            it = 1 if idx == 1 else p.l_lev
            nf = (p.l_lev + p.l_uni) * idx + p.l_uni
            ni = idx + 1 + (p.k + 1)
            mac = 2 * it * idx + 2 * p.l_uni * idx + 2 * p.l_uni
            ph2 += (nf + ni) * fft + mac * 8 * H
        out = {"phase1": ph1 / 1e9, "phase2": ph2 / 1e9}
    elif p.scheme in (0, 1):
        if p.scheme == 0:
            ph1 = p.n * ((2 * p.l_gsw + 2) * fft + 4 * p.l_gsw * 8 * H + 2 * 6 * H)
        else:
            ph1 = p.d * ((2 * p.l_gsw + 2) * fft + p.ell * (4 * p.l_gsw + 2) * 8 * H)
        out = {"phase1": ph1 / 1e9, "phase2": 0.0}
    else:
        tot = 0.0
        for idx in range(1, p.k + 1):
            nf = 2 * p.l_uni * (idx + 1)
            ni = (idx + 1) + (p.k + 1)
            mac = 4 * p.l_uni * (idx + 1)
            tot += p.n * ((nf + ni) * fft + mac * 8 * H + (p.k + 1) * 6 * H)
        out = {"phase1": tot / 1e9, "phase2": 0.0}
    out["total"] = out["phase1"] + out["phase2"]
    # key switch: k * 3/4 * N_ks * f * (n+1) * 4 bytes of ksk rows gathered per gate
    nks = p.N - p.n if p.is_block else p.N
    frac = 0.75 if not p.is_block else 0.75
    out["ks_bytes"] = p.k * frac * nks * p.f * (p.n + 1) * 4
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.lines = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def make_inputs(ks, batch: int, rank: int):
    """Two full-support MK-LWE encryptions of uniform random bits per gate."""
    rng = np.random.default_rng(GATE_SEED + rank)
    m1 = rng.integers(0, 2, batch)
    m2 = rng.integers(0, 2, batch)
    base = GATE_SEED + rank * 10_000_000
    c1 = ks.encrypt_batch(m1, base)
    c2 = ks.encrypt_batch(m2, base + 5_000_000)
    return m1.astype(bool), m2.astype(bool), c1, c2


def cpu_baseline(ks, c1, c2, budget_s: float = 12.0, max_threads: int = 0) -> dict:
    """The CPU oracle (a C port of the reference's algorithm: `julia` is not installed) on this host's cores,
    on a bounded sample of the same workload."""
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    from mktfhe_b200 import params as P
    p = ks.params
    orc = O.Oracle(p, ks.brk, ks.ksk, ks.rlk if p.scheme in (P.KMS, P.KMS_BLOCK) else None,
                   ks.pubb if p.is_mk else None, ks.crs_fft)
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the num_threads clause overrides it)
    threads = len(os.sched_getaffinity(0)) if max_threads <= 0 else max_threads
    n0 = min(threads, c1.shape[0])
    t = time.perf_counter()
    orc.gate_batch(0, c1[:n0], c2[:n0], threads)
    t0 = time.perf_counter() - t
    reps = int(max(1, min(budget_s / max(t0, 1e-3), c1.shape[0] // max(n0, 1))))
    n = n0 * reps
    t = time.perf_counter()
    out = orc.gate_batch(0, c1[:n], c2[:n], threads)
    dt = time.perf_counter() - t
    return {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} MK-NAND gates of the same inputs, {dt:.1f} s, C oracle -O2 OpenMP over gates"}, out, n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="kms2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="gates per GPU per step (default: the workload's)")
    ap.add_argument("--mode", default="fast", choices=["fast", "strict"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--keygen", default=None, choices=["host", "device"], help="where the evaluation keys are generated (default: per workload)")
    ap.add_argument("--no-also", action="store_true", help="skip the sub-lines for the other BASELINE configurations")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner) go to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())

    from mktfhe_b200 import params as P
    from mktfhe_b200.keys import KeySet
    pname, dbatch = WORKLOADS[args.workload]
    p = P.ALL[pname]
    batch = args.batch or dbatch
    cfg = {"workload": f"{pname} MK-NAND, {batch} gates per GPU per step, full-support inputs",
           "params": pname, "batch_per_gpu": batch, "parties": p.k,
           "l2_policy": "working set > L2: keys + per-step accumulators exceed 126 MB, no flush needed",
           "key_seed": KEY_SEED, "gate_seed": GATE_SEED}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        ks = KeySet(p, seed=KEY_SEED, nthreads=len(os.sched_getaffinity(0)))
        nsample = min(batch, 2048)
        _, _, c1, c2 = make_inputs(ks, nsample, 0)
        vals = []
        info = None
        for it in range(args.warmup + args.steps):
            info, _, n = cpu_baseline(ks, c1, c2, budget_s=float(os.environ.get("MKTFHE_REF_BUDGET_S", "8.0")))
            if it >= args.warmup:
                vals.append(info["value"])
        v = float(np.mean(vals)) if vals else info["value"]
        info["value"] = v
        line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * batch / v, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "cpu_baseline": info,
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "reference arm = C port of the reference algorithm (oracle/); Julia is not installed on this image"}
        emit(line)
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = {"rank": rank, "world": world, "local_rank": local_rank, "torch": torch, "dist": dist, "mode": args.mode, "keygen": args.keygen}
    line, rc = run_workload(ctx, args.workload, batch, args.steps, args.warmup, headline=True,
                            want_cpu_baseline=(world == 1 and not args.no_cpu_baseline))
    # the other configurations BASELINE.json names, each as a full sub-line (fewer steps: they are reported, the headline is timed
    # to the contract).  Under torchrun every rank runs them too, so the scaling record carries KMS32 and KMS8block at N GPUs.
    if args.workload == "kms2" and not args.no_also and not args.batch:
        also = []
        for wl in ALSO:
            sub, _ = run_workload(ctx, wl, WORKLOADS[wl][1], ALSO_STEPS, ALSO_WARMUP, headline=False, want_cpu_baseline=False)
            if rank == 0:
                also.append(sub)
        if rank == 0:
            line["also"] = also
            # a sub-line that fails its decrypt check carries "valid": false and a null value itself; the exit code follows the
            # headline alone, so one marginal configuration cannot void the headline measurement
            line["also_invalid"] = [sub["config"]["params"] for sub in also if not sub.get("valid", False)]
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return rc


# BASELINE.json configs[0] (CGGI), configs[2] (KMS 8-party block, 16384 gates over 8 GPUs = 2048 per GPU), configs[3] (CCS 16-party),
# configs[4] (KMS 32-party); "kms8" is the k = 8 point of BASELINE.json's metric ("k=2/8/32") without block keys
ALSO = ("cggi", "kms8block", "kms8", "ccs16", "kms32")
if os.environ.get("MKTFHE_BENCH_ALSO"):          # a shorter list for a quick run, e.g. MKTFHE_BENCH_ALSO=cggi
    ALSO = tuple(w for w in os.environ["MKTFHE_BENCH_ALSO"].split(",") if w in WORKLOADS)
# Where the evaluation keys come from.  "host": libmktfhe_host.so on rank 0, pinned upload, NCCL broadcast to the other ranks (the
# reference's flow: keys are built on the host).  "device": generated on every GPU from the seed (byte-identical, no PCIe / NVLink).
# The headline keeps the host path; the 10.6 GB key set of KMS32party takes 15 s + 3 s that way and 2 s on the device.
KEYGEN = {"kms2": "host", "kms8block": "host"}
ALSO_STEPS, ALSO_WARMUP = 3, 3
# Fraction of the sampled outputs that must decrypt correctly for the line to count (exit code 3 and "valid": false otherwise).
# CCS16party / KMS32party sit at the decision margin in the reference algorithm itself: the CPU oracle fails 1.6 % of KMS32party
# gates and ~5 % of CCS16party gates on the same inputs (tests/golden/failrate_*.npz).
# KMS8party: output noise 2^26.87 against the 2^29 margin = 4.4 sigma, i.e. about 1e-5 failures per gate in the reference algorithm
# itself; with 256 checked gates on each of 8 ranks one miss must not void the sub-line (a broken kernel fails half of them).
MIN_OK = {"kms32": 0.93, "kms32block": 0.93, "ccs16": 0.85, "kms8": 0.99}
SMEM_BYTES_PER_CLK_PER_SM = 128


def traffic_record():
    """dram bytes per launch from the committed ncu capture, valid only for the kernel sources it was taken with."""
    import hashlib
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        rec = json.load(open(path))
    except Exception:
        return {}, "no capture (profiles/traffic.json)"
    h = hashlib.sha256()
    for f in rec.get("sources", []):
        try:
            h.update(open(os.path.join(ROOT, f), "rb").read())
        except OSError:
            return {}, "capture lists a missing source file"
    if h.hexdigest() != rec.get("sources_sha256"):
        return {}, "capture is stale: kernel sources changed since profiles/traffic.json was taken"
    return rec.get("workloads", {}), f"ncu dram__bytes_read.sum + dram__bytes_write.sum per launch ({rec.get('capture', 'profiles/')})"


def run_workload(ctx, workload, batch, steps, warmup, headline, want_cpu_baseline):
    torch, dist = ctx["torch"], ctx["dist"]
    rank, world, local_rank = ctx["rank"], ctx["world"], ctx["local_rank"]
    from mktfhe_b200 import params as P
    from mktfhe_b200.scheme import MODE_FAST, MODE_STRICT
    from mktfhe_b200 import dist as mkdist
    pname, _ = WORKLOADS[workload]
    p = P.ALL[pname]
    cfg = {"workload": f"{pname} MK-NAND, {batch} gates per GPU per step, full-support inputs",
           "params": pname, "batch_per_gpu": batch, "parties": p.k,
           "l2_policy": "working set > L2: keys + per-step accumulators exceed 126 MB, no flush needed",
           "key_seed": KEY_SEED, "gate_seed": GATE_SEED}

    t_k = time.perf_counter()
    scheme, ks, key_times = mkdist.setup_replicated(p, KEY_SEED, local_rank, rank, world, timings=True,
                                                     keygen=ctx.get("keygen") or KEYGEN.get(workload, "device"))
    keygen_s = time.perf_counter() - t_k
    want_mode = MODE_FAST if ctx["mode"] == "fast" else MODE_STRICT
    try:
        scheme.set_mode(want_mode)
    except Exception:
        pass
    mode_name = "fast" if scheme.mode == MODE_FAST else "strict"

    m1, m2, c1, c2 = make_inputs(ks, batch, rank)
    lw = p.lwe_words
    # pinned host buffers for the end-to-end leg, device-resident copies for the kernel leg
    h1 = torch.from_numpy(c1.view(np.int32)).pin_memory()
    h2 = torch.from_numpy(c2.view(np.int32)).pin_memory()
    hout = torch.empty_like(h1).pin_memory()
    d1, d2 = h1.cuda(), h2.cuda()
    dout = torch.empty_like(d1)
    stream = torch.cuda.ExternalStream(scheme.stream)

    def step_dev():
        scheme.gate_dev(0, d1.data_ptr(), d2.data_ptr(), dout.data_ptr(), batch)

    def step_e2e():
        from mktfhe_b200 import _lib
        rc = _lib.lib().mktfhe_gate_batch(scheme._h, 0, h1.data_ptr(), h2.data_ptr(), hout.data_ptr(), batch)
        if rc != 0:
            raise RuntimeError("mktfhe_gate_batch failed")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        scheme.sync()

    def timed(fn, nsteps):
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        barrier()
        with torch.cuda.stream(stream):
            ev0.record(stream)
            for _ in range(nsteps):
                fn()
            ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(warmup):
        step_dev()
    scheme.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    ms_dev = timed(step_dev, steps)
    stage_ms, launches = scheme.last_stage_ms()
    clocks = sampler.stop()

    # correctness of what was timed: decrypt a sample of the outputs
    res = dout.cpu().numpy().view(np.uint32)
    ncheck = min(batch, 256)
    want = ~(m1[:ncheck] & m2[:ncheck])
    ok = int(np.sum(ks.decrypt_batch(res[:ncheck]) == want))
    ph = ks.phase_batch(res[:ncheck]).astype(np.int64)
    perr = (ph - np.where(want, 1 << 29, 7 << 29)) & 0xFFFFFFFF      # output phase error on Torus32 (decision margin 2^29)
    perr = np.where(perr >= (1 << 31), perr - (1 << 32), perr).astype(np.float64)
    perr_std_log2 = float(np.log2(np.std(perr) + 1.0))

    for _ in range(min(warmup, 1)):
        step_e2e()
    ms_e2e = timed(step_e2e, steps)
    ok_e2e = int(np.sum(ks.decrypt_batch(hout.numpy().view(np.uint32)[:ncheck]) == want))

    total_gates = batch * world * steps
    value = total_gates / (ms_dev * 1e-3)
    e2e_value = total_gates / (ms_e2e * 1e-3)
    need = MIN_OK.get(workload, 1.0) * ncheck
    valid = ok >= need and ok_e2e >= need
    if world > 1:                                  # a wrong result on any rank invalidates the line
        t = torch.tensor([1.0 if valid else 0.0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        valid = bool(t.item() > 0.5)

    line = {"metric": METRIC, "value": value if valid else None, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_dev / steps, "ms_per_gate": ms_dev / steps / batch, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "mode": mode_name, "clocks": clocks, "valid": valid,
            "e2e": {"value": e2e_value if valid else None, "unit": UNIT, "h2d_bytes_per_step": 2 * batch * lw * 4, "d2h_bytes_per_step": batch * lw * 4,
                    "ms_per_step": ms_e2e / steps},
            "gpu_launches": launches * steps,
            "decrypt_check": {"checked": ncheck, "ok_device_leg": ok, "ok_e2e_leg": ok_e2e, "required": int(np.ceil(need)),
                              "phase_error_std_log2": perr_std_log2,
                              "note": "decision margin 2^29; CCS16 / KMS32 sets sit close to it in the reference algorithm itself (DESIGN.md)"},
            "stage_ms_last_step": stage_ms, "keygen_and_upload_s": keygen_s, "key_setup": key_times}
    if not valid:
        line["invalid"] = f"decrypt check failed: {ok}/{ncheck} (device leg), {ok_e2e}/{ncheck} (end to end), {int(np.ceil(need))} required"

    if rank == 0:
        # roofline of the dominant kernel (phase 1: FP64 FFT + pointwise MAC), measured live
        alg = algorithmic_gflop_per_gate(p)
        traffic, traffic_note = traffic_record()
        dom_ms = stage_ms["phase1"]
        peak = scheme.dfma_peak_tflops()
        ach = alg["phase1"] * batch / (dom_ms * 1e-3) / 1e3 if dom_ms > 0 else 0.0
        line["roofline"] = {"bound": "fp64", "kernel": "blind rotation (FP64 transforms + pointwise MAC): fastw::k_phase1_w / fast::k_phase1_tma<3> / fastw32::k_cggi_w / fast32::k_rgsw_tma<3> / k_ccs_fast",
                            "achieved": ach, "peak": peak,
                            "unit": "TFLOP/s", "frac": ach / peak if peak else None, "traffic": traffic.get(workload, {}).get("phase1"),
                            "traffic_source": traffic_note,
                            "peak_source": "measured in this run: register-only DFMA loop, 16 chains x 16 warps/SM (MEASURED_PEAKS.json has no FP64 entry); nominal 37.2",
                            "algorithmic_gflop_per_gate": alg, "kernel_ms_per_launch": dom_ms}
        try:
            hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            src = "measured (MEASURED_PEAKS.json)"
        except Exception:
            hbm, src = 6650.0, "fallback (B200_PROFILING.md)"
        # Key switch: every gate adds k * 3/4 * N_ks * f rows of (n + 1) words.  The tiled kernel brings a candidate row into shared
        # memory once per 32 gates (TMA) and each gate reads the row its digit selects from there, so the algorithmic bytes move
        # through SHARED memory (128 B/clk/SM), not DRAM: that is the pipe the kernel is bound by (ncu: shared pipe 60 %).
        ks_ms = stage_ms["keyswitch"]
        sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        smem_peak = SMEM_BYTES_PER_CLK_PER_SM * sm_count * sm_mhz * 1e6 / 1e9
        ks_ach = alg["ks_bytes"] * batch / (ks_ms * 1e-3) / 1e9 if ks_ms > 0 else 0.0
        ks_dram = traffic.get(workload, {}).get("keyswitch")
        line["roofline_keyswitch"] = {"bound": "shared-memory", "achieved": ks_ach, "peak": smem_peak, "unit": "GB/s", "frac": ks_ach / smem_peak,
                                      "peak_source": f"128 B/clk/SM x {sm_count} SMs x {sm_mhz:.0f} MHz (SM clock sampled during the run)",
                                      "traffic": ks_dram, "traffic_source": traffic_note,
                                      "hbm": {"achieved": (ks_dram / (ks_ms * 1e-3) / 1e9) if (ks_dram and ks_ms > 0) else None, "peak": hbm, "peak_source": src,
                                              "note": "DRAM side: measured bytes / kernel time; rows are fetched once per 32-gate tile"},
                                      "kernel_ms_per_launch": ks_ms}
        if want_cpu_baseline:
            info, _, _ = cpu_baseline(ks, c1, c2)
            line["cpu_baseline"] = info
    scheme.close()
    del d1, d2, dout, h1, h2, hout
    torch.cuda.empty_cache()
    return line, (0 if valid else 3)


if __name__ == "__main__":
    sys.exit(main())
