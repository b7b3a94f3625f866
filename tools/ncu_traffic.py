#!/usr/bin/env python
"""profiles/traffic.json from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... --csv` launch lists.

    python tools/ncu_traffic.py kms2=profiles/r2_traffic_kms2.csv [kms32=...] > profiles/traffic.json

Per workload: DRAM bytes (read + write) of one launch of the blind-rotation kernel ("phase1"), of FAST phase 2 and of the key
switch, taken from the LAST launch of each kernel in the capture.  The record carries the SHA-256 of the kernel sources it was
captured with; bench.py reports `roofline.traffic` only while that hash still matches (a stale capture reads as null)."""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["mktfhe_b200/csrc/kernels_fast.cuh", "mktfhe_b200/csrc/kernels_fast_w.cuh", "mktfhe_b200/csrc/kernels_fast32.cuh",
           "mktfhe_b200/csrc/kernels_fast32_w.cuh",
           "mktfhe_b200/csrc/keyswitch.cuh", "mktfhe_b200/csrc/capi.cu"]
CLASSES = (("phase1", ("k_phase1", "k_rgsw_tm", "k_cggi_w", "k_ccs_fast", "k_rgsw_blindrotate", "k_ccs_blindrotate")),
           ("phase2", ("k_phase2", "k_kms_phase2")), ("keyswitch", ("k_keyswitch",)))


def sources_hash():
    h = hashlib.sha256()
    for f in SOURCES:
        h.update(open(os.path.join(ROOT, f), "rb").read())
    return h.hexdigest()


def parse(path):
    per = {}
    with open(path, newline="") as fh:
        rows = [r for r in csv.reader(fh) if len(r) >= 15 and r[0].isdigit()]
    for r in rows:
        kid, kernel, metric, val = int(r[0]), r[4], r[12], float(r[14].replace(",", ""))
        per.setdefault(kid, {"kernel": kernel})[metric] = val
    out = {}
    for kid in sorted(per):
        d = per[kid]
        for cls, pats in CLASSES:
            if any(pt in d["kernel"] for pt in pats) and "dram__bytes_read.sum" in d:
                out[cls] = int(d["dram__bytes_read.sum"] + d.get("dram__bytes_write.sum", 0))
                out[cls + "_kernel"] = d["kernel"]
                if "lts__t_bytes.sum" in d:
                    out[cls + "_l2_bytes"] = int(d["lts__t_bytes.sum"])
    return out


def main():
    rec = {"sources": SOURCES, "sources_sha256": sources_hash(), "capture": ", ".join(a.split("=", 1)[1] for a in sys.argv[1:]), "workloads": {}}
    for a in sys.argv[1:]:
        name, path = a.split("=", 1)
        rec["workloads"][name] = parse(path)
    json.dump(rec, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
