#!/usr/bin/env python
"""Opcode histogram per kernel from `cuobjdump -sass` (evidence for profiles/: LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk,
SYNCS = mbarrier, DFMA/DADD/DMUL = FP64 pipe, STL/LDL = spills).   python tools/sass_hist.py lib.so [regex] """
import collections, re, subprocess, sys
lib = sys.argv[1]; pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn = None; hist = {}; maxreg = {}
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = m.group(1); hist[fn] = collections.Counter(); maxreg[fn] = 0; continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and fn:
        hist[fn][m.group(1)] += 1
        for r in re.findall(r"\bR(\d+)\b", ln): maxreg[fn] = max(maxreg[fn], int(r))
for fn, h in hist.items():
    dem = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip() or fn
    if pat and not pat.search(dem): continue
    tot = sum(h.values())
    print(f"== {dem}: {tot} instructions, highest register R{maxreg[fn]}")
    print("   " + "  ".join(f"{k}={v}" for k, v in h.most_common(28)))
    print("   tcgen05/TMA/mbarrier: " + "  ".join(f"{k}={h.get(k,0)}" for k in ("LDTM", "STTM", "UBLKCP", "SYNCS", "UTCBAR", "USETMAXREG", "SHFL", "STL", "LDL", "BAR")))
