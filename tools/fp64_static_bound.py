#!/usr/bin/env python
"""Static (unweighted) version of tools/fp64_issue_bound.py on `cuobjdump -sass` text of one kernel: average FP64 issue cycles
per instruction under the measured register-source model.   python tools/fp64_static_bound.py lib.so kernel_regex"""
import re, subprocess, sys
txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
pat = re.compile(sys.argv[2])
on = False; prev = {}; n = c2 = c = n3 = 0
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        dem = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout
        on = bool(pat.search(dem)); prev = {}; continue
    if not on: continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)\s+(.*?);", ln)
    if not m: continue
    op = m.group(1).split('.')[0]
    if op not in ("DFMA", "DADD", "DMUL"):
        prev = {}; continue
    ops = [o.strip() for o in m.group(2).split(',')][1:]
    regs = []; now = {}
    for slot, o in enumerate(ops):
        mm = re.match(r"[-|]?(R\d+)(\.reuse)?\|?$", o)
        if mm:
            if prev.get(slot) != mm.group(1): regs.append(mm.group(1))
            if mm.group(2): now[slot] = mm.group(1)
    prev = now
    k = max(2, len(set(regs))); n += 1; c += k; c2 += 2; n3 += k >= 3
print(f"{n} FP64 instructions, {100*n3/max(n,1):.1f} % with >= 3 fresh register sources, average issue cycles {c/max(n,1):.3f} (x{c/max(c2,1):.3f} of 2)")
