#!/usr/bin/env python
"""Lower bound on the FP64 issue cycles of a kernel from an ncu source-page CSV (SASS + executed counts).
Model measured with tools/fp64_issue.cu on B200: a DFMA/DADD/DMUL occupies its scheduler's FP64 path for
max(2, number of DISTINCT 64-bit register sources that are not served by the operand-reuse cache) cycles
(three distinct register sources: 3.05-3.3 cycles measured; constant-bank / immediate / reused operands are free).
    ncu -i rep.ncu-rep --page source --csv > src.csv ;  python tools/fp64_issue_bound.py src.csv [sm_cycles] """
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
prev_reuse = {}
tot2 = tot = 0.0; n3 = nall = 0
for r in rows[2:]:
    src = r[idx['Source']]
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?(DFMA|DADD|DMUL)\S*\s+(.*?);?$", src.strip())
    ex = float(r[idx['Instructions Executed']])
    if not m:
        prev_reuse = {}
        continue
    ops = [o.strip() for o in m.group(2).split(',')][1:]          # sources
    regs = []; reuse_now = {}
    for slot, o in enumerate(ops):
        mm = re.match(r"[-|]?(R\d+)(\.reuse)?\|?$", o)
        if mm and mm.group(1) != 'RZ':
            if prev_reuse.get(slot) != mm.group(1):
                regs.append(mm.group(1))
            if mm.group(2): reuse_now[slot] = mm.group(1)
    prev_reuse = reuse_now
    c = max(2, len(set(regs)))
    tot += c * ex; tot2 += 2 * ex; nall += ex
    if c >= 3: n3 += ex
print(f"FP64 warp-instructions {nall:.3e}; with >= 3 fresh register sources {100 * n3 / nall:.1f} %")
print(f"issue cycles at 2/instr {tot2:.3e}; with the register-source limit {tot:.3e}  (x{tot / tot2:.3f})")
if len(sys.argv) > 2:
    cyc = float(sys.argv[2]) * 148 * 4
    print(f"of the {cyc:.3e} scheduler-cycles elapsed: nominal pipe {100 * tot2 / cyc:.1f} %, register-limited bound {100 * tot / cyc:.1f} %")
