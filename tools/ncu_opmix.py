#!/usr/bin/env python3
"""usage: ncu -i rep --page source --csv --kernel-name regex:X | tools/ncu_opmix.py [cells_per_warp_unit]
dynamic opcode mix of one kernel, normalised by N warp-level work units"""
import csv, sys, collections
unit = float(sys.argv[1]) if len(sys.argv) > 1 else 1048576.0
rows = list(csv.reader(sys.stdin))
h = rows[1]; ie = h.index("Instructions Executed"); si = h.index("Source")
c = collections.Counter()
for r in rows[2:]:
    if len(r) <= ie: continue
    toks = r[si].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    if not r[ie].isdigit(): break      # a second view (source lines) follows the SASS view
    c[op.split(".")[0].rstrip(";")] += int(r[ie])
print("total per unit %.1f" % (sum(c.values()) / unit))
print(" ".join(f"{op}:{n/unit:.1f}" for op, n in c.most_common(40)))
