#!/usr/bin/env python3
"""usage: ncu_opmix.py <source-page.csv> [n_warp_iterations]: executed warp-instructions and stall samples per opcode."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ia, isrc, iex, ismp = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
ex = collections.Counter(); smp = collections.Counter(); stall = collections.defaultdict(collections.Counter)
first = True
for r in rows[2:]:
    if r and r[0] == 'Kernel Name': break  # only the first profiled launch
    if len(r) <= iex or not r[iex].isdigit(): continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
    if not m: continue
    op = m.group(2)
    ex[op] += int(r[iex] or 0); smp[op] += int(r[ismp] or 0)
    for i in stall_cols:
        v = int(r[i] or 0)
        if v: stall[op][h[i]] += v
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
tot = sum(ex.values()); ts = sum(smp.values())
print(f"total warp-instructions {tot} ({tot/div:.1f} per unit), samples {ts}")
for op, n in ex.most_common(28):
    top = ", ".join(f"{k[6:]} {v}" for k, v in stall[op].most_common(3))
    print(f"{op:28s} {n:12d} {n/div:9.1f}  samples {100*smp[op]/ts:5.1f}%  [{top}]")
