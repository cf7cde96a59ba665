#!/usr/bin/env python3
"""usage: ncu_opmix_all.py <source-page.csv[.gz]>: opcode mix (executed warp-instructions, stall-sample share) of EVERY profiled
launch in an `ncu --page source --csv` export."""
import csv, sys, collections, re, gzip
op = gzip.open if sys.argv[1].endswith('.gz') else open
rows = list(csv.reader(op(sys.argv[1], 'rt')))
i = 0
while i < len(rows):
    if not rows[i] or rows[i][0] != 'Kernel Name':
        i += 1; continue
    name = rows[i][1]; h = rows[i + 1]
    isrc, iex, ismp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    ex = collections.Counter(); smp = collections.Counter()
    i += 2
    while i < len(rows) and not (rows[i] and rows[i][0] == 'Kernel Name'):
        r = rows[i]; i += 1
        if len(r) <= iex or not r[iex].isdigit(): continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
        if not m: continue
        ex[m.group(2)] += int(r[iex] or 0); smp[m.group(2)] += int(r[ismp] or 0)
    tot = sum(ex.values()) or 1; ts = sum(smp.values()) or 1
    print(f"\n== opcode mix, launch `{name}`: {tot} warp-instructions executed")
    for o, n in ex.most_common(18):
        print(f"   {o:28s} {n:12d}  {100*n/tot:5.1f} % of instructions   {100*smp[o]/ts:5.1f} % of stall samples")
