#!/usr/bin/env python3
"""usage: sass_inventory.py [lib.so] — static instruction counts per kernel from cuobjdump -sass: the mnemonics that prove the
Blackwell-native path (UTCIMMA = tcgen05.mma, UTMALDG = TMA load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier)."""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "sumcheck_b200/libsumcheck_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCIMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "REDUX", "IMAD.WIDE", "IMAD.HI", "LDS.128", "STS.128", "ATOMS", "ATOMG", "RED", "ELECT"]
fn, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(.*\)$", "", fn).replace("void ", "").replace("(int)", "")
        counts[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn:
        op = m.group(2)
        counts[fn]["total"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[fn][k] += 1
print(f"cuobjdump -sass {lib} (sm_100a), static instruction counts per kernel.")
print("UTCIMMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor (TMA load), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier,")
print("REDUX = redux.sync, ATOMS / RED = shared / global atomics (the anti-diagonal sums of the contraction epilogue).\n")
for fn, c in counts.items():
    print(f"{fn:58s} total {c['total']:6d}  " + "  ".join(f"{k} {c[k]}" for k in KEYS if c[k]))
