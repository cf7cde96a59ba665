"""One-shot e2e timing (sc_ml_prove_oneshot from pageable tables), nv=24: python tools/e2e_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench as B
import sumcheck_b200 as sc
from sumcheck_b200.synth import synth_table_fast
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 24
tabs, prods, _ = B.ml_inputs(3, nv, synth_table_fast)
poly = sc.ListOfProductsOfPolynomials.new(nv)
for c, ix in prods:
    poly.add_product([tabs[j] for j in ix], c)
out = np.zeros((nv, 4, 4), dtype=np.uint64)
for _ in range(3):
    sc.MLSumcheck.prove_into(poly, out)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); sc.MLSumcheck.prove_into(poly, out); ts.append((time.perf_counter() - t0) * 1e3)
print(f"e2e oneshot nv={nv}: median {sorted(ts)[5]:.2f} ms  min {min(ts):.2f}")
