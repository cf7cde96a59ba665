"""Workload for ncu captures: python tools/prof_run.py <config 1..5> [proofs]  — the same synthetic inputs as bench.py,
`proofs` resident proofs (default 4: 3 warm-up + 1), nothing else (no e2e legs, no CPU baseline)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench as B
import sumcheck_b200 as sc
from sumcheck_b200.synth import synth_table_fast
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
kind, nv, n_products, m = B.CONFIGS[cfg]
if kind == "gkr":
    idx, val, f2, f3, g = B.gkr_inputs(nv, synth_table_fast)
    f1 = sc.SparseMultilinearExtension(3 * nv, idx, val)
    for _ in range(reps):
        sc.GKRRoundSumcheck.prove(sc.Blake2b512Rng.setup(), f1, f2, f3, g)
else:
    tabs, prods, _ = B.ml_inputs(cfg, nv, synth_table_fast)
    poly = sc.ListOfProductsOfPolynomials.new(nv)
    for c, ix in prods:
        poly.add_product([tabs[j] for j in ix], c)
    st = sc.IPForMLSumcheck.prover_init(poly)
    ev = np.zeros((nv, m + 1, 4), dtype=np.uint64)
    for _ in range(reps):
        st.reset()
        st.prove_into(sc.Blake2b512Rng.setup(), ev)
    print("launches per proof", st.launch_count(), "tc rounds", st.tc_round_count(), "resident rounds", st.resident_round_count())
