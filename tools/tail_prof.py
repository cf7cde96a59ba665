import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sumcheck_b200 as sc
from sumcheck_b200.synth import synth_table_fast
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 14
tabs = [synth_table_fast(1 << nv, 7 + j) for j in range(3)]
poly = sc.ListOfProductsOfPolynomials.new(nv)
poly.add_product(tabs, synth_table_fast(1, 99)[0])
st = sc.IPForMLSumcheck.prover_init(poly)
ev = np.zeros((nv, 4, 4), dtype=np.uint64)
for i in range(3):
    st.reset()
    if i == 2: os.environ["SC_TAIL_PROF"] = "1"
    st.prove_into(sc.Blake2b512Rng.setup(), ev)
