"""Debug aid: per-round / per-point comparison of the contraction kernels with the oracle at a small size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ.setdefault("SC_TC_MIN_PAIRS", "128")
os.environ.setdefault("SC_RES_MAX_PAIRS", "64")
import numpy as np
import sumcheck_b200 as sc
from oracle import oracle as orc

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 9
nprod = int(sys.argv[2]) if len(sys.argv) > 2 else 1
m = int(sys.argv[3]) if len(sys.argv) > 3 else 3
tabs = [orc.synth_table(1 << nv, 100 + j) for j in range(m * nprod)]
coeffs = orc.synth_table(nprod, 999)
prods = [(coeffs[k], list(range(m * k, m * k + m))) for k in range(nprod)]
poly = sc.ListOfProductsOfPolynomials.new(nv)
for c, ix in prods:
    poly.add_product([tabs[j] for j in ix], c)
st = sc.IPForMLSumcheck.prover_init(poly)
got = np.zeros((nv, m + 1, 4), dtype=np.uint64)
st.prove_into(sc.Blake2b512Rng.setup(), got)
want, _, _ = orc.ml_prove(orc.Poly(nv, tabs, prods))
for i in range(nv):
    print("round", i + 1, "pairs", 1 << (nv - 1 - i), [bool(np.array_equal(got[i, t], want[i, t])) for t in range(m + 1)])
print("gemm rounds", st.gemm_round_count(), "tc", st.tc_round_count(), "launches", st.launch_count())
print("OK" if np.array_equal(got, want) else "MISMATCH")
