"""Per-round kernel times of one resident proof: python tools/deg_bench.py <nv> <multiplicands> [products]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sumcheck_b200 as sc
from sumcheck_b200.synth import synth_table_fast
nv, m = int(sys.argv[1]), int(sys.argv[2])
k = int(sys.argv[3]) if len(sys.argv) > 3 else 1
tabs = [synth_table_fast(1 << nv, 0x77000 + j) for j in range(m * k)]
poly = sc.ListOfProductsOfPolynomials.new(nv)
for i in range(k):
    poly.add_product(tabs[m * i:m * i + m], synth_table_fast(1, 0x77F00 + i)[0])
st = sc.IPForMLSumcheck.prover_init(poly)
ev = np.zeros((nv, m + 1, 4), dtype=np.uint64)
ts = []
for rep in range(6):
    st.reset()
    t0 = time.perf_counter()
    st.prove_into(sc.Blake2b512Rng.setup(), ev)
    ts.append((time.perf_counter() - t0) * 1e3)
st.set_timing(True)
st.reset()
st.prove_into(sc.Blake2b512Rng.setup(), ev)
rm = st.round_times_ms()
T = m * k
print(f"nv={nv} m={m} products={k}: proof {min(ts[2:]):.3f} ms wall; round_ms {[round(float(x), 4) for x in rm[:9]]}; "
      f"round 1 {32 * T * (1 << nv) / (rm[0] * 1e-3) / 1e12:.2f} TB/s, round 2 {32 * T * 3 * (1 << (nv - 1)) / (rm[1] * 1e-3) / 1e12:.2f} TB/s; "
      f"gemm rounds {st.gemm_round_count()}, launches {st.launch_count()}")
