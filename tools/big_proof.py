"""A proof larger than BASELINE's: python tools/big_proof.py [nv=26] — one product of three tables; rounds 1 and 2 are split over several
launches of the contraction kernels (s32 accumulator head-room, gemm_sum.cuh MAX_ITEMS_PER_CTA); checked bit for bit against the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sumcheck_b200 as sc
from oracle import oracle as orc
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 26
tabs = [orc.synth_table(1 << nv, 500 + j) for j in range(3)]
coeff = orc.synth_table(1, 599)[0]
poly = sc.ListOfProductsOfPolynomials.new(nv)
poly.add_product(tabs, coeff)
st = sc.IPForMLSumcheck.prover_init(poly)
got = np.zeros((nv, 4, 4), dtype=np.uint64)
for rep in range(3):
    st.reset()
    t0 = time.perf_counter()
    st.prove_into(sc.Blake2b512Rng.setup(), got)
    t = time.perf_counter() - t0
orc.set_threads(os.cpu_count() or 1)
want, _, _ = orc.ml_prove(orc.Poly(nv, tabs, [(coeff, [0, 1, 2])]))
byts = 32 * 3 * (4 * (1 << nv) - 6)
print(f"nv={nv}: {t * 1e3:.3f} ms per proof, {byts / t / 1e12:.2f} TB/s of algorithmic bytes over the whole proof, {st.launch_count()} launches, "
      f"{st.gemm_round_count()} rounds on the contraction kernels, bit-exact vs oracle: {np.array_equal(got, want)}")
