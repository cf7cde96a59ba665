"""Per-round breakdown of the resident rounds (SC_RES_PROF=1): python tools/res_prof.py [nv] [torch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sumcheck_b200 as sc
from sumcheck_b200.synth import synth_table_fast
nv = int(sys.argv[1]) if len(sys.argv) > 1 else 20
use_torch = len(sys.argv) > 2
tabs = [synth_table_fast(1 << nv, 0x5C0300 + j) for j in range(3)]
poly = sc.ListOfProductsOfPolynomials.new(nv)
poly.add_product(tabs, synth_table_fast(1, 0x5C03FF)[0])
st = sc.IPForMLSumcheck.prover_init(poly)
if use_torch:
    import torch
    st.set_stream(torch.cuda.current_stream().cuda_stream)
ev = np.zeros((nv, 4, 4), dtype=np.uint64)
for rep in range(6):
    st.reset()
    t0 = time.perf_counter()
    st.prove_into(sc.Blake2b512Rng.setup(), ev)
    print(f"proof {rep}: {(time.perf_counter() - t0) * 1e3:.3f} ms wall, {st.launch_count()} launches, {st.resident_round_count()} resident rounds", file=sys.stderr)
st.set_timing(True)
st.reset()
st.prove_into(sc.Blake2b512Rng.setup(), ev)
print("round_ms", [round(float(x), 4) for x in st.round_times_ms()], file=sys.stderr)
st.set_timing(False)
os.environ["SC_RES_PROF"] = "1"
st.reset()
st.prove_into(sc.Blake2b512Rng.setup(), ev)
