"""Times every BASELINE.json config on one GPU (tables resident in HBM, whole proof incl. transcript) next to the CPU
oracle on all host cores, and checks the proofs agree.  One JSON line per config.  bench.py stays the headline metric."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sumcheck_b200 as sc
from sumcheck_b200.synth import synth_table_fast
from oracle import oracle as orc

CONFIGS = [  # (name, cfg id, nv, n_products, multiplicands per product)
    ("config 1: nv=12, 1 product of degree 2", 1, 12, 1, 2),
    ("config 2: nv=20, degree 3", 2, 20, 1, 3),
    ("config 3: nv=24, degree 3 (G=1)", 3, 24, 1, 3),
    ("config 4: nv=22, 4 products each degree 4", 4, 22, 4, 4),
]
only = sys.argv[1:]
orc.set_threads(os.cpu_count() or 1)
for name, cfg, nv, n_products, m in CONFIGS:
    if only and str(cfg) not in only:
        continue
    T = n_products * m
    tabs = [synth_table_fast(1 << nv, 0x5C0000 + 0x100 * cfg + j) for j in range(T)]
    coeffs = synth_table_fast(n_products, 0x5C00FF + 0x100 * cfg)
    poly = sc.ListOfProductsOfPolynomials.new(nv)
    for k in range(n_products):
        poly.add_product(tabs[k * m:(k + 1) * m], coeffs[k])
    st = sc.IPForMLSumcheck.prover_init(poly)
    ev = np.zeros((nv, m + 1, 4), dtype=np.uint64)
    for _ in range(3):
        st.reset(); st.prove_into(sc.Blake2b512Rng.setup(), ev)
    K = 10
    t0 = time.perf_counter()
    for _ in range(K):
        st.reset(); st.prove_into(sc.Blake2b512Rng.setup(), ev)
    gpu_ms = (time.perf_counter() - t0) / K * 1e3
    opoly = orc.Poly(nv, tabs, [(coeffs[k], list(range(k * m, (k + 1) * m))) for k in range(n_products)])
    t0 = time.perf_counter()
    want, _, _ = orc.ml_prove(opoly)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    fs = (m + 1) * ((1 << nv) - 1)
    alg = 32 * T * (4 * (1 << nv) - 6)
    print(json.dumps({"config": name, "gpu_ms_per_proof": round(gpu_ms, 3), "field_sums_per_s": fs / (gpu_ms * 1e-3),
                      "algorithmic_GBps": alg / (gpu_ms * 1e-3) / 1e9, "cpu_oracle_ms": round(cpu_ms, 1), "cpu_threads": os.cpu_count(),
                      "parity": "bit-exact" if np.array_equal(ev, want) else "MISMATCH"}), flush=True)
    del st, poly, tabs
