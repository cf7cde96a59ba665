"""BASELINE config 5: GKRRoundSumcheck prove, dim=18, f1 with 2^18 nonzeros — GPU (through the C ABI, host buffers in,
proof out) vs the CPU oracle on all cores.  Prints one JSON line.  Not the headline metric; see bench.py for that."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sumcheck_b200 as sc
from sumcheck_b200.synth import synth_table_fast
from oracle import oracle as orc

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 18
n = 1 << dim
f2, f3 = synth_table_fast(n, 0x5C0500), synth_table_fast(n, 0x5C0501)
g = synth_table_fast(dim, 0x5C0502)
val = synth_table_fast(n, 0x5C0503)
rng = np.random.default_rng(0x5C0504)
idx = np.unique(rng.integers(0, 1 << (3 * dim), size=n + 4096, dtype=np.uint64))[:n].copy()
rng.shuffle(idx)
f1 = sc.SparseMultilinearExtension(3 * dim, idx, val[:idx.shape[0]].copy())
for _ in range(3):
    proof = sc.GKRRoundSumcheck.prove(sc.Blake2b512Rng.setup(), f1, f2, f3, g)
t0 = time.perf_counter()
K = 10
for _ in range(K):
    proof = sc.GKRRoundSumcheck.prove(sc.Blake2b512Rng.setup(), f1, f2, f3, g)
gpu_ms = (time.perf_counter() - t0) / K * 1e3
orc.set_threads(os.cpu_count() or 1)
t0 = time.perf_counter()
m1, m2, _, _ = orc.gkr_prove(orc.Rng(), dim, f1.indices, f1.values, f2, f3, g)
cpu_ms = (time.perf_counter() - t0) * 1e3
ok = np.array_equal(np.stack([m.evaluations for m in proof.phase1_sumcheck_msgs]), m1) and \
     np.array_equal(np.stack([m.evaluations for m in proof.phase2_sumcheck_msgs]), m2)
print(json.dumps({"workload": f"GKRRoundSumcheck prove dim={dim}, {idx.shape[0]} nonzeros (BASELINE config 5), host buffers in, proof out",
                  "gpu_ms_per_proof_e2e": gpu_ms, "cpu_oracle_ms": cpu_ms, "cpu_threads": os.cpu_count(), "parity": "bit-exact" if ok else "MISMATCH"}))
