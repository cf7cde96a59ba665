"""BASELINE config 5 and SURVEY §8 f-3: GKRRoundSumcheck prove, dim=18, f1 with 2^18 nonzeros — one layer per call vs L layers
in ONE call (sc_gkr_prove_batch), host buffers in, proofs out; the CPU oracle on all cores for one layer.  One JSON line.
usage: python tools/gkr_bench.py [dim] [L]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sumcheck_b200 as sc
from sumcheck_b200.synth import synth_table_fast
from oracle import oracle as orc

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 18
L = int(sys.argv[2]) if len(sys.argv) > 2 else 16
n = 1 << dim


def layer(seed):
    f2, f3 = synth_table_fast(n, seed), synth_table_fast(n, seed + 1)
    g = synth_table_fast(dim, seed + 2)
    val = synth_table_fast(n, seed + 3)
    rng = np.random.default_rng(seed + 4)
    idx = np.unique(rng.integers(0, 1 << (3 * dim), size=n + 4096, dtype=np.uint64))[:n].copy()
    rng.shuffle(idx)
    return sc.SparseMultilinearExtension(3 * dim, idx, val[:idx.shape[0]].copy()), f2, f3, g


layers = [layer(0x5C0500 + 16 * l) for l in range(L)]
f1s, f2s, f3s, gs = map(list, zip(*layers))


def separate():
    return [sc.GKRRoundSumcheck.prove(sc.Blake2b512Rng.setup(), f1s[l], f2s[l], f3s[l], gs[l]) for l in range(L)]


def batched():
    return sc.GKRRoundSumcheck.prove_batch([sc.Blake2b512Rng.setup() for _ in range(L)], f1s, f2s, f3s, gs)


def timeit(fn, K=5):
    for _ in range(2):
        out = fn()
    t0 = time.perf_counter()
    for _ in range(K):
        out = fn()
    return (time.perf_counter() - t0) / K * 1e3, out


sep_ms, a = timeit(separate)
bat_ms, b = timeit(batched)
same = all(np.array_equal(x.evaluations, y.evaluations) for pa, pb in zip(a, b)
           for x, y in zip(pa.phase1_sumcheck_msgs + pa.phase2_sumcheck_msgs, pb.phase1_sumcheck_msgs + pb.phase2_sumcheck_msgs))
orc.set_threads(os.cpu_count() or 1)
t0 = time.perf_counter()
m1, m2, _, _ = orc.gkr_prove(orc.Rng(), dim, f1s[0].indices, f1s[0].values, f2s[0], f3s[0], gs[0])
cpu_ms = (time.perf_counter() - t0) * 1e3
ok = same and np.array_equal(np.stack([m.evaluations for m in b[0].phase1_sumcheck_msgs]), m1) and \
     np.array_equal(np.stack([m.evaluations for m in b[0].phase2_sumcheck_msgs]), m2)
h2d = sum(x.nbytes for x in (f1s[0].indices, f1s[0].values, f2s[0], f3s[0], gs[0]))
print(json.dumps({"workload": f"GKRRoundSumcheck prove dim={dim}, {f1s[0].indices.shape[0]} nonzeros per layer (BASELINE config 5 shape), {L} layers, pageable host buffers in, proofs out",
                  "separate_calls_ms": sep_ms, "per_layer_ms_separate": sep_ms / L, "one_batched_call_ms": bat_ms, "per_layer_ms_batched": bat_ms / L,
                  "speedup": sep_ms / bat_ms, "h2d_MiB_per_layer": h2d / 2**20, "upload_floor_ms_at_55GBps": L * h2d / 55e9 * 1e3,
                  "cpu_oracle_ms_one_layer": cpu_ms, "cpu_threads": os.cpu_count(), "parity": "bit-exact (batched == separate == oracle)" if ok else "MISMATCH"}))
