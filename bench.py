#!/usr/bin/env python3
"""bench.py — MLSumcheck prover field-sums/sec (BLS12-381 Fr, nv=24, deg=3), the metric BASELINE.json names.

One "step" = one whole proof: MLSumcheck::prove_as_subprotocol's round loop (ml_sumcheck/mod.rs:59-64) for all nv rounds —
fused fold+sum kernels, the resident small-round kernel, per-round Blake2b transcript and challenge sampling.
  value     : field-sums/s with the tables already resident in HBM when the timed region starts; CUDA events
              on the launching stream around every step; `ms_per_step` = mean of the K steps, the median is in `config`
  e2e       : the same metric through the DROP-IN call — sc_ml_prove_oneshot = MLSumcheck::prove(&poly) (mod.rs:42-45):
              caller tables in PAGEABLE host memory (a Rust Vec<F>), create + upload + prove + destroy inside every step
  e2e_reuse : a caller that keeps one handle and pinned buffers (sc_prover_load_tables + sc_ml_prove)
--config N selects a BASELINE.json config (1..5; default 3 = the headline).  Prints ONE JSON line (rank 0).
`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) on the same workload; the Rust
reference itself cannot be built in this image (no rustc).  Every line carries `parity`: the proof is compared with the
oracle's bit for bit, and the run fails if they differ.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "field-sums/s"
# BASELINE.json configs: (nv or dim, products, multiplicands per product)
CONFIGS = {1: ("ml", 12, 1, 2), 2: ("ml", 20, 1, 3), 3: ("ml", 24, 1, 3), 4: ("ml", 22, 4, 4), 5: ("gkr", 18, 1, 2)}


def metric_name(cfg):
    kind, nv, n_products, m = CONFIGS[cfg]
    if cfg == 3:
        return "MLSumcheck prover field-sums/sec (BLS12-381 Fr, nv=24, deg=3)"
    if kind == "gkr":
        return f"GKRRoundSumcheck prover field-sums/sec (BLS12-381 Fr, dim={nv}, two phases of deg 2)"
    return f"MLSumcheck prover field-sums/sec (BLS12-381 Fr, nv={nv}, {n_products} product(s) of deg {m})"


def workload_name(cfg, nv=None):
    """The SAME string in both arms (the driver compares them)."""
    kind, nv0, n_products, m = CONFIGS[cfg]
    nv = nv or nv0
    if kind == "gkr":
        return f"GKRRoundSumcheck prove dim={nv}, f1 with 2^{nv} nonzeros over {3 * nv} variables, f2/f3 dense (BASELINE config 5), whole proof incl. transcript"
    return (f"MLSumcheck prove nv={nv} deg={m} T={n_products * m}, {n_products} product(s) (BASELINE config {cfg}), "
            "whole proof incl. transcript")


def field_sums(nv, d, phases=1):
    return phases * (d + 1) * ((1 << nv) - 1)  # sum_i (d+1) * 2^(nv-i)


def algorithmic_bytes(nv, T, rnd):
    """round 1 reads T*N*32; round i>=2 reads T*2^(nv-i+2)*32 and writes T*2^(nv-i+1)*32."""
    N = 1 << nv
    return 32 * T * (N if rnd == 1 else 3 * (1 << (nv - rnd + 1)))


class ClockSampler:
    """nvidia-smi sampler (B200_PROFILING.md clocks line) running during the timed region."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.th.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- synthetic inputs
def ml_inputs(cfg, nv, gen, first=0, count=None, pinned=False):
    """table j of config c uses seed 0x5C0000 + 0x100*c + j, coefficients 0x5C00FF + 0x100*c."""
    import numpy as np
    _, _, n_products, m = CONFIGS[cfg]
    T = n_products * m
    count = (1 << nv) if count is None else count
    tabs, keep = [], []
    for j in range(T):
        if pinned:
            import torch
            h = torch.empty((count, 4), dtype=torch.int64, pin_memory=True)
            keep.append(h)
            a = h.numpy().view(np.uint64)
        else:
            a = np.empty((count, 4), dtype=np.uint64)  # pageable, like a Rust Vec<F>
        gen(count, 0x5C0000 + 0x100 * cfg + j, out=a, first=first)
        tabs.append(a)
    coeffs = gen(n_products, 0x5C00FF + 0x100 * cfg)
    prods = [(coeffs[k], list(range(k * m, (k + 1) * m))) for k in range(n_products)]
    return tabs, prods, keep


def gkr_inputs(dim, gen):
    import numpy as np
    n = 1 << dim
    f2, f3 = gen(n, 0x5C0500), gen(n, 0x5C0501)
    g = gen(dim, 0x5C0502)
    val = gen(n, 0x5C0503)
    rng = np.random.default_rng(0x5C0504)
    idx = np.unique(rng.integers(0, 1 << (3 * dim), size=n + 4096, dtype=np.uint64))[:n].copy()
    rng.shuffle(idx)
    return idx, val[:idx.shape[0]].copy(), f2, f3, g


# ---------------------------------------------------------------------------------------------------- CPU arm
def oracle_gen():
    """Input generator of the CPU arm: the oracle's own (the same counter-based stream as the product's helper)."""
    from oracle import oracle as orc

    def gen(n, seed, out=None, first=0):
        assert first == 0
        t = orc.synth_table(n, seed)
        if out is not None:
            out[:] = t
            return out
        return t
    return gen


def run_reference(args, cfg, nv):
    """CPU arm: the oracle's restatement of the reference schedule (prover.rs:85-148), all host threads."""
    import numpy as np
    from oracle import oracle as orc
    synth_table_fast = oracle_gen()  # nothing of the product is loaded on this arm
    kind, _, n_products, m = CONFIGS[cfg]
    cores = os.cpu_count() or 1
    orc.set_threads(cores)
    if kind == "gkr":
        idx, val, f2, f3, g = gkr_inputs(nv, synth_table_fast)
        run = lambda: orc.gkr_prove(orc.Rng(), nv, idx, val, f2, f3, g)[:2]
        fs, d = field_sums(nv, 2, phases=2), 2
    else:
        tabs, prods, _ = ml_inputs(cfg, nv, synth_table_fast)
        poly = orc.Poly(nv, tabs, prods)
        run = lambda: orc.ml_prove(poly)[0]
        fs, d = field_sums(nv, m), m
    warm = max(args.warmup, 1) if nv <= 20 else min(args.warmup, 1)  # a full nv=24 CPU proof takes seconds: one warm-up
    for _ in range(warm):
        run()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        out = run()
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    v = fs / dt
    sample = (f"full workload: one {'GKRRoundSumcheck' if kind == 'gkr' else 'MLSumcheck'}::prove per step, {cores} threads "
              "(OpenMP restatement of the rayon schedule)")
    line = {"metric": metric_name(cfg), "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u64x4 Montgomery (mod p, 255-bit)", "data": "synthetic",
            "config": {"workload": workload_name(cfg, nv), "median_ms_per_step": sorted(times)[len(times) // 2] * 1e3},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    return line, out


def one_thread_cpu(cfg, nv):
    """1-thread number of the CPU restatement on a bounded sample (the algorithm is linear in 2^nv): BASELINE.md §3.2."""
    from oracle import oracle as orc
    synth_table_fast = oracle_gen()
    kind, _, n_products, m = CONFIGS[cfg]
    nv_s = min(nv, 19 if kind == "ml" else 16)
    orc.set_threads(1)
    try:
        if kind == "gkr":
            idx, val, f2, f3, g = gkr_inputs(nv_s, synth_table_fast)
            t0 = time.perf_counter()
            orc.gkr_prove(orc.Rng(), nv_s, idx, val, f2, f3, g)
            fs = field_sums(nv_s, 2, phases=2)
        else:
            tabs, prods, _ = ml_inputs(cfg, nv_s, synth_table_fast)
            poly = orc.Poly(nv_s, tabs, prods)
            t0 = time.perf_counter()
            orc.ml_prove(poly)
            fs = field_sums(nv_s, m)
        dt = time.perf_counter() - t0
    finally:
        orc.set_threads(os.cpu_count() or 1)
    return {"value": fs / dt, "unit": UNIT, "cores": 1, "sample": f"one proof of the same shape at nv={nv_s} ({dt * 1e3:.0f} ms)"}


# ---------------------------------------------------------------------------------------------------- GPU arm
def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def median(xs):
    s = sorted(xs)
    return s[len(s) // 2]


def bench_ml(args, cfg, nv, dev):
    import numpy as np
    import torch

    import sumcheck_b200 as sc
    from sumcheck_b200.synth import synth_table_fast
    _, _, n_products, m = CONFIGS[cfg]
    d, T, N = m, n_products * m, 1 << nv
    tabs, prods, _ = ml_inputs(cfg, nv, synth_table_fast)  # pageable: what MLSumcheck::prove(&poly) is handed
    poly = sc.ListOfProductsOfPolynomials.new(nv)
    for c, ix in prods:
        poly.add_product([tabs[j] for j in ix], c)
    t0 = time.perf_counter()
    st = sc.IPForMLSumcheck.prover_init(poly, device=dev)  # first call: context, kernels, bounce slots, pool threads
    first_init_ms = (time.perf_counter() - t0) * 1e3
    stream = torch.cuda.current_stream()
    st.set_stream(stream.cuda_stream)
    evals = np.zeros((nv, d + 1, 4), dtype=np.uint64)

    def prove_resident():
        st.reset()
        st.prove_into(sc.Blake2b512Rng.setup(), evals)

    for _ in range(args.warmup):
        prove_resident()
    sampler = ClockSampler(dev)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    torch.cuda.synchronize()
    ev[0].record(stream)
    for k in range(args.steps):
        prove_resident()
        ev[k + 1].record(stream)
    torch.cuda.synchronize()
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    ms_step = ev[0].elapsed_time(ev[args.steps]) / args.steps
    launches = st.launch_count()
    n_tc, n_res, n_gemm = st.tc_round_count(), st.resident_round_count(), st.gemm_round_count()
    gemm = n_gemm > 0  # degree-3 products: the sum over the hypercube runs as a tensor-core contraction (csrc/gemm_sum.cuh)
    # per-round kernel times (CUDA events on the launching stream) from extra, separately instrumented steps
    round_ms = np.zeros(nv, dtype=np.float64)
    st.set_timing(True)
    prove_resident()  # (discarded) instrumented rounds run the ordinary build of the fold kernel: load it before timing
    for _ in range(args.steps):
        prove_resident()
        round_ms += st.round_times_ms()
    st.set_timing(False)
    round_ms /= args.steps
    first = evals.copy()

    # ---- e2e: the drop-in call, pageable tables, create + upload + prove + destroy every step
    e2e_out = np.zeros_like(evals)
    for _ in range(2):
        sc.MLSumcheck.prove_into(poly, e2e_out, device=dev)
    torch.cuda.synchronize()
    e2e_ms = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        sc.MLSumcheck.prove_into(poly, e2e_out, device=dev)
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
    assert np.array_equal(first, e2e_out), "resident and one-shot proofs differ"
    # ---- e2e_reuse: one handle, pinned caller buffers
    pin, _, keep = ml_inputs(cfg, nv, synth_table_fast, pinned=True)
    st.load_tables(pin)
    st.prove_into(sc.Blake2b512Rng.setup(), evals)
    reuse_ms = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        st.load_tables(pin)
        st.prove_into(sc.Blake2b512Rng.setup(), evals)
        reuse_ms.append((time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop()
    assert np.array_equal(first, evals), "resident and handle-reuse proofs differ"

    fs = field_sums(nv, d)
    peak, peak_src = peak_hbm()
    total_bytes = 32 * T * (4 * N - 6)
    # dominant kernel: the fold rounds on sck::round_tc_kernel<d> (TMA-staged tables, fix_variables on the tensor cores, fused
    # with the d-point sum; P(1) from the claim) = rounds 2..1+n_tc, aggregated over its launches.  Round 1 runs
    # round1_tma_kernel, the remaining rounds share ONE launch of resident_kernel (latency-bound).
    tc_rounds = list(range(2, 2 + n_tc))
    if tc_rounds:
        dom_bytes = sum(algorithmic_bytes(nv, T, i) for i in tc_rounds)
        dom_ms = float(sum(round_ms[i - 1] for i in tc_rounds))
        dom_name = (f"gsum::gemm_fold_kernel<{2 if d == 4 else 3}, {d}> (TMA + tcgen05.mma fix_variables, {3 if d == 3 else 6} plain products per pair, "
                    f"tcgen05.mma contraction over the pairs), rounds 2..{1 + n_tc} aggregated") if gemm else \
                   f"sck::round_tc_kernel<{d}> (TMA + tcgen05.mma fold fused with the sum), rounds 2..{1 + n_tc} aggregated"
    else:  # small configs: everything after round 1 is the resident launch
        dom_bytes = sum(algorithmic_bytes(nv, T, i) for i in range(2, nv + 1))
        dom_ms = float(round_ms[1:].sum())
        dom_name = f"sck::resident_kernel<{d}> (all fold rounds in one launch; latency-bound)"
    ach = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    res_ms = float(round_ms[1 + n_tc:].sum())
    e2e_med, reuse_med = median(e2e_ms), median(reuse_ms)
    line = {
        "metric": metric_name(cfg), "value": fs / (ms_step * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32x8 Montgomery (mod p, 255-bit)", "data": "synthetic",
        "config": {"workload": workload_name(cfg, nv),
                   "cache": f"inputs {T * N * 32 / 2**20:.0f} MiB" + (" > 126 MB L2, re-read from HBM every step" if T * N * 32 > 126e6 else
                                                                  " (fits the 126 MB L2: this config is latency-bound, see roofline.note)"),
                   "median_ms_per_step": median(step_ms), "first_prover_init_ms": first_init_ms, "proofs_per_s": 1e3 / ms_step,
                   "hypercube_points_per_s": N / (ms_step * 1e-3), "kernel_ms_per_step": float(round_ms.sum()),
                   "round_ms": [round(float(x), 4) for x in round_ms],
                   "rounds": ({"gemm_round1_kernel": 1, "gemm_fold_kernel": int(n_tc), "resident_kernel (one launch)": int(n_res)} if gemm else
                              {"round1_tma_kernel": 1, "round_tc_kernel": int(n_tc), "resident_kernel (one launch)": int(n_res)})},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "kernel": dom_name, "algorithmic_bytes": dom_bytes, "kernel_ms": dom_ms, "peak_source": peak_src,
                     "whole_proof": {"algorithmic_bytes": total_bytes, "achieved": total_bytes / (ms_step * 1e-3) / 1e9,
                                     "frac": total_bytes / (ms_step * 1e-3) / 1e9 / peak,
                                     "frac_of_8TBps_nominal": total_bytes / (ms_step * 1e-3) / 8e12},
                     "round1": {"kernel": f"gsum::gemm_round1_kernel<3, {d}>" if gemm else f"sck::round1_tma_kernel<{d + 1}>", "achieved": algorithmic_bytes(nv, T, 1) / (round_ms[0] * 1e-3) / 1e9,
                                "frac": algorithmic_bytes(nv, T, 1) / (round_ms[0] * 1e-3) / 1e9 / peak},
                     "resident_rounds_ms": res_ms,
                     "note": "traffic: see profiles/ (ncu dram bytes per launch); not re-measured inside bench.py"},
        "e2e": {"value": fs / (e2e_med * 1e-3), "unit": UNIT, "h2d_bytes_per_step": T * N * 32,
                "d2h_bytes_per_step": nv * (d + 1) * 32, "ms_per_step": e2e_med,
                "mean_ms_per_step": sum(e2e_ms) / len(e2e_ms),
                "call": "sc_ml_prove_oneshot (= MLSumcheck::prove(&poly)): pageable caller tables, create + upload + prove + destroy per step"},
        "e2e_reuse": {"value": fs / (reuse_med * 1e-3), "unit": UNIT, "ms_per_step": reuse_med,
                      "call": "sc_prover_load_tables from pinned buffers + sc_ml_prove on one handle"},
        "gpu_launches": int(launches) * args.steps, "clocks": clocks,
    }
    return line, first


def bench_gkr(args, dim, dev):
    import numpy as np
    import torch

    import sumcheck_b200 as sc
    from sumcheck_b200.synth import synth_table_fast
    idx, val, f2, f3, g = gkr_inputs(dim, synth_table_fast)
    f1 = sc.SparseMultilinearExtension(3 * dim, idx, val)

    def prove():
        return sc.GKRRoundSumcheck.prove(sc.Blake2b512Rng.setup(), f1, f2, f3, g, device=dev)

    for _ in range(args.warmup):
        proof = prove()
    sampler = ClockSampler(dev)
    sampler.start()
    torch.cuda.synchronize()
    ms = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        proof = prove()
        ms.append((time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop()
    med, mean = median(ms), sum(ms) / len(ms)
    fs = field_sums(dim, 2, phases=2)
    peak, peak_src = peak_hbm()
    n = 1 << dim
    alg = 2 * 32 * 2 * (4 * n - 6)  # two phases of a T = 2 sumcheck
    h2d = idx.nbytes + val.nbytes + f2.nbytes + f3.nbytes + g.nbytes
    got = (np.stack([m_.evaluations for m_ in proof.phase1_sumcheck_msgs]), np.stack([m_.evaluations for m_ in proof.phase2_sumcheck_msgs]))
    line = {
        "metric": metric_name(5), "value": fs / (mean * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": mean, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u32x8 Montgomery (mod p, 255-bit)", "data": "synthetic",
        "config": {"workload": workload_name(5, dim), "median_ms_per_step": med,
                   "cache": f"{h2d / 2**20:.0f} MiB of inputs uploaded every step (the entry point takes host buffers); 16 MiB per phase fits L2: latency-bound",
                   "note": "value == e2e: GKRRoundSumcheck::prove has no resident-input form in the reference API"},
        "roofline": {"bound": "hbm", "achieved": alg / (mean * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg / (mean * 1e-3) / 1e9 / peak, "traffic": None, "algorithmic_bytes": alg, "peak_source": peak_src,
                     "kernel": "whole call: initialisers (eq tables, 64-bit atomic scatters), 2 x (round1 + resident_kernel<2>); 2*dim sequential rounds: latency-bound"},
        "e2e": {"value": fs / (med * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 2 * dim * 2 * 64,
                "ms_per_step": med, "call": "sc_gkr_prove (= GKRRoundSumcheck::prove): pageable host buffers in, proof out"},
        "gpu_launches": None, "clocks": clocks,
    }
    return line, got


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--nv", type=int, default=None, help="override the config's nv / dim (experiments)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--single-process", action="store_true", help="N > 1 GPUs driven by ONE process (sc_prover_create_multi)")
    args = ap.parse_args()
    cfg = args.config
    kind, nv0, n_products, m = CONFIGS[cfg]
    nv = args.nv or nv0
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank == 0:
            line, _ = run_reference(args, cfg, nv)
            print(json.dumps(line), flush=True)
        return

    if world > 1 or (args.gpus > 1 and args.single_process):
        from sumcheck_b200 import multi
        if kind != "ml" or n_products != 1:
            raise SystemExit("the sharded bench runs the single-product MLSumcheck configs (BASELINE config 3)")
        return multi.bench_main(args, cfg, nv, sys.modules[__name__])

    import numpy as np
    import torch
    args.warmup = max(args.warmup, 3)  # timing rules: at least 3 warm-up steps, whatever was asked for
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    if kind == "gkr":
        line, got = bench_gkr(args, nv, dev)
    else:
        line, got = bench_ml(args, cfg, nv, dev)
    ok = True
    if not args.no_cpu_baseline:
        ref, want = run_reference(argparse.Namespace(gpus=1, steps=1, warmup=0), cfg, nv)
        line["cpu_baseline"] = ref["cpu_baseline"]
        line["cpu_baseline"]["one_thread"] = one_thread_cpu(cfg, nv)
        ok = all(np.array_equal(a, b) for a, b in zip(got, want)) if kind == "gkr" else np.array_equal(got, want)
        line["parity"] = "bit-exact vs oracle" if ok else "MISMATCH vs oracle"
    print(json.dumps(line), flush=True)
    if not ok:
        raise SystemExit("parity check failed: the CUDA proof differs from the oracle's")


if __name__ == "__main__":
    main()
