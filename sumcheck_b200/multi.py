"""Multi-GPU host plumbing: one process per GPU (torchrun / torch.multiprocessing), torch.distributed only carries the
128-byte communicator id; the data path (per-round all-gather of the d+1 partial sums, tail gather) runs inside
libsumcheck_b200.so over NCCL/NVLink (csrc/capi_multi.inc).  See SURVEY.md §8e / DESIGN.md "Multi-GPU"."""
import ctypes as C
import json
import os
import time

import numpy as np

from . import api, capi
from .api import ProverState, _check, _p32, _p64


def shard_range(nv, world, rank):
    """Elements of every table owned by `rank`: the HIGH log2(world) bits of the hypercube index select the rank."""
    assert world >= 1 and world & (world - 1) == 0, "world size must be a power of two"
    n = (1 << nv) // world
    assert n >= 2, "nv too small to shard"
    return rank * n, (rank + 1) * n


def broadcast_unique_id(dist, rank, device=None):
    """Rank 0 asks the library for a communicator id and shares it through torch.distributed (gloo or nccl)."""
    import torch
    buf = np.zeros(capi_id_bytes(), dtype=np.uint8)
    if rank == 0:
        _check(capi.lib().sc_comm_get_unique_id(buf.ctypes.data_as(capi.U8P)))
    t = torch.from_numpy(buf)
    if device is not None and dist.get_backend() == "nccl":
        t = t.to(device)
    dist.broadcast(t, src=0)
    return t.cpu().numpy().copy()


def capi_id_bytes():
    return 128


class Comm:
    def __init__(self, uid, rank, world, device):
        self.rank, self.world, self.device = rank, world, device
        h = C.c_void_p()
        uid = np.ascontiguousarray(uid, dtype=np.uint8)
        _check(capi.lib().sc_comm_create(C.byref(h), uid.ctypes.data_as(capi.U8P), rank, world, device))
        self._h = h

    def close(self):
        if self._h:
            capi.lib().sc_comm_destroy(self._h)
            self._h = None


def prover_init_sharded(comm, nv, shard_tables, products, keep=None):
    """IPForMLSumcheck::prover_init over this rank's shard.  products: [(coeff[4], [table indices])]."""
    coeffs = np.ascontiguousarray(np.stack([np.asarray(c, dtype=np.uint64) for c, _ in products]))
    offs, idx = [0], []
    for _, ix in products:
        idx.extend(ix)
        offs.append(len(idx))
    offsets, indices = np.array(offs, dtype=np.uint32), np.array(idx, dtype=np.uint32)
    tabs = (C.c_void_p * len(shard_tables))(*[t.ctypes.data for t in shard_tables])
    h = C.c_void_p()
    _check(capi.lib().sc_prover_create_sharded(C.byref(h), comm._h, nv, len(shard_tables), tabs, len(products), _p64(coeffs),
                                               _p32(offsets), _p32(indices)))
    st = ProverState(h)
    st._keep = (shard_tables, keep)
    return st


def ml_prove_sharded(comm, nv, shard_tables, products):
    """MLSumcheck::prove on a sharded polynomial; every rank returns the same evals[nv, d+1, 4]."""
    st = prover_init_sharded(comm, nv, shard_tables, products)
    d = st.max_multiplicands
    evals = np.zeros((nv, d + 1, 4), dtype=np.uint64)
    st.prove_into(api.Blake2b512Rng.setup(), evals)
    return evals, st


def bench_main(args, cfg, nv, B):
    """bench.py body for N > 1: strong scaling of ONE nv-variable proof over N GPUs.  Launched by torchrun (one process per
    GPU, WORLD_SIZE > 1) or, with --single-process, as one process driving all N GPUs (sc_prover_create_multi).
    B = the bench module (workload names, field_sums, ClockSampler ...)."""
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return _bench_torchrun(args, cfg, nv, B)
    return _bench_single_process(args, cfg, nv, B)


def _oracle_proof(cfg, nv, B):
    """The unsharded CPU oracle on the same synthetic inputs (all host threads): what every bench line is compared with."""
    from oracle import oracle as orc
    from .synth import synth_table_fast
    orc.set_threads(os.cpu_count() or 1)
    tabs, prods, _ = B.ml_inputs(cfg, nv, synth_table_fast)
    t0 = time.perf_counter()
    want = orc.ml_prove(orc.Poly(nv, tabs, prods))[0]
    return want, time.perf_counter() - t0


def _line(B, cfg, nv, d, T, world, args, ms_step, step_ms, round_ms, launches, e2e_ms, e2e_call, h2d, clocks, how, extra):
    fs = B.field_sums(nv, d)
    peak, src = B.peak_hbm()
    nv_l = nv - (world.bit_length() - 1)
    fold_bytes = sum(B.algorithmic_bytes(nv_l, T, i) for i in range(2, nv_l + 1))  # per GPU
    fold_ms = float(round_ms[1:nv_l].sum())
    ach = fold_bytes / (fold_ms * 1e-3) / 1e9 if fold_ms > 0 else 0.0
    total_bytes = 32 * T * (4 * (1 << nv) - 6)
    line = {
        "metric": B.metric_name(cfg), "value": fs / (ms_step * 1e-3), "unit": B.UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32x8 Montgomery (mod p, 255-bit)", "data": "synthetic",
        "config": {"workload": B.workload_name(cfg, nv), "sharding": f"tables sharded by the high {world.bit_length() - 1} hypercube bits over {world} GPUs; " + how,
                   "cache": f"per-GPU shard {T * ((1 << nv) // world) * 32 / 2**20:.0f} MiB, re-read from HBM every step",
                   "median_ms_per_step": B.median(step_ms), "round_ms_rank0": [round(float(x), 4) for x in round_ms]},
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "kernel": "sck::round_tc_kernel<3> / resident_kernel<3> on rank 0's shard (fold rounds up to the switch, incl. the fused peer-memory exchange), per GPU",
                     "peak_source": src,
                     "whole_proof": {"algorithmic_bytes": total_bytes, "achieved_per_gpu": total_bytes / world / (ms_step * 1e-3) / 1e9,
                                     "frac": total_bytes / world / (ms_step * 1e-3) / 1e9 / peak}},
        "e2e": {"value": fs / (B.median(e2e_ms) * 1e-3), "unit": B.UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": nv * (d + 1) * 32 * world, "ms_per_step": B.median(e2e_ms), "call": e2e_call},
        "gpu_launches": int(launches) * args.steps * world, "clocks": clocks,
    }
    line.update(extra)
    return line


def _bench_single_process(args, cfg, nv, B):
    import torch

    from .synth import synth_table_fast
    world = args.gpus
    _, _, n_products, m = B.CONFIGS[cfg]
    d, T = m, n_products * m
    tabs, prods, _ = B.ml_inputs(cfg, nv, synth_table_fast)  # full tables, pageable
    poly = api.ListOfProductsOfPolynomials.new(nv)
    for c, ix in prods:
        poly.add_product([tabs[j] for j in ix], c)
    devs = list(range(world))
    st = api.IPForMLSumcheck.prover_init(poly, device=devs)
    evals = np.zeros((nv, d + 1, 4), dtype=np.uint64)

    def prove():
        st.reset()
        st.prove_into(api.Blake2b512Rng.setup(), evals)

    for _ in range(max(args.warmup, 3)):
        prove()
    sampler = B.ClockSampler(0)
    sampler.start()
    for dv in devs:
        torch.cuda.synchronize(dv)
    step_ms = []
    for _ in range(args.steps):  # the call is synchronous (returns with the proof on the host): wall clock around it
        t0 = time.perf_counter()
        prove()
        step_ms.append((time.perf_counter() - t0) * 1e3)
    ms_step = sum(step_ms) / len(step_ms)
    st.set_timing(True)
    prove()
    st.set_timing(False)
    round_ms = st.round_times_ms().astype(np.float64)
    launches = st.launch_count()
    first = evals.copy()
    e2e_ms, out = [], np.zeros_like(evals)
    for k in range(args.steps + 1):  # the drop-in call over N GPUs: create (every rank uploads its slice) + prove + destroy
        t0 = time.perf_counter()
        s2 = api.IPForMLSumcheck.prover_init(poly, device=devs)
        s2.prove_into(api.Blake2b512Rng.setup(), out)
        del s2
        if k:
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop()
    extra = {}
    ok = np.array_equal(first, out)
    if not args.no_cpu_baseline:
        want, cpu_s = _oracle_proof(cfg, nv, B)
        ok = ok and np.array_equal(first, want)
        extra["cpu_baseline"] = {"value": B.field_sums(nv, d) / cpu_s, "unit": B.UNIT, "cores": os.cpu_count(), "kind": "port",
                                 "sample": "full workload, one proof, all host threads (OpenMP restatement of the rayon schedule)"}
        extra["parity"] = "bit-exact vs oracle" if ok else "MISMATCH vs oracle"
    line = _line(B, cfg, nv, d, T, world, args, ms_step, step_ms, round_ms, launches, e2e_ms,
                 "sc_prover_create_multi + sc_ml_prove + destroy: pageable caller tables, one process, one host thread per GPU",
                 T * (1 << nv) * 32, clocks, "ONE process, one host thread per GPU, exchange of the d partial sums fused into the round kernels (peer memory)", extra)
    print(json.dumps(line), flush=True)
    if not ok:
        raise SystemExit("parity check failed")


def _bench_torchrun(args, cfg, nv, B):
    import torch
    import torch.distributed as dist

    from .synth import synth_table_fast
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = int(os.environ.get("LOCAL_RANK", rank))
    _, _, n_products, m = B.CONFIGS[cfg]
    d, T = m, n_products * m
    torch.cuda.set_device(dev)
    # every rank stages its pageable shard through its own copy pool: share the host cores between the ranks of this node
    os.environ.setdefault("SC_COPY_THREADS", str(max(1, (os.cpu_count() or 1) // world)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    comm = Comm(broadcast_unique_id(dist, rank, torch.device("cuda", dev)), rank, world, dev)
    lo, hi = shard_range(nv, world, rank)
    n_loc = hi - lo
    # this rank's shard of the same synthetic tables the 1-GPU run uses (counter-based generator: any slice is cheap)
    tabs, prods, _ = B.ml_inputs(cfg, nv, synth_table_fast, first=lo, count=n_loc)  # pageable
    st = prover_init_sharded(comm, nv, tabs, prods)
    stream = torch.cuda.current_stream()
    st.set_stream(stream.cuda_stream)
    evals = np.zeros((nv, d + 1, 4), dtype=np.uint64)

    def prove():
        st.reset()
        st.prove_into(api.Blake2b512Rng.setup(), evals)

    def prove_e2e():
        st.load_tables(tabs)
        st.prove_into(api.Blake2b512Rng.setup(), evals)

    for _ in range(max(args.warmup, 3)):
        prove()
    sampler = B.ClockSampler(dev)
    if rank == 0:
        sampler.start()

    def timed(fn):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        dist.barrier()
        torch.cuda.synchronize()
        ev[0].record(stream)
        for k in range(args.steps):
            fn()
            ev[k + 1].record(stream)
        torch.cuda.synchronize()
        per = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
        t = torch.tensor([ev[0].elapsed_time(ev[args.steps]) / args.steps] + per, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
        t = t.cpu().tolist()
        return t[0], t[1:]

    ms_step, step_ms = timed(prove)
    st.set_timing(True)
    prove()
    st.set_timing(False)
    round_ms = st.round_times_ms().astype(np.float64)
    launches = st.launch_count()
    first = evals.copy()
    prove_e2e()
    _, e2e_ms = timed(prove_e2e)
    # the same with PINNED caller buffers (a caller that allocates its tables with cudaHostAlloc): no host-side staging copy
    pin, _, keep = B.ml_inputs(cfg, nv, synth_table_fast, first=lo, count=n_loc, pinned=True)

    def prove_reuse():
        st.load_tables(pin)
        st.prove_into(api.Blake2b512Rng.setup(), evals)

    prove_reuse()
    _, reuse_ms = timed(prove_reuse)
    same = torch.tensor(np.frombuffer(first.tobytes(), dtype=np.int64).copy(), device=f"cuda:{dev}")
    ref = same.clone()
    dist.broadcast(ref, src=0)
    agree = torch.tensor([int(torch.equal(same, ref)) * int(np.array_equal(first, evals))], device=f"cuda:{dev}")
    dist.all_reduce(agree, op=dist.ReduceOp.MIN)
    ok = bool(agree.item())
    if rank == 0:
        clocks = sampler.stop()
        extra = {}
        if not args.no_cpu_baseline:  # rank 0 proves the UNSHARDED polynomial on the CPU oracle and compares bit for bit
            want, cpu_s = _oracle_proof(cfg, nv, B)
            ok = ok and np.array_equal(first, want)
            extra["cpu_baseline"] = {"value": B.field_sums(nv, d) / cpu_s, "unit": B.UNIT, "cores": os.cpu_count(), "kind": "port",
                                     "sample": "full workload, one proof, all host threads (OpenMP restatement of the rayon schedule)"}
            extra["parity"] = ("bit-exact vs oracle (all ranks hold rank 0's proof)" if ok else "MISMATCH vs oracle")
        else:
            extra["parity"] = "ranks agree (oracle comparison skipped)" if ok else "RANKS DISAGREE"
        line = _line(B, cfg, nv, d, T, world, args, ms_step, step_ms, round_ms, launches, e2e_ms,
                     "sc_prover_load_tables (every rank re-uploads its pageable shard) + sc_ml_prove on one sharded handle per rank",
                     T * n_loc * 32 * world, clocks,
                     "one process per GPU, per-round exchange of the d partial sums fused into the round kernels (NVLink peer memory)", extra)
        line["e2e"]["note"] = ("pageable shards are staged through pinned bounce slots by host threads: every rank's staging copy shares the "
                               "node's host DRAM bandwidth, so this number does not scale with N on a 16-core host; e2e_reuse uploads from pinned buffers")
        line["e2e_reuse"] = {"value": B.field_sums(nv, d) / (B.median(reuse_ms) * 1e-3), "unit": B.UNIT, "ms_per_step": B.median(reuse_ms),
                             "call": "sc_prover_load_tables from PINNED shard buffers + sc_ml_prove on one sharded handle per rank"}
        print(json.dumps(line), flush=True)
    comm.close()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        raise SystemExit("parity check failed")
