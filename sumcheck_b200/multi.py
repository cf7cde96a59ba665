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


def bench_main(args, nv, d, T, METRIC, UNIT, field_sums, algorithmic_bytes, ClockSampler):
    """bench.py body for N > 1 (launched by torchrun): strong scaling of ONE nv-variable proof over N GPUs."""
    import torch
    import torch.distributed as dist

    from .synth import synth_table_fast
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    comm = Comm(broadcast_unique_id(dist, rank, torch.device("cuda", dev)), rank, world, dev)
    lo, hi = shard_range(nv, world, rank)
    n_loc = hi - lo
    # this rank's shard of the same synthetic tables the 1-GPU run uses (counter-based generator: any slice is cheap)
    host = [torch.empty((n_loc, 4), dtype=torch.int64, pin_memory=True) for _ in range(T)]
    tabs = [h.numpy().view(np.uint64) for h in host]
    for j in range(T):
        synth_table_fast(n_loc, 0x5C0300 + j, out=tabs[j], first=lo)
    coeff = synth_table_fast(1, 0x5C03FF)[0]
    st = prover_init_sharded(comm, nv, tabs, [(coeff, list(range(T)))])
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
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # max over ranks
        return float(t.item())

    ms_step = timed(prove)
    st.set_timing(True)
    prove()
    st.set_timing(False)
    round_ms = st.round_times_ms().astype(np.float64)
    launches = st.launch_count()
    first = evals.copy()
    prove_e2e()
    ms_e2e = timed(prove_e2e)
    same = torch.tensor(np.frombuffer(first.tobytes(), dtype=np.int64).copy(), device=f"cuda:{dev}")
    ref = same.clone()
    dist.broadcast(ref, src=0)
    agree = torch.tensor([int(torch.equal(same, ref))], device=f"cuda:{dev}")
    dist.all_reduce(agree, op=dist.ReduceOp.MIN)
    if rank == 0:
        clocks = sampler.stop()
        fs = field_sums(nv, d)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        pk = os.path.join(root, "MEASURED_PEAKS.json")
        peak, src = (json.load(open(pk))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if os.path.exists(pk) else (6650.0, "fallback")
        nv_l = nv - (world.bit_length() - 1)
        fold_bytes = sum(algorithmic_bytes(nv_l, T, i) for i in range(2, nv_l + 1))  # per GPU
        fold_ms = float(round_ms[1:nv_l].sum())
        ach = fold_bytes / (fold_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": fs / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32x8 Montgomery (mod p, 255-bit)", "data": "synthetic",
            "config": {"workload": f"MLSumcheck prove nv={nv} deg={d} T={T}, 1 product (BASELINE config 3), tables sharded by the high "
                                   f"{world.bit_length() - 1} hypercube bits over {world} GPUs, per-round exchange of the d+1 partial sums fused into the round kernel (NVLink peer memory)",
                       "cache": f"per-GPU shard {T * n_loc * 32 / 2**20:.0f} MiB, re-read from HBM every step",
                       "round_ms_rank0": [round(float(x), 4) for x in round_ms], "ranks_agree": bool(agree.item())},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                         "kernel": "sck::round_tc_kernel<3> / round_kernel<3,true> on rank 0's shard (TMA + tcgen05.mma fold for rounds with >= 2^14 pairs per shard; incl. the fused peer-memory exchange), sharded rounds aggregated",
                         "peak_source": src},
            "e2e": {"value": fs / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": T * n_loc * 32 * world,
                    "d2h_bytes_per_step": nv * (d + 1) * 32 * 2 * world, "ms_per_step": ms_e2e},
            "gpu_launches": int(launches) * args.steps * world, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    comm.close()
    dist.destroy_process_group()
