"""GPU tests for the N>1 path (-m gpu, needs >= 2 devices; skipped on a 1-GPU box): the sharded CUDA prover
(sc_prover_create_sharded: NCCL all-gather of the partial sums inside the library) against the unsharded oracle."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as orc
        from sumcheck_b200 import multi
        comm = multi.Comm(multi.broadcast_unique_id(dist, rank), rank, world, rank)
        results = []
        for nv, n_products, m, seed in [(world.bit_length(), 1, 2, 4), (6, 1, 3, 1), (9, 2, 2, 2), (12, 3, 4, 3), (16, 1, 3, 5), (10, 1, 5, 6), (9, 1, 7, 7)]:
            T = n_products * m
            tabs = [orc.synth_table(1 << nv, seed * 100 + j) for j in range(T)]
            coeffs = orc.synth_table(n_products, seed * 100 + 99)
            prods = [(coeffs[k], list(range(k * m, (k + 1) * m))) for k in range(n_products)]
            lo, hi = multi.shard_range(nv, world, rank)
            evals, st = multi.ml_prove_sharded(comm, nv, [np.ascontiguousarray(t[lo:hi]) for t in tabs], prods)
            want, rand, _ = orc.ml_prove(orc.Poly(nv, tabs, prods))
            ok = bool(np.array_equal(evals, want)) and bool(np.array_equal(st.randomness, rand))
            # second proof on the same handle (sub-prover reuse)
            st.reset()
            import sumcheck_b200 as sc
            ev2 = np.zeros_like(evals)
            st.prove_into(sc.Blake2b512Rng.setup(), ev2)
            results.append(ok and bool(np.array_equal(ev2, want)))
            del st
        comm.close()
        q.put((rank, results))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_prover_matches_oracle(world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29700 + world, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    for rank, results in out:
        assert all(results), f"rank {rank}: {results}"
