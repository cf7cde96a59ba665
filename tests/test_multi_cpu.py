"""CPU tests for the N>1 path (gloo, world_size 2): shard partition, id broadcast plumbing, and the sharded schedule
(tests/sharded_schedule.py, oracle as compute) reproducing the unsharded proof bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_by_high_bits():
    from sumcheck_b200.multi import shard_range
    for nv, world in [(4, 2), (6, 4), (10, 8), (3, 1)]:
        seen = []
        for r in range(world):
            lo, hi = shard_range(nv, world, r)
            assert lo % 2 == 0 and (hi - lo) == (1 << nv) // world      # pairs (2b, 2b+1) never straddle ranks
            assert (lo >> (nv - (world.bit_length() - 1))) == r or world == 1
            seen += list(range(lo, hi))
        assert seen == list(range(1 << nv))
    with pytest.raises(AssertionError):
        shard_range(4, 3, 0)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch
        from oracle import oracle as orc
        from sharded_schedule import sharded_prove
        # id broadcast plumbing (the 128-byte communicator id travels over torch.distributed)
        buf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = torch.arange(128, dtype=torch.uint8)
        dist.broadcast(buf, src=0)
        assert buf.tolist() == list(range(128))
        results = []
        for nv, n_products, m, seed in [(5, 1, 3, 1), (6, 2, 2, 2), (4, 3, 4, 3), (world.bit_length(), 1, 2, 4)]:
            T = n_products * m
            tabs = [orc.synth_table(1 << nv, seed * 100 + j) for j in range(T)]
            coeffs = orc.synth_table(n_products, seed * 100 + 99)
            prods = [(coeffs[k], list(range(k * m, (k + 1) * m))) for k in range(n_products)]
            got = sharded_prove(nv, tabs, prods)
            want, _, _ = orc.ml_prove(orc.Poly(nv, tabs, prods))
            results.append(bool(np.array_equal(got, want)))
        q.put((rank, results))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_schedule_matches_unsharded_oracle(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    for rank, results in out:
        assert all(results), f"rank {rank}: {results}"
