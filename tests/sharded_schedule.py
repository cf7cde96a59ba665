"""Test infrastructure: the multi-GPU schedule of SURVEY §8e / csrc/capi_multi.inc restated over torch.distributed with
the ORACLE as the per-rank compute, so the sharding maths (high-bit partition, partial sums add, tail gather) can be
checked on CPU with gloo.  Mirrors sharded_round() step for step."""
import numpy as np
import torch
import torch.distributed as dist

from oracle import oracle as orc
from oracle import pymodel as pm
from sumcheck_b200.multi import shard_range


def fr_sum(elems):
    """sum mod p of Montgomery-limb elements (addition is linear in Montgomery form)"""
    x = sum(sum(int(e[i]) << (64 * i) for i in range(4)) for e in elems) % pm.P
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def all_gather_u64(arr):
    t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int64))
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return [o.numpy().view(np.uint64) for o in outs]


def sharded_prove(nv, tables, products):
    """Every rank passes the FULL tables (test convenience) and uses only its shard. Returns evals[nv, d+1, 4]."""
    rank, world = dist.get_rank(), dist.get_world_size()
    g = world.bit_length() - 1
    nv_l = nv - g
    lo, hi = shard_range(nv, world, rank)
    shard = [np.ascontiguousarray(t[lo:hi]) for t in tables]
    local = orc.Prover(orc.Poly(nv_l, shard, products))
    d = local.d
    rng = orc.Rng()
    rng.feed_poly_info(d, nv)                      # GLOBAL PolynomialInfo (ml_sumcheck/mod.rs:54)
    evals, r, sub = [], None, None
    for i in range(1, nv + 1):
        if i <= nv_l:                              # sharded rounds: local partial sums, all-gather, sum mod p
            part = local.prove_round(r)
            parts = all_gather_u64(part)
            msg = np.stack([fr_sum([pp[t] for pp in parts]) for t in range(d + 1)])
        else:
            if i == nv_l + 1:                      # tail: fold the last pair on r, gather one element per rank and table
                folded = np.stack([orc.dense_fix_variable(local.table(j), r)[0] for j in range(len(shard))])
                gathered = all_gather_u64(folded)  # [rank][table]
                sub_tabs = [np.stack([gathered[q][j] for q in range(world)]) for j in range(len(shard))]
                sub = orc.Prover(orc.Poly(g, sub_tabs, products))
                msg = sub.prove_round(None)
            else:
                msg = sub.prove_round(r)
        rng.feed_prover_msg(msg)
        evals.append(msg)
        r = rng.sample_fr()
    return np.stack(evals)
