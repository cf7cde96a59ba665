"""Shared test helpers: random instances in both representations (python ints for the big-int model, Montgomery
limb arrays for the C interfaces)."""
import random

import numpy as np

from oracle import pymodel as pm


def limbs(x):
    return np.array(pm.to_mont_limbs(x), dtype=np.uint64)


def table_limbs(vals):
    return np.array([pm.to_mont_limbs(v) for v in vals], dtype=np.uint64).reshape(len(vals), 4)


def ints(arr):
    """[..., 4] Montgomery limb array -> nested list of canonical ints"""
    a = np.asarray(arr, dtype=np.uint64)
    if a.ndim == 1:
        return pm.from_mont_limbs(a)
    return [ints(x) for x in a]


def random_instance(seed, nv, n_products, mult_range, shared=False, n_shared_tables=5):
    """Mirrors ml_sumcheck/test.rs:44-62 random_list_of_products (fresh tables per product), or with shared=True the
    test_shared_reference shape (test.rs:215-252): products index a small pool, repeats inside a product allowed."""
    rnd = random.Random(seed)
    tables, products = [], []
    if shared:
        tables = [[rnd.randrange(pm.P) for _ in range(1 << nv)] for _ in range(n_shared_tables)]
        for _ in range(n_products):
            m = rnd.randrange(mult_range[0], mult_range[1])
            products.append((rnd.randrange(pm.P), [rnd.randrange(n_shared_tables) for _ in range(m)]))
        used = sorted({j for _, ix in products for j in ix})  # flattened list only holds tables that were added
        remap = {j: i for i, j in enumerate(used)}
        tables = [tables[j] for j in used]
        products = [(c, [remap[j] for j in ix]) for c, ix in products]
    else:
        for _ in range(n_products):
            m = rnd.randrange(mult_range[0], mult_range[1])
            ix = []
            for _ in range(m):
                tables.append([rnd.randrange(pm.P) for _ in range(1 << nv)])
                ix.append(len(tables) - 1)
            products.append((rnd.randrange(pm.P), ix))
    return tables, products


def to_poly(orc, nv, tables, products):
    return orc.Poly(nv, [table_limbs(t) for t in tables], [(limbs(c), ix) for c, ix in products])


def random_gkr(seed, dim, nnz=None):
    rnd = random.Random(seed)
    nnz = (1 << dim) if nnz is None else nnz
    idxs = set()
    while len(idxs) < nnz:
        idxs.add(rnd.randrange(1 << (3 * dim)))
    f1 = {i: rnd.randrange(1, pm.P) for i in sorted(idxs)}
    f2 = [rnd.randrange(pm.P) for _ in range(1 << dim)]
    f3 = [rnd.randrange(pm.P) for _ in range(1 << dim)]
    g = [rnd.randrange(pm.P) for _ in range(dim)]
    return f1, f2, f3, g


def gkr_arrays(f1, f2, f3, g):
    idx = np.array(list(f1.keys()), dtype=np.uint64)
    val = table_limbs(list(f1.values()))
    return idx, val, table_limbs(f2), table_limbs(f3), table_limbs(g)
