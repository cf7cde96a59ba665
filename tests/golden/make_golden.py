"""Generates tests/golden/*.json from the independent Python big-int model (oracle/pymodel.py).

The reference holds no golden vectors and cannot be built here (Rust; no rustc), so these
fixtures are produced by the model, NOT by the reference: they pin the C oracle and the CUDA path to a frozen
answer so that a later change to either is caught.  Run from the repo root:  python tests/golden/make_golden.py
Inputs are stored explicitly (hex of canonical integers) so the fixtures do not depend on any RNG implementation.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pymodel as pm  # noqa: E402
from helpers import random_gkr, random_instance  # noqa: E402

hx = lambda v: format(v, "x")


def ml_case(name, seed, nv, n_products, mult_range, shared=False, pre_feed=b""):
    tables, products = random_instance(seed, nv, n_products, mult_range, shared)
    rng = pm.Blake2b512Rng()
    if pre_feed:
        rng.feed(pre_feed)
    msgs, randomness, final_tables = pm.ml_prove(nv, tables, products, rng)
    return {
        "name": name, "kind": "ml", "nv": nv, "pre_feed": pre_feed.hex(),
        "tables": [[hx(v) for v in t] for t in tables],
        "products": [[hx(c), ix] for c, ix in products],
        "msgs": [[hx(v) for v in m] for m in msgs],
        "randomness": [hx(r) for r in randomness],
        "final_tables": [[hx(v) for v in t] for t in final_tables],
        "proof_bytes": pm.ser_proof(msgs).hex(),
        "sum": hx(pm.true_sum(nv, tables, products)),
    }


def gkr_case(name, seed, dim, nnz=None):
    f1, f2, f3, g = random_gkr(seed, dim, nnz)
    m1, m2, u, v = pm.gkr_prove(f1, f2, f3, g, pm.Blake2b512Rng())
    return {
        "name": name, "kind": "gkr", "dim": dim,
        "f1": [[hx(i), hx(val)] for i, val in f1.items()],
        "f2": [hx(x) for x in f2], "f3": [hx(x) for x in f3], "g": [hx(x) for x in g],
        "phase1": [[hx(x) for x in m] for m in m1], "phase2": [[hx(x) for x in m] for m in m2],
        "u": [hx(x) for x in u], "v": [hx(x) for x in v],
        "sum": hx(pm.gkr_sum_naive(f1, f2, f3, g)),
    }


CASES = [
    lambda: ml_case("cfg1_shape_nv6_deg2", 1001, 6, 1, (2, 3)),
    lambda: ml_case("cfg2_shape_nv6_deg3", 1002, 6, 1, (3, 4)),
    lambda: ml_case("cfg4_shape_nv5_4x4", 1003, 5, 4, (4, 5)),
    lambda: ml_case("shared_tables_nv5", 1004, 5, 5, (1, 4), shared=True),
    lambda: ml_case("trivial_nv1_wide", 1005, 1, 5, (4, 13)),
    lambda: ml_case("subprotocol_prefed_nv4", 1006, 4, 2, (2, 5), pre_feed=b"Test Trivial Works"),
    lambda: gkr_case("gkr_dim4", 2001, 4),
    lambda: gkr_case("gkr_dim5_sparse", 2002, 5, 11),
]

if __name__ == "__main__":
    for mk in CASES:
        c = mk()
        with open(os.path.join(HERE, c["name"] + ".json"), "w") as f:
            json.dump(c, f, separators=(",", ":"))
        print("wrote", c["name"])
