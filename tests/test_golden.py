"""CPU tests: the C oracle against the committed golden fixtures (tests/golden/, made by make_golden.py)."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import pymodel as pm
from helpers import ints, limbs, table_limbs

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.json")))


def load(path):
    with open(path) as f:
        return json.load(f)


def h2i(x):
    return int(x, 16)


@pytest.mark.parametrize("path", [p for p in GOLDEN if load(p)["kind"] == "ml"], ids=os.path.basename)
def test_oracle_ml_golden(orc, path):
    c = load(path)
    tables = [[h2i(v) for v in t] for t in c["tables"]]
    products = [(h2i(cf), ix) for cf, ix in c["products"]]
    poly = orc.Poly(c["nv"], [table_limbs(t) for t in tables], [(limbs(cf), ix) for cf, ix in products])
    rng = orc.Rng()
    if c["pre_feed"]:
        rng.feed_bytes(bytes.fromhex(c["pre_feed"]))
    evals, rand, fin = orc.ml_prove(poly, rng)
    assert orc.serialize_proof(evals).hex() == c["proof_bytes"]
    assert ints(rand) == [h2i(r) for r in c["randomness"]]
    assert ints(fin) == [[h2i(v) for v in t] for t in c["final_tables"]]
    assert (ints(evals[0][0]) + ints(evals[0][1])) % pm.P == h2i(c["sum"])


@pytest.mark.parametrize("path", [p for p in GOLDEN if load(p)["kind"] == "gkr"], ids=os.path.basename)
def test_oracle_gkr_golden(orc, path):
    c = load(path)
    idx = np.array([h2i(i) for i, _ in c["f1"]], dtype=np.uint64)
    val = table_limbs([h2i(v) for _, v in c["f1"]])
    f2, f3, g = (table_limbs([h2i(x) for x in c[k]]) for k in ("f2", "f3", "g"))
    m1, m2, u, v = orc.gkr_prove(orc.Rng(), c["dim"], idx, val, f2, f3, g)
    assert ints(m1) == [[h2i(x) for x in m] for m in c["phase1"]]
    assert ints(m2) == [[h2i(x) for x in m] for m in c["phase2"]]
    assert ints(u) == [h2i(x) for x in c["u"]] and ints(v) == [h2i(x) for x in c["v"]]
