"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (sumcheck_b200 -> libsumcheck_b200.so), against
the CPU oracle on identical inputs, against the committed golden fixtures, and at BASELINE.json's full sizes.
Bar: bit-exact (integer field arithmetic) — every comparison is exact equality of limbs / bytes."""
import glob
import json
import os
import random

import numpy as np
import pytest

import sumcheck_b200 as sc
from oracle import pymodel as pm
from helpers import gkr_arrays, limbs, random_gkr, random_instance, table_limbs

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.json")))


def h2i(x):
    return int(x, 16)


def build_poly(nv, tables, products):
    """tables: list of [2^nv,4] arrays; products: [(coeff[4], [idx])] -> (product ListOfProducts, same for the oracle)."""
    poly = sc.ListOfProductsOfPolynomials.new(nv)
    for c, ix in products:
        poly.add_product([tables[j] for j in ix], c)
    return poly


def both_polys(orc, nv, tables_int, products_int):
    tabs = [table_limbs(t) for t in tables_int]
    prods = [(limbs(c), ix) for c, ix in products_int]
    opoly = orc.Poly(nv, tabs, prods)
    return build_poly(nv, opoly.tables, prods), opoly


def assert_same_proof(orc, poly, opoly, pre_feed=b""):
    rng, orng = sc.Blake2b512Rng.setup(), orc.Rng()
    if pre_feed:
        rng.feed(pre_feed)
        orng.feed_bytes(pre_feed)
    proof, state = sc.MLSumcheck.prove_as_subprotocol(rng, poly)
    evals, rand, fin = orc.ml_prove(opoly, orng)
    got = np.stack([m.evaluations for m in proof])
    assert np.array_equal(got, evals)
    assert sc.MLSumcheck.serialize_proof(proof) == orc.serialize_proof(evals)
    assert np.array_equal(state.randomness, rand)                      # ProverState.randomness (test.rs:119)
    # the 2-entry tables left in ProverState; the product flattens tables in first-use order (data_structures.rs:84-92)
    order = [next(k for k, t in enumerate(opoly.tables) if t is ft) for ft in poly.flattened_ml_extensions]
    for j, t in enumerate(state.flattened_ml_extensions):
        assert np.array_equal(t, fin[order[j]])
    # the transcripts stay in lock-step afterwards
    assert rng.next_u64() == orng.next_u64()
    return got, rand


def test_device_is_b200_class():
    import torch
    assert torch.cuda.is_available()
    assert sc.lib().sc_device_count() >= 1
    assert torch.cuda.get_device_capability(0)[0] >= 10


@pytest.mark.parametrize("path", [p for p in GOLDEN if json.load(open(p))["kind"] == "ml"], ids=os.path.basename)
def test_ml_golden(orc, path):
    c = json.load(open(path))
    tables = [table_limbs([h2i(v) for v in t]) for t in c["tables"]]
    poly = build_poly(c["nv"], tables, [(limbs(h2i(cf)), ix) for cf, ix in c["products"]])
    rng = sc.Blake2b512Rng.setup()
    if c["pre_feed"]:
        rng.feed(bytes.fromhex(c["pre_feed"]))
    proof, state = sc.MLSumcheck.prove_as_subprotocol(rng, poly)
    assert sc.MLSumcheck.serialize_proof(proof).hex() == c["proof_bytes"]
    assert [pm.from_mont_limbs(r) for r in state.randomness] == [h2i(r) for r in c["randomness"]]
    assert pm.from_mont_limbs(sc.MLSumcheck.extract_sum(proof)) == h2i(c["sum"])


@pytest.mark.parametrize("path", [p for p in GOLDEN if json.load(open(p))["kind"] == "gkr"], ids=os.path.basename)
def test_gkr_golden(path):
    c = json.load(open(path))
    dim = c["dim"]
    f1 = sc.SparseMultilinearExtension(3 * dim, np.array([h2i(i) for i, _ in c["f1"]], dtype=np.uint64),
                                       table_limbs([h2i(v) for _, v in c["f1"]]))
    f2, f3, g = (table_limbs([h2i(x) for x in c[k]]) for k in ("f2", "f3", "g"))
    proof, u, v = sc.GKRRoundSumcheck.prove(sc.Blake2b512Rng.setup(), f1, f2, f3, g, return_challenges=True)
    to_int = lambda msgs: [[pm.from_mont_limbs(e) for e in m.evaluations] for m in msgs]
    assert to_int(proof.phase1_sumcheck_msgs) == [[h2i(x) for x in m] for m in c["phase1"]]
    assert to_int(proof.phase2_sumcheck_msgs) == [[h2i(x) for x in m] for m in c["phase2"]]
    assert [pm.from_mont_limbs(x) for x in u] == [h2i(x) for x in c["u"]]
    assert [pm.from_mont_limbs(x) for x in v] == [h2i(x) for x in c["v"]]
    assert pm.from_mont_limbs(proof.extract_sum()) == h2i(c["sum"])


@pytest.mark.parametrize("nv,n_products,mult_range,shared", [
    (1, 5, (4, 13), False),    # test_trivial_polynomial (ml_sumcheck/test.rs:122-144): d up to 12 -> chunked eval points
    (2, 3, (1, 4), False),
    (3, 2, (6, 8), False),     # d+1 in {7,8}: 5+2 / 5+3 launches per round
    (5, 1, (6, 7), False),     # d+1 = 7
    (7, 5, (4, 9), False),     # test_normal_polynomial shape (test.rs:145-167) at a smaller nv
    (8, 3, (3, 4), False),     # test_extract_sum shape (test.rs:206-213)
    (8, 5, (1, 4), True),      # test_shared_reference shape (test.rs:215-269): shared tables, repeats, short products
    (9, 1, (1, 2), False),     # a single table, d = 1
    (12, 1, (2, 3), False),    # BASELINE config 1: nv=12, 1 product of degree 2
    (12, 5, (4, 9), False),    # test_normal_polynomial at its real size
])
def test_ml_prove_matches_oracle(orc, nv, n_products, mult_range, shared):
    tables, products = random_instance(500 + 7 * nv + n_products, nv, n_products, mult_range, shared)
    poly, opoly = both_polys(orc, nv, tables, products)
    assert len(poly.flattened_ml_extensions) == len(opoly.tables)      # de-duplication (test.rs:254,258)
    got, rand = assert_same_proof(orc, poly, opoly, pre_feed=b"Test Trivial Works" if nv % 2 else b"")
    if nv <= 9:
        assert pm.from_mont_limbs(sc.MLSumcheck.extract_sum([sc.ProverMsg(got[0])])) == pm.true_sum(nv, tables, products)


def test_interactive_rounds_state_machine_and_edge_challenges(orc):
    """test_protocol (test.rs:77-97): arbitrary challenges incl. 0, 1, p-1; tables compared after every round; the
    reference's panics (prover.rs:79-81, 90-92, 96-98)."""
    nv = 6
    tables, products = random_instance(77, nv, 3, (2, 5))
    poly, opoly = both_polys(orc, nv, tables, products)
    st, ost = sc.IPForMLSumcheck.prover_init(poly), orc.Prover(opoly)
    with pytest.raises(sc.Panic) as e:
        sc.IPForMLSumcheck.prove_round(sc.IPForMLSumcheck.prover_init(poly), sc.VerifierMsg(limbs(5)))
    assert e.value.code == -2
    rnd = random.Random(9)
    chal = [0, 1, pm.P - 1, 2, rnd.randrange(pm.P), rnd.randrange(pm.P)]
    v_msg = None
    for i in range(nv):
        m = sc.IPForMLSumcheck.prove_round(st, v_msg)
        om = ost.prove_round(None if v_msg is None else v_msg.randomness)
        assert np.array_equal(m.evaluations, om)
        assert st.round == i + 1
        for j, ft in enumerate(poly.flattened_ml_extensions):
            k = next(k for k, t in enumerate(opoly.tables) if t is ft)
            assert np.array_equal(st.table(j), ost.table(k))
        v_msg = sc.VerifierMsg(limbs(chal[i]))
    with pytest.raises(sc.Panic) as e:
        sc.IPForMLSumcheck.prove_round(st, None)
    assert e.value.code == -3
    with pytest.raises(sc.Panic) as e:
        sc.IPForMLSumcheck.prove_round(st, v_msg)
    assert e.value.code == -4


def test_special_values(orc):
    """Tables made of 0, 1, p-1 and coefficient 1 / p-1 (SURVEY §8d edge set)."""
    nv = 5
    rnd = random.Random(3)
    pool = [0, 1, pm.P - 1, 2, pm.P - 2]
    tables = [[rnd.choice(pool) for _ in range(1 << nv)] for _ in range(3)]
    products = [(1, [0, 1, 2]), (pm.P - 1, [2, 2]), (0, [1])]
    poly, opoly = both_polys(orc, nv, tables, products)
    assert_same_proof(orc, poly, opoly)


def test_generic_feedable_rng_path(orc):
    """prove_as_subprotocol with a FeedableRNG that is not the built-in one: per-round C calls, host-side loop."""
    class Wrapped:  # same stream as Blake2b512Rng, but opaque to the library
        def __init__(self): self.inner = sc.Blake2b512Rng.setup(); self.state = self.inner.state
        def feed(self, b): self.inner.feed(b)
    nv = 6
    tables, products = random_instance(31, nv, 2, (2, 4))
    poly, opoly = both_polys(orc, nv, tables, products)
    proof, _ = sc.MLSumcheck.prove_as_subprotocol(Wrapped(), poly)
    evals, _, _ = orc.ml_prove(opoly)
    assert np.array_equal(np.stack([m.evaluations for m in proof]), evals)


@pytest.mark.parametrize("nv,n_products,m,pre", [(3, 1, 3, b""), (11, 1, 3, b""), (12, 1, 2, b"12345678"), (13, 2, 2, b""),
                                                  (14, 1, 5, b"x" * 24), (12, 1, 1, b""), (12, 1, 3, b"Test Trivial Works")])
def test_fused_tail_device_transcript(orc, monkeypatch, nv, n_products, m, pre):
    """SC_TAIL=1: the last rounds run in ONE launch with the Blake2b transcript on the device.  An
    18-byte pre-feed leaves the hash buffer unaligned, which must silently keep the host transcript."""
    monkeypatch.setenv("SC_TAIL", "1")
    T = n_products * m
    tabs = [orc.synth_table(1 << nv, 4200 + 10 * nv + j) for j in range(T)]
    coeffs = orc.synth_table(n_products, 4299 + nv)
    prods = [(coeffs[k], list(range(k * m, (k + 1) * m))) for k in range(n_products)]
    opoly = orc.Poly(nv, tabs, prods)
    assert_same_proof(orc, build_poly(nv, opoly.tables, prods), opoly, pre_feed=pre)


@pytest.mark.parametrize("nv,n_products,mult_range,shared", [(1, 3, (2, 5), False), (7, 4, (1, 6), True), (12, 5, (4, 9), False)])
def test_prove_verify_evaluate_on_device(orc, nv, n_products, mult_range, shared):
    """The reference's own acceptance test (test_polynomial, ml_sumcheck/test.rs:64-75) with every step on the device:
    prove -> verify -> poly.evaluate(subclaim.point) == subclaim.expected_evaluation; each step also equals the oracle."""
    tables, products = random_instance(900 + nv, nv, n_products, mult_range, shared)
    poly, opoly = both_polys(orc, nv, tables, products)
    proof = sc.MLSumcheck.prove(poly)
    asserted = sc.MLSumcheck.extract_sum(proof)
    sub = sc.MLSumcheck.verify(poly.info(), asserted, proof)
    evals = np.stack([m.evaluations for m in proof])
    opoint, oexp = orc.ml_verify(nv, poly.max_multiplicands, asserted, evals)
    assert np.array_equal(sub.point, opoint) and np.array_equal(sub.expected_evaluation, oexp)
    got = poly.evaluate(sub.point)
    assert np.array_equal(got, sub.expected_evaluation)          # "wrong subclaim" check of the reference
    assert np.array_equal(got, orc.poly_evaluate(opoly, opoint))
    wrong = limbs((pm.from_mont_limbs(asserted) + 1) % pm.P)
    with pytest.raises(sc.SumcheckError) as e:                    # Error::Reject (verifier.rs:109-113)
        sc.MLSumcheck.verify(poly.info(), wrong, proof)
    assert e.value.code == -6


# ------------------------------------------------------------------------------------------- resident rounds
@pytest.mark.parametrize("nv,n_products,mult_range,shared,max_pairs", [
    (2, 1, (2, 3), False, None),      # one resident round (round 2, a single pair)
    (9, 1, (3, 4), False, None),      # every fold round resident: 2 CTAs -> 1 CTA
    (13, 1, (3, 4), False, None),     # 32 CTAs at the start, the grid shrinks round by round
    (15, 1, (2, 3), False, None),     # 128 CTAs, one pair per thread
    (17, 1, (3, 4), False, 1 << 16),  # round 2 (2^15 pairs) and round 1 on the launch-per-round kernels, then resident with 2^15... pairs
    (17, 1, (3, 4), False, 1 << 20),  # all fold rounds resident: several pairs per thread (grid-stride loop)
    (12, 5, (1, 6), True, None),      # shared tables, repeats, short products
    (12, 3, (2, 5), False, 128),      # hand-over in the middle: rounds up to 256 pairs by launches, then one resident CTA
    (11, 1, (5, 6), False, None),     # d = 5: five summed points
    (11, 2, (6, 8), False, None),     # d + 1 > 6: no claim shortcut -> no resident rounds (must still agree)
])
def test_resident_rounds(orc, monkeypatch, nv, n_products, mult_range, shared, max_pairs):
    """Rounds served by the resident kernel (csrc/resident_kernel.cuh: ONE cooperative launch for all small rounds, fold
    constants / raw sums exchanged with the host transcript through mapped memory) against the oracle, the launch-per-round
    path (SC_NO_RESIDENT=1), the folded tables left in ProverState, a second proof on the same handle, and the counters."""
    if max_pairs is not None:
        monkeypatch.setenv("SC_RES_MAX_PAIRS", str(max_pairs))
    tables, products = random_instance(9100 + 17 * nv + n_products, nv, n_products, mult_range, shared)
    poly, opoly = both_polys(orc, nv, tables, products)
    got, rand = assert_same_proof(orc, poly, opoly)   # proof bytes, randomness, final 2-entry tables
    import ctypes as C
    st = sc.IPForMLSumcheck.prover_init(poly)
    ev = np.zeros((nv, poly.max_multiplicands + 1, 4), dtype=np.uint64)
    for rep in range(2):
        rng = sc.Blake2b512Rng.setup()
        ev[:] = 0
        assert sc.lib().sc_ml_prove(st._h, C.byref(rng.state), ev.ctypes.data_as(sc.capi.U64P), None) == 0
        assert np.array_equal(ev, got)
        lim = max_pairs if max_pairs is not None else 1 << 16
        want = sum(1 for i in range(2, nv + 1) if (1 << (nv - i)) <= lim) if poly.max_multiplicands <= 5 else 0
        assert st.resident_round_count() == want
        st.reset()
    monkeypatch.setenv("SC_NO_RESIDENT", "1")
    st2 = sc.IPForMLSumcheck.prover_init(poly)
    rng = sc.Blake2b512Rng.setup()
    ev2 = np.zeros_like(ev)
    assert sc.lib().sc_ml_prove(st2._h, C.byref(rng.state), ev2.ctypes.data_as(sc.capi.U64P), None) == 0
    assert st2.resident_round_count() == 0
    assert np.array_equal(ev2, got)


def test_resident_rounds_special_values(orc):
    """0 / 1 / p-1 tables and a zero coefficient through the resident rounds."""
    nv = 10
    rnd = random.Random(12)
    pool = [0, 1, pm.P - 1, 2, pm.P - 2, (1 << 248) - 1, pm.P >> 1]
    tables = [[rnd.choice(pool) for _ in range(1 << nv)] for _ in range(3)]
    products = [(pm.P - 1, [0, 1, 2]), (1, [2, 2]), (0, [1])]
    poly, opoly = both_polys(orc, nv, tables, products)
    assert_same_proof(orc, poly, opoly)


def test_resident_abandoned_proof_does_not_hang(orc):
    """A handle destroyed / reset while nothing is in flight, and many handles in a row (recycled pinned blocks carry other
    handles' sequence numbers — they are cleared at creation): proofs stay correct."""
    nv = 8
    for k in range(6):
        tabs = [orc.synth_table(1 << nv, 300 + 10 * k + j) for j in range(2)]
        prods = [(orc.synth_table(1, 399 + k)[0], [0, 1])]
        poly = build_poly(nv, tabs, prods)
        proof = sc.MLSumcheck.prove(poly)
        want = orc.ml_prove(orc.Poly(nv, tabs, prods))[0]
        assert np.array_equal(np.stack([m.evaluations for m in proof]), want)


# ------------------------------------------------------------------------------------------- tensor-core fold rounds
@pytest.mark.parametrize("nv,n_products,mult_range,shared", [
    (9, 1, (3, 4), False),     # smallest shape with a 128-pair fold round (round 2 of nv=9): one tile, one CTA
    (10, 1, (1, 2), False),    # single multiplicand
    (11, 1, (2, 3), False),
    (12, 1, (5, 6), False),    # d = 5: five summed points per launch
    (12, 5, (1, 6), True),     # shared tables, repeats inside a product, short products: tiles staged once per use
    (13, 3, (2, 5), False),    # several products, different lengths
    (12, 2, (6, 8), False),    # d+1 > 6: no claim shortcut, rounds stay on the plain kernels (must still agree)
])
def test_tensor_core_fold_rounds(orc, monkeypatch, nv, n_products, mult_range, shared):
    """Fold rounds on the TMA + tcgen05.mma kernel (csrc/tc_round.cuh), forced down to 128-pair rounds with
    SC_TC_MIN_PAIRS so that small shapes cover one tile per CTA, several tiles per CTA and grid > tiles.  Same bytes as
    the oracle, the same folded tables, and identical to the plain kernels (SC_NO_TC=1)."""
    monkeypatch.setenv("SC_TC_MIN_PAIRS", "128")
    monkeypatch.setenv("SC_RES_MAX_PAIRS", "64")   # the resident kernel would otherwise take every round of these small shapes
    tables, products = random_instance(7000 + 13 * nv + n_products, nv, n_products, mult_range, shared)
    poly, opoly = both_polys(orc, nv, tables, products)
    got, rand = assert_same_proof(orc, poly, opoly)
    st = sc.IPForMLSumcheck.prover_init(poly)
    import ctypes as C
    ev = np.zeros((nv, poly.max_multiplicands + 1, 4), dtype=np.uint64)
    rng = sc.Blake2b512Rng.setup()
    assert sc.lib().sc_ml_prove(st._h, C.byref(rng.state), ev.ctypes.data_as(sc.capi.U64P), None) == 0
    assert np.array_equal(ev, got)
    want_tc = max(0, nv - 8) if poly.max_multiplicands <= 5 else 0   # rounds 2 .. nv-7 have >= 128 output pairs
    assert st.tc_round_count() == want_tc
    monkeypatch.setenv("SC_NO_TC", "1")
    st2 = sc.IPForMLSumcheck.prover_init(poly)
    rng = sc.Blake2b512Rng.setup()
    ev2 = np.zeros_like(ev)
    assert sc.lib().sc_ml_prove(st2._h, C.byref(rng.state), ev2.ctypes.data_as(sc.capi.U64P), None) == 0
    assert st2.tc_round_count() == 0
    assert np.array_equal(ev2, ev)
    # the message finished on the device (publish_round: coefficient, claim, canonical forms) instead of on the host
    monkeypatch.delenv("SC_NO_TC")
    monkeypatch.setenv("SC_NO_HOST_POST", "1")
    st3 = sc.IPForMLSumcheck.prover_init(poly)
    rng = sc.Blake2b512Rng.setup()
    ev3 = np.zeros_like(ev)
    assert sc.lib().sc_ml_prove(st3._h, C.byref(rng.state), ev3.ctypes.data_as(sc.capi.U64P), None) == 0
    assert st3.tc_round_count() == 0   # the TMA / tensor-core kernels sum at the alternative points, which only the host converts back
    assert np.array_equal(ev3, ev)


def test_tensor_core_fold_special_values_and_edge_challenges(orc, monkeypatch):
    """0 / 1 / p-1 tables (all-0x00 and near-all-0xff bytes in the u8 operand) and challenges 0, 1, p-1 (constants
    matrix of zeros / of 2^(8k) mod p) through the interactive API; folded tables compared after every round."""
    monkeypatch.setenv("SC_TC_MIN_PAIRS", "128")
    nv = 10
    rnd = random.Random(11)
    pool = [0, 1, pm.P - 1, 2, pm.P - 2, (1 << 248) - 1, pm.P >> 1]
    tables = [[rnd.choice(pool) for _ in range(1 << nv)] for _ in range(3)]
    products = [(pm.P - 1, [0, 1, 2]), (1, [2, 2]), (0, [1])]
    poly, opoly = both_polys(orc, nv, tables, products)
    st, ost = sc.IPForMLSumcheck.prover_init(poly), orc.Prover(opoly)
    chal = [0, 1, pm.P - 1, rnd.randrange(pm.P)]
    v_msg = None
    for i in range(5):
        m = sc.IPForMLSumcheck.prove_round(st, v_msg)
        om = ost.prove_round(None if v_msg is None else v_msg.randomness)
        assert np.array_equal(m.evaluations, om)
        for j, ft in enumerate(poly.flattened_ml_extensions):
            k = next(k for k, t in enumerate(opoly.tables) if t is ft)
            assert np.array_equal(st.table(j), ost.table(k))
        v_msg = sc.VerifierMsg(limbs(chal[i % len(chal)]))
    assert st.tc_round_count() == 2   # rounds 2 and 3 of nv=10 (256 and 128 output pairs)


@pytest.mark.parametrize("nv,n_products,shared,max_tiles,m", [
    (8, 1, False, None, 3),     # one tile: round 1 only
    (9, 1, False, None, 3),     # round 1 = two tiles, round 2 = one tile (fold)
    (11, 1, True, None, 3),     # one product over a shared pool: repeated tables inside the product
    (13, 1, False, None, 3),    # more tiles than one group handles at once
    (12, 3, False, None, 3),    # three products of three fresh tables each: coefficients pre-scaled into tables
    (13, 2, False, 3, 3),       # a round split over several launches (the s32 accumulators' head-room), 3 tiles per launch
    (12, 4, True, None, 3),     # shared tables between products: nothing to pre-scale, must fall back and still agree
    (8, 1, False, None, 4),     # products of FOUR tables: both operands of the contraction are plain products (192 x 192)
    (10, 1, True, None, 4),
    (13, 1, False, None, 4),
    (12, 4, False, None, 4),    # the shape of BASELINE config 4: four products of four fresh tables
    (13, 2, False, 5, 4),
    (8, 1, False, None, 2),     # products of TWO tables (GKR phases): no per-pair multiplication, round 1 is TMA + MMA only
    (12, 1, False, None, 2),
    (11, 1, True, None, 2),     # the same table twice
    (13, 3, False, None, 2),
    (13, 2, False, 3, 2),
])
def test_contraction_rounds(orc, monkeypatch, nv, n_products, shared, max_tiles, m):
    """Products of three (or four) tables on the tensor-core contraction kernels (csrc/gemm_sum.cuh): plain products per pair, the
    sum over the pairs as a u8 x u8 -> s32 tcgen05.mma, six (nine) big integers to the host.  Same bytes as the oracle, the same
    folded tables (assert_same_proof), and identical to the previous kernels (SC_NO_GEMM=1)."""
    import ctypes as C
    monkeypatch.setenv("SC_TC_MIN_PAIRS", "128")
    monkeypatch.setenv("SC_RES_MAX_PAIRS", "64")
    if max_tiles:
        monkeypatch.setenv("SC_GEMM_MAX_TILES", str(max_tiles))
    tables, products = random_instance(9100 + 17 * nv + n_products + 1000 * m, nv, n_products, (m, m + 1), shared)
    poly, opoly = both_polys(orc, nv, tables, products)
    got, rand = assert_same_proof(orc, poly, opoly)
    st = sc.IPForMLSumcheck.prover_init(poly)
    ev = np.zeros((nv, m + 1, 4), dtype=np.uint64)
    rng = sc.Blake2b512Rng.setup()
    assert sc.lib().sc_ml_prove(st._h, C.byref(rng.state), ev.ctypes.data_as(sc.capi.U64P), None) == 0
    assert np.array_equal(ev, got)
    if not (shared and n_products > 1):
        assert st.gemm_round_count() == nv - 7   # rounds 1 .. nv-7 have >= 128 pairs
        assert st.tc_round_count() == nv - 8     # ... of which the fold rounds also count as tensor-core fold rounds
    monkeypatch.setenv("SC_NO_GEMM", "1")
    st2 = sc.IPForMLSumcheck.prover_init(poly)
    rng = sc.Blake2b512Rng.setup()
    ev2 = np.zeros_like(ev)
    assert sc.lib().sc_ml_prove(st2._h, C.byref(rng.state), ev2.ctypes.data_as(sc.capi.U64P), None) == 0
    assert st2.gemm_round_count() == 0
    assert np.array_equal(ev2, ev)


@pytest.mark.parametrize("m", [2, 3, 4])
def test_contraction_rounds_special_values_and_edge_challenges(orc, monkeypatch, m):
    """0 / 1 / p-1 / all-0xff-byte tables and the challenges 0, 1, p-1 through the interactive API (sc_prove_round), folded tables
    compared with the oracle after every round: extreme bytes in both MMA operands, extreme carries in the plain products."""
    monkeypatch.setenv("SC_TC_MIN_PAIRS", "128")
    nv = 11
    rnd = random.Random(12)
    pool = [0, 1, pm.P - 1, 2, pm.P - 2, (1 << 248) - 1, pm.P >> 1]
    tables = [[rnd.choice(pool) for _ in range(1 << nv)] for _ in range(m)]
    tables[m - 1] = [pm.P - 1] * (1 << nv)   # Montgomery form of p-1 and near-maximal products everywhere
    products = [(pm.P - 1, list(range(m)))]
    poly, opoly = both_polys(orc, nv, tables, products)
    st, ost = sc.IPForMLSumcheck.prover_init(poly), orc.Prover(opoly)
    chal = [0, 1, pm.P - 1, rnd.randrange(pm.P)]
    v_msg = None
    for i in range(6):
        m = sc.IPForMLSumcheck.prove_round(st, v_msg)
        om = ost.prove_round(None if v_msg is None else v_msg.randomness)
        assert np.array_equal(m.evaluations, om)
        for j, ft in enumerate(poly.flattened_ml_extensions):
            k = next(k for k, t in enumerate(opoly.tables) if t is ft)
            assert np.array_equal(st.table(j), ost.table(k))
        v_msg = sc.VerifierMsg(limbs(chal[i % len(chal)]))
    assert st.gemm_round_count() == 4   # rounds 1..4 of nv=11 (1024 .. 128 pairs)


@pytest.mark.parametrize("nv,n_products,m", [(12, 1, 3), (13, 2, 3), (12, 1, 4), (12, 1, 2)])
def test_fold_rounds_launched_ahead_of_their_challenge(orc, monkeypatch, nv, n_products, m):
    """Inside a whole-proof call the next large fold round is launched right behind the current one and receives its challenge
    through mapped memory (gemm_prelaunch).  Same proof as the oracle's, as with SC_NO_PRELAUNCH=1, as through caller-driven rounds
    (which never launch ahead), and the same number of launches; a second proof on the handle after reset."""
    import ctypes as C
    monkeypatch.setenv("SC_TC_MIN_PAIRS", "128")
    monkeypatch.setenv("SC_RES_MAX_PAIRS", "64")
    tables, products = random_instance(9900 + nv + m, nv, n_products, (m, m + 1), False)
    poly, opoly = both_polys(orc, nv, tables, products)
    want = orc.ml_prove(opoly)[0]
    st = sc.IPForMLSumcheck.prover_init(poly)
    ev = np.zeros((nv, m + 1, 4), dtype=np.uint64)
    for rep in range(2):
        st.reset()
        ev[:] = 0
        st.prove_into(sc.Blake2b512Rng.setup(), ev)
        assert np.array_equal(ev, want)
        assert st.gemm_round_count() == nv - 7
    launches = st.launch_count()
    # caller-driven rounds on the same handle
    st.reset()
    ost = orc.Prover(opoly)
    v = None
    for i in range(nv):
        msg = sc.IPForMLSumcheck.prove_round(st, v)
        assert np.array_equal(msg.evaluations, ost.prove_round(None if v is None else v.randomness))
        v = sc.VerifierMsg(limbs(7 + i))
    monkeypatch.setenv("SC_NO_PRELAUNCH", "1")   # read once per process: only effective if nothing launched ahead before
    st2 = sc.IPForMLSumcheck.prover_init(poly)
    ev2 = np.zeros_like(ev)
    st2.prove_into(sc.Blake2b512Rng.setup(), ev2)
    assert np.array_equal(ev2, want) and st2.launch_count() == launches


@pytest.mark.parametrize("n_products,m", [(1, 3), (1, 4), (2, 3), (1, 2)])
def test_contraction_rounds_on_borrowed_device_tables(orc, monkeypatch, n_products, m):
    """sc_prover_create_device: the caller's tables already live in HBM and are borrowed, never written.  One product: the
    contraction kernels read them through descriptors built on the caller's pointers; several products: nothing can be
    pre-scaled in place (the tables are not ours), so the rounds fall back — both must agree with the oracle, twice (reset)."""
    import ctypes as C
    import torch
    monkeypatch.setenv("SC_TC_MIN_PAIRS", "128")
    monkeypatch.setenv("SC_RES_MAX_PAIRS", "64")
    nv = 11
    tables, products = random_instance(12000 + n_products + 10 * m, nv, n_products, (m, m + 1), False)
    poly, opoly = both_polys(orc, nv, tables, products)
    want = orc.ml_prove(opoly)[0]
    tabs = [np.ascontiguousarray(t) for t in opoly.tables]
    dev = [torch.from_numpy(t.view(np.int64)).cuda() for t in tabs]
    before = [d.clone() for d in dev]
    ptrs = (C.c_void_p * len(dev))(*[d.data_ptr() for d in dev])
    coeffs = np.ascontiguousarray(np.stack([limbs(c) for c, _ in products]))
    offsets = np.array([0] + list(np.cumsum([len(ix) for _, ix in products])), dtype=np.uint32)
    indices = np.array([j for _, ix in products for j in ix], dtype=np.uint32)
    h = C.c_void_p()
    L = sc.lib()
    assert L.sc_prover_create_device(C.byref(h), nv, len(dev), ptrs, n_products, coeffs.ctypes.data_as(sc.capi.U64P),
                                     offsets.ctypes.data_as(sc.capi.U32P), indices.ctypes.data_as(sc.capi.U32P), 0) == 0
    try:
        for rep in range(2):
            assert L.sc_prover_reset(h) == 0
            ev = np.zeros((nv, m + 1, 4), dtype=np.uint64)
            rng = sc.Blake2b512Rng.setup()
            assert L.sc_ml_prove(h, C.byref(rng.state), ev.ctypes.data_as(sc.capi.U64P), None) == 0
            assert np.array_equal(ev, want)
        assert (L.sc_prover_gemm_round_count(h) == nv - 7) == (n_products == 1)
        torch.cuda.synchronize()
        for d, b in zip(dev, before):
            assert torch.equal(d, b)   # borrowed tables are never written
    finally:
        L.sc_prover_destroy(h)


def test_reset_reproves_identically(orc):
    nv = 10
    tabs = [orc.synth_table(1 << nv, 900 + j) for j in range(3)]
    prods = [(orc.synth_table(1, 77)[0], [0, 1, 2])]
    poly = build_poly(nv, tabs, prods)
    st = sc.IPForMLSumcheck.prover_init(poly)
    import ctypes as C
    outs = []
    for _ in range(3):
        rng = sc.Blake2b512Rng.setup()
        ev = np.zeros((nv, 4, 4), dtype=np.uint64)
        assert sc.lib().sc_ml_prove(st._h, C.byref(rng.state), ev.ctypes.data_as(sc.capi.U64P), None) == 0
        outs.append(ev)
        st.reset()
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    assert np.array_equal(outs[0], orc.ml_prove(orc.Poly(nv, tabs, prods))[0])


def test_prescaled_coefficients_reload_and_export(orc):
    """Several products: prover_init multiplies each coefficient into a table only that product uses (csrc/sumcheck.cu
    prescale_tables).  The proof, the tables ProverState exposes after every round (coefficient divided out again) and a
    re-proof after sc_prover_load_tables must all equal the oracle's; shared tables and a zero coefficient keep the
    in-kernel scaling."""
    nv = 9
    rnd = random.Random(21)
    tables = [[rnd.randrange(pm.P) for _ in range(1 << nv)] for _ in range(6)]
    # product 0: private tables 0,1; product 1: private table 2 + shared 5; product 2: only shared tables (5, 5);
    # product 3: zero coefficient with a private table 3; product 4: single private multiplicand 4
    products = [(rnd.randrange(1, pm.P), [0, 1]), (rnd.randrange(1, pm.P), [5, 2]), (rnd.randrange(1, pm.P), [5, 5]), (0, [3, 5]),
                (pm.P - 1, [4])]
    poly, opoly = both_polys(orc, nv, tables, products)
    got, _ = assert_same_proof(orc, poly, opoly)
    st, ost = sc.IPForMLSumcheck.prover_init(poly), orc.Prover(opoly)
    v_msg = None
    for i in range(4):
        m = sc.IPForMLSumcheck.prove_round(st, v_msg)
        assert np.array_equal(m.evaluations, ost.prove_round(None if v_msg is None else v_msg.randomness))
        for j, ft in enumerate(poly.flattened_ml_extensions):
            k = next(k for k, t in enumerate(opoly.tables) if t is ft)
            assert np.array_equal(st.table(j), ost.table(k))
        v_msg = sc.VerifierMsg(limbs(rnd.randrange(pm.P)))
    import ctypes as C
    st.load_tables(poly.flattened_ml_extensions)      # fresh upload of the caller's (unscaled) tables: scaled again
    ev = np.zeros_like(got)
    rng = sc.Blake2b512Rng.setup()
    assert sc.lib().sc_ml_prove(st._h, C.byref(rng.state), ev.ctypes.data_as(sc.capi.U64P), None) == 0
    assert np.array_equal(ev, got)


def test_allocation_cache_reuse_and_release(orc):
    """Handles hand their slab / pinned block / stream to the next handle: stale contents must never leak into a proof."""
    L = sc.lib()
    for rep in range(3):
        for nv in (7, 10, 7):
            tabs = [orc.synth_table(1 << nv, 9100 + 17 * rep + j) for j in range(3)]
            prods = [(orc.synth_table(1, 9200 + rep)[0], [0, 1, 2])]
            proof = sc.MLSumcheck.prove(build_poly(nv, tabs, prods))      # create -> prove -> destroy
            assert np.array_equal(np.stack([m.evaluations for m in proof]), orc.ml_prove(orc.Poly(nv, tabs, prods))[0])
        L.sc_release_cached_memory()


def test_pipelined_upload_first_round(orc):
    """sc_prover_load_tables at nv >= 21: chunked H2D on a copy stream with round 1 summed chunk by chunk behind it.  The
    proof after the upload, a proof after reset (round 1 recomputed normally) and the oracle must agree — on the NEW tables."""
    import ctypes as C
    nv = 21
    old = [orc.synth_table(1 << nv, 31000 + j) for j in range(3)]
    new = [orc.synth_table(1 << nv, 32000 + j) for j in range(3)]
    coeff = orc.synth_table(1, 33000)[0]
    st = sc.IPForMLSumcheck.prover_init(build_poly(nv, old, [(coeff, [0, 1, 2])]))
    orc.set_threads(os.cpu_count() or 1)
    try:
        want = orc.ml_prove(orc.Poly(nv, new, [(coeff, [0, 1, 2])]))[0]
    finally:
        orc.set_threads(1)
    for rep in range(2):
        st.load_tables(new)
        ev = np.zeros((nv, 4, 4), dtype=np.uint64)
        st.prove_into(sc.Blake2b512Rng.setup(), ev)
        assert np.array_equal(ev, want)
        per_round = sum(1 for i in range(2, nv + 1) if (1 << (nv - i)) > (1 << 16))   # fold rounds with a launch of their own
        assert st.launch_count() == 8 + per_round + 1   # 8 chunk launches of round 1, the large fold rounds, ONE resident launch
    st.reset()
    ev2 = np.zeros_like(ev)
    st.prove_into(sc.Blake2b512Rng.setup(), ev2)
    assert np.array_equal(ev2, want) and st.launch_count() == 1 + per_round + 1


# ------------------------------------------------------------------------------------------- BASELINE.json sizes
def synth_poly(orc, cfg, nv, n_products, m):
    """SURVEY §8d synthetic inputs: table j of config c uses seed 0x5C0000 + 0x100*c + j, coefficients 0x5C00FF + 0x100*c."""
    T = n_products * m
    tabs = [orc.synth_table(1 << nv, 0x5C0000 + 0x100 * cfg + j) for j in range(T)]
    coeffs = orc.synth_table(n_products, 0x5C00FF + 0x100 * cfg)
    prods = [(coeffs[k], list(range(k * m, (k + 1) * m))) for k in range(n_products)]
    return tabs, prods


@pytest.mark.parametrize("cfg,nv,n_products,m", [
    (1, 12, 1, 2),   # config 1 (bit-exact check)
    (2, 20, 1, 3),   # config 2: nv=20 deg 3
    (4, 18, 4, 4),   # config 4 shape at nv=18 (full nv=22 below)
])
def test_baseline_configs_bit_exact(orc, cfg, nv, n_products, m):
    tabs, prods = synth_poly(orc, cfg, nv, n_products, m)
    orc.set_threads(os.cpu_count() or 1)
    try:
        opoly = orc.Poly(nv, tabs, prods)
        assert_same_proof(orc, build_poly(nv, opoly.tables, prods), opoly)
    finally:
        orc.set_threads(1)


@pytest.mark.parametrize("cfg,nv,n_products,m", [
    (3, 24, 1, 3),   # config 3 per-GPU shape at G=1: nv=24 deg 3 (1.5 GiB of tables)
    (4, 22, 4, 4),   # config 4: nv=22, 4 products of degree 4 (2 GiB of tables)
])
def test_full_size_configs(orc, cfg, nv, n_products, m):
    """Full BASELINE sizes: bit-exact against the multi-threaded oracle AND the size-independent relations the
    reference's own tests use (verify accepts; poly.evaluate(point) == expected_evaluation, test.rs:64-75)."""
    tabs, prods = synth_poly(orc, cfg, nv, n_products, m)
    poly, opoly = build_poly(nv, tabs, prods), orc.Poly(nv, tabs, prods)
    proof = sc.MLSumcheck.prove(poly)
    got = np.stack([pm_.evaluations for pm_ in proof])
    claimed = sc.MLSumcheck.extract_sum(proof)
    point, expected = orc.ml_verify(nv, m, claimed, got)               # oracle verifier accepts the GPU proof
    assert np.array_equal(orc.poly_evaluate(opoly, point), expected)   # subclaim holds on the real polynomial
    orc.set_threads(os.cpu_count() or 1)
    try:
        evals, _, _ = orc.ml_prove(opoly)
    finally:
        orc.set_threads(1)
    assert np.array_equal(got, evals)
    st = sc.IPForMLSumcheck.prover_init(poly)
    for i in range(3):
        sc.IPForMLSumcheck.prove_round(st, None if i == 0 else sc.VerifierMsg(limbs(3 + i)))
    assert st.tc_round_count() == 2   # the large fold rounds run on the TMA + tensor-core kernel


# ------------------------------------------------------------------------------------------- GKR
@pytest.mark.parametrize("dim,nnz", [(1, None), (2, 3), (3, None), (6, None), (9, None), (7, 5), (10, 3000)])
def test_gkr_matches_oracle(orc, dim, nnz):
    """gkr test.rs:71-88 (test_small dim=9, test_extract dim=6) + sparse/dense extremes."""
    f1, f2, f3, g = random_gkr(60 + dim, dim, nnz)
    idx, val, a2, a3, ag = gkr_arrays(f1, f2, f3, g)
    f1s = sc.SparseMultilinearExtension(3 * dim, idx, val)
    proof, u, v = sc.GKRRoundSumcheck.prove(sc.Blake2b512Rng.setup(), f1s, a2, a3, ag, return_challenges=True)
    m1, m2, ou, ov = orc.gkr_prove(orc.Rng(), dim, idx, val, a2, a3, ag)
    assert np.array_equal(np.stack([m.evaluations for m in proof.phase1_sumcheck_msgs]), m1)
    assert np.array_equal(np.stack([m.evaluations for m in proof.phase2_sumcheck_msgs]), m2)
    assert np.array_equal(u, ou) and np.array_equal(v, ov)
    claimed = proof.extract_sum()
    if dim <= 7:
        assert pm.from_mont_limbs(claimed) == pm.gkr_sum_naive(f1, f2, f3, g)       # test_extract
    vu, vv, exp = orc.gkr_verify(orc.Rng(), dim, m1, m2, claimed)                  # verify accepts
    assert orc.gkr_verify_subclaim(dim, idx, val, a2, a3, ag, vu, vv, exp)         # verify_subclaim true


def test_gkr_phase_initialisers(orc):
    """initialize_phase_one / initialize_phase_two / start_phase{1,2}_sumcheck as free functions (mod.rs:22-82)."""
    dim = 7
    f1, f2, f3, g = random_gkr(99, dim, 300)
    # force collisions: several nonzeros sharing (x, y) but differing in z
    f1 = dict(f1)
    keys = list(f1.keys())[:40]
    for k in keys:
        f1[(k & ~((1 << dim) - 1)) | ((k + 1) & ((1 << dim) - 1))] = f1[k]
    idx, val, a2, a3, ag = gkr_arrays(f1, f2, f3, g)
    f1s = sc.SparseMultilinearExtension(3 * dim, idx, val)
    h_g, f1_g = sc.initialize_phase_one(f1s, a3, ag)
    oh, oi, ov = orc.gkr_initialize_phase_one(dim, idx, val, a3, ag)
    assert np.array_equal(h_g, oh)
    assert np.array_equal(f1_g.indices, oi) and np.array_equal(f1_g.values, ov)     # merged, BTreeMap order
    u = table_limbs([random.Random(5).randrange(pm.P) for _ in range(dim)])
    f1_gu = sc.initialize_phase_two(f1_g, u)
    assert np.array_equal(f1_gu, orc.gkr_initialize_phase_two(dim, oi, ov, u))
    # start_phase1: first round message of 1*(h_g*f2)
    st = sc.start_phase1_sumcheck(h_g, a2)
    m = sc.IPForMLSumcheck.prove_round(st, None)
    op = orc.Prover(orc.Poly(dim, [oh, a2], [(sc.api.FR_ONE, [0, 1])]))
    assert np.array_equal(m.evaluations, op.prove_round())
    # start_phase2: tables are (f1_gu, f2_u * f3)
    f2_u = orc.dense_evaluate(a2, u)
    st2 = sc.start_phase2_sumcheck(f1_gu, a3, f2_u)
    scaled = np.stack([orc.fr_op("mul", f2_u, x) for x in a3])
    assert np.array_equal(st2.table(1), scaled)
    m2 = sc.IPForMLSumcheck.prove_round(st2, None)
    op2 = orc.Prover(orc.Poly(dim, [f1_gu, scaled], [(sc.api.FR_ONE, [0, 1])]))
    assert np.array_equal(m2.evaluations, op2.prove_round())


def test_gkr_config5_dim18(orc):
    """BASELINE config 5: GKRRoundSumcheck prove, dim=18, f1 with 2^18 nonzeros over 54 variables."""
    dim = 18
    n = 1 << dim
    f2, f3 = orc.synth_table(n, 0x5C0500), orc.synth_table(n, 0x5C0501)
    g = orc.synth_table(dim, 0x5C0502)
    val = orc.synth_table(n, 0x5C0503)
    rng = np.random.default_rng(0x5C0504)
    idx = np.unique(rng.integers(0, 1 << (3 * dim), size=n + 4096, dtype=np.uint64))[:n].copy()
    rng.shuffle(idx)
    f1s = sc.SparseMultilinearExtension(3 * dim, idx, val[:idx.shape[0]].copy())
    proof, u, v = sc.GKRRoundSumcheck.prove(sc.Blake2b512Rng.setup(), f1s, f2, f3, g, return_challenges=True)
    m1, m2, ou, ov = orc.gkr_prove(orc.Rng(), dim, f1s.indices, f1s.values, f2, f3, g)
    assert np.array_equal(np.stack([m.evaluations for m in proof.phase1_sumcheck_msgs]), m1)
    assert np.array_equal(np.stack([m.evaluations for m in proof.phase2_sumcheck_msgs]), m2)
    vu, vv, exp = orc.gkr_verify(orc.Rng(), dim, m1, m2, proof.extract_sum())
    assert orc.gkr_verify_subclaim(dim, f1s.indices, f1s.values, f2, f3, g, vu, vv, exp)


@pytest.mark.parametrize("dim,L", [(4, 3), (9, 5), (13, 4)])
def test_gkr_batch_matches_separate_proofs(orc, dim, L):
    """sc_gkr_prove_batch: L layers in one call (rounds of all layers issued before the first is collected) == L separate
    oracle proofs, bit for bit, including the transcripts afterwards; different sparsities per layer."""
    rnd = random.Random(1234 + dim)
    f1s, f2s, f3s, gs, want = [], [], [], [], []
    for l in range(L):
        f1, f2, f3, g = random_gkr(500 + 10 * dim + l, dim, nnz=rnd.choice([1, 1 << (dim - 1), 1 << dim]))
        idx, val, f2a, f3a, ga = gkr_arrays(f1, f2, f3, g)
        f1s.append(sc.SparseMultilinearExtension(3 * dim, idx, val)); f2s.append(f2a); f3s.append(f3a); gs.append(ga)
        org = orc.Rng()
        org.feed_bytes(bytes([l]))
        m1, m2, _, _ = orc.gkr_prove(org, dim, idx, val, f2a, f3a, ga)
        want.append((m1, m2, org.next_u64()))
    rngs = []
    for l in range(L):
        r = sc.Blake2b512Rng.setup()
        r.feed(bytes([l]))
        rngs.append(r)
    proofs = sc.GKRRoundSumcheck.prove_batch(rngs, f1s, f2s, f3s, gs)
    for l in range(L):
        assert np.array_equal(np.stack([m.evaluations for m in proofs[l].phase1_sumcheck_msgs]), want[l][0]), f"layer {l} phase 1"
        assert np.array_equal(np.stack([m.evaluations for m in proofs[l].phase2_sumcheck_msgs]), want[l][1]), f"layer {l} phase 2"
        assert rngs[l].next_u64() == want[l][2]


def test_gkr_generic_feedable_rng(orc):
    """GKRRoundSumcheck::prove<R: FeedableRNG> with an rng the library cannot see into (gkr mod.rs:93): the reference's loop
    runs on the host side of the mirror, every step (phase initialisers, prove_round, f2.evaluate(u)) on the device."""
    class Wrapped:  # same stream as Blake2b512Rng, but opaque to the library
        def __init__(self): self.inner = sc.Blake2b512Rng.setup(); self.state = self.inner.state
        def feed(self, b): self.inner.feed(b)
    for dim, nnz in [(3, 5), (7, 1 << 7)]:
        f1, f2, f3, g = random_gkr(900 + dim, dim, nnz=nnz)
        idx, val, f2a, f3a, ga = gkr_arrays(f1, f2, f3, g)
        proof, u, v = sc.GKRRoundSumcheck.prove(Wrapped(), sc.SparseMultilinearExtension(3 * dim, idx, val), f2a, f3a, ga, return_challenges=True)
        m1, m2, ou, ov = orc.gkr_prove(orc.Rng(), dim, idx, val, f2a, f3a, ga)
        assert np.array_equal(np.stack([m.evaluations for m in proof.phase1_sumcheck_msgs]), m1)
        assert np.array_equal(np.stack([m.evaluations for m in proof.phase2_sumcheck_msgs]), m2)
        assert np.array_equal(u, ou) and np.array_equal(v, ov)
