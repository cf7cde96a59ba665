"""CPU tests (-m "not gpu"): pin the C oracle against the independent Python big-int model, published BLAKE2b
vectors, and the reference's own acceptance relations."""
import hashlib
import random

import numpy as np
import pytest

from oracle import pymodel as pm
from helpers import gkr_arrays, ints, limbs, random_gkr, random_instance, table_limbs, to_poly


def test_field_constants_and_ops(orc):
    assert pm.P.bit_length() == 255
    rnd = random.Random(1)
    edge = [0, 1, 2, pm.P - 1, pm.P - 2, (1 << 254), pm.R % pm.P]
    vals = edge + [rnd.randrange(pm.P) for _ in range(40)]
    for a in vals:
        assert orc.fr_to_int(limbs(a)) == a
        assert orc.fr_to_bytes(limbs(a)) == pm.ser_fr(a)
        assert np.array_equal(orc.fr_from_int(a), limbs(a))
        for b in vals[:12]:
            assert ints(orc.fr_op("add", limbs(a), limbs(b))) == (a + b) % pm.P
            assert ints(orc.fr_op("sub", limbs(a), limbs(b))) == (a - b) % pm.P
            assert ints(orc.fr_op("mul", limbs(a), limbs(b))) == (a * b) % pm.P
    for a in vals[1:]:
        assert ints(orc.fr_op("inv", limbs(a))) == pow(a, -1, pm.P)
    assert ints(orc.fr_from_u64(7)) == 7


def test_blake2b_known_answers(orc):
    # RFC 7693 appendix A: BLAKE2b-512("abc")
    abc = bytes.fromhex(
        "ba80a53f981c4d0d6a2797b69f12f6e94c212f14685ac4b74b12bb6fdbffa2d1"
        "7d87c5392aab792dc252d5de4533cc9518d38aa8dbf1925ab92386edd4009923")
    assert orc.blake2b512(b"abc") == abc
    # empty string (published)
    empty = bytes.fromhex(
        "786a02f742015903c6c6fd852552d272912f4740e15847618a86e217f71f5419"
        "d25e1031afee585313896444934eb04b903a685b1448b755d56f701afe9be2ce")
    assert orc.blake2b512(b"") == empty
    rnd = random.Random(2)
    for n in [1, 63, 64, 65, 127, 128, 129, 255, 256, 257, 1000, 4096]:
        data = bytes(rnd.randrange(256) for _ in range(n))
        assert orc.blake2b512(data) == hashlib.blake2b(data, digest_size=64).digest()


def test_rng_matches_model_and_is_deterministic(orc):
    """rng.rs:113-175 scenario (feed/sample interleaving incl. the 127/128/777-byte unaligned fills), checked
    byte-for-byte against the independent model instead of only run-to-run."""
    rnd = random.Random(3)
    msgs = [bytes(rnd.randrange(256) for _ in range(128)) for _ in range(7)]

    def seq(r, feed, fill, sample):
        out = []
        feed(r, msgs[0]); out.append(sample(r)); out.append(sample(r))
        feed(r, msgs[1]); feed(r, msgs[2]); out.append(sample(r))
        feed(r, msgs[3]); out.append(sample(r)); out.append(sample(r))
        feed(r, msgs[4]); feed(r, msgs[5]); feed(r, msgs[6])
        out.append(sample(r)); out.append(sample(r))
        b1 = fill(r, 127); feed(r, b1); b2 = fill(r, 128); b3 = fill(r, 777)
        assert b2[:64] != b3[:64]
        out.append(sample(r)); feed(r, b3); out.append(sample(r))
        out += [b1, b2, b3, fill(r, 0), fill(r, 64), fill(r, 8)]
        return out

    a = seq(orc.Rng(), lambda r, b: r.feed_bytes(b), lambda r, n: r.fill_bytes(n), lambda r: ints(r.sample_fr()))
    b = seq(pm.Blake2b512Rng(), lambda r, b: r.feed(b), lambda r, n: r.fill_bytes(n), lambda r: r.sample_fr())
    c = seq(orc.Rng(), lambda r, b: r.feed_bytes(b), lambda r, n: r.fill_bytes(n), lambda r: ints(r.sample_fr()))
    assert a == b == c
    assert len(set(x for x in a if isinstance(x, int))) == 9  # "Producing same element" check, rng.rs:149


def test_next_u64_is_hash_chain(orc):
    # each next_u64 = first 8 bytes of H(absorbed so far), then absorb the whole 64-byte digest
    r = orc.Rng()
    r.feed_bytes(b"xyz")
    absorbed = b"xyz"
    for _ in range(4):
        d = hashlib.blake2b(absorbed, digest_size=64).digest()
        assert r.next_u64() == int.from_bytes(d[:8], "little")
        absorbed += d


@pytest.mark.parametrize("nv,n_products,mult_range,shared", [
    (1, 5, (4, 13), False),   # test_trivial_polynomial (test.rs:122-144)
    (5, 5, (4, 9), False),    # test_normal_polynomial shape at a python-sized nv
    (4, 3, (3, 4), False),    # test_extract_sum shape
    (6, 5, (1, 4), True),     # test_shared_reference shape: shared tables, repeats, products shorter than d
    (7, 1, (2, 3), False),    # BASELINE config 1 shape (1 product of degree 2)
    (6, 1, (3, 4), False),    # BASELINE config 2/3 shape (1 product of degree 3)
])
def test_ml_prove_matches_model_and_verifies(orc, nv, n_products, mult_range, shared):
    tables, products = random_instance(100 + nv, nv, n_products, mult_range, shared)
    poly = to_poly(orc, nv, tables, products)
    evals, rand, fin = orc.ml_prove(poly)
    msgs, randomness, final_tables = pm.ml_prove(nv, tables, products)
    assert ints(evals) == msgs
    assert ints(rand) == randomness
    assert ints(fin) == final_tables
    assert orc.serialize_proof(evals) == pm.ser_proof(msgs)
    # the reference's acceptance relations (test.rs:64-75, 206-213)
    asserted = pm.true_sum(nv, tables, products)
    assert (msgs[0][0] + msgs[0][1]) % pm.P == asserted                      # extract_sum
    point, expected = orc.ml_verify(nv, poly.d, limbs(asserted), evals)
    mpoint, mexpected = pm.ml_verify(nv, poly.d, asserted, msgs)
    assert ints(point) == mpoint == randomness                                 # test.rs:119
    assert ints(expected) == mexpected
    assert ints(orc.poly_evaluate(poly, point)) == mexpected == pm.poly_evaluate(tables, products, mpoint)
    # a wrong claim is rejected
    with pytest.raises(orc.OraclePanic):
        orc.ml_verify(nv, poly.d, limbs((asserted + 1) % pm.P), evals)


def test_interactive_rounds_and_state_machine(orc):
    """test_protocol (test.rs:77-97) with an arbitrary challenge source + the panics of prover.rs:50-52,79-81,90-98."""
    nv = 5
    tables, products = random_instance(7, nv, 3, (2, 5))
    poly = to_poly(orc, nv, tables, products)
    pr, mp = orc.Prover(poly), pm.Prover(nv, tables, products)
    rnd = random.Random(9)
    with pytest.raises(orc.OraclePanic) as e:
        orc.Prover(poly).prove_round(limbs(5))
    assert e.value.code == -2
    r = None
    for i in range(nv):
        m = pr.prove_round(None if r is None else limbs(r))
        assert ints(m) == mp.prove_round(r)
        for j in range(len(tables)):
            assert ints(pr.table(j)) == mp.tables[j]
        r = [0, 1, pm.P - 1, rnd.randrange(pm.P), rnd.randrange(pm.P)][i]  # edge challenges 0, 1, p-1
    with pytest.raises(orc.OraclePanic) as e:
        pr.prove_round(None)
    assert e.value.code == -3
    with pytest.raises(orc.OraclePanic) as e:
        pr.prove_round(limbs(3))
    assert e.value.code == -4
    with pytest.raises(orc.OraclePanic) as e:   # zero_polynomial_should_error (test.rs:187-204)
        orc.Prover(orc.Poly(0, [table_limbs([5])], [(limbs(1), [0])]))
    assert e.value.code == -1


def test_transcript_sensitivity(orc):
    """test_normal_polynomial_different_transcripts_fails (test.rs:168-186)."""
    nv = 4
    tables, products = random_instance(11, nv, 2, (2, 4))
    poly = to_poly(orc, nv, tables, products)
    asserted = pm.true_sum(nv, tables, products)
    prng, vrng, bad = orc.Rng(), orc.Rng(), orc.Rng()
    prng.feed_bytes(b"Test Trivial Works"); vrng.feed_bytes(b"Test Trivial Works"); bad.feed_bytes(b"Test Trivial Fails")
    evals, rand, _ = orc.ml_prove(poly, prng)
    point, exp = orc.ml_verify(nv, poly.d, limbs(asserted), evals, vrng)
    assert np.array_equal(point, rand)
    assert np.array_equal(orc.poly_evaluate(poly, point), exp)
    try:
        point2, exp2 = orc.ml_verify(nv, poly.d, limbs(asserted), evals, bad)
        assert not np.array_equal(orc.poly_evaluate(poly, point2), exp2)
    except orc.OraclePanic:
        pass


def test_fix_variable_and_eq(orc):
    rnd = random.Random(5)
    t = [rnd.randrange(pm.P) for _ in range(64)]
    for r in [0, 1, pm.P - 1, rnd.randrange(pm.P)]:
        assert ints(orc.dense_fix_variable(table_limbs(t), limbs(r))) == pm.fix_variable(t, r)
    pt = [rnd.randrange(pm.P) for _ in range(6)]
    assert ints(orc.dense_evaluate(table_limbs(t), table_limbs(pt))) == pm.dense_evaluate(t, pt)
    assert ints(orc.precompute_eq(table_limbs(pt))) == pm.eq_table(pt)
    assert sum(pm.eq_table(pt)) % pm.P == 1


@pytest.mark.parametrize("dim,nnz", [(3, None), (5, None), (4, 3), (6, 200)])
def test_gkr_matches_model_and_verifies(orc, dim, nnz):
    """gkr test.rs:57-88: prove -> verify -> verify_subclaim; extract_sum == naive sum."""
    f1, f2, f3, g = random_gkr(40 + dim, dim, nnz)
    idx, val, a2, a3, ag = gkr_arrays(f1, f2, f3, g)
    m1, m2, u, v = orc.gkr_prove(orc.Rng(), dim, idx, val, a2, a3, ag)
    q1, q2, qu, qv = pm.gkr_prove(f1, f2, f3, g, pm.Blake2b512Rng())
    assert ints(m1) == q1 and ints(m2) == q2 and ints(u) == qu and ints(v) == qv
    claimed = pm.gkr_sum_naive(f1, f2, f3, g)
    assert (q1[0][0] + q1[0][1]) % pm.P == claimed                        # GKRProof::extract_sum
    vu, vv, exp = orc.gkr_verify(orc.Rng(), dim, m1, m2, limbs(claimed))
    pu, pv, pexp = pm.gkr_verify(dim, q1, q2, claimed, pm.Blake2b512Rng())
    assert ints(vu) == pu == qu and ints(vv) == pv == qv and ints(exp) == pexp
    assert orc.gkr_verify_subclaim(dim, idx, val, a2, a3, ag, vu, vv, exp)
    assert pm.gkr_verify_subclaim(f1, f2, f3, g, pu, pv, pexp)
    # phase initialisers individually (mod.rs:22-42, 57-63)
    h_g, gi, gv = orc.gkr_initialize_phase_one(dim, idx, val, a3, ag)
    f1_g = pm.sparse_fix_low(f1, g)
    assert dict(zip([int(i) for i in gi], ints(gv))) == f1_g
    f1_gu = orc.gkr_initialize_phase_two(dim, gi, gv, u)
    sp = pm.sparse_fix_low(f1_g, qu)
    assert ints(f1_gu) == [sp.get(y, 0) for y in range(1 << dim)]


def test_threads_do_not_change_results(orc):
    nv = 12
    rnd = np.random.default_rng(0)
    tabs = [orc.synth_table(1 << nv, 0x5C0000 + j) for j in range(3)]
    poly = orc.Poly(nv, tabs, [(orc.synth_table(1, 77)[0], [0, 1, 2])])
    orc.set_threads(1)
    e1 = orc.ml_prove(poly)
    orc.set_threads(4)
    e4 = orc.ml_prove(poly)
    orc.set_threads(1)
    for a, b in zip(e1, e4):
        assert np.array_equal(a, b)


def test_synth_table_is_reduced_and_counter_based(orc):
    t = orc.synth_table(4096, 123)
    assert all(pm.from_mont_limbs(x) < pm.P for x in t[:64])
    assert np.array_equal(t[:100], orc.synth_table(100, 123))
    assert not np.array_equal(t[:100], orc.synth_table(100, 124))
