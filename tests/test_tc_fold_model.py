"""CPU model of the tensor-core fold (sumcheck_b200/csrc/tc_fold.cuh): the algebra the kernel relies on, checked with
Python integers so that it is pinned without a GPU.
  * fix_variables as a byte-matrix product: new == sum_k x_k * C_k (mod p) over the 64 bytes of a pair, and the column
    sums S_j = sum_k x_k * C_k[j] stay below 2^22 (s32 accumulators, u8 x u8 products);
  * columns_to_fr: the carry of the byte-spaced columns into 32-bit limbs and the 16-bit Barrett step
    q = floor(floor(V / 2^238) * floor(2^270 / p) / 2^32), for which V - q*p must lie in [0, 2p) for EVERY V < 64*255*p."""
import random

from oracle import pymodel as pm

P = pm.P
MU_270 = (1 << 270) // P
M32 = (1 << 32) - 1


def constants(r):
    """C_k, k = 0..63, as in tcf::build_bmat: (1-r)*2^(8k) for the bytes of old[2b], r*2^(8(k-32)) for old[2b+1]."""
    return [((1 - r) % P) * (1 << (8 * k)) % P for k in range(32)] + [r * (1 << (8 * k)) % P for k in range(32)]


def column_sums(pair_bytes, C):
    return [sum(pair_bytes[k] * ((C[k] >> (8 * j)) & 0xFF) for k in range(64)) for j in range(32)]


def columns_to_fr(S):
    """Line-by-line model of tcf::columns_to_fr with explicit 32/64-bit wrap-around."""
    limbs, hi = [], 0
    for i in range(8):
        t0 = (S[4 * i] + (S[4 * i + 1] << 8)) & M32
        t1 = (S[4 * i + 2] + (S[4 * i + 3] << 8)) & M32
        assert t0 == S[4 * i] + (S[4 * i + 1] << 8) and t1 == S[4 * i + 2] + (S[4 * i + 3] << 8)   # < 2^31: no wrap
        w = t0 + (t1 << 16) + hi
        assert w < 1 << 64
        limbs.append(w & M32)
        hi = w >> 32
    l8 = hi
    assert l8 < 1 << 14
    x = ((limbs[7] >> 14) | (l8 << 18)) & M32
    assert x == (sum(l << (32 * i) for i, l in enumerate(limbs)) + (l8 << 256)) >> 238   # floor(V / 2^238) fits 32 bits
    q = (x * MU_270) >> 32
    V = sum(l << (32 * i) for i, l in enumerate(limbs)) + (l8 << 256)
    rem = V - q * P
    assert 0 <= rem < 2 * P, "Barrett bound violated"
    assert rem < 1 << 256                                                                        # limb 8 of the difference is 0
    return rem - P if rem >= P else rem


def test_mu_constant():
    assert MU_270 == 0x8D54


def test_fold_is_a_byte_matrix_product():
    rnd = random.Random(5)
    for r in [0, 1, P - 1, rnd.randrange(P), rnd.randrange(P)]:
        C = constants(r)
        for _ in range(20):
            a, b = rnd.choice([0, 1, P - 1, rnd.randrange(P)]), rnd.choice([0, 1, P - 1, rnd.randrange(P)])
            xb = list(a.to_bytes(32, "little")) + list(b.to_bytes(32, "little"))
            S = column_sums(xb, C)
            assert max(S) < 1 << 22
            V = sum(s << (8 * j) for j, s in enumerate(S))
            assert V < 64 * 255 * P
            want = (a + r * (b - a)) % P
            assert V % P == want
            assert columns_to_fr(S) == want


def test_barrett_bound_on_extreme_columns():
    rnd = random.Random(6)
    smax = 64 * 255 * 255
    # every column at its maximum exceeds 64*255*p (real constants are < p, so their top bytes are small); the bound only
    # has to hold for V < 64*255*p, so build extreme V below that limit in column form
    for _ in range(2000):
        V = rnd.choice([64 * 255 * P - 1 - rnd.randrange(1 << 200), rnd.randrange(64 * 255 * P), P * rnd.randrange(1, 16320) - 1,
                        P * rnd.randrange(1, 16320)])
        # spread V over byte columns with random (valid) carries pushed down into the columns
        S = [(V >> (8 * j)) & 0xFF for j in range(31)] + [V >> 248]
        for j in range(31):
            give = min(S[j + 1], rnd.randrange(0, 1 << 13))
            if S[j] + 256 * give <= smax:
                S[j + 1] -= give
                S[j] += 256 * give
        assert all(0 <= s <= smax for s in S[:31]) and S[31] < 1 << 22
        assert sum(s << (8 * j) for j, s in enumerate(S)) == V
        assert columns_to_fr(S) == V % P
