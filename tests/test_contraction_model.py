"""CPU model of the tensor-core contraction rounds (csrc/gemm_sum.cuh), in Python integers: the byte-matrix identity the
u8 x u8 -> s32 MMA relies on, the s32 head-room of the accumulators, the anti-diagonal carry of the epilogue, and the host
half (csrc/host_fr.h gemm_finish through sc_fr_contraction_finish) against the big-integer model of prove_round
(prover.rs:110-148).  No GPU."""
import random

import numpy as np

from oracle import pymodel as pm

R = 1 << 256


def mont(x):
    return x * R % pm.P


def sides(tables_pair_values, k):
    """The blocks one side of the split contributes for ONE pair: k = 1: (a, b); k = 2: (a a', b b', (a+b)(a'+b')) as PLAIN
    integers of Montgomery-form table values."""
    if k == 1:
        (a, b), = tables_pair_values
        return [a, b]
    (a, b), (a2, b2) = tables_pair_values
    return [a * a2, b * b2, (a + b) * (a2 + b2)]


def device_model(x_blocks, y_blocks, bx, by):
    """What the kernels compute for one block pair over all pairs: D[u][v] = sum_b x_b[u] * y_b[v] over the BYTES, then the
    anti-diagonal sums E[k] = sum_{u+v=k} D[u][v], then the byte-serial carry into the integer Z = sum_k 2^(8k) E[k]."""
    D = np.zeros((bx, by), dtype=object)
    for x, y in zip(x_blocks, y_blocks):
        xb = np.frombuffer(int(x).to_bytes(bx, "little"), dtype=np.uint8).astype(object)
        yb = np.frombuffer(int(y).to_bytes(by, "little"), dtype=np.uint8).astype(object)
        D += np.outer(xb, yb)
    assert max(int(v) for v in D.flat) < 2 ** 31 or len(x_blocks) > 33025  # the s32 head-room for <= 1024 K-steps of 32 pairs
    E = [0] * (bx + by - 1)
    for u in range(bx):
        for v in range(by):
            E[u + v] += int(D[u][v])
    acc, out = 0, 0
    for k in range(bx + by + 8):
        if k < len(E):
            acc += E[k]
        out |= (acc & 0xFF) << (8 * k)
        acc >>= 8
    assert acc == 0
    return out


def test_byte_matrix_identity_and_carry():
    rnd = random.Random(5)
    for bx, by in [(64, 32), (64, 64)]:
        n = 37
        xs = [rnd.choice([0, 1, (1 << (8 * bx)) - 1, rnd.getrandbits(8 * bx)]) for _ in range(n)]
        ys = [rnd.choice([0, 1, (1 << (8 * by)) - 1, rnd.getrandbits(8 * by)]) for _ in range(n)]
        assert device_model(xs, ys, bx, by) == sum(x * y for x, y in zip(xs, ys))


def test_accumulator_headroom():
    """A CTA adds at most 256 items x 4 K-steps x 32 pairs into one s32 accumulator (gsum::MAX_ITEMS_PER_CTA)."""
    assert 256 * 4 * 32 * 255 * 255 < 2 ** 31


def limbs32(x, n):
    return [(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def test_host_finish_matches_prove_round_model():
    """Random tables, nv = 4; products of 2, 3 and 4 tables: the integers Z[i][j] built from plain products of Montgomery-form
    values (as the device does, here via the byte model for one shape and directly for the others) go through
    sc_fr_contraction_finish and must give the unscaled round polynomial sum_b prod_j ((1-t) a_j + t b_j), t = 0..d."""
    import ctypes as C
    from sumcheck_b200 import capi
    L = capi.lib()
    rnd = random.Random(77)
    nv = 4
    for kx, ky, use_bytes in [(1, 1, False), (2, 1, True), (2, 2, False), (2, 1, False)]:
        m = kx + ky
        for trial in range(3):
            pool = [0, 1, pm.P - 1, pm.P - 2] if trial == 0 else None
            tabs = [[(rnd.choice(pool) if pool else rnd.randrange(pm.P)) for _ in range(1 << nv)] for _ in range(m)]
            pairs = range(1 << (nv - 1))
            xb = [sides([(mont(tabs[j][2 * b]), mont(tabs[j][2 * b + 1])) for j in range(kx)], kx) for b in pairs]
            yb = [sides([(mont(tabs[kx + j][2 * b]), mont(tabs[kx + j][2 * b + 1])) for j in range(ky)], ky) for b in pairs]
            nx, ny = len(xb[0]), len(yb[0])
            bx, by = (32 if kx == 1 else 64), (32 if ky == 1 else 64)
            n_limbs = (bx + by) // 4 + 2
            z = []
            for i in range(nx):
                for j in range(ny):
                    xs, ys = [x[i] for x in xb], [y[j] for y in yb]
                    v = device_model(xs, ys, bx, by) if use_bytes else sum(x * y for x, y in zip(xs, ys))
                    z += limbs32(v, n_limbs)
            zz = np.array(z, dtype=np.uint32)
            out = np.zeros((m + 1, 4), dtype=np.uint64)
            assert L.sc_fr_contraction_finish(zz.ctypes.data_as(C.c_void_p), n_limbs, kx, ky, out.ctypes.data_as(capi.U64P)) == 0
            for t in range(m + 1):
                want = 0
                for b in pairs:
                    term = 1
                    for j in range(m):
                        term = term * ((1 - t) * tabs[j][2 * b] + t * tabs[j][2 * b + 1]) % pm.P
                    want = (want + term) % pm.P
                assert pm.from_mont_limbs(out[t]) == want, (kx, ky, trial, t)
    assert L.sc_fr_contraction_finish(None, 26, 2, 1, None) != 0


# ---- the work split of a persistent CTA (gemm_sum.cuh Split): three producer warps and the compute groups each derive ring slots and
# barrier phases from (n, j, g) alone, so the arithmetic has to describe ONE consistent sequence for every grid / item count
class _Split:
    def __init__(self, items, G, MM, grid, block):
        self.G, self.MM = G, MM
        self.stride, self.first = grid * G, block * G
        last_g = self.first + G - 1
        self.n_min = (items - last_g + self.stride - 1) // self.stride if last_g < items else 0
        self.rem = 0
        for g in range(G - 1):
            f = self.first + g
            ni = (items - f + self.stride - 1) // self.stride if f < items else 0
            self.rem += 1 if ni > self.n_min else 0

    def n_items(self, g):
        return self.n_min + (1 if g < self.rem else 0)

    def steps(self):
        return self.n_min + (1 if self.rem else 0)

    def groups(self, n):
        return self.G if n < self.n_min else self.rem

    def unit(self, n, j, g):
        return self.MM * self.G * n + j * self.groups(n) + g

    def item(self, n, g):
        return self.first + g + n * self.stride


def test_work_split_is_one_consistent_sequence():
    for G, MM in [(3, 3), (2, 4), (3, 2), (1, 2)]:
        for items in list(range(1, 40)) + [148 * G, 148 * G + 1, 1000, 4096]:
            grid = min(148, (items + G - 1) // G)   # gemm.cu grid_for
            seen = set()
            for block in range(grid):
                sp = _Split(items, G, MM, grid, block)
                assert sp.n_items(0) >= 1                      # every launched CTA initialises its accumulators
                walked = []                                    # the order the producer warps walk the units
                for n in range(sp.steps()):
                    for j in range(MM):
                        for g in range(G):
                            if g < sp.groups(n):
                                walked.append((n, j, g))
                # dense unit numbers in walking order = what the compute groups compute for their own (n, j, g)
                assert [sp.unit(*u) for u in walked] == list(range(len(walked)))
                assert len(walked) == MM * sum(sp.n_items(g) for g in range(G))
                for g in range(G):
                    assert all(n < sp.n_items(g) for (n, j, gg) in walked if gg == g)
                    for n in range(sp.n_items(g)):
                        w = sp.item(n, g)
                        assert w < items and w not in seen
                        seen.add(w)
            assert seen == set(range(items))                   # every item exactly once over the grid
