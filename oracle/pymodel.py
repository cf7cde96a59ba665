"""ORACLE — TEST INFRASTRUCTURE ONLY.  Independent Python big-int model of the reference's prover path.

Second, independently written restatement (plain ``int`` arithmetic mod p, ``hashlib.blake2b``) used to pin the C
oracle (oracle/sumcheck_oracle.c): the two must agree byte for byte.  PARITY UNPINNED against the Rust reference
itself — it cannot be built here and holds no golden vectors.  Pure-Python loops: small cases only.

Field elements are canonical integers in [0, p) inside this module; ``to_mont``/``from_mont`` convert to the
4 x u64 Montgomery limbs that cross the C interfaces.  Citations are relative to /root/reference.
"""
import hashlib

P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
R = (1 << 256) % P
R_INV = pow(R, -1, P)
MASK64 = (1 << 64) - 1


def to_mont_limbs(x):
    m = x * R % P
    return [(m >> (64 * i)) & MASK64 for i in range(4)]


def from_mont_limbs(limbs):
    m = sum(int(l) << (64 * i) for i, l in enumerate(limbs))
    assert m < P
    return m * R_INV % P


def ser_fr(x):
    """ark-serialize of Fr: canonical integer, 32 bytes LE (no flag bits for a 255-bit modulus)."""
    return int(x).to_bytes(32, "little")


def ser_prover_msg(evals):
    """prover.rs:13-17 ProverMsg{evaluations: Vec<F>}: u64 LE length then the elements."""
    return len(evals).to_bytes(8, "little") + b"".join(ser_fr(e) for e in evals)


def ser_poly_info(max_multiplicands, num_variables):
    """data_structures.rs:47-55 PolynomialInfo: two usize as u64 LE, declaration order."""
    return max_multiplicands.to_bytes(8, "little") + num_variables.to_bytes(8, "little")


def ser_proof(msgs):
    """Proof<F> = Vec<ProverMsg<F>> (ml_sumcheck/mod.rs:22)."""
    return len(msgs).to_bytes(8, "little") + b"".join(ser_prover_msg(m) for m in msgs)


class Blake2b512Rng:
    """src/rng.rs:22-81."""

    def __init__(self):  # setup(), rng.rs:30-34
        self.h = hashlib.blake2b(digest_size=64)

    def feed(self, data: bytes):  # rng.rs:36-41, msg already serialised
        self.h.update(data)

    def fill_bytes(self, n):  # rng.rs:61-80
        out = self.h.copy().digest()
        dest = bytearray()
        dp = 0
        while len(dest) < n:
            dest.append(out[dp])
            dp += 1
            if dp == 64:
                self.h.update(out)
                out = self.h.copy().digest()
                dp = 0
        self.h.update(out)
        return bytes(dest)

    def next_u64(self):  # rng.rs:51-55
        return int.from_bytes(self.fill_bytes(8), "little")

    def sample_fr(self):
        """verifier.rs:128-132 -> ark-ff Fp::rand: 4 x next_u64 as limbs, clear top bit, reject >= p; the accepted
        limbs are the Montgomery representation, so the canonical value is limbs * R^-1."""
        while True:
            limbs = [self.next_u64() for _ in range(4)]
            limbs[3] &= MASK64 >> 1
            m = sum(l << (64 * i) for i, l in enumerate(limbs))
            if m < P:
                return m * R_INV % P


def fix_variable(table, r):
    """ark-poly DenseMultilinearExtension::fix_variables(&[r]): new[b] = old[2b] + r (old[2b+1] - old[2b])."""
    return [(table[2 * b] + r * (table[2 * b + 1] - table[2 * b])) % P for b in range(len(table) // 2)]


class Prover:
    """prover.rs:19-153.  products: list of (coefficient, [table indices])."""

    def __init__(self, nv, tables, products):
        if nv == 0:
            raise ValueError("Attempt to prove a constant.")  # prover.rs:50-52
        self.nv = nv
        self.tables = [list(t) for t in tables]
        self.products = [(c, list(ix)) for c, ix in products]
        self.max_multiplicands = max(len(ix) for _, ix in products)
        self.round = 0
        self.randomness = []

    def prove_round(self, r=None):
        if r is not None:
            if self.round == 0:
                raise ValueError("first round should be prover first.")
            self.randomness.append(r)
            self.tables = [fix_variable(t, r) for t in self.tables]
        elif self.round > 0:
            raise ValueError("verifier message is empty")
        self.round += 1
        if self.round > self.nv:
            raise ValueError("Prover is not active")
        d = self.max_multiplicands
        sums = [0] * (d + 1)
        for b in range(1 << (self.nv - self.round)):
            for c, ix in self.products:
                prod = [c] * (d + 1)
                for j in ix:
                    start = self.tables[j][2 * b]
                    step = self.tables[j][2 * b + 1] - start
                    for t in range(d + 1):
                        prod[t] = prod[t] * (start + t * step) % P
                for t in range(d + 1):
                    sums[t] = (sums[t] + prod[t]) % P
        return sums


def ml_prove(nv, tables, products, rng=None):
    """MLSumcheck::prove / prove_as_subprotocol (ml_sumcheck/mod.rs:42-70). Returns (msgs, randomness, final tables)."""
    rng = rng or Blake2b512Rng()
    pr = Prover(nv, tables, products)
    rng.feed(ser_poly_info(pr.max_multiplicands, nv))
    msgs, r = [], None
    for _ in range(nv):
        m = pr.prove_round(r)
        rng.feed(ser_prover_msg(m))
        msgs.append(m)
        r = rng.sample_fr()
    pr.randomness.append(r)
    return msgs, pr.randomness, pr.tables


def interpolate(evals, x):
    """verifier.rs:139-251 — value of the degree<=len-1 interpolant through (i, evals[i]) at x."""
    n = len(evals)
    res = 0
    for i in range(n):
        num, den = 1, 1
        for j in range(n):
            if j != i:
                num = num * (x - j) % P
                den = den * (i - j) % P
        res = (res + evals[i] * num * pow(den, -1, P)) % P
    return res


def check_subclaim(msgs, randomness, claimed):
    """verifier.rs:90-121."""
    expected = claimed
    for m, r in zip(msgs, randomness):
        if (m[0] + m[1]) % P != expected:
            raise ValueError("Prover message is not consistent with the claim.")
        expected = interpolate(m, r)
    return expected


def ml_verify(nv, d, claimed, msgs, rng=None):
    """MLSumcheck::verify_as_subprotocol (mod.rs:84-100). Returns (point, expected_evaluation)."""
    rng = rng or Blake2b512Rng()
    rng.feed(ser_poly_info(d, nv))
    point = []
    for m in msgs:
        rng.feed(ser_prover_msg(m))
        point.append(rng.sample_fr())
    return point, check_subclaim(msgs, point, claimed)


def dense_evaluate(table, point):
    t = list(table)
    for r in point:
        t = fix_variable(t, r)
    return t[0]


def poly_evaluate(tables, products, point):
    """data_structures.rs:99-109."""
    tv = [dense_evaluate(t, point) for t in tables]
    s = 0
    for c, ix in products:
        pr = c
        for j in ix:
            pr = pr * tv[j] % P
        s = (s + pr) % P
    return s


def true_sum(nv, tables, products):
    s = 0
    for b in range(1 << nv):
        for c, ix in products:
            pr = c
            for j in ix:
                pr = pr * tables[j][b] % P
            s = (s + pr) % P
    return s


# ---------------------------------------------------------------- GKR round sumcheck
def eq_table(g):
    """ark-poly precompute_eq: eq[b] = prod_j (g_j if bit j of b else 1 - g_j)."""
    out = [1]
    for gj in g:
        out = [e * (1 - gj) % P for e in out] + [e * gj % P for e in out]
    return out


def sparse_fix_low(f, pt):
    """SparseMultilinearExtension::fix_variables: fixes the low len(pt) index bits. f: dict idx -> value."""
    k = len(pt)
    eq = eq_table(pt)
    out = {}
    for idx, v in f.items():
        ni = idx >> k
        out[ni] = (out.get(ni, 0) + eq[idx & ((1 << k) - 1)] * v) % P
    return out


def gkr_prove(f1, f2, f3, g, rng):
    """gkr_round_sumcheck/mod.rs:93-139. f1: dict over 3*dim vars (g|x|y, LSB first)."""
    dim = len(g)
    f1_g = sparse_fix_low(f1, g)
    h_g = [0] * (1 << dim)
    for xy, v in f1_g.items():
        if v != 0:
            x, y = xy & ((1 << dim) - 1), xy >> dim
            h_g[x] = (h_g[x] + v * f3[y]) % P

    def run(a, b):
        pr = Prover(dim, [a, b], [(1, [0, 1])])
        msgs, ch, r = [], [], None
        for _ in range(dim):
            m = pr.prove_round(r)
            rng.feed(ser_prover_msg(m))
            msgs.append(m)
            r = rng.sample_fr()
            ch.append(r)
        return msgs, ch

    m1, u = run(h_g, f2)
    f1_gu_sparse = sparse_fix_low(f1_g, u)
    f1_gu = [f1_gu_sparse.get(y, 0) for y in range(1 << dim)]
    f2_u = dense_evaluate(f2, u)
    f3_f2u = [f2_u * v % P for v in f3]
    m2, v = run(f1_gu, f3_f2u)
    return m1, m2, u, v


def gkr_verify(dim, m1, m2, claimed, rng):
    """gkr_round_sumcheck/mod.rs:147-192."""
    u = []
    for m in m1:
        rng.feed(ser_prover_msg(m))
        u.append(rng.sample_fr())
    e1 = check_subclaim(m1, u, claimed)
    v = []
    for m in m2:
        rng.feed(ser_prover_msg(m))
        v.append(rng.sample_fr())
    e2 = check_subclaim(m2, v, e1)
    return u, v, e2


def gkr_verify_subclaim(f1, f2, f3, g, u, v, expected):
    """gkr data_structures.rs:33-56."""
    f1e = sparse_fix_low(f1, list(g) + list(u) + list(v)).get(0, 0)
    return f1e * dense_evaluate(f2, u) * dense_evaluate(f3, v) % P == expected


def gkr_sum_naive(f1, f2, f3, g):
    """gkr test.rs:24-45."""
    dim = len(g)
    f1_g = sparse_fix_low(f1, g)
    s = 0
    for xy, val in f1_g.items():
        x, y = xy & ((1 << dim) - 1), xy >> dim
        s = (s + val * f2[x] * f3[y]) % P
    return s
