"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/liboracle.so (the CPU restatement).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Field elements are numpy uint64 arrays whose last axis is the 4 Montgomery limbs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
U64P = C.POINTER(C.c_uint64)
U32P = C.POINTER(C.c_uint32)
U8P = C.POINTER(C.c_uint8)

ERR_NAMES = {0: "ok", -1: "Attempt to prove a constant.", -2: "first round should be prover first.",
             -3: "verifier message is empty", -4: "Prover is not active", -5: "bad input", -6: "Reject"}


class OraclePanic(Exception):
    """The reference would panic!/Err here (code = ORC_ERR_*)."""

    def __init__(self, code):
        super().__init__(ERR_NAMES.get(code, str(code)))
        self.code = code


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_rng_setup.restype = C.c_void_p
        L.orc_rng_clone.restype = C.c_void_p
        L.orc_rng_clone.argtypes = [C.c_void_p]
        L.orc_rng_free.argtypes = [C.c_void_p]
        L.orc_rng_feed_bytes.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.orc_rng_feed_poly_info.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.orc_rng_feed_prover_msg.argtypes = [C.c_void_p, U64P, C.c_size_t]
        L.orc_rng_fill_bytes.argtypes = [C.c_void_p, U8P, C.c_size_t]
        L.orc_rng_next_u64.restype = C.c_uint64
        L.orc_rng_next_u64.argtypes = [C.c_void_p]
        L.orc_rng_sample_fr.argtypes = [C.c_void_p, U64P]
        L.orc_prover_init.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32, C.POINTER(U64P), C.c_uint32, U64P,
                                      U32P, U32P]
        L.orc_prove_round.argtypes = [C.c_void_p, U64P, U64P]
        L.orc_prover_degree.argtypes = [C.c_void_p]
        L.orc_prover_degree.restype = C.c_uint32
        L.orc_prover_table_len.argtypes = [C.c_void_p]
        L.orc_prover_table_len.restype = C.c_uint64
        L.orc_prover_copy_table.argtypes = [C.c_void_p, C.c_uint32, U64P]
        L.orc_prover_free.argtypes = [C.c_void_p]
        L.orc_ml_prove.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(U64P), C.c_uint32, U64P, U32P, U32P,
                                   U64P, U64P, U64P]
        L.orc_serialize_proof.argtypes = [U64P, C.c_uint32, C.c_uint32, U8P]
        L.orc_serialize_proof.restype = C.c_size_t
        L.orc_ml_verify.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, U64P, U64P, U64P, U64P]
        L.orc_interpolate_uni_poly.argtypes = [U64P, U64P, C.c_uint32, U64P]
        L.orc_poly_evaluate.argtypes = [U64P, C.c_uint32, C.c_uint32, C.POINTER(U64P), C.c_uint32, U64P, U32P, U32P,
                                        U64P]
        L.orc_dense_fix_variable.argtypes = [U64P, U64P, C.c_uint64, U64P]
        L.orc_dense_evaluate.argtypes = [U64P, U64P, C.c_uint32, U64P]
        L.orc_precompute_eq.argtypes = [U64P, U64P, C.c_uint32]
        L.orc_gkr_initialize_phase_one.argtypes = [C.c_uint32, C.c_size_t, U64P, U64P, U64P, U64P, U64P, U64P, U64P]
        L.orc_gkr_initialize_phase_one.restype = C.c_size_t
        L.orc_gkr_initialize_phase_two.argtypes = [C.c_uint32, C.c_size_t, U64P, U64P, U64P, U64P]
        L.orc_gkr_prove.argtypes = [C.c_void_p, C.c_uint32, C.c_size_t, U64P, U64P, U64P, U64P, U64P, U64P, U64P,
                                    U64P, U64P]
        L.orc_gkr_verify.argtypes = [C.c_void_p, C.c_uint32, U64P, U64P, U64P, U64P, U64P, U64P]
        L.orc_gkr_verify_subclaim.argtypes = [C.c_uint32, C.c_size_t, U64P, U64P, U64P, U64P, U64P, U64P, U64P, U64P]
        L.orc_synth_table.argtypes = [U64P, C.c_uint64, C.c_uint64]
        L.orc_blake2b512.argtypes = [C.c_char_p, C.c_size_t, U8P]
        L.orc_set_threads.argtypes = [C.c_int]
        for f in ("orc_fr_add", "orc_fr_sub", "orc_fr_mul"):
            getattr(L, f).argtypes = [U64P, U64P, U64P]
        L.orc_fr_inv.argtypes = [U64P, U64P]
        L.orc_fr_from_u64.argtypes = [U64P, C.c_uint64]
        L.orc_fr_from_canonical.argtypes = [U64P, U64P]
        L.orc_fr_to_canonical.argtypes = [U64P, U64P]
        L.orc_fr_to_bytes.argtypes = [U8P, U64P]
        _LIB = L
    return _LIB


def p64(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(U64P)


def p32(a):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(U32P)


def _fr():
    return np.zeros(4, dtype=np.uint64)


def set_threads(n):
    lib().orc_set_threads(int(n))


def fr_op(name, a, b=None):
    out = _fr()
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if b is None:
        getattr(lib(), "orc_fr_" + name)(p64(out), p64(a))
    else:
        getattr(lib(), "orc_fr_" + name)(p64(out), p64(a), p64(np.ascontiguousarray(b, dtype=np.uint64)))
    return out


def fr_from_u64(v):
    out = _fr()
    lib().orc_fr_from_u64(p64(out), v)
    return out


def fr_from_int(x):
    """canonical python int -> Montgomery limbs"""
    c = np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
    return fr_op("from_canonical", c)


def fr_to_int(a):
    """Montgomery limbs -> canonical python int"""
    c = fr_op("to_canonical", a)
    return sum(int(c[i]) << (64 * i) for i in range(4))


def fr_to_bytes(a):
    out = np.zeros(32, dtype=np.uint8)
    lib().orc_fr_to_bytes(out.ctypes.data_as(U8P), p64(np.ascontiguousarray(a, dtype=np.uint64)))
    return out.tobytes()


def blake2b512(data: bytes):
    out = np.zeros(64, dtype=np.uint8)
    lib().orc_blake2b512(data, len(data), out.ctypes.data_as(U8P))
    return out.tobytes()


def synth_table(n_elems, seed):
    out = np.zeros((n_elems, 4), dtype=np.uint64)
    lib().orc_synth_table(p64(out), n_elems, seed)
    return out


class Rng:
    """Blake2b512Rng (src/rng.rs:22-81)."""

    def __init__(self, handle=None):
        self.h = handle if handle is not None else lib().orc_rng_setup()

    def clone(self):
        return Rng(lib().orc_rng_clone(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_rng_free(self.h)
            self.h = None

    def feed_bytes(self, b: bytes):
        lib().orc_rng_feed_bytes(self.h, b, len(b))

    def feed_poly_info(self, max_multiplicands, nv):
        lib().orc_rng_feed_poly_info(self.h, max_multiplicands, nv)

    def feed_prover_msg(self, evals):
        evals = np.ascontiguousarray(evals, dtype=np.uint64)
        lib().orc_rng_feed_prover_msg(self.h, p64(evals), evals.shape[0])

    def fill_bytes(self, n):
        out = np.zeros(max(n, 1), dtype=np.uint8)
        lib().orc_rng_fill_bytes(self.h, out.ctypes.data_as(U8P), n)
        return out[:n].tobytes()

    def next_u64(self):
        return int(lib().orc_rng_next_u64(self.h))

    def sample_fr(self):
        out = _fr()
        lib().orc_rng_sample_fr(self.h, p64(out))
        return out


class Poly:
    """ListOfProductsOfPolynomials flattened to what crosses the C ABI (data_structures.rs:25-35)."""

    def __init__(self, nv, tables, products):
        self.nv = nv
        self.tables = [np.ascontiguousarray(t, dtype=np.uint64) for t in tables]
        for t in self.tables:
            assert t.shape == (1 << nv, 4)
        self.products = products
        self.coeffs = np.ascontiguousarray(np.stack([np.asarray(c, dtype=np.uint64) for c, _ in products])
                                           if products else np.zeros((0, 4), dtype=np.uint64))
        offs, idx = [0], []
        for _, ix in products:
            idx.extend(ix)
            offs.append(len(idx))
        self.offsets = np.array(offs, dtype=np.uint32)
        self.indices = np.array(idx if idx else [0], dtype=np.uint32)
        self.d = max((len(ix) for _, ix in products), default=0)
        self._tabs = (U64P * max(len(self.tables), 1))(*[p64(t) for t in self.tables])

    def cargs(self):
        return (C.c_uint32(self.nv), C.c_uint32(len(self.tables)), self._tabs, C.c_uint32(len(self.products)),
                p64(self.coeffs) if len(self.products) else None, p32(self.offsets), p32(self.indices))


class Prover:
    """IPForMLSumcheck::{prover_init, prove_round} (prover.rs:49,74)."""

    def __init__(self, poly: Poly):
        self.poly = poly
        h = C.c_void_p()
        rc = lib().orc_prover_init(C.byref(h), *poly.cargs())
        if rc:
            raise OraclePanic(rc)
        self.h = h
        self.d = lib().orc_prover_degree(h)

    def prove_round(self, r=None):
        out = np.zeros((self.d + 1, 4), dtype=np.uint64)
        rp = p64(np.ascontiguousarray(r, dtype=np.uint64)) if r is not None else None
        rc = lib().orc_prove_round(self.h, rp, p64(out))
        if rc:
            raise OraclePanic(rc)
        return out

    def table(self, j):
        n = lib().orc_prover_table_len(self.h)
        out = np.zeros((n, 4), dtype=np.uint64)
        lib().orc_prover_copy_table(self.h, j, p64(out))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_prover_free(self.h)
            self.h = None


def ml_prove(poly: Poly, rng: Rng = None):
    """MLSumcheck::prove (rng None) / prove_as_subprotocol. Returns (evals[nv,d+1,4], randomness[nv,4], final[T,2,4])."""
    nv, d, T = poly.nv, poly.d, len(poly.tables)
    evals = np.zeros((max(nv, 1), d + 1, 4), dtype=np.uint64)
    rand = np.zeros((max(nv, 1), 4), dtype=np.uint64)
    fin = np.zeros((max(T, 1), 2, 4), dtype=np.uint64)
    rc = lib().orc_ml_prove(rng.h if rng else None, *poly.cargs(), p64(evals), p64(rand), p64(fin))
    if rc:
        raise OraclePanic(rc)
    return evals[:nv], rand[:nv], fin[:T]


def serialize_proof(evals):
    nv, dp1 = evals.shape[0], evals.shape[1]
    out = np.zeros(8 + nv * (8 + 32 * dp1), dtype=np.uint8)
    n = lib().orc_serialize_proof(p64(np.ascontiguousarray(evals)), nv, dp1 - 1, out.ctypes.data_as(U8P))
    assert n == out.size
    return out.tobytes()


def ml_verify(nv, d, claimed_sum, evals, rng: Rng = None):
    """MLSumcheck::verify / verify_as_subprotocol -> (point[nv,4], expected[4])."""
    point = np.zeros((max(nv, 1), 4), dtype=np.uint64)
    exp = _fr()
    rc = lib().orc_ml_verify(rng.h if rng else None, nv, d, p64(np.ascontiguousarray(claimed_sum, dtype=np.uint64)),
                             p64(np.ascontiguousarray(evals)), p64(point), p64(exp))
    if rc:
        raise OraclePanic(rc)
    return point[:nv], exp


def poly_evaluate(poly: Poly, point):
    out = _fr()
    lib().orc_poly_evaluate(p64(out), *poly.cargs(), p64(np.ascontiguousarray(point, dtype=np.uint64)))
    return out


def dense_fix_variable(table, r):
    table = np.ascontiguousarray(table, dtype=np.uint64)
    out = np.zeros((table.shape[0] // 2, 4), dtype=np.uint64)
    lib().orc_dense_fix_variable(p64(out), p64(table), table.shape[0], p64(np.ascontiguousarray(r, dtype=np.uint64)))
    return out


def dense_evaluate(table, point):
    table = np.ascontiguousarray(table, dtype=np.uint64)
    nv = int(table.shape[0]).bit_length() - 1
    out = _fr()
    lib().orc_dense_evaluate(p64(out), p64(table), nv, p64(np.ascontiguousarray(point, dtype=np.uint64)))
    return out


def precompute_eq(g):
    g = np.ascontiguousarray(g, dtype=np.uint64)
    out = np.zeros((1 << g.shape[0], 4), dtype=np.uint64)
    lib().orc_precompute_eq(p64(out), p64(g), g.shape[0])
    return out


def gkr_initialize_phase_one(dim, f1_idx, f1_val, f3, g):
    nnz = len(f1_idx)
    h_g = np.zeros((1 << dim, 4), dtype=np.uint64)
    gi = np.zeros(max(nnz, 1), dtype=np.uint64)
    gv = np.zeros((max(nnz, 1), 4), dtype=np.uint64)
    n = lib().orc_gkr_initialize_phase_one(dim, nnz, p64(f1_idx), p64(f1_val), p64(f3), p64(g), p64(h_g), p64(gi), p64(gv))
    return h_g, gi[:n].copy(), gv[:n].copy()


def gkr_initialize_phase_two(dim, f1g_idx, f1g_val, u):
    out = np.zeros((1 << dim, 4), dtype=np.uint64)
    lib().orc_gkr_initialize_phase_two(dim, len(f1g_idx), p64(np.ascontiguousarray(f1g_idx)),
                                       p64(np.ascontiguousarray(f1g_val)), p64(np.ascontiguousarray(u)), p64(out))
    return out


def gkr_prove(rng: Rng, dim, f1_idx, f1_val, f2, f3, g):
    """GKRRoundSumcheck::prove -> (phase1[dim,3,4], phase2[dim,3,4], u[dim,4], v[dim,4])."""
    m1 = np.zeros((dim, 3, 4), dtype=np.uint64)
    m2 = np.zeros((dim, 3, 4), dtype=np.uint64)
    u = np.zeros((dim, 4), dtype=np.uint64)
    v = np.zeros((dim, 4), dtype=np.uint64)
    rc = lib().orc_gkr_prove(rng.h, dim, len(f1_idx), p64(f1_idx), p64(f1_val), p64(f2), p64(f3), p64(g), p64(m1),
                             p64(m2), p64(u), p64(v))
    if rc:
        raise OraclePanic(rc)
    return m1, m2, u, v


def gkr_verify(rng: Rng, dim, m1, m2, claimed_sum):
    u = np.zeros((dim, 4), dtype=np.uint64)
    v = np.zeros((dim, 4), dtype=np.uint64)
    exp = _fr()
    rc = lib().orc_gkr_verify(rng.h, dim, p64(np.ascontiguousarray(m1)), p64(np.ascontiguousarray(m2)),
                              p64(np.ascontiguousarray(claimed_sum, dtype=np.uint64)), p64(u), p64(v), p64(exp))
    if rc:
        raise OraclePanic(rc)
    return u, v, exp


def gkr_verify_subclaim(dim, f1_idx, f1_val, f2, f3, g, u, v, expected):
    return bool(lib().orc_gkr_verify_subclaim(dim, len(f1_idx), p64(f1_idx), p64(f1_val), p64(f2), p64(f3), p64(g),
                                              p64(np.ascontiguousarray(u)), p64(np.ascontiguousarray(v)),
                                              p64(np.ascontiguousarray(expected, dtype=np.uint64))))
