"""Host-side mirror of the reference's prover API (same names, argument meaning and error behaviour), over the C ABI.

Reference (paths relative to arkworks-rs/sumcheck):
  ListOfProductsOfPolynomials / PolynomialInfo   src/ml_sumcheck/data_structures.rs:25-109
  IPForMLSumcheck.prover_init / prove_round       src/ml_sumcheck/protocol/prover.rs:49,74
  MLSumcheck.prove / prove_as_subprotocol         src/ml_sumcheck/mod.rs:42,50
  Blake2b512Rng (FeedableRNG)                     src/rng.rs:11-81
  GKRRoundSumcheck.prove + phase initialisers     src/gkr_round_sumcheck/mod.rs:22-139

Field elements are numpy ``uint64`` arrays whose last axis holds the 4 little-endian Montgomery limbs of a
BLS12-381 Fr element — the memory layout of ark-ff's ``Fr``.  A dense multilinear extension is a ``[2^nv, 4]`` array.
Where the reference panics this module raises ``Panic``; where it returns ``Err`` it raises ``SumcheckError``.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import U8P, U32P, U64P, RngState

P_MODULUS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
FR_ONE = np.array([0x00000001FFFFFFFE, 0x5884B7FA00034802, 0x998C4FEFECBC4FF5, 0x1824B159ACC5056F], dtype=np.uint64)


class Panic(Exception):
    """The reference panics here (misuse of the prover state machine, prover.rs:50-52,79-81,90-92,96-98)."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class SumcheckError(Exception):
    """crate::Error (src/error.rs:7-19) — device/driver failures map to Error::OtherError."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def _check(rc):
    if rc == 0:
        return
    msg = capi.lib().sc_last_error().decode()
    if -4 <= rc <= -1:
        raise Panic(rc, msg)
    raise SumcheckError(rc, msg)


def _p64(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(U64P)


def _p32(a):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(U32P)


def _elems(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


# ---------------------------------------------------------------------------------------------- transcript
class Blake2b512Rng:
    """FeedableRNG over Blake2b-512 (src/rng.rs:22-81); state is the plain-data sc_blake2b512_rng."""

    def __init__(self):
        self.state = RngState()
        capi.lib().sc_rng_setup(C.byref(self.state))

    @classmethod
    def setup(cls):  # rng.rs:30-34
        return cls()

    def feed(self, msg):
        """rng.rs:36-41.  `msg` is bytes already in ark-serialize form, or an object with serialize_uncompressed()."""
        b = msg if isinstance(msg, (bytes, bytearray)) else msg.serialize_uncompressed()
        capi.lib().sc_rng_feed_bytes(C.byref(self.state), bytes(b), len(b))

    def fill_bytes(self, n):  # rng.rs:57-80
        out = np.zeros(max(n, 1), dtype=np.uint8)
        capi.lib().sc_rng_fill_bytes(C.byref(self.state), out.ctypes.data_as(U8P), n)
        return out[:n].tobytes()

    def next_u64(self):  # rng.rs:51-55
        return int(capi.lib().sc_rng_next_u64(C.byref(self.state)))

    def next_u32(self):  # rng.rs:45-49
        return int.from_bytes(self.fill_bytes(4), "little")


# ---------------------------------------------------------------------------------------------- data structures
class PolynomialInfo:
    """data_structures.rs:47-55."""

    def __init__(self, max_multiplicands, num_variables):
        self.max_multiplicands = max_multiplicands
        self.num_variables = num_variables

    def serialize_uncompressed(self):
        return self.max_multiplicands.to_bytes(8, "little") + self.num_variables.to_bytes(8, "little")


class ListOfProductsOfPolynomials:
    """data_structures.rs:25-96.  Tables are de-duplicated by object identity, like the reference's Rc pointers."""

    def __init__(self, num_variables):
        self.max_multiplicands = 0
        self.num_variables = num_variables
        self.products = []                 # [(coefficient[4], [indices])]
        self.flattened_ml_extensions = []  # unique tables
        self._lookup = {}                  # id(table) -> index   (raw_pointers_lookup_table)

    @classmethod
    def new(cls, num_variables):  # data_structures.rs:59-67
        return cls(num_variables)

    def info(self):  # data_structures.rs:39-44
        return PolynomialInfo(self.max_multiplicands, self.num_variables)

    def add_product(self, product, coefficient):  # data_structures.rs:71-96
        product = list(product)
        assert len(product) > 0
        self.max_multiplicands = max(self.max_multiplicands, len(product))
        indexed = []
        for m in product:
            assert m.dtype == np.uint64 and m.shape == (1 << self.num_variables, 4), \
                "product has a multiplicand with wrong number of variables"
            key = id(m)
            if key in self._lookup:
                indexed.append(self._lookup[key])
            else:
                self._lookup[key] = len(self.flattened_ml_extensions)
                self.flattened_ml_extensions.append(np.ascontiguousarray(m))
                if self.flattened_ml_extensions[-1] is not m:  # keep identity stable for later look-ups
                    self._lookup[id(self.flattened_ml_extensions[-1])] = self._lookup[key]
                indexed.append(self._lookup[key])
        self.products.append((_elems(coefficient), indexed))

    def evaluate(self, point, device=0):  # data_structures.rs:99-109, on the device
        coeffs, offsets, indices = self._csr()
        T = len(self.flattened_ml_extensions)
        tabs = (C.c_void_p * max(T, 1))(*[t.ctypes.data for t in self.flattened_ml_extensions])
        out = np.zeros(4, dtype=np.uint64)
        _check(capi.lib().sc_poly_evaluate(self.num_variables, T, tabs, len(self.products), _p64(coeffs), _p32(offsets),
                                           _p32(indices), _p64(_elems(point)), device, _p64(out)))
        return out

    # what crosses the C ABI
    def _csr(self):
        coeffs = np.ascontiguousarray(np.stack([c for c, _ in self.products])) if self.products else np.zeros((0, 4), np.uint64)
        offs, idx = [0], []
        for _, ix in self.products:
            idx.extend(ix)
            offs.append(len(idx))
        return coeffs, np.array(offs, dtype=np.uint32), np.array(idx if idx else [0], dtype=np.uint32)


class ProverMsg:
    """prover.rs:14-17."""

    def __init__(self, evaluations):
        self.evaluations = evaluations  # [d+1, 4]

    def serialize_uncompressed(self):
        n = self.evaluations.shape[0]
        b = _serialize_msgs(self.evaluations.reshape(1, n, 4))
        return b[8:]  # strip the outer Vec<ProverMsg> length


class SubClaim:
    """verifier.rs:29-34."""

    def __init__(self, point, expected_evaluation):
        self.point = point
        self.expected_evaluation = expected_evaluation


class VerifierMsg:
    """verifier.rs:11-15."""

    def __init__(self, randomness):
        self.randomness = _elems(randomness)


def _serialize_msgs(evals):
    """ark-serialize of Vec<ProverMsg<F>> for evals[nv, d+1, 4]."""
    evals = np.ascontiguousarray(evals, dtype=np.uint64)
    nv, dp1 = evals.shape[0], evals.shape[1]
    L = capi.lib()
    n = L.sc_serialize_proof(_p64(evals), nv, dp1 - 1, None)
    out = np.zeros(n, dtype=np.uint8)
    got = L.sc_serialize_proof(_p64(evals), nv, dp1 - 1, out.ctypes.data_as(U8P))
    if got != n:
        raise SumcheckError(-10, "sc_serialize_proof failed: " + L.sc_last_error().decode())
    return out.tobytes()


class ProverState:
    """prover.rs:19-33, resident in HBM.  Public fields of the reference are properties here."""

    def __init__(self, handle, poly=None):
        self._h = handle
        self._poly = poly  # keeps host tables alive / gives list_of_products

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            capi.lib().sc_prover_destroy(h)

    @property
    def round(self):
        return int(capi.lib().sc_prover_round(self._h))

    @property
    def num_vars(self):
        return int(capi.lib().sc_prover_num_vars(self._h))

    @property
    def max_multiplicands(self):
        return int(capi.lib().sc_prover_max_multiplicands(self._h))

    @property
    def list_of_products(self):
        return list(self._poly.products) if self._poly is not None else None

    @property
    def randomness(self):
        L = capi.lib()
        n = L.sc_prover_randomness(self._h, None, 0)
        out = np.zeros((max(n, 1), 4), dtype=np.uint64)
        L.sc_prover_randomness(self._h, _p64(out), n)
        return out[:n]

    def table(self, j):
        L = capi.lib()
        ln = C.c_uint64()
        _check(L.sc_prover_table(self._h, j, None, 0, C.byref(ln)))
        out = np.zeros((ln.value, 4), dtype=np.uint64)
        _check(L.sc_prover_table(self._h, j, _p64(out), ln.value, C.byref(ln)))
        return out

    @property
    def flattened_ml_extensions(self):
        n = len(self._poly.flattened_ml_extensions) if self._poly is not None else 2
        return [self.table(j) for j in range(n)]

    def reset(self):
        _check(capi.lib().sc_prover_reset(self._h))

    def load_tables(self, tables):
        tabs = (C.c_void_p * len(tables))(*[t.ctypes.data for t in tables])
        _check(capi.lib().sc_prover_load_tables(self._h, tabs))

    def set_stream(self, cuda_stream):
        _check(capi.lib().sc_prover_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def prove_into(self, rng, evals):
        """sc_ml_prove on this (round-0) handle into a preallocated [nv, d+1, 4] array."""
        _check(capi.lib().sc_ml_prove(self._h, C.byref(rng.state), _p64(evals), None))

    def set_timing(self, enabled):
        _check(capi.lib().sc_prover_set_timing(self._h, 1 if enabled else 0))

    def round_times_ms(self):
        nv = self.num_vars
        out = np.zeros(nv, dtype=np.float32)
        capi.lib().sc_prover_round_times_ms(self._h, out.ctypes.data_as(capi.F32P), nv)
        return out

    def launch_count(self):
        return int(capi.lib().sc_prover_launch_count(self._h))

    def resident_round_count(self):
        """Rounds served by the resident kernel (one cooperative launch for all small rounds) since creation/reset."""
        return int(capi.lib().sc_prover_resident_round_count(self._h))

    def gemm_round_count(self):
        """Rounds that ran on the tensor-core contraction kernels (degree-3 products) since creation/reset."""
        return int(capi.lib().sc_prover_gemm_round_count(self._h))

    def tc_round_count(self):
        """Fold rounds that ran on the TMA + tensor-core kernel since creation/reset."""
        return int(capi.lib().sc_prover_tc_round_count(self._h))


def _create(poly, device):
    coeffs, offsets, indices = poly._csr()
    T = len(poly.flattened_ml_extensions)
    tabs = (C.c_void_p * max(T, 1))(*[t.ctypes.data for t in poly.flattened_ml_extensions])
    h = C.c_void_p()
    if isinstance(device, (list, tuple)):  # several GPUs driven by this one process (sc_prover_create_multi)
        ids = (C.c_int * len(device))(*device)
        _check(capi.lib().sc_prover_create_multi(C.byref(h), poly.num_variables, T, tabs, len(poly.products),
                                                 _p64(coeffs) if len(poly.products) else None, _p32(offsets), _p32(indices), ids, len(device)))
        return ProverState(h, poly)
    _check(capi.lib().sc_prover_create(C.byref(h), poly.num_variables, T, tabs, len(poly.products),
                                       _p64(coeffs) if len(poly.products) else None, _p32(offsets), _p32(indices), device))
    return ProverState(h, poly)


class IPForMLSumcheck:
    """protocol/mod.rs:10-13 marker + prover.rs / verifier.rs:128 entry points on the prover path."""

    @staticmethod
    def prover_init(polynomial, device=0):  # prover.rs:49-69
        """`device` is a CUDA device index, or a list of them: the tables are then sharded by the high hypercube bits over those
        GPUs (one host thread per GPU inside the library) and the state behaves like a single-GPU one."""
        return _create(polynomial, device)

    @staticmethod
    def prove_round(prover_state, v_msg):  # prover.rs:74-153
        d = prover_state.max_multiplicands
        out = np.zeros((d + 1, 4), dtype=np.uint64)
        r = _p64(v_msg.randomness) if v_msg is not None else None
        _check(capi.lib().sc_prove_round(prover_state._h, r, _p64(out)))
        return ProverMsg(out)

    @staticmethod
    def sample_round(rng):  # verifier.rs:128-132
        out = np.zeros(4, dtype=np.uint64)
        capi.lib().sc_rng_sample_fr(C.byref(rng.state), _p64(out))
        return VerifierMsg(out)


def _fr_add_mont(a, b):
    """a + b mod p on Montgomery limbs (addition is representation-agnostic)."""
    x = sum(int(a[i]) << (64 * i) for i in range(4)) + sum(int(b[i]) << (64 * i) for i in range(4))
    x %= P_MODULUS
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


class MLSumcheck:
    """src/ml_sumcheck/mod.rs:19-70 (prover side)."""

    @staticmethod
    def extract_sum(proof):  # mod.rs:26-28
        return _fr_add_mont(proof[0].evaluations[0], proof[0].evaluations[1])

    @staticmethod
    def prove(polynomial, device=0):  # mod.rs:42-45
        rng = Blake2b512Rng.setup()
        proof, _ = MLSumcheck.prove_as_subprotocol(rng, polynomial, device)
        return proof

    @staticmethod
    def prove_into(polynomial, evals, device=0, randomness=None):
        """MLSumcheck::prove through the one-call C entry point (sc_ml_prove_oneshot: create + upload + prove + destroy)
        into a preallocated [nv, d+1, 4] array — what the Rust wrapper's `prove(&poly)` binds."""
        coeffs, offsets, indices = polynomial._csr()
        T = len(polynomial.flattened_ml_extensions)
        tabs = (C.c_void_p * max(T, 1))(*[t.ctypes.data for t in polynomial.flattened_ml_extensions])
        _check(capi.lib().sc_ml_prove_oneshot(polynomial.num_variables, T, tabs, len(polynomial.products), _p64(coeffs), _p32(offsets),
                                              _p32(indices), device, _p64(evals), _p64(randomness) if randomness is not None else None))
        return evals

    @staticmethod
    def prove_as_subprotocol(fs_rng, polynomial, device=0):  # mod.rs:50-70
        if isinstance(fs_rng, Blake2b512Rng):
            state = _create(polynomial, device)
            nv, d = polynomial.num_variables, state.max_multiplicands
            evals = np.zeros((nv, d + 1, 4), dtype=np.uint64)
            _check(capi.lib().sc_ml_prove(state._h, C.byref(fs_rng.state), _p64(evals), None))
            return [ProverMsg(evals[i]) for i in range(nv)], state
        # generic FeedableRNG: the round loop stays on the host, one C call per round
        fs_rng.feed(polynomial.info().serialize_uncompressed())
        state = IPForMLSumcheck.prover_init(polynomial, device)
        v_msg, msgs = None, []
        for _ in range(polynomial.num_variables):
            pm = IPForMLSumcheck.prove_round(state, v_msg)
            fs_rng.feed(pm.serialize_uncompressed())
            msgs.append(pm)
            v_msg = IPForMLSumcheck.sample_round(fs_rng)
        # mod.rs:65-67 pushes the last challenge without folding
        _check(capi.lib().sc_prover_push_randomness(state._h, _p64(v_msg.randomness)))
        return msgs, state

    @staticmethod
    def verify(polynomial_info, claimed_sum, proof, device=0):  # mod.rs:73-80
        return MLSumcheck.verify_as_subprotocol(Blake2b512Rng.setup(), polynomial_info, claimed_sum, proof, device)

    @staticmethod
    def verify_as_subprotocol(fs_rng, polynomial_info, claimed_sum, proof, device=0):  # mod.rs:84-100
        nv, d = polynomial_info.num_variables, polynomial_info.max_multiplicands
        if len(proof) < nv:
            raise Panic(-4, "proof is incomplete")  # mod.rs:93
        evals = np.ascontiguousarray(np.stack([m.evaluations for m in proof[:nv]]))
        point = np.zeros((nv, 4), dtype=np.uint64)
        expected = np.zeros(4, dtype=np.uint64)
        _check(capi.lib().sc_ml_verify(C.byref(fs_rng.state), nv, d, _p64(_elems(claimed_sum)), _p64(evals), device, _p64(point),
                                       _p64(expected)))
        return SubClaim(point, expected)

    @staticmethod
    def serialize_proof(proof):
        return _serialize_msgs(np.stack([m.evaluations for m in proof]))


# ---------------------------------------------------------------------------------------------- GKR
class SparseMultilinearExtension:
    """ark-poly SparseMultilinearExtension as (unique indices, values)."""

    def __init__(self, num_vars, indices, values):
        self.num_vars = num_vars
        self.indices = np.ascontiguousarray(indices, dtype=np.uint64)
        self.values = _elems(values).reshape(-1, 4)
        assert self.indices.shape[0] == self.values.shape[0]


class GKRProof:
    """gkr_round_sumcheck/data_structures.rs:9-20."""

    def __init__(self, phase1_sumcheck_msgs, phase2_sumcheck_msgs):
        self.phase1_sumcheck_msgs = phase1_sumcheck_msgs
        self.phase2_sumcheck_msgs = phase2_sumcheck_msgs

    def extract_sum(self):
        return _fr_add_mont(self.phase1_sumcheck_msgs[0].evaluations[0], self.phase1_sumcheck_msgs[0].evaluations[1])


def _dim_of(table):
    n = table.shape[0]
    assert n & (n - 1) == 0
    return n.bit_length() - 1


def initialize_phase_one(f1, f3, g, device=0):
    """mod.rs:22-42 -> (h_g dense, f1 fixed at g as SparseMultilinearExtension)."""
    f3, g = _elems(f3), _elems(g)
    dim = _dim_of(f3)
    assert f1.num_vars == dim * 3 and g.shape[0] == dim
    nnz = f1.indices.shape[0]
    h_g = np.zeros((1 << dim, 4), dtype=np.uint64)
    gi = np.zeros(max(nnz, 1), dtype=np.uint64)
    gv = np.zeros((max(nnz, 1), 4), dtype=np.uint64)
    n_g = C.c_uint64()
    _check(capi.lib().sc_gkr_initialize_phase_one(dim, nnz, _p64(f1.indices), _p64(f1.values), _p64(f3), _p64(g), device,
                                                  _p64(h_g), _p64(gi), _p64(gv), C.byref(n_g)))
    return h_g, SparseMultilinearExtension(2 * dim, gi[:n_g.value].copy(), gv[:n_g.value].copy())


def initialize_phase_two(f1_g, u, device=0):
    """mod.rs:57-63 -> f1 fixed at g||u as a dense table."""
    u = _elems(u)
    dim = u.shape[0]
    assert dim * 2 == f1_g.num_vars
    out = np.zeros((1 << dim, 4), dtype=np.uint64)
    _check(capi.lib().sc_gkr_initialize_phase_two(dim, f1_g.indices.shape[0], _p64(f1_g.indices), _p64(f1_g.values), _p64(u),
                                                  device, _p64(out)))
    return out


def start_phase1_sumcheck(h_g, f2, device=0):
    """mod.rs:45-54."""
    h_g, f2 = _elems(h_g), _elems(f2)
    dim = _dim_of(h_g)
    assert _dim_of(f2) == dim
    h = C.c_void_p()
    _check(capi.lib().sc_gkr_start_phase1_sumcheck(C.byref(h), dim, _p64(h_g), _p64(f2), device))
    return ProverState(h)


def start_phase2_sumcheck(f1_gu, f3, f2_u, device=0):
    """mod.rs:66-82."""
    f1_gu, f3, f2_u = _elems(f1_gu), _elems(f3), _elems(f2_u)
    dim = _dim_of(f1_gu)
    assert _dim_of(f3) == dim
    h = C.c_void_p()
    _check(capi.lib().sc_gkr_start_phase2_sumcheck(C.byref(h), dim, _p64(f1_gu), _p64(f3), _p64(f2_u), device))
    return ProverState(h)


class GKRRoundSumcheck:
    """src/gkr_round_sumcheck/mod.rs:85-139 (prover side)."""

    @staticmethod
    def prove(rng, f1, f2, f3, g, device=0, return_challenges=False):
        f2, f3, g = _elems(f2), _elems(f3), _elems(g)
        dim = _dim_of(f2)
        assert f1.num_vars == 3 * dim and _dim_of(f3) == dim  # mod.rs:100-101
        m1 = np.zeros((dim, 3, 4), dtype=np.uint64)
        m2 = np.zeros((dim, 3, 4), dtype=np.uint64)
        u = np.zeros((dim, 4), dtype=np.uint64)
        v = np.zeros((dim, 4), dtype=np.uint64)
        if not isinstance(rng, Blake2b512Rng):
            # generic R: FeedableRNG (mod.rs:93): the reference's own loop, every step on the device — the four phase
            # helpers, prove_round per round, f2.evaluate(u); the transcript stays with the caller's rng object
            h_g, f1_g = initialize_phase_one(f1, f3, g, device)                     # mod.rs:106
            ps = start_phase1_sumcheck(h_g, f2, device)                            # mod.rs:107
            vm, msgs1, u_l = None, [], []
            for _ in range(dim):                                                    # mod.rs:111-119
                pm = IPForMLSumcheck.prove_round(ps, vm)
                rng.feed(pm.serialize_uncompressed())
                msgs1.append(pm)
                vm = IPForMLSumcheck.sample_round(rng)
                u_l.append(vm.randomness)
            u = np.stack(u_l)
            f1_gu = initialize_phase_two(f1_g, u, device)                           # mod.rs:121
            f2_poly = ListOfProductsOfPolynomials.new(dim)
            f2_poly.add_product([f2], FR_ONE)
            f2_u = f2_poly.evaluate(u, device)                                      # f2.evaluate(&u), mod.rs:122
            ps = start_phase2_sumcheck(f1_gu, f3, f2_u, device)
            vm, msgs2, v_l = None, [], []
            for _ in range(dim):                                                    # mod.rs:126-133
                pm = IPForMLSumcheck.prove_round(ps, vm)
                rng.feed(pm.serialize_uncompressed())
                msgs2.append(pm)
                vm = IPForMLSumcheck.sample_round(rng)
                v_l.append(vm.randomness)
            proof = GKRProof(msgs1, msgs2)
            return (proof, u, np.stack(v_l)) if return_challenges else proof
        _check(capi.lib().sc_gkr_prove(C.byref(rng.state), dim, f1.indices.shape[0], _p64(f1.indices), _p64(f1.values),
                                       _p64(f2), _p64(f3), _p64(g), device, _p64(m1), _p64(m2), _p64(u), _p64(v)))
        proof = GKRProof([ProverMsg(m1[i]) for i in range(dim)], [ProverMsg(m2[i]) for i in range(dim)])
        return (proof, u, v) if return_challenges else proof


    @staticmethod
    def prove_batch(rngs, f1s, f2s, f3s, gs, device=0):
        """L independent GKRRoundSumcheck::prove calls of one dim in one library call (sc_gkr_prove_batch): the uploads overlap
        with the initialisers and every sumcheck round is issued for all layers before the first result is collected.  Each
        returned GKRProof is bit-identical to a separate prove() with the same rng state."""
        L = len(f1s)
        assert L >= 1 and len(rngs) == len(f2s) == len(f3s) == len(gs) == L
        f2s, f3s, gs = [_elems(x) for x in f2s], [_elems(x) for x in f3s], [_elems(x) for x in gs]
        dim = _dim_of(f2s[0])
        for f1, f2, f3, g in zip(f1s, f2s, f3s, gs):
            assert _dim_of(f2) == dim and _dim_of(f3) == dim and f1.num_vars == 3 * dim and g.shape[0] == dim
            if not isinstance(rngs[0], Blake2b512Rng):
                raise SumcheckError(-5, "the batched device path binds the concrete Blake2b512Rng")
        states = (RngState * L)(*[r.state for r in rngs])
        nnz = np.array([f1.indices.shape[0] for f1 in f1s], dtype=np.uint64)
        m1 = [np.zeros((dim, 3, 4), dtype=np.uint64) for _ in range(L)]
        m2 = [np.zeros((dim, 3, 4), dtype=np.uint64) for _ in range(L)]
        vp = lambda arrs: (C.c_void_p * L)(*[a.ctypes.data for a in arrs])
        _check(capi.lib().sc_gkr_prove_batch(L, states, dim, _p64(nnz), vp([f.indices for f in f1s]), vp([f.values for f in f1s]), vp(f2s),
                                             vp(f3s), vp(gs), device, vp(m1), vp(m2), None, None))
        for r, st in zip(rngs, states):  # the transcripts advanced inside the call
            C.memmove(C.byref(r.state), C.byref(st), C.sizeof(RngState))
        return [GKRProof([ProverMsg(a[i]) for i in range(dim)], [ProverMsg(b[i]) for i in range(dim)]) for a, b in zip(m1, m2)]
