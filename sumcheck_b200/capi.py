"""ctypes declarations for libsumcheck_b200.so — one entry per symbol of include/sumcheck_b200.h."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

U64P = C.POINTER(C.c_uint64)
U32P = C.POINTER(C.c_uint32)
U8P = C.POINTER(C.c_uint8)
F32P = C.POINTER(C.c_float)


class RngState(C.Structure):
    """sc_blake2b512_rng (plain data)."""
    _fields_ = [("h", C.c_uint64 * 8), ("t", C.c_uint64 * 2), ("buf", C.c_uint8 * 128), ("buflen", C.c_uint64)]


# name -> (restype, argtypes); the CPU test checks this table against the header.
SIGNATURES = {
    "sc_last_error": (C.c_char_p, []),
    "sc_device_count": (C.c_int, []),
    "sc_rng_setup": (None, [C.POINTER(RngState)]),
    "sc_rng_feed_bytes": (None, [C.POINTER(RngState), C.c_char_p, C.c_size_t]),
    "sc_rng_fill_bytes": (None, [C.POINTER(RngState), U8P, C.c_size_t]),
    "sc_rng_next_u64": (C.c_uint64, [C.POINTER(RngState)]),
    "sc_rng_sample_fr": (None, [C.POINTER(RngState), U64P]),
    "sc_prover_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.c_uint32, U64P,
                                   U32P, U32P, C.c_int]),
    "sc_prover_create_device": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.c_uint32,
                                          U64P, U32P, U32P, C.c_int]),
    "sc_prover_destroy": (None, [C.c_void_p]),
    "sc_prover_reset": (C.c_int, [C.c_void_p]),
    "sc_prover_load_tables": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "sc_prover_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sc_prove_round": (C.c_int, [C.c_void_p, U64P, U64P]),
    "sc_prover_max_multiplicands": (C.c_uint32, [C.c_void_p]),
    "sc_prover_num_vars": (C.c_uint32, [C.c_void_p]),
    "sc_prover_round": (C.c_uint32, [C.c_void_p]),
    "sc_prover_randomness": (C.c_uint32, [C.c_void_p, U64P, C.c_uint32]),
    "sc_prover_push_randomness": (C.c_int, [C.c_void_p, U64P]),
    "sc_prover_table": (C.c_int, [C.c_void_p, C.c_uint32, U64P, C.c_uint64, U64P]),
    "sc_ml_prove": (C.c_int, [C.c_void_p, C.POINTER(RngState), U64P, U64P]),
    "sc_ml_prove_oneshot": (C.c_int, [C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.c_uint32, U64P, U32P, U32P, C.c_int,
                                      U64P, U64P]),
    "sc_serialize_proof": (C.c_size_t, [U64P, C.c_uint32, C.c_uint32, U8P]),
    "sc_synth_table": (None, [U64P, C.c_uint64, C.c_uint64]),
    "sc_synth_table_at": (None, [U64P, C.c_uint64, C.c_uint64, C.c_uint64]),
    "sc_prover_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "sc_prover_round_times_ms": (C.c_uint32, [C.c_void_p, F32P, C.c_uint32]),
    "sc_prover_launch_count": (C.c_uint64, [C.c_void_p]),
    "sc_prover_tc_round_count": (C.c_uint64, [C.c_void_p]),
    "sc_prover_resident_round_count": (C.c_uint64, [C.c_void_p]),
    "sc_prover_gemm_round_count": (C.c_uint64, [C.c_void_p]),
    "sc_release_cached_memory": (None, []),
    "sc_fr_interpolate": (C.c_int, [U64P, C.c_uint32, U64P, U64P]),
    "sc_fr_contraction_finish": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, U64P]),
    "sc_poly_evaluate": (C.c_int, [C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.c_uint32, U64P, U32P, U32P, U64P, C.c_int, U64P]),
    "sc_ml_verify": (C.c_int, [C.POINTER(RngState), C.c_uint32, C.c_uint32, U64P, U64P, C.c_int, U64P, U64P]),
    "sc_comm_get_unique_id": (C.c_int, [U8P]),
    "sc_comm_create": (C.c_int, [C.POINTER(C.c_void_p), U8P, C.c_int, C.c_int, C.c_int]),
    "sc_comm_destroy": (None, [C.c_void_p]),
    "sc_prover_create_sharded": (C.c_int, [C.POINTER(C.c_void_p), C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p),
                                           C.c_uint32, U64P, U32P, U32P]),
    "sc_prover_create_multi": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.c_uint32, U64P, U32P,
                                         U32P, C.POINTER(C.c_int), C.c_uint32]),
    "sc_gkr_initialize_phase_one": (C.c_int, [C.c_uint32, C.c_uint64, U64P, U64P, U64P, U64P, C.c_int, U64P, U64P, U64P, U64P]),
    "sc_gkr_initialize_phase_two": (C.c_int, [C.c_uint32, C.c_uint64, U64P, U64P, U64P, C.c_int, U64P]),
    "sc_gkr_start_phase1_sumcheck": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32, U64P, U64P, C.c_int]),
    "sc_gkr_start_phase2_sumcheck": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint32, U64P, U64P, U64P, C.c_int]),
    "sc_gkr_prove": (C.c_int, [C.POINTER(RngState), C.c_uint32, C.c_uint64, U64P, U64P, U64P, U64P, U64P, C.c_int, U64P, U64P,
                               U64P, U64P]),
    "sc_gkr_prove_batch": (C.c_int, [C.c_uint32, C.POINTER(RngState), C.c_uint32, U64P, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
}


def lib_path():
    # SC_LIB selects an alternative build of the SAME library (kernel-variant experiments under tools/); never a fallback
    return os.environ.get("SC_LIB") or os.path.join(_HERE, "libsumcheck_b200.so")


def lib():
    """Load the CUDA extension.  Fails loudly when it has not been built: there is no fallback implementation."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  sumcheck_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB
