"""sumcheck_b200 — B200-native sumcheck prover path behind the reference's prover API.

Host-side mirror (Python over ctypes) of the arkworks-rs/sumcheck prover surface:
``MLSumcheck.prove / prove_as_subprotocol``, ``IPForMLSumcheck.prover_init / prove_round``,
``GKRRoundSumcheck.prove`` and the GKR phase initialisers.  Every call goes through the C ABI of
``libsumcheck_b200.so`` (include/sumcheck_b200.h); there is no CPU implementation behind it.
"""
from .api import (  # noqa: F401
    Blake2b512Rng,
    GKRProof,
    GKRRoundSumcheck,
    IPForMLSumcheck,
    ListOfProductsOfPolynomials,
    MLSumcheck,
    Panic,
    PolynomialInfo,
    ProverMsg,
    ProverState,
    SparseMultilinearExtension,
    SubClaim,
    SumcheckError,
    VerifierMsg,
    initialize_phase_one,
    initialize_phase_two,
    start_phase1_sumcheck,
    start_phase2_sumcheck,
)
from .capi import lib, lib_path  # noqa: F401
