/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's prover path, the transcript and the (out-of-scope, but
 * needed as a self-check) verifier.  Never linked into / called from the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may use it.
 *
 * PARITY UNPINNED: the reference (arkworks-rs/sumcheck @ /root/reference) cannot be built here (no
 * rustc/cargo; its arithmetic lives in the un-vendored, un-pinned crates ark-ff / ark-poly /
 * ark-serialize / ark-std / blake2) and it holds NO golden vectors or known-answer tests for this path
 * (SURVEY.md §4, §8c).  What pins this oracle instead: (i) an independent Python big-int model
 * (oracle/pymodel.py) must agree byte for byte; (ii) BLAKE2b against RFC 7693 vectors and hashlib;
 * (iii) the reference's own acceptance relations (prove -> verify accepts, poly.evaluate(point) ==
 * expected_evaluation, extract_sum == true sum) hold on every test input.
 *
 * All field elements cross this interface as uint64_t[4]: little-endian limbs, Montgomery form R=2^256,
 * fully reduced — the in-memory layout of ark-ff's Fp<MontBackend<FrConfig,4>,4>.
 */
#ifndef SUMCHECK_ORACLE_H
#define SUMCHECK_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* misuse codes = the reference's panics */
#define ORC_OK 0
#define ORC_ERR_CONSTANT (-1)        /* prover.rs:50-52  "Attempt to prove a constant."          */
#define ORC_ERR_FIRST_ROUND_MSG (-2) /* prover.rs:79-81  "first round should be prover first."   */
#define ORC_ERR_MISSING_MSG (-3)     /* prover.rs:90-92  "verifier message is empty"             */
#define ORC_ERR_NOT_ACTIVE (-4)      /* prover.rs:96-98  "Prover is not active"                  */
#define ORC_ERR_BAD_INPUT (-5)       /* data_structures.rs:78,82 asserts (empty product, bad index) */
#define ORC_ERR_REJECT (-6)          /* verifier.rs:109-113 Error::Reject                        */

/* ---- field helpers (ark-ff Fp ops) ---- */
void orc_fr_add(uint64_t out[4], const uint64_t a[4], const uint64_t b[4]);
void orc_fr_sub(uint64_t out[4], const uint64_t a[4], const uint64_t b[4]);
void orc_fr_mul(uint64_t out[4], const uint64_t a[4], const uint64_t b[4]);
void orc_fr_inv(uint64_t out[4], const uint64_t a[4]);
void orc_fr_from_u64(uint64_t out[4], uint64_t v);
void orc_fr_from_canonical(uint64_t out[4], const uint64_t canonical[4]);
void orc_fr_to_canonical(uint64_t out[4], const uint64_t a[4]);
void orc_fr_to_bytes(uint8_t out[32], const uint64_t a[4]); /* ark-serialize of Fr */

/* ---- Blake2b512Rng (src/rng.rs:22-81) ---- */
typedef struct orc_rng orc_rng;
orc_rng *orc_rng_setup(void);                                    /* rng.rs:30-34 */
orc_rng *orc_rng_clone(const orc_rng *);
void orc_rng_free(orc_rng *);
void orc_rng_feed_bytes(orc_rng *, const uint8_t *bytes, size_t n);          /* rng.rs:36-41 with msg already serialised */
void orc_rng_feed_poly_info(orc_rng *, uint64_t max_multiplicands, uint64_t num_variables); /* data_structures.rs:47-55 */
void orc_rng_feed_prover_msg(orc_rng *, const uint64_t *evals, size_t n_evals);             /* prover.rs:13-17 */
void orc_rng_fill_bytes(orc_rng *, uint8_t *dest, size_t n);                 /* rng.rs:61-80 */
uint64_t orc_rng_next_u64(orc_rng *);                                        /* rng.rs:51-55 */
void orc_rng_sample_fr(orc_rng *, uint64_t out[4]);                          /* verifier.rs:128-132 -> ark-ff Fp::rand */
void orc_blake2b512(const uint8_t *in, size_t n, uint8_t out[64]);

/* ---- IPForMLSumcheck prover (src/ml_sumcheck/protocol/prover.rs) ----
 * Polynomial = ListOfProductsOfPolynomials flattened (data_structures.rs:25-35): T unique tables of
 * 2^nv elements; products in CSR form: coeffs[n], offsets[n+1], indices[offsets[n]]. */
typedef struct orc_prover orc_prover;
void orc_set_threads(int n); /* 1 = serial (reference without `parallel`); >1 = the rayon schedule with n threads */
int orc_prover_init(orc_prover **out, uint32_t nv, uint32_t n_tables, const uint64_t *const *tables,
                    uint32_t n_products, const uint64_t *coeffs, const uint32_t *offsets, const uint32_t *indices);
int orc_prove_round(orc_prover *, const uint64_t *r_or_null, uint64_t *evals_out /* (d+1)*4 */);
uint32_t orc_prover_degree(const orc_prover *);       /* max_multiplicands */
uint32_t orc_prover_round(const orc_prover *);
uint64_t orc_prover_table_len(const orc_prover *);    /* current length of every table */
void orc_prover_copy_table(const orc_prover *, uint32_t j, uint64_t *out); /* flattened_ml_extensions[j].evaluations */
void orc_prover_free(orc_prover *);

/* MLSumcheck::prove_as_subprotocol (src/ml_sumcheck/mod.rs:50-70).  rng==NULL: MLSumcheck::prove (fresh setup()).
 * evals_out: nv*(d+1)*4 u64; randomness_out (optional): nv*4 u64 = ProverState.randomness;
 * final_tables_out (optional): T*2*4 u64 = the 2-entry tables left in ProverState. */
int orc_ml_prove(orc_rng *rng, uint32_t nv, uint32_t n_tables, const uint64_t *const *tables, uint32_t n_products,
                 const uint64_t *coeffs, const uint32_t *offsets, const uint32_t *indices, uint64_t *evals_out,
                 uint64_t *randomness_out, uint64_t *final_tables_out);
/* ark-serialize of Proof<F> = Vec<ProverMsg<F>>: u64 count, then per msg u64 len + (d+1) x 32 B.  Returns bytes written. */
size_t orc_serialize_proof(const uint64_t *evals, uint32_t nv, uint32_t d, uint8_t *out);

/* MLSumcheck::verify_as_subprotocol (mod.rs:84-100) + check_and_generate_subclaim (verifier.rs:90-121). */
int orc_ml_verify(orc_rng *rng, uint32_t nv, uint32_t d, const uint64_t claimed_sum[4], const uint64_t *evals,
                  uint64_t *point_out /* nv*4 */, uint64_t expected_out[4]);
void orc_interpolate_uni_poly(uint64_t out[4], const uint64_t *p_i, uint32_t len, const uint64_t eval_at[4]); /* verifier.rs:139 */
/* ListOfProductsOfPolynomials::evaluate (data_structures.rs:99-109) */
void orc_poly_evaluate(uint64_t out[4], uint32_t nv, uint32_t n_tables, const uint64_t *const *tables,
                       uint32_t n_products, const uint64_t *coeffs, const uint32_t *offsets, const uint32_t *indices,
                       const uint64_t *point);
/* DenseMultilinearExtension::fix_variables(&[r]) (ark-poly) : out has len/2 elements */
void orc_dense_fix_variable(uint64_t *out, const uint64_t *in, uint64_t len, const uint64_t r[4]);
void orc_dense_evaluate(uint64_t out[4], const uint64_t *table, uint32_t nv, const uint64_t *point);

/* ---- GKRRoundSumcheck (src/gkr_round_sumcheck/mod.rs) ----
 * f1: sparse MLE over 3*dim variables given as nnz unique (index, value) pairs; index bit layout g | x | y,
 * least-significant first (test.rs:47-55).  f2, f3: dense 2^dim. */
void orc_precompute_eq(uint64_t *out /* 2^dim*4 */, const uint64_t *g, uint32_t dim);
/* initialize_phase_one (mod.rs:22-42): h_g dense 2^dim; f1_g sparse (sorted unique): returns its nnz */
size_t orc_gkr_initialize_phase_one(uint32_t dim, size_t nnz, const uint64_t *f1_idx, const uint64_t *f1_val,
                                    const uint64_t *f3, const uint64_t *g, uint64_t *h_g_out,
                                    uint64_t *f1g_idx_out /* nnz */, uint64_t *f1g_val_out /* nnz*4 */);
/* initialize_phase_two (mod.rs:57-63) */
void orc_gkr_initialize_phase_two(uint32_t dim, size_t nnz_g, const uint64_t *f1g_idx, const uint64_t *f1g_val,
                                  const uint64_t *u, uint64_t *f1_gu_out /* 2^dim*4 */);
/* GKRRoundSumcheck::prove (mod.rs:93-139); msgs: dim*3*4 u64 per phase; u/v (optional): dim*4 */
int orc_gkr_prove(orc_rng *rng, uint32_t dim, size_t nnz, const uint64_t *f1_idx, const uint64_t *f1_val,
                  const uint64_t *f2, const uint64_t *f3, const uint64_t *g, uint64_t *phase1_out,
                  uint64_t *phase2_out, uint64_t *u_out, uint64_t *v_out);
/* GKRRoundSumcheck::verify (mod.rs:147-192) */
int orc_gkr_verify(orc_rng *rng, uint32_t dim, const uint64_t *phase1, const uint64_t *phase2,
                   const uint64_t claimed_sum[4], uint64_t *u_out, uint64_t *v_out, uint64_t expected_out[4]);
/* GKRRoundSumcheckSubClaim::verify_subclaim (data_structures.rs:33-56): 1 = true */
int orc_gkr_verify_subclaim(uint32_t dim, size_t nnz, const uint64_t *f1_idx, const uint64_t *f1_val,
                            const uint64_t *f2, const uint64_t *f3, const uint64_t *g, const uint64_t *u,
                            const uint64_t *v, const uint64_t expected[4]);

/* Deterministic synthetic inputs: SplitMix64(seed) -> 4 limbs, mask top bit, reject >= p,
 * used directly as the Montgomery representation. */
void orc_synth_table(uint64_t *out, uint64_t n_elems, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
