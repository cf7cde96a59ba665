/* sumcheck_b200 — C ABI of the B200-native sumcheck prover path (libsumcheck_b200.so).
 *
 * Drop-in boundary for the PROVER path of arkworks-rs/sumcheck (file:line below are relative to the reference
 * tree).  The reference is pure Rust with #![forbid(unsafe_code)] (src/lib.rs:1) and no FFI of its own, so these
 * entry points are what a sibling wrapper crate exposing the same names would bind (INTEGRATION.md shows the
 * `extern "C"` block).  Everything is plain pointers and sizes; no CUDA or torch types appear.
 *
 * Field elements: BLS12-381 Fr, `uint64_t[4]` little-endian limbs in Montgomery form (R = 2^256), fully reduced —
 * exactly the in-memory layout of ark-ff's `Fp<MontBackend<FrConfig,4>,4>`, so `&[Fr]` can be passed as-is.
 *
 * Polynomial (ListOfProductsOfPolynomials, src/ml_sumcheck/data_structures.rs:25-35) crosses as:
 *   tables[n_tables]      host pointers to flattened_ml_extensions[j].evaluations (2^nv elements each)
 *   coeffs[n_products]    products[k].0
 *   offsets[n_products+1], indices[offsets[n_products]]   CSR of products[k].1 (indices into tables; repeats allowed)
 *
 * Conventions: every function returns SC_OK (0) or a negative SC_ERR_* code; sc_last_error() gives a thread-local
 * message.  SC_ERR_PANIC_* are the reference's panics: the Rust wrapper turns them back into panic!().
 * A handle is single-owner and not thread-safe (the reference's input container is !Send).  Calls are synchronous.
 */
#ifndef SUMCHECK_B200_H
#define SUMCHECK_B200_H
#include <stddef.h>
#include <stdint.h>

/* The library is built with -fvisibility=hidden: only the sc_* entry points below are exported. */
#if defined(__GNUC__)
#define SC_API __attribute__((visibility("default")))
#else
#define SC_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define SC_OK 0
#define SC_ERR_PANIC_CONSTANT (-1)        /* prover.rs:50-52  "Attempt to prove a constant."        */
#define SC_ERR_PANIC_FIRST_ROUND_MSG (-2) /* prover.rs:79-81  "first round should be prover first." */
#define SC_ERR_PANIC_MISSING_MSG (-3)     /* prover.rs:90-92  "verifier message is empty"           */
#define SC_ERR_PANIC_NOT_ACTIVE (-4)      /* prover.rs:96-98  "Prover is not active"                */
#define SC_ERR_BAD_INPUT (-5)             /* data_structures.rs:78,82 asserts; gkr mod.rs:28-29,100-101 asserts */
#define SC_ERR_REJECT (-6)                /* verifier.rs:109-113 Error::Reject (sc_ml_verify only) */
#define SC_ERR_CUDA (-10)                 /* CUDA runtime failure -> Error::OtherError (src/error.rs:19) */
#define SC_ERR_NO_DEVICE (-11)            /* no usable sm_100 device: the library never falls back to the CPU */
#define SC_ERR_COMM (-12)                 /* multi-GPU exchange failure */

SC_API const char *sc_last_error(void);
SC_API int sc_device_count(void); /* number of visible CUDA devices (0 or negative code when none) */

/* ------------------------------------------------------------------------------------------------------------
 * Blake2b512Rng (src/rng.rs:22-81) as plain data, so the wrapper crate's FeedableRNG type can live on either side.
 * sc_rng_sample_fr = IPForMLSumcheck::sample_round (verifier.rs:128-132 -> ark-ff Fp::rand). */
typedef struct sc_blake2b512_rng {
    uint64_t h[8];
    uint64_t t[2];
    uint8_t buf[128];
    uint64_t buflen;
} sc_blake2b512_rng;
SC_API void sc_rng_setup(sc_blake2b512_rng *rng);                                   /* rng.rs:30-34 */
SC_API void sc_rng_feed_bytes(sc_blake2b512_rng *rng, const uint8_t *b, size_t n);  /* rng.rs:36-41 (bytes = serialize_uncompressed(msg)) */
SC_API void sc_rng_fill_bytes(sc_blake2b512_rng *rng, uint8_t *dest, size_t n);     /* rng.rs:61-80 */
SC_API uint64_t sc_rng_next_u64(sc_blake2b512_rng *rng);                            /* rng.rs:51-55 */
SC_API void sc_rng_sample_fr(sc_blake2b512_rng *rng, uint64_t out[4]);              /* verifier.rs:128-132 */

/* ------------------------------------------------------------------------------------------------------------
 * ProverState (src/ml_sumcheck/protocol/prover.rs:19-33) resident in HBM. */
typedef struct sc_prover sc_prover;

/* IPForMLSumcheck::prover_init (prover.rs:49-69): deep-copies every table to device `device` (the caller's buffers
 * are never written).  nv == 0 -> SC_ERR_PANIC_CONSTANT. */
SC_API int sc_prover_create(sc_prover **out, uint32_t nv, uint32_t n_tables, const uint64_t *const *tables,
                     uint32_t n_products, const uint64_t *coeffs, const uint32_t *offsets, const uint32_t *indices,
                     int device);
/* Same, but the tables are ALREADY device pointers on `device` (e.g. produced by an earlier GPU stage); they are
 * read, never written, and must stay alive until the handle is destroyed or reset no longer needs them. */
SC_API int sc_prover_create_device(sc_prover **out, uint32_t nv, uint32_t n_tables, const uint64_t *const *d_tables,
                            uint32_t n_products, const uint64_t *coeffs, const uint32_t *offsets,
                            const uint32_t *indices, int device);
SC_API void sc_prover_destroy(sc_prover *p);
/* Rewind to round 0 on the tables given at creation (they are kept pristine in HBM); for repeated proofs. */
SC_API int sc_prover_reset(sc_prover *p);

/* Replace the resident tables with new host data of the same shape (H2D into the pristine copies) and rewind to
 * round 0: the per-proof upload of a caller that proves many polynomials of one shape with one handle.  For large
 * single-product polynomials (nv >= 21, d + 1 <= 5) the upload is pipelined: the tables arrive in 8 chunks on a copy stream
 * and round 1 — which needs no challenge — is summed chunk by chunk behind them, so the first prove_round / sc_ml_prove
 * after this call finds its first message ready (sc_prover_reset discards it).  Returns when the caller's buffers may
 * be reused.  SC_NO_EAGER_R1=1 disables the pipelining. */
SC_API int sc_prover_load_tables(sc_prover *p, const uint64_t *const *tables);
/* Run this handle's kernels and copies on the caller's CUDA stream (a cudaStream_t passed as void*; NULL restores
 * the handle's own stream), so the caller can order / time the work with its own events. */
SC_API int sc_prover_set_stream(sc_prover *p, void *cuda_stream);

/* IPForMLSumcheck::prove_round (prover.rs:74-153).  r_or_null = Some(VerifierMsg{randomness}) / None.
 * evals_out receives ProverMsg.evaluations: (max_multiplicands+1) x 4 u64, P(0)..P(d). */
SC_API int sc_prove_round(sc_prover *p, const uint64_t *r_or_null, uint64_t *evals_out);

SC_API uint32_t sc_prover_max_multiplicands(const sc_prover *p); /* ProverState.max_multiplicands */
SC_API uint32_t sc_prover_num_vars(const sc_prover *p);          /* ProverState.num_vars          */
SC_API uint32_t sc_prover_round(const sc_prover *p);             /* ProverState.round             */
/* ProverState.randomness (prover.rs:21): copies min(len, cap) elements, returns len. */
SC_API uint32_t sc_prover_randomness(const sc_prover *p, uint64_t *out, uint32_t cap);
/* `prover_state.randomness.push(r)` (ml_sumcheck/mod.rs:65-67): records the final challenge WITHOUT folding — for
 * wrappers that run the round loop themselves; sc_ml_prove does it internally. */
SC_API int sc_prover_push_randomness(sc_prover *p, const uint64_t r[4]);
/* ProverState.flattened_ml_extensions[j].evaluations at the current round (length 2^(nv-round+1), or 2^nv at
 * round 0): copies it to `out` (host) and stores its length in *len_out. */
SC_API int sc_prover_table(const sc_prover *p, uint32_t j, uint64_t *out, uint64_t cap_elems, uint64_t *len_out);

/* MLSumcheck::prove_as_subprotocol (src/ml_sumcheck/mod.rs:50-70) on a prover at round 0: feeds PolynomialInfo,
 * runs all rounds with the transcript, pushes the last challenge.  `rng` is updated in place (as &mut fs_rng).
 * evals_out: nv*(d+1)*4 u64 = Proof<F>; randomness_out (nullable): nv*4 u64 = ProverState.randomness. */
SC_API int sc_ml_prove(sc_prover *p, sc_blake2b512_rng *rng, uint64_t *evals_out, uint64_t *randomness_out);

/* MLSumcheck::prove (mod.rs:42-45): fresh Blake2b512Rng::setup(), host tables in, proof out. */
SC_API int sc_ml_prove_oneshot(uint32_t nv, uint32_t n_tables, const uint64_t *const *tables, uint32_t n_products,
                        const uint64_t *coeffs, const uint32_t *offsets, const uint32_t *indices, int device,
                        uint64_t *evals_out, uint64_t *randomness_out);

/* ark-serialize bytes of Proof<F> = Vec<ProverMsg<F>> (what `proof.serialize_uncompressed` yields): returns the
 * byte count 8 + nv*(8 + 32*(d+1)); writes when out != NULL. */
SC_API size_t sc_serialize_proof(const uint64_t *evals, uint32_t nv, uint32_t d, uint8_t *out);

/* Synthetic-input helper for benches and smoke runs (not part of the reference surface): fills `out` with n_elems
 * uniform Fr elements (Montgomery limbs) from a counter-based SplitMix64 stream — see sumcheck_b200/synth.py. */
SC_API void sc_synth_table(uint64_t *out, uint64_t n_elems, uint64_t seed);
/* elements [first_elem, first_elem + n_elems) of the same stream (a shard of a table) */
SC_API void sc_synth_table_at(uint64_t *out, uint64_t first_elem, uint64_t n_elems, uint64_t seed);

/* Enable (1) / disable (0, default) per-round CUDA-event timing of sc_ml_prove / sc_gkr phases on this handle. */
SC_API int sc_prover_set_timing(sc_prover *p, int enabled);
/* Per-round device timings of the last sc_ml_prove on this handle (ms, CUDA events on the launching stream):
 * copies min(nv, cap) values, returns nv. Kernel-only; excludes transcript/host time. */
SC_API uint32_t sc_prover_round_times_ms(const sc_prover *p, float *out, uint32_t cap);
/* Number of kernels launched by this handle since creation/reset. */
SC_API uint64_t sc_prover_launch_count(const sc_prover *p);
/* How many of those were fold rounds run by the TMA + tensor-core kernel (tcgen05.mma fix_variables; rounds with at
 * least SC_TC_MIN_PAIRS output pairs, default 2^14).  Environment: SC_NO_TC=1 keeps every round on the plain kernels,
 * SC_TC_MIN_PAIRS=<n> moves the threshold (tests use 128 to cover the path at small sizes). */
SC_API uint64_t sc_prover_tc_round_count(const sc_prover *p);
/* ... and how many rounds were served by the resident kernel: sc_ml_prove / sc_gkr_prove run every round with at most
 * SC_RES_MAX_PAIRS output pairs (default 2^16) inside ONE cooperative launch that stays on the GPU and exchanges the fold
 * constants / raw sums with the host transcript through mapped memory (no launch per round).  SC_NO_RESIDENT=1 disables it;
 * sc_prove_round (caller-driven rounds) never uses it. */
SC_API uint64_t sc_prover_resident_round_count(const sc_prover *p);
/* ... and how many rounds ran on the tensor-core CONTRACTION kernels (csrc/gemm_sum.cuh): for lists whose products all have
 * two (GKR phases), three (BASELINE configs 2 and 3) or four (config 4) multiplicands the sum over the hypercube of prover.rs:110-148 is a u8 x u8 -> s32
 * tcgen05.mma over the bytes of per-pair plain products; rounds with at least SC_TC_MIN_PAIRS pairs.  SC_NO_GEMM=1 disables it. */
SC_API uint64_t sc_prover_gemm_round_count(const sc_prover *p);

/* Handles return their device slab (<= 256 MiB), pinned result block and stream to a small per-device cache that later
 * handles reuse (allocation latency dominates small proofs such as the two phases of a GKR round).  This frees it;
 * SC_NO_ALLOC_CACHE=1 disables caching. */
SC_API void sc_release_cached_memory(void);

/* interpolate_uni_poly (src/ml_sumcheck/protocol/verifier.rs:139-251): the value at r of the polynomial of degree
 * n_evals-1 through (j, evals[j]), j = 0..n_evals-1 (n_evals <= 33).  Host-side scalar arithmetic (no GPU needed); the
 * same routine finishes every prover round: the device delivers the summed points and P(1) = P_prev(r) - P(0). */
SC_API int sc_fr_interpolate(const uint64_t *evals, uint32_t n_evals, const uint64_t r[4], uint64_t out[4]);

/* The host half of a tensor-core contraction round (csrc/gemm_sum.cuh, csrc/host_fr.h gemm_finish): a product of kx + ky tables
 * (kx, ky in {1, 2}) is split into an X side and a Y side; a side of one table contributes the values (a, b) = (table[2b],
 * table[2b+1]), a side of two tables the plain products (a a', b b', (a+b)(a'+b')).  z = the nx*ny integers sum_b X_i * Y_j
 * (n_limbs 32-bit limbs each, row-major, plain products of Montgomery-form operands) as the device delivers them; out
 * receives P(0..kx+ky), Montgomery form, before any coefficient: the unscaled round polynomial of prover.rs:110-148.
 * Host-side scalar arithmetic (no GPU needed) — exposed so that it can be checked against the big-integer model on CPU. */
SC_API int sc_fr_contraction_finish(const uint32_t *z, uint32_t n_limbs, uint32_t kx, uint32_t ky, uint64_t *out);

/* ------------------------------------------------------------------------------------------------------------
 * GKRRoundSumcheck (src/gkr_round_sumcheck/mod.rs).  f1: SparseMultilinearExtension over 3*dim variables as nnz
 * (index, value) pairs with unique indices (its BTreeMap); index bits are g | x | y, least significant first
 * (gkr test.rs:47-55).  f2, f3: dense, 2^dim elements.  g: dim elements. */

/* initialize_phase_one (mod.rs:22-42).  h_g_out: 2^dim elements.  f1_g (f1 fixed at g, a sparse MLE over 2*dim
 * variables) is returned as sorted unique (index,value) pairs: capacity nnz each; *nnz_g_out = its length. */
SC_API int sc_gkr_initialize_phase_one(uint32_t dim, uint64_t nnz, const uint64_t *f1_idx, const uint64_t *f1_val,
                                const uint64_t *f3, const uint64_t *g, int device, uint64_t *h_g_out,
                                uint64_t *f1g_idx_out, uint64_t *f1g_val_out, uint64_t *nnz_g_out);
/* initialize_phase_two (mod.rs:57-63): f1_g fixed at u, as a dense 2^dim table. */
SC_API int sc_gkr_initialize_phase_two(uint32_t dim, uint64_t nnz_g, const uint64_t *f1g_idx, const uint64_t *f1g_val,
                                const uint64_t *u, int device, uint64_t *f1_gu_out);
/* start_phase1_sumcheck (mod.rs:45-54) / start_phase2_sumcheck (mod.rs:66-82): ProverState for 1*(a*b) resp.
 * 1*(f1_gu * (f2_u*f3)). */
SC_API int sc_gkr_start_phase1_sumcheck(sc_prover **out, uint32_t dim, const uint64_t *h_g, const uint64_t *f2, int device);
SC_API int sc_gkr_start_phase2_sumcheck(sc_prover **out, uint32_t dim, const uint64_t *f1_gu, const uint64_t *f3,
                                 const uint64_t f2_u[4], int device);
/* GKRRoundSumcheck::prove (mod.rs:93-139) with the concrete Blake2b512Rng.  phase1_out/phase2_out: dim*3*4 u64
 * (GKRProof.phase{1,2}_sumcheck_msgs); u_out/v_out nullable: dim*4 u64. */
SC_API int sc_gkr_prove(sc_blake2b512_rng *rng, uint32_t dim, uint64_t nnz, const uint64_t *f1_idx, const uint64_t *f1_val,
                 const uint64_t *f2, const uint64_t *f3, const uint64_t *g, int device, uint64_t *phase1_out,
                 uint64_t *phase2_out, uint64_t *u_out, uint64_t *v_out);

/* n_layers independent GKRRoundSumcheck::prove calls of one dim in ONE call (the layers of a GKR circuit
 * whose inputs are already known, or a batch of circuits): every array argument holds one pointer (rngs: one state) per layer,
 * with the meaning it has in sc_gkr_prove.  The uploads of the layers overlap with the initialisers of the previous ones, and
 * every sumcheck round is issued for all layers before the first result is collected, so the per-round latency is shared by
 * the batch.  Each layer's outputs are bit-identical to a separate sc_gkr_prove call with the same rng state. */
SC_API int sc_gkr_prove_batch(uint32_t n_layers, sc_blake2b512_rng *rngs, uint32_t dim, const uint64_t *nnz,
                              const uint64_t *const *f1_idx, const uint64_t *const *f1_val, const uint64_t *const *f2,
                              const uint64_t *const *f3, const uint64_t *const *g, int device, uint64_t *const *phase1_out,
                              uint64_t *const *phase2_out, uint64_t *const *u_out, uint64_t *const *v_out);

/* ------------------------------------------------------------------------------------------------------------
 * Verifier-side counterparts (what the reference's tests call right after proving). */
/* ListOfProductsOfPolynomials::evaluate (src/ml_sumcheck/data_structures.rs:99-109): the polynomial at `point`
 * (nv elements); every table is folded nv times on the device. */
SC_API int sc_poly_evaluate(uint32_t nv, uint32_t n_tables, const uint64_t *const *tables, uint32_t n_products,
                     const uint64_t *coeffs, const uint32_t *offsets, const uint32_t *indices, const uint64_t *point,
                     int device, uint64_t out[4]);
/* MLSumcheck::verify_as_subprotocol (src/ml_sumcheck/mod.rs:84-100) with check_and_generate_subclaim
 * (verifier.rs:90-121).  evals: nv*(d+1)*4 u64 (the Proof).  Returns SC_ERR_REJECT when a round's P(0)+P(1) does not
 * match; else SubClaim.point -> point_out (nv*4, nullable) and SubClaim.expected_evaluation -> expected_out. */
SC_API int sc_ml_verify(sc_blake2b512_rng *rng, uint32_t nv, uint32_t d, const uint64_t claimed_sum[4], const uint64_t *evals,
                 int device, uint64_t *point_out, uint64_t expected_out[4]);

/* ------------------------------------------------------------------------------------------------------------
 * Multi-GPU (one process per GPU).  The reference has no distributed path; this is the B200-native extension the
 * north star asks for: every table is sharded by the HIGH log2(n_ranks) bits of the hypercube index, folds and sums
 * stay local, and the only per-round exchange is the d+1 partial sums over NVLink.  All ranks call the same
 * functions collectively and receive identical outputs. */
#define SC_COMM_ID_BYTES 128
typedef struct sc_comm sc_comm;
/* Rank 0 creates an id and shares it with the other ranks through any channel the host application has. */
SC_API int sc_comm_get_unique_id(uint8_t id_out[SC_COMM_ID_BYTES]);
SC_API int sc_comm_create(sc_comm **out, const uint8_t id[SC_COMM_ID_BYTES], int rank, int n_ranks, int device);
SC_API void sc_comm_destroy(sc_comm *c);
/* prover_init over a sharded polynomial: `nv` is the GLOBAL number of variables; shard_tables[j] holds elements
 * [rank*2^nv/n_ranks, (rank+1)*2^nv/n_ranks) of table j.  The handle then works with sc_prove_round / sc_ml_prove
 * exactly like a single-GPU one (sc_prover_table returns this rank's shard while the rounds are still sharded). */
SC_API int sc_prover_create_sharded(sc_prover **out, sc_comm *comm, uint32_t nv, uint32_t n_tables,
                             const uint64_t *const *shard_tables, uint32_t n_products, const uint64_t *coeffs,
                             const uint32_t *offsets, const uint32_t *indices);


/* Single-process multi-GPU: the same sharded prover driven by ONE host process — what MLSumcheck::prove(&poly)
 * (src/ml_sumcheck/mod.rs:42) needs to use several GPUs from one Rust process.  `tables` are the FULL host tables (2^nv
 * elements each); rank r uploads elements [r*2^nv/n, (r+1)*2^nv/n) to device_ids[r] from its own host thread.  n_devices
 * must be a power of two; the devices need peer access to each other (NVLink/NVSwitch), and a device may be listed more
 * than once (ranks then share it).  The returned handle works with every sc_prover_* / sc_prove_round / sc_ml_prove call
 * (sc_prover_set_stream excepted); n_devices == 1 gives an ordinary handle. */
SC_API int sc_prover_create_multi(sc_prover **out, uint32_t nv, uint32_t n_tables, const uint64_t *const *tables,
                                  uint32_t n_products, const uint64_t *coeffs, const uint32_t *offsets, const uint32_t *indices,
                                  const int *device_ids, uint32_t n_devices);

#ifdef __cplusplus
}
#endif
#endif
