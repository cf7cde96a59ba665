/* ORACLE — TEST INFRASTRUCTURE ONLY (see sumcheck_oracle.h for the header note: PARITY UNPINNED).
 * Every function cites the reference file:line (relative to /root/reference) it restates. */
#include "sumcheck_oracle.h"
#include "blake2b.h"
#include "fr.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ field helpers */
fr_t fr_pow(const fr_t *a, const fr_t *e) {
    fr_t acc = FR_ONE;
    for (int i = 255; i >= 0; i--) {
        acc = fr_mul(&acc, &acc);
        if ((e->l[i / 64] >> (i % 64)) & 1) acc = fr_mul(&acc, a);
    }
    return acc;
}
fr_t fr_inv(const fr_t *a) { /* Fermat: a^(p-2) */
    fr_t e = FR_P;
    e.l[0] -= 2; /* p's low limb is ...0001, no borrow past limb 0?  0xffffffff00000001-2 = 0xfffffffeffffffff: fine */
    return fr_pow(a, &e);
}
void fr_to_bytes(uint8_t out[32], const fr_t *a) {
    fr_t c = fr_to_canonical(a);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(c.l[i] >> (8 * j));
}
#define FR(p) ((const fr_t *)(p))
void orc_fr_add(uint64_t o[4], const uint64_t a[4], const uint64_t b[4]) { *(fr_t *)o = fr_add(FR(a), FR(b)); }
void orc_fr_sub(uint64_t o[4], const uint64_t a[4], const uint64_t b[4]) { *(fr_t *)o = fr_sub(FR(a), FR(b)); }
void orc_fr_mul(uint64_t o[4], const uint64_t a[4], const uint64_t b[4]) { *(fr_t *)o = fr_mul(FR(a), FR(b)); }
void orc_fr_inv(uint64_t o[4], const uint64_t a[4]) { *(fr_t *)o = fr_inv(FR(a)); }
void orc_fr_from_u64(uint64_t o[4], uint64_t v) { *(fr_t *)o = fr_from_u64(v); }
void orc_fr_from_canonical(uint64_t o[4], const uint64_t c[4]) { *(fr_t *)o = fr_from_canonical(FR(c)); }
void orc_fr_to_canonical(uint64_t o[4], const uint64_t a[4]) { *(fr_t *)o = fr_to_canonical(FR(a)); }
void orc_fr_to_bytes(uint8_t out[32], const uint64_t a[4]) { fr_to_bytes(out, FR(a)); }

/* ------------------------------------------------------------------ Blake2b512Rng: src/rng.rs:22-81 */
struct orc_rng {
    blake2b_state st; /* rng.rs:24 current_digest */
};
orc_rng *orc_rng_setup(void) { /* rng.rs:30-34 */
    orc_rng *r = (orc_rng *)malloc(sizeof(orc_rng));
    blake2b_init(&r->st);
    return r;
}
orc_rng *orc_rng_clone(const orc_rng *r) {
    orc_rng *c = (orc_rng *)malloc(sizeof(orc_rng));
    *c = *r;
    return c;
}
void orc_rng_free(orc_rng *r) { free(r); }
void orc_rng_feed_bytes(orc_rng *r, const uint8_t *b, size_t n) { blake2b_update(&r->st, b, n); } /* rng.rs:39 */
static void put_u64(uint8_t *o, uint64_t v) {
    for (int j = 0; j < 8; j++) o[j] = (uint8_t)(v >> (8 * j));
}
void orc_rng_feed_poly_info(orc_rng *r, uint64_t max_mult, uint64_t nv) {
    /* data_structures.rs:47-55: derive(CanonicalSerialize) = fields in order, usize as u64 LE */
    uint8_t b[16];
    put_u64(b, max_mult);
    put_u64(b + 8, nv);
    orc_rng_feed_bytes(r, b, 16);
}
void orc_rng_feed_prover_msg(orc_rng *r, const uint64_t *evals, size_t n) {
    /* prover.rs:13-17: ProverMsg{evaluations: Vec<F>} = u64 LE length, then n x 32 B canonical LE */
    uint8_t *b = (uint8_t *)malloc(8 + 32 * n);
    put_u64(b, (uint64_t)n);
    for (size_t i = 0; i < n; i++) fr_to_bytes(b + 8 + 32 * i, FR(evals + 4 * i));
    orc_rng_feed_bytes(r, b, 8 + 32 * n);
    free(b);
}
void orc_rng_fill_bytes(orc_rng *r, uint8_t *dest, size_t n) { /* rng.rs:61-80, statement by statement */
    uint8_t output[64];
    blake2b_final(&r->st, output); /* :62-63 digest = clone; output = finalize */
    size_t ptr = 0, digest_ptr = 0;
    while (ptr < n) {
        dest[ptr] = output[digest_ptr];
        ptr++;
        digest_ptr++;
        if (digest_ptr == 64) { /* :71-76 */
            blake2b_update(&r->st, output, 64);
            blake2b_final(&r->st, output);
            digest_ptr = 0;
        }
    }
    blake2b_update(&r->st, output, 64); /* :78 */
}
uint64_t orc_rng_next_u64(orc_rng *r) { /* rng.rs:51-55 */
    uint8_t t[8];
    orc_rng_fill_bytes(r, t, 8);
    uint64_t v = 0;
    for (int j = 7; j >= 0; j--) v = (v << 8) | t[j];
    return v;
}
void orc_rng_sample_fr(orc_rng *r, uint64_t out[4]) {
    /* verifier.rs:128-132 sample_round = F::rand(rng).  ark-ff (external): loop { 4 x next_u64 -> limbs in
     * order; clear the top 256-255 = 1 bit; accept if < p; the limbs ARE the Montgomery representation }. */
    for (;;) {
        fr_t t;
        for (int i = 0; i < 4; i++) t.l[i] = orc_rng_next_u64(r);
        t.l[3] &= 0xffffffffffffffffULL >> 1;
        if (!fr_geq_p(&t)) {
            memcpy(out, &t, 32);
            return;
        }
    }
}
void orc_blake2b512(const uint8_t *in, size_t n, uint8_t out[64]) {
    blake2b_state s;
    blake2b_init(&s);
    blake2b_update(&s, in, n);
    blake2b_final(&s, out);
}

/* ------------------------------------------------------------------ ProverState: prover.rs:19-33 */
struct orc_prover {
    uint32_t nv, n_tables, n_products, max_mult, round;
    uint64_t len;       /* current table length */
    fr_t **tables;      /* flattened_ml_extensions (deep copies, prover.rs:55-59) */
    fr_t *coeffs;       /* list_of_products[k].0 */
    uint32_t *offsets;  /* CSR of list_of_products[k].1 */
    uint32_t *indices;
    fr_t *randomness;   /* prover.rs:21 */
    uint32_t n_rand;
};
static int g_threads = 1;
void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

int orc_prover_init(orc_prover **out, uint32_t nv, uint32_t T, const uint64_t *const *tables, uint32_t n,
                    const uint64_t *coeffs, const uint32_t *offsets, const uint32_t *indices) {
    *out = NULL;
    if (nv == 0) return ORC_ERR_CONSTANT; /* prover.rs:50-52 */
    uint32_t maxm = 0;
    for (uint32_t k = 0; k < n; k++) {
        if (offsets[k + 1] <= offsets[k]) return ORC_ERR_BAD_INPUT; /* data_structures.rs:78 assert!(!product.is_empty()) */
        uint32_t m = offsets[k + 1] - offsets[k];
        if (m > maxm) maxm = m; /* data_structures.rs:79 */
        for (uint32_t j = offsets[k]; j < offsets[k + 1]; j++)
            if (indices[j] >= T) return ORC_ERR_BAD_INPUT;
    }
    orc_prover *p = (orc_prover *)calloc(1, sizeof(orc_prover));
    p->nv = nv; p->n_tables = T; p->n_products = n; p->max_mult = maxm; p->round = 0;
    p->len = (uint64_t)1 << nv;
    p->tables = (fr_t **)calloc(T, sizeof(fr_t *));
    for (uint32_t j = 0; j < T; j++) { /* deep copy: prover.rs:55-59 */
        p->tables[j] = (fr_t *)malloc(p->len * sizeof(fr_t));
        memcpy(p->tables[j], tables[j], p->len * sizeof(fr_t));
    }
    p->coeffs = (fr_t *)malloc((n ? n : 1) * sizeof(fr_t));
    memcpy(p->coeffs, coeffs, n * sizeof(fr_t));
    p->offsets = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));
    memcpy(p->offsets, offsets, (n + 1) * sizeof(uint32_t));
    p->indices = (uint32_t *)malloc((offsets[n] ? offsets[n] : 1) * sizeof(uint32_t));
    memcpy(p->indices, indices, offsets[n] * sizeof(uint32_t));
    p->randomness = (fr_t *)malloc((nv + 1) * sizeof(fr_t));
    p->n_rand = 0;
    *out = p;
    return ORC_OK;
}
void orc_prover_free(orc_prover *p) {
    if (!p) return;
    for (uint32_t j = 0; j < p->n_tables; j++) free(p->tables[j]);
    free(p->tables); free(p->coeffs); free(p->offsets); free(p->indices); free(p->randomness);
    free(p);
}
uint32_t orc_prover_degree(const orc_prover *p) { return p->max_mult; }
uint32_t orc_prover_round(const orc_prover *p) { return p->round; }
uint64_t orc_prover_table_len(const orc_prover *p) { return p->len; }
void orc_prover_copy_table(const orc_prover *p, uint32_t j, uint64_t *out) { memcpy(out, p->tables[j], p->len * sizeof(fr_t)); }

/* ark-poly DenseMultilinearExtension::fix_variables(&[r]) (external; called at prover.rs:88):
 * copies the table (`to_vec`), folds in place new[b] = old[2b] + r*(old[2b+1]-old[2b]), copies the half. */
static fr_t *dense_fix_variable_alloc(const fr_t *old, uint64_t len, const fr_t *r) {
    fr_t *poly = (fr_t *)malloc(len * sizeof(fr_t));
    memcpy(poly, old, len * sizeof(fr_t));
    for (uint64_t b = 0; b < len / 2; b++) {
        fr_t left = poly[b << 1], right = poly[(b << 1) + 1];
        fr_t d = fr_sub(&right, &left);
        d = fr_mul(r, &d);
        poly[b] = fr_add(&left, &d);
    }
    fr_t *res = (fr_t *)malloc((len / 2 ? len / 2 : 1) * sizeof(fr_t));
    memcpy(res, poly, (len / 2) * sizeof(fr_t));
    free(poly);
    return res;
}
void orc_dense_fix_variable(uint64_t *out, const uint64_t *in, uint64_t len, const uint64_t r[4]) {
    fr_t *res = dense_fix_variable_alloc(FR(in), len, FR(r));
    memcpy(out, res, (len / 2) * sizeof(fr_t));
    free(res);
}
void orc_dense_evaluate(uint64_t out[4], const uint64_t *table, uint32_t nv, const uint64_t *point) {
    /* ark-poly Polynomial::evaluate = fix_variables(point)[0]: variable 0 first */
    uint64_t len = (uint64_t)1 << nv;
    fr_t *cur = (fr_t *)malloc(len * sizeof(fr_t));
    memcpy(cur, table, len * sizeof(fr_t));
    for (uint32_t i = 0; i < nv; i++) {
        const fr_t *r = FR(point + 4 * i);
        for (uint64_t b = 0; b < len / 2; b++) {
            fr_t d = fr_sub(&cur[2 * b + 1], &cur[2 * b]);
            d = fr_mul(r, &d);
            cur[b] = fr_add(&cur[2 * b], &d);
        }
        len /= 2;
    }
    memcpy(out, &cur[0], 32);
    free(cur);
}

/* one index b of the sum loop body: prover.rs:114-129 */
static inline void sum_body(const orc_prover *p, uint64_t b, uint32_t degree, fr_t *products_sum, fr_t *product) {
    for (uint32_t k = 0; k < p->n_products; k++) {
        for (uint32_t t = 0; t <= degree; t++) product[t] = p->coeffs[k]; /* :116 product.fill(coefficient) */
        for (uint32_t jj = p->offsets[k]; jj < p->offsets[k + 1]; jj++) {
            const fr_t *table = p->tables[p->indices[jj]];
            fr_t start = table[b << 1];                       /* :119 */
            fr_t step = fr_sub(&table[(b << 1) + 1], &start); /* :120 */
            for (uint32_t t = 0; t <= degree; t++) {          /* :121-124 */
                product[t] = fr_mul(&product[t], &start);
                start = fr_add(&start, &step);
            }
        }
        for (uint32_t t = 0; t <= degree; t++) products_sum[t] = fr_add(&products_sum[t], &product[t]); /* :126-128 */
    }
}

int orc_prove_round(orc_prover *p, const uint64_t *r_or_null, uint64_t *evals_out) { /* prover.rs:74-153 */
    if (r_or_null) {
        if (p->round == 0) return ORC_ERR_FIRST_ROUND_MSG; /* :79-81 */
        p->randomness[p->n_rand++] = *FR(r_or_null);       /* :82 */
        const fr_t r = p->randomness[p->round - 1];        /* :85-86 */
        /* :87-89 cfg_iter_mut! over tables: rayon parallelism is ACROSS tables only */
#ifdef _OPENMP
#pragma omp parallel for num_threads(g_threads) schedule(dynamic, 1) if (g_threads > 1)
#endif
        for (uint32_t j = 0; j < p->n_tables; j++) {
            fr_t *nt = dense_fix_variable_alloc(p->tables[j], p->len, &r);
            free(p->tables[j]);
            p->tables[j] = nt;
        }
        p->len /= 2;
    } else if (p->round > 0) {
        return ORC_ERR_MISSING_MSG; /* :90-92 */
    }
    p->round += 1;                                   /* :94 */
    if (p->round > p->nv) return ORC_ERR_NOT_ACTIVE; /* :96-98 */
    const uint32_t i = p->round, nv = p->nv, degree = p->max_mult; /* :100-102 */
    const uint64_t n_b = (uint64_t)1 << (nv - i);
    fr_t *total = (fr_t *)calloc(degree + 1, sizeof(fr_t));
    if (g_threads <= 1) { /* :104-105, :110-132 serial fold */
        fr_t *product = (fr_t *)calloc(degree + 1, sizeof(fr_t));
        for (uint64_t b = 0; b < n_b; b++) sum_body(p, b, degree, total, product);
        free(product);
    } else { /* :106-107,:110 cfg_into_iter!(0..1<<(nv-i), 1<<10): chunks of >= 1024 indices; :138-148 reduce */
        const uint64_t chunk = 1 << 10;
        const uint64_t n_chunks = (n_b + chunk - 1) / chunk;
#ifdef _OPENMP
#pragma omp parallel num_threads(g_threads)
#endif
        {
            fr_t *ps = (fr_t *)calloc(degree + 1, sizeof(fr_t));
            fr_t *product = (fr_t *)calloc(degree + 1, sizeof(fr_t));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1) nowait
#endif
            for (uint64_t c = 0; c < n_chunks; c++) {
                uint64_t b1 = (c + 1) * chunk < n_b ? (c + 1) * chunk : n_b;
                for (uint64_t b = c * chunk; b < b1; b++) sum_body(p, b, degree, ps, product);
            }
#ifdef _OPENMP
#pragma omp critical
#endif
            for (uint32_t t = 0; t <= degree; t++) total[t] = fr_add(&total[t], &ps[t]);
            free(ps);
            free(product);
        }
    }
    memcpy(evals_out, total, (degree + 1) * sizeof(fr_t)); /* :150-152 */
    free(total);
    return ORC_OK;
}

int orc_ml_prove(orc_rng *rng_in, uint32_t nv, uint32_t T, const uint64_t *const *tables, uint32_t n,
                 const uint64_t *coeffs, const uint32_t *offsets, const uint32_t *indices, uint64_t *evals_out,
                 uint64_t *randomness_out, uint64_t *final_tables_out) {
    /* MLSumcheck::prove (mod.rs:42-45) when rng_in == NULL, else prove_as_subprotocol (mod.rs:50-70) */
    orc_prover *p;
    /* mod.rs:54 feeds poly.info() BEFORE prover_init (:56) can panic on nv == 0; the panic aborts either way */
    int rc = orc_prover_init(&p, nv, T, tables, n, coeffs, offsets, indices);
    if (rc) return rc;
    orc_rng *rng = rng_in ? rng_in : orc_rng_setup();
    const uint32_t d = p->max_mult;
    orc_rng_feed_poly_info(rng, d, nv); /* :54 */
    fr_t r;
    int have_r = 0;
    for (uint32_t i = 0; i < nv; i++) { /* :59-64 */
        uint64_t *msg = evals_out + (size_t)i * (d + 1) * 4;
        rc = orc_prove_round(p, have_r ? r.l : NULL, msg);
        if (rc) break;
        orc_rng_feed_prover_msg(rng, msg, d + 1); /* :61 */
        orc_rng_sample_fr(rng, r.l);              /* :63 */
        have_r = 1;
    }
    if (!rc) {
        p->randomness[p->n_rand++] = r; /* :65-67 */
        if (randomness_out) memcpy(randomness_out, p->randomness, (size_t)nv * sizeof(fr_t));
        if (final_tables_out)
            for (uint32_t j = 0; j < T; j++) memcpy(final_tables_out + (size_t)j * 8, p->tables[j], 2 * sizeof(fr_t));
    }
    orc_prover_free(p);
    if (!rng_in) orc_rng_free(rng);
    return rc;
}

size_t orc_serialize_proof(const uint64_t *evals, uint32_t nv, uint32_t d, uint8_t *out) {
    /* Proof<F> = Vec<ProverMsg<F>> (mod.rs:22): u64 LE count, then each ProverMsg */
    size_t o = 0;
    put_u64(out + o, nv); o += 8;
    for (uint32_t i = 0; i < nv; i++) {
        put_u64(out + o, d + 1); o += 8;
        for (uint32_t t = 0; t <= d; t++) { fr_to_bytes(out + o, FR(evals + ((size_t)i * (d + 1) + t) * 4)); o += 32; }
    }
    return o;
}

/* ------------------------------------------------------------------ verifier (self-check only) */
void orc_interpolate_uni_poly(uint64_t out[4], const uint64_t *p_i, uint32_t len, const uint64_t eval_at[4]) {
    /* verifier.rs:139-251 computes sum_i p_i * prod_{j!=i}(x-j)/(i-j) with three integer-width branches;
     * all are the same exact field value, restated here as plain Lagrange. */
    const fr_t *x = FR(eval_at);
    fr_t res = FR_ZERO;
    for (uint32_t i = 0; i < len; i++) {
        fr_t num = FR_ONE, den = FR_ONE;
        fr_t fi = fr_from_u64(i);
        for (uint32_t j = 0; j < len; j++) {
            if (j == i) continue;
            fr_t fj = fr_from_u64(j);
            fr_t a = fr_sub(x, &fj), b = fr_sub(&fi, &fj);
            num = fr_mul(&num, &a);
            den = fr_mul(&den, &b);
        }
        fr_t di = fr_inv(&den);
        fr_t term = fr_mul(&num, &di);
        term = fr_mul(&term, FR(p_i + 4 * i));
        res = fr_add(&res, &term);
    }
    memcpy(out, &res, 32);
}

static int check_and_generate_subclaim(uint32_t nv, uint32_t d, const uint64_t *evals, const fr_t *randomness,
                                       const fr_t *asserted_sum, fr_t *expected_out) {
    /* verifier.rs:90-121 */
    fr_t expected = *asserted_sum;
    for (uint32_t i = 0; i < nv; i++) {
        const uint64_t *ev = evals + (size_t)i * (d + 1) * 4;
        fr_t s = fr_add(FR(ev), FR(ev + 4));
        if (!fr_eq(&s, &expected)) return ORC_ERR_REJECT; /* :109-113 */
        orc_interpolate_uni_poly(expected.l, ev, d + 1, randomness[i].l); /* :114 */
    }
    *expected_out = expected;
    return ORC_OK;
}

int orc_ml_verify(orc_rng *rng_in, uint32_t nv, uint32_t d, const uint64_t claimed_sum[4], const uint64_t *evals,
                  uint64_t *point_out, uint64_t expected_out[4]) {
    /* MLSumcheck::verify (mod.rs:73-80) when rng_in == NULL, else verify_as_subprotocol (mod.rs:84-100) */
    orc_rng *rng = rng_in ? rng_in : orc_rng_setup();
    orc_rng_feed_poly_info(rng, d, nv); /* :90 */
    fr_t *rand = (fr_t *)malloc((nv ? nv : 1) * sizeof(fr_t));
    for (uint32_t i = 0; i < nv; i++) { /* :92-97 -> verifier.rs:54-83 */
        orc_rng_feed_prover_msg(rng, evals + (size_t)i * (d + 1) * 4, d + 1);
        orc_rng_sample_fr(rng, rand[i].l);
    }
    fr_t expected;
    int rc = check_and_generate_subclaim(nv, d, evals, rand, FR(claimed_sum), &expected);
    if (!rc) {
        if (point_out) memcpy(point_out, rand, (size_t)nv * sizeof(fr_t));
        memcpy(expected_out, &expected, 32);
    }
    free(rand);
    if (!rng_in) orc_rng_free(rng);
    return rc;
}

void orc_poly_evaluate(uint64_t out[4], uint32_t nv, uint32_t T, const uint64_t *const *tables, uint32_t n,
                       const uint64_t *coeffs, const uint32_t *offsets, const uint32_t *indices, const uint64_t *point) {
    /* data_structures.rs:99-109 */
    fr_t *tv = (fr_t *)malloc((T ? T : 1) * sizeof(fr_t));
    for (uint32_t j = 0; j < T; j++) orc_dense_evaluate(tv[j].l, tables[j], nv, point);
    fr_t sum = FR_ZERO;
    for (uint32_t k = 0; k < n; k++) {
        fr_t prod = FR_ONE;
        for (uint32_t jj = offsets[k]; jj < offsets[k + 1]; jj++) prod = fr_mul(&prod, &tv[indices[jj]]);
        prod = fr_mul(FR(coeffs + 4 * k), &prod);
        sum = fr_add(&sum, &prod);
    }
    memcpy(out, &sum, 32);
    free(tv);
}

/* ------------------------------------------------------------------ GKR round sumcheck */
void orc_precompute_eq(uint64_t *out, const uint64_t *g, uint32_t dim) {
    /* ark-poly sparse.rs precompute_eq (external): eq[b] = prod_j (b_j ? g_j : 1-g_j), bit j of b <-> g[j] */
    fr_t *dp = (fr_t *)out;
    if (dim == 0) { dp[0] = FR_ONE; return; }
    dp[0] = fr_sub(&FR_ONE, FR(g));
    dp[1] = *FR(g);
    for (uint32_t i = 1; i < dim; i++) {
        for (uint64_t b = 0; b < ((uint64_t)1 << i); b++) {
            fr_t prev = dp[b];
            dp[b + ((uint64_t)1 << i)] = fr_mul(&prev, FR(g + 4 * i));
            dp[b] = fr_sub(&prev, &dp[b + ((uint64_t)1 << i)]);
        }
    }
}

typedef struct { uint64_t idx; fr_t v; } sp_entry;
static int sp_cmp(const void *a, const void *b) {
    uint64_t x = ((const sp_entry *)a)->idx, y = ((const sp_entry *)b)->idx;
    return x < y ? -1 : (x > y ? 1 : 0);
}
/* ark-poly SparseMultilinearExtension::fix_variables(pt) (external; called at gkr mod.rs:31,62): fixes the LOW
 * len(pt) index bits: out[idx >> k] += eq(pt)[idx & (2^k-1)] * v.  (ark-poly walks pt in windows; the exact field
 * result is the same.)  Output: sorted unique entries (its BTreeMap), zero-valued entries kept.  Returns count. */
static size_t sparse_fix_low(size_t nnz, const uint64_t *idx, const fr_t *val, const fr_t *pt, uint32_t k,
                             uint64_t *idx_out, fr_t *val_out) {
    fr_t *eq = (fr_t *)malloc(((size_t)1 << k) * sizeof(fr_t));
    orc_precompute_eq(eq->l, pt->l, k);
    sp_entry *e = (sp_entry *)malloc((nnz ? nnz : 1) * sizeof(sp_entry));
    const uint64_t mask = (k >= 64) ? ~0ULL : (((uint64_t)1 << k) - 1);
    for (size_t i = 0; i < nnz; i++) {
        e[i].idx = idx[i] >> k;
        e[i].v = fr_mul(&eq[idx[i] & mask], &val[i]);
    }
    qsort(e, nnz, sizeof(sp_entry), sp_cmp);
    size_t n_out = 0;
    for (size_t i = 0; i < nnz; i++) {
        if (n_out && idx_out[n_out - 1] == e[i].idx) {
            val_out[n_out - 1] = fr_add(&val_out[n_out - 1], &e[i].v);
        } else {
            idx_out[n_out] = e[i].idx;
            val_out[n_out] = e[i].v;
            n_out++;
        }
    }
    free(e);
    free(eq);
    return n_out;
}

size_t orc_gkr_initialize_phase_one(uint32_t dim, size_t nnz, const uint64_t *f1_idx, const uint64_t *f1_val,
                                    const uint64_t *f3, const uint64_t *g, uint64_t *h_g_out, uint64_t *f1g_idx_out,
                                    uint64_t *f1g_val_out) {
    /* gkr_round_sumcheck/mod.rs:22-42 */
    size_t n_g = sparse_fix_low(nnz, f1_idx, FR(f1_val), FR(g), dim, f1g_idx_out, (fr_t *)f1g_val_out); /* :31 */
    fr_t *a_hg = (fr_t *)h_g_out;
    memset(a_hg, 0, ((size_t)1 << dim) * sizeof(fr_t)); /* :30 */
    const uint64_t mask = ((uint64_t)1 << dim) - 1;
    for (size_t i = 0; i < n_g; i++) { /* :32-38 */
        const fr_t *v = FR(f1g_val_out + 4 * i);
        if (!fr_is_zero(v)) {
            uint64_t xy = f1g_idx_out[i];
            uint64_t x = xy & mask, y = xy >> dim;
            fr_t t = fr_mul(v, FR(f3 + 4 * y));
            a_hg[x] = fr_add(&a_hg[x], &t);
        }
    }
    return n_g;
}

void orc_gkr_initialize_phase_two(uint32_t dim, size_t nnz_g, const uint64_t *f1g_idx, const uint64_t *f1g_val,
                                  const uint64_t *u, uint64_t *f1_gu_out) {
    /* mod.rs:57-63: f1_g.fix_variables(u).to_dense_multilinear_extension() */
    uint64_t *idx = (uint64_t *)malloc((nnz_g ? nnz_g : 1) * sizeof(uint64_t));
    fr_t *val = (fr_t *)malloc((nnz_g ? nnz_g : 1) * sizeof(fr_t));
    size_t n = sparse_fix_low(nnz_g, f1g_idx, FR(f1g_val), FR(u), dim, idx, val);
    fr_t *dense = (fr_t *)f1_gu_out;
    memset(dense, 0, ((size_t)1 << dim) * sizeof(fr_t));
    for (size_t i = 0; i < n; i++) dense[idx[i]] = val[i];
    free(idx);
    free(val);
}

static int run_sumcheck_2tables(orc_rng *rng, uint32_t dim, const fr_t *a, const fr_t *b, uint64_t *msgs_out,
                                fr_t *challenges_out) {
    /* start_phase{1,2}_sumcheck (mod.rs:45-54, 66-82): poly = 1 * (a * b); then the loop at mod.rs:111-119 / 126-133:
     * prove_round, rng.feed(&pm), sample_round.  NB no PolynomialInfo is fed. */
    const uint64_t *tabs[2] = {a->l, b->l};
    uint32_t offsets[2] = {0, 2}, indices[2] = {0, 1};
    orc_prover *p;
    int rc = orc_prover_init(&p, dim, 2, tabs, 1, FR_ONE.l, offsets, indices);
    if (rc) return rc;
    fr_t r;
    int have_r = 0;
    for (uint32_t i = 0; i < dim; i++) {
        uint64_t *msg = msgs_out + (size_t)i * 3 * 4;
        rc = orc_prove_round(p, have_r ? r.l : NULL, msg);
        if (rc) break;
        orc_rng_feed_prover_msg(rng, msg, 3);
        orc_rng_sample_fr(rng, r.l);
        have_r = 1;
        challenges_out[i] = r;
    }
    orc_prover_free(p);
    return rc;
}

int orc_gkr_prove(orc_rng *rng, uint32_t dim, size_t nnz, const uint64_t *f1_idx, const uint64_t *f1_val,
                  const uint64_t *f2, const uint64_t *f3, const uint64_t *g, uint64_t *phase1_out, uint64_t *phase2_out,
                  uint64_t *u_out, uint64_t *v_out) {
    /* mod.rs:93-139 */
    const size_t N = (size_t)1 << dim;
    fr_t *h_g = (fr_t *)malloc(N * sizeof(fr_t));
    uint64_t *f1g_idx = (uint64_t *)malloc((nnz ? nnz : 1) * sizeof(uint64_t));
    fr_t *f1g_val = (fr_t *)malloc((nnz ? nnz : 1) * sizeof(fr_t));
    size_t n_g = orc_gkr_initialize_phase_one(dim, nnz, f1_idx, f1_val, f3, g, h_g->l, f1g_idx, f1g_val->l); /* :106 */
    fr_t *u = (fr_t *)malloc((dim ? dim : 1) * sizeof(fr_t)), *v = (fr_t *)malloc((dim ? dim : 1) * sizeof(fr_t));
    int rc = run_sumcheck_2tables(rng, dim, h_g, FR(f2), phase1_out, u); /* :107-119 */
    if (!rc) {
        fr_t *f1_gu = (fr_t *)malloc(N * sizeof(fr_t));
        orc_gkr_initialize_phase_two(dim, n_g, f1g_idx, f1g_val->l, u->l, f1_gu->l); /* :121 */
        fr_t f2_u;
        orc_dense_evaluate(f2_u.l, f2, dim, u->l); /* :122 f2.evaluate(&u) */
        fr_t *f3_f2u = (fr_t *)malloc(N * sizeof(fr_t));
        for (size_t i = 0; i < N; i++) f3_f2u[i] = fr_mul(&f2_u, FR(f3 + 4 * i)); /* :71-75 zero += (f2_u, f3) */
        rc = run_sumcheck_2tables(rng, dim, f1_gu, f3_f2u, phase2_out, v);         /* :122-133 */
        free(f1_gu);
        free(f3_f2u);
    }
    if (!rc && u_out) memcpy(u_out, u, dim * sizeof(fr_t));
    if (!rc && v_out) memcpy(v_out, v, dim * sizeof(fr_t));
    free(h_g); free(f1g_idx); free(f1g_val); free(u); free(v);
    return rc;
}

int orc_gkr_verify(orc_rng *rng, uint32_t dim, const uint64_t *phase1, const uint64_t *phase2,
                   const uint64_t claimed_sum[4], uint64_t *u_out, uint64_t *v_out, uint64_t expected_out[4]) {
    /* mod.rs:147-192; max_multiplicands = 2 both phases */
    fr_t *u = (fr_t *)malloc((dim ? dim : 1) * sizeof(fr_t)), *v = (fr_t *)malloc((dim ? dim : 1) * sizeof(fr_t));
    for (uint32_t i = 0; i < dim; i++) { /* :158-162 */
        orc_rng_feed_prover_msg(rng, phase1 + (size_t)i * 12, 3);
        orc_rng_sample_fr(rng, u[i].l);
    }
    fr_t e1, e2;
    int rc = check_and_generate_subclaim(dim, 2, phase1, u, FR(claimed_sum), &e1); /* :163 */
    if (!rc) {
        for (uint32_t i = 0; i < dim; i++) { /* :170-174 */
            orc_rng_feed_prover_msg(rng, phase2 + (size_t)i * 12, 3);
            orc_rng_sample_fr(rng, v[i].l);
        }
        rc = check_and_generate_subclaim(dim, 2, phase2, v, &e1, &e2); /* :175-178 */
    }
    if (!rc) {
        if (u_out) memcpy(u_out, u, dim * sizeof(fr_t));
        if (v_out) memcpy(v_out, v, dim * sizeof(fr_t));
        memcpy(expected_out, &e2, 32);
    }
    free(u); free(v);
    return rc;
}

int orc_gkr_verify_subclaim(uint32_t dim, size_t nnz, const uint64_t *f1_idx, const uint64_t *f1_val,
                            const uint64_t *f2, const uint64_t *f3, const uint64_t *g, const uint64_t *u,
                            const uint64_t *v, const uint64_t expected[4]) {
    /* gkr data_structures.rs:33-56: f1.evaluate(g|u|v) * f2.evaluate(u) * f3.evaluate(v) == expected */
    uint64_t *i1 = (uint64_t *)malloc((nnz ? nnz : 1) * 8), *i2 = (uint64_t *)malloc((nnz ? nnz : 1) * 8);
    fr_t *v1 = (fr_t *)malloc((nnz ? nnz : 1) * sizeof(fr_t)), *v2 = (fr_t *)malloc((nnz ? nnz : 1) * sizeof(fr_t));
    size_t n1 = sparse_fix_low(nnz, f1_idx, FR(f1_val), FR(g), dim, i1, v1);
    size_t n2 = sparse_fix_low(n1, i1, v1, FR(u), dim, i2, v2);
    size_t n3 = sparse_fix_low(n2, i2, v2, FR(v), dim, i1, v1);
    fr_t f1e = FR_ZERO;
    for (size_t i = 0; i < n3; i++) if (i1[i] == 0) f1e = v1[i];
    fr_t f2e, f3e;
    orc_dense_evaluate(f2e.l, f2, dim, u);
    orc_dense_evaluate(f3e.l, f3, dim, v);
    fr_t actual = fr_mul(&f1e, &f2e);
    actual = fr_mul(&actual, &f3e);
    free(i1); free(i2); free(v1); free(v2);
    return fr_eq(&actual, FR(expected));
}

/* ------------------------------------------------------------------ synthetic inputs (generator ours) */
static inline uint64_t sm64_mix(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
void orc_synth_table(uint64_t *out, uint64_t n_elems, uint64_t seed) {
    /* counter-based SplitMix64: element e, attempt k (k<8), limb i uses counter ((e*8+k)*4+i+1).  Accept the first
     * attempt whose value (top bit masked) is < p; if all 8 fail (p ~ 6e-9) mask the top TWO bits of attempt 7. */
    const uint64_t GAMMA = 0x9e3779b97f4a7c15ULL;
    for (uint64_t e = 0; e < n_elems; e++) {
        fr_t t;
        for (uint64_t k = 0; k < 8; k++) {
            for (uint64_t i = 0; i < 4; i++) t.l[i] = sm64_mix(seed + GAMMA * (((e * 8 + k) * 4) + i + 1));
            t.l[3] &= 0x7fffffffffffffffULL;
            if (!fr_geq_p(&t)) break;
            if (k == 7) t.l[3] &= 0x3fffffffffffffffULL;
        }
        memcpy(out + 4 * e, &t, 32);
    }
}
