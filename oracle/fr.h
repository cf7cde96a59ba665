/* ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or called from the product path
 * (sumcheck_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use anything under oracle/.
 *
 * BLS12-381 scalar field Fr, restating what the reference gets from the un-vendored crate `ark-ff`
 * (Cargo.toml:19,63 of /root/reference — semver 0.4.0 patched to un-pinned git master of
 * arkworks-rs/algebra; no Cargo.lock, so NO pinned version exists).  PARITY UNPINNED: the reference
 * holds no golden vectors for this path; this restatement is cross-checked
 * against an independent Python big-int model (oracle/pymodel.py) instead.
 *
 * Representation (ark-ff `Fp<MontBackend<FrConfig,4>,4>`): 4 x u64 little-endian limbs, Montgomery
 * form with R = 2^256, always fully reduced to [0,p).
 */
#ifndef ORACLE_FR_H
#define ORACLE_FR_H
#include <stdint.h>
#include <string.h>

typedef struct { uint64_t l[4]; } fr_t;

static const fr_t FR_P = {{0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL}};
static const fr_t FR_ONE = {{0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL}}; /* R mod p */
static const fr_t FR_R2 = {{0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL}};  /* R^2 mod p */
static const fr_t FR_ZERO = {{0, 0, 0, 0}};
#define FR_INV64 0xfffffffeffffffffULL /* -p^{-1} mod 2^64 */

typedef unsigned __int128 u128;

static inline int fr_geq_p(const fr_t *a) {
    for (int i = 3; i >= 0; i--) {
        if (a->l[i] > FR_P.l[i]) return 1;
        if (a->l[i] < FR_P.l[i]) return 0;
    }
    return 1;
}
static inline int fr_is_zero(const fr_t *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fr_eq(const fr_t *a, const fr_t *b) { return memcmp(a, b, sizeof(fr_t)) == 0; }

static inline void fr_sub_p(fr_t *a) {
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a->l[i] - FR_P.l[i] - borrow;
        a->l[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
}
static inline fr_t fr_add(const fr_t *a, const fr_t *b) {
    fr_t r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a->l[i] + b->l[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
    }
    /* 2p < 2^256: no carry out of limb 3 */
    if (fr_geq_p(&r)) fr_sub_p(&r);
    return r;
}
static inline fr_t fr_sub(const fr_t *a, const fr_t *b) {
    fr_t r;
    u128 borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a->l[i] - b->l[i] - borrow;
        r.l[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
    if (borrow) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)r.l[i] + FR_P.l[i];
            r.l[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    return r;
}
static inline fr_t fr_neg(const fr_t *a) { return fr_sub(&FR_ZERO, a); }

/* Montgomery product a*b*R^{-1} mod p — CIOS, 4 x 64-bit limbs (what ark-ff's non-asm MontBackend computes). */
static inline fr_t fr_mul(const fr_t *a, const fr_t *b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * FR_INV64;
        c = (u128)m * FR_P.l[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * FR_P.l[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    fr_t r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || fr_geq_p(&r)) fr_sub_p(&r);
    return r;
}
/* canonical integer (4 limbs, < p) -> Montgomery */
static inline fr_t fr_from_canonical(const fr_t *c) { return fr_mul(c, &FR_R2); }
/* Montgomery -> canonical integer */
static inline fr_t fr_to_canonical(const fr_t *a) {
    fr_t one = {{1, 0, 0, 0}};
    return fr_mul(a, &one);
}
static inline fr_t fr_from_u64(uint64_t v) {
    fr_t c = {{v, 0, 0, 0}};
    return fr_from_canonical(&c);
}
fr_t fr_pow(const fr_t *a, const fr_t *e_canonical);
fr_t fr_inv(const fr_t *a);
/* ark-serialize CanonicalSerialize for Fp (no flags for Fr): 32 bytes LE of the canonical integer. */
void fr_to_bytes(uint8_t out[32], const fr_t *a);
#endif
