// BLAKE2b-512 (RFC 7693, unkeyed) + the reference's hash-chain PRG `Blake2b512Rng`
// (/root/reference/src/rng.rs:22-81), usable from host and device code.
// The transcript normally lives on the host (the drop-in keeps Fiat-Shamir in the caller's language); the
// __device__ build is what the fused tail kernel uses so small rounds need no host round-trip.
#pragma once
#include <cstddef>
#include <cstdint>

#ifdef __CUDACC__
#define SC_HD __host__ __device__
#else
#define SC_HD
#endif

namespace b2 {

// Plain-data hasher state: this struct IS the C-ABI type sc_blake2b512_rng (include/sumcheck_b200.h).
struct State {
    uint64_t h[8];
    uint64_t t[2];
    uint8_t buf[128];
    uint64_t buflen;
};

SC_HD inline uint64_t rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }

SC_HD inline uint64_t iv(int i) {
    const uint64_t IV[8] = {0x6a09e667f3bcc908ULL, 0xbb67ae8584caa73bULL, 0x3c6ef372fe94f82bULL, 0xa54ff53a5f1d36f1ULL,
                            0x510e527fade682d1ULL, 0x9b05688c2b3e6c1fULL, 0x1f83d9abfb41bd6bULL, 0x5be0cd19137e2179ULL};
    return IV[i];
}

#define B2_G(a, b, c, d, x, y)      \
    do {                            \
        a = a + b + (x);            \
        d = rotr64(d ^ a, 32);      \
        c = c + d;                  \
        b = rotr64(b ^ c, 24);      \
        a = a + b + (y);            \
        d = rotr64(d ^ a, 16);      \
        c = c + d;                  \
        b = rotr64(b ^ c, 63);      \
    } while (0)

#define B2_ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    do {                                                                                \
        B2_G(v0, v4, v8, v12, m[s0], m[s1]);                                            \
        B2_G(v1, v5, v9, v13, m[s2], m[s3]);                                            \
        B2_G(v2, v6, v10, v14, m[s4], m[s5]);                                           \
        B2_G(v3, v7, v11, v15, m[s6], m[s7]);                                           \
        B2_G(v0, v5, v10, v15, m[s8], m[s9]);                                           \
        B2_G(v1, v6, v11, v12, m[s10], m[s11]);                                         \
        B2_G(v2, v7, v8, v13, m[s12], m[s13]);                                          \
        B2_G(v3, v4, v9, v14, m[s14], m[s15]);                                          \
    } while (0)

// One compression F(h, block, t, last).  Message schedule unrolled so that m[] stays in registers on the device.
SC_HD inline void compress(uint64_t h[8], const uint8_t block[128], uint64_t t0, uint64_t t1, bool last) {
    uint64_t m[16];
    for (int i = 0; i < 16; i++) {
        uint64_t w = 0;
        for (int j = 7; j >= 0; j--) w = (w << 8) | block[8 * i + j];
        m[i] = w;
    }
    uint64_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    uint64_t v8 = iv(0), v9 = iv(1), v10 = iv(2), v11 = iv(3), v12 = iv(4) ^ t0, v13 = iv(5) ^ t1,
             v14 = last ? ~iv(6) : iv(6), v15 = iv(7);
    B2_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    B2_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    B2_ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4);
    B2_ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8);
    B2_ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13);
    B2_ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9);
    B2_ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11);
    B2_ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10);
    B2_ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5);
    B2_ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0);
    B2_ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    B2_ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3);
    h[0] ^= v0 ^ v8;  h[1] ^= v1 ^ v9;  h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}

SC_HD inline void init(State* s) {  // Blake2b512::new(): rng.rs:30-34
    for (int i = 0; i < 8; i++) s->h[i] = iv(i);
    s->h[0] ^= 0x01010000ULL ^ 64;
    s->t[0] = s->t[1] = 0;
    for (int i = 0; i < 128; i++) s->buf[i] = 0;
    s->buflen = 0;
}

SC_HD inline void update(State* s, const uint8_t* in, size_t n) {  // Digest::update: rng.rs:39
    while (n > 0) {
        if (s->buflen == 128) {
            s->t[0] += 128;
            if (s->t[0] < 128) s->t[1]++;
            compress(s->h, s->buf, s->t[0], s->t[1], false);
            s->buflen = 0;
        }
        size_t take = 128 - (size_t)s->buflen;
        if (take > n) take = n;
        for (size_t i = 0; i < take; i++) s->buf[s->buflen + i] = in[i];
        s->buflen += take;
        in += take;
        n -= take;
    }
}

SC_HD inline void finalize_copy(const State* s, uint8_t out[64]) {  // digest.clone().finalize(): rng.rs:62-63
    uint64_t h[8];
    uint8_t blk[128];
    for (int i = 0; i < 8; i++) h[i] = s->h[i];
    uint64_t t0 = s->t[0] + s->buflen, t1 = s->t[1] + (t0 < s->buflen ? 1 : 0);
    for (int i = 0; i < 128; i++) blk[i] = (uint64_t)i < s->buflen ? s->buf[i] : 0;
    compress(h, blk, t0, t1, true);
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(h[i] >> (8 * j));
}

// RngCore::try_fill_bytes: rng.rs:61-80, statement for statement.
SC_HD inline void fill_bytes(State* s, uint8_t* dest, size_t n) {
    uint8_t output[64];
    finalize_copy(s, output);
    size_t ptr = 0, digest_ptr = 0;
    while (ptr < n) {
        dest[ptr] = output[digest_ptr];
        ptr++;
        digest_ptr++;
        if (digest_ptr == 64) {
            update(s, output, 64);
            finalize_copy(s, output);
            digest_ptr = 0;
        }
    }
    update(s, output, 64);
}

SC_HD inline uint64_t next_u64(State* s) {  // rng.rs:51-55
    uint8_t t[8];
    fill_bytes(s, t, 8);
    uint64_t v = 0;
    for (int j = 7; j >= 0; j--) v = (v << 8) | t[j];
    return v;
}

// IPForMLSumcheck::sample_round (verifier.rs:128-132) = F::rand(rng) of ark-ff (external crate): draw 4 u64 limbs in
// order, clear the top 256-255 = 1 bit, accept when < p; the accepted limbs are used AS the Montgomery representation.
SC_HD inline void sample_fr(State* s, uint64_t out[4]) {
    const uint64_t P[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
    for (;;) {
        for (int i = 0; i < 4; i++) out[i] = next_u64(s);
        out[3] &= 0x7fffffffffffffffULL;
        bool lt = false;
        for (int i = 3; i >= 0; i--) {
            if (out[i] < P[i]) { lt = true; break; }
            if (out[i] > P[i]) break;
        }
        if (lt) return;
    }
}

SC_HD inline void put_u64(uint8_t* o, uint64_t v) {
    for (int j = 0; j < 8; j++) o[j] = (uint8_t)(v >> (8 * j));
}

}  // namespace b2

// ---------------------------------------------------------------------------------------------------------------
// Word-oriented device transcript for the fused tail kernel (kernels.cuh, tail_kernel): every message of
// MLSumcheck::prove is a whole number of 64-bit words (PolynomialInfo 16 B, ProverMsg 8+32(d+1) B, digests 64 B), so the
// 128-byte block buffer is kept as 16 u64 words in shared memory and one thread drives it.  Only usable when the
// state handed over by the host has buflen % 8 == 0 (otherwise the host keeps the transcript).
#ifdef __CUDACC__
namespace b2w {

struct WState {       // lives in shared memory
    uint64_t h[8];
    uint64_t t0;      // bytes compressed so far (low word; a sumcheck transcript never reaches 2^64 bytes)
    uint64_t buf[16];
    uint32_t nwords;  // words currently in buf (0..16)
    uint64_t fh[8];   // scratch for finalisation: chaining value copy
    uint64_t fblk[16];  //                         zero-padded last block
};

__constant__ uint8_t SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};

// One compression, ROLLED over the 12 rounds and never inlined: the tail kernel calls it ~7 times per protocol round
// from a single thread, so what matters is that its ~300 instructions stay hot in the instruction cache (the fully
// unrolled form is ~40 KB of straight-line code and ran 10x slower because every fetch missed).
// m: the 16 message words (shared memory; indexed through SIGMA).
static __device__ __noinline__ void compress_words(uint64_t* h, const uint64_t* m, uint64_t t0, bool last) {
    using b2::rotr64;
    uint64_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    uint64_t v8 = b2::iv(0), v9 = b2::iv(1), v10 = b2::iv(2), v11 = b2::iv(3), v12 = b2::iv(4) ^ t0, v13 = b2::iv(5),
             v14 = last ? ~b2::iv(6) : b2::iv(6), v15 = b2::iv(7);
#pragma unroll 1
    for (int r = 0; r < 12; r++) {
        const uint8_t* s = SIGMA[r];
        B2_G(v0, v4, v8, v12, m[s[0]], m[s[1]]);
        B2_G(v1, v5, v9, v13, m[s[2]], m[s[3]]);
        B2_G(v2, v6, v10, v14, m[s[4]], m[s[5]]);
        B2_G(v3, v7, v11, v15, m[s[6]], m[s[7]]);
        B2_G(v0, v5, v10, v15, m[s[8]], m[s[9]]);
        B2_G(v1, v6, v11, v12, m[s[10]], m[s[11]]);
        B2_G(v2, v7, v8, v13, m[s[12]], m[s[13]]);
        B2_G(v3, v4, v9, v14, m[s[14]], m[s[15]]);
    }
    h[0] ^= v0 ^ v8;  h[1] ^= v1 ^ v9;  h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}

__device__ inline void from_state(WState* w, const b2::State* s) {  // requires s->buflen % 8 == 0
    for (int i = 0; i < 8; i++) w->h[i] = s->h[i];
    w->t0 = s->t[0];
    w->nwords = (uint32_t)(s->buflen / 8);
    for (int i = 0; i < 16; i++) {
        uint64_t v = 0;
        for (int j = 7; j >= 0; j--) v = (v << 8) | s->buf[8 * i + j];
        w->buf[i] = v;
    }
}
__device__ inline void to_state(const WState* w, b2::State* s) {
    for (int i = 0; i < 8; i++) s->h[i] = w->h[i];
    s->t[0] = w->t0;
    s->t[1] = 0;
    s->buflen = (uint64_t)w->nwords * 8;
    for (int i = 0; i < 16; i++)
        for (int j = 0; j < 8; j++) s->buf[8 * i + j] = (i < (int)w->nwords) ? (uint8_t)(w->buf[i] >> (8 * j)) : 0;
}

static __device__ __noinline__ void absorb_word(WState* w, uint64_t x) {  // Digest::update, one word
    if (w->nwords == 16) {
        w->t0 += 128;
        compress_words(w->h, w->buf, w->t0, false);
        w->nwords = 0;
    }
    w->buf[w->nwords++] = x;
}

// rng.rs:51-55 + 61-80 for an 8-byte request: out = first word of H(state); then absorb the whole 64-byte digest
static __device__ __noinline__ uint64_t next_u64(WState* w) {
    for (int i = 0; i < 8; i++) w->fh[i] = w->h[i];
    for (int i = 0; i < 16; i++) w->fblk[i] = (i < (int)w->nwords) ? w->buf[i] : 0;
    compress_words(w->fh, w->fblk, w->t0 + (uint64_t)w->nwords * 8, true);
    for (int i = 0; i < 8; i++) absorb_word(w, w->fh[i]);
    return w->fh[0];
}

// verifier.rs:128-132 -> ark-ff Fp::rand (see b2::sample_fr)
static __device__ __noinline__ void sample_fr(WState* w, uint64_t out[4]) {
    const uint64_t P[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
    for (;;) {
        for (int i = 0; i < 4; i++) out[i] = next_u64(w);
        out[3] &= 0x7fffffffffffffffULL;
        bool lt = false;
        for (int i = 3; i >= 0; i--) {
            if (out[i] < P[i]) { lt = true; break; }
            if (out[i] > P[i]) break;
        }
        if (lt) return;
    }
}

}  // namespace b2w
#endif
