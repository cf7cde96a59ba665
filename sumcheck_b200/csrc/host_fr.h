// Scalar BLS12-381 Fr arithmetic on the HOST, for the O(d^2) per-round glue around the device sums — what
// prove_round does after its parallel reduce (prover.rs:138-153) plus the P(1) = P_prev(r) - P(0) identity
// (verifier.rs:109-114, interpolate_uni_poly at verifier.rs:139-251).  The device delivers the d (or d+1) raw sums of
// a round; scaling by a deferred coefficient, the claim from the previous message and the canonical (serialised) form
// are ~25 multiplications — 1 us here against ~7 us of dependent latency on a single GPU warp.  Not a CPU path for the
// protocol: tables never come here.  4 x u64 limbs, Montgomery R = 2^256 — the layout of ark-ff's Fp<MontBackend<_,4>,4>.
#pragma once
#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

namespace hfr {

struct F {
    uint64_t l[4];
};

static const uint64_t P[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static const uint64_t INV = 0xfffffffeffffffffULL;  // -p^-1 mod 2^64
static const F ONE = {{0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL}};  // R mod p
static const F R2 = {{0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL}};   // R^2 mod p

inline bool geq_p(const uint64_t (&a)[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > P[i]) return true;
        if (a[i] < P[i]) return false;
    }
    return true;
}
inline void sub_p(uint64_t (&a)[4]) {
    unsigned __int128 bw = 0;
    for (int i = 0; i < 4; i++) {
        unsigned __int128 d = (unsigned __int128)a[i] - P[i] - (uint64_t)bw;
        a[i] = (uint64_t)d;
        bw = (d >> 64) & 1;
    }
}
inline F add(const F& a, const F& b) {
    F r;
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (unsigned __int128)a.l[i] + b.l[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq_p(r.l)) sub_p(r.l);  // a + b < 2p < 2^256: no carry out
    return r;
}
inline F sub(const F& a, const F& b) {
    F r;
    unsigned __int128 bw = 0;
    for (int i = 0; i < 4; i++) {
        unsigned __int128 d = (unsigned __int128)a.l[i] - b.l[i] - (uint64_t)bw;
        r.l[i] = (uint64_t)d;
        bw = (d >> 64) & 1;
    }
    if (bw) {
        unsigned __int128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (unsigned __int128)r.l[i] + P[i];
            r.l[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    return r;
}
// Montgomery product a*b*2^-256 mod p (CIOS, 64-bit digits)
inline F mul(const F& a, const F& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        unsigned __int128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (unsigned __int128)a.l[j] * b.l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        const uint64_t m = t[0] * INV;
        c = (unsigned __int128)m * P[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (unsigned __int128)m * P[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    F r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_p(r.l)) sub_p(r.l);
    return r;
}
inline F from_u64(uint64_t k) {
    F raw = {{k, 0, 0, 0}};
    return mul(raw, R2);
}
inline F to_canonical(const F& a) {  // the integer ark-serialize writes
    F one_int = {{1, 0, 0, 0}};
    return mul(a, one_int);
}
inline F inverse(const F& a) {  // a^(p-2)
    static const uint64_t E[4] = {0xfffffffeffffffffULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
    F acc = ONE;
    for (int i = 255; i >= 0; i--) {
        acc = mul(acc, acc);
        if ((E[i >> 6] >> (i & 63)) & 1) acc = mul(acc, a);
    }
    return acc;
}

// V * 2^-256 mod p for an unreduced 17-limb (32-bit) integer V = V0 + V1 * 2^256 + v16 * 2^512 — the sum of plain 256x256-bit
// products of Montgomery-form operands that the resident kernel delivers (kernels.cuh block_sum_wide); the device's
// fr::wide_reduce computes the same field element.  V0 * R^-1 = mul(V0, 1), V1 * 2^256 * R^-1 = V1, v16 * 2^512 * R^-1 = mul(v16, R^2).
inline F reduce_wide17(const uint32_t* v) {
    F v0, v1, top = {{v[16], 0, 0, 0}};
    for (int i = 0; i < 4; i++) {
        v0.l[i] = (uint64_t)v[2 * i] | ((uint64_t)v[2 * i + 1] << 32);
        v1.l[i] = (uint64_t)v[8 + 2 * i] | ((uint64_t)v[8 + 2 * i + 1] << 32);
    }
    for (int k = 0; k < 2; k++) {  // 2^256 < 3p
        if (geq_p(v0.l)) sub_p(v0.l);
        if (geq_p(v1.l)) sub_p(v1.l);
    }
    const F one_int = {{1, 0, 0, 0}};
    return add(add(mul(v0, one_int), v1), mul(top, R2));
}

// Lagrange weights w_j = 1 / prod_{k != j} (j - k) for the nodes 0..d (cached per degree; d <= 32)
inline const std::vector<F>& lagrange_weights(uint32_t d) {
    static std::vector<F> cache[33];
    static std::mutex mu;  // handles may live on different host threads
    std::lock_guard<std::mutex> lk(mu);
    std::vector<F>& w = cache[d];
    if (w.empty()) {
        std::vector<F> fresh(d + 1);
        for (uint32_t j = 0; j <= d; j++) {
            F den = ONE;
            for (uint32_t k = 0; k <= d; k++)
                if (k != j) den = mul(den, sub(from_u64(j), from_u64(k)));
            fresh[j] = inverse(den);
        }
        w.swap(fresh);
    }
    return w;
}

// Alternative evaluation points -> the message P(0..d) (kernels.cuh consume_pair_acc ALT).  The device sums the product
// polynomial at the finite points F_d = the first d of (0, 1, -1, 2, -2) and delivers c_d, the coefficient of t^d ("the point
// at infinity").  R(t) = P(t) - c_d t^d has degree < d and is known on the d points of F_d, so
//     P(t) = sum_f L_f(t) R(f) + c_d t^d,    L_f(t) = prod_{g != f} (t - g) / (f - g)
// for t = 0..d — exact field arithmetic, hence the element the reference computes.  fin[i] = P(F_d[i]).  d <= 5.
inline F from_i64(long long k) {
    return k >= 0 ? from_u64((uint64_t)k) : sub(F{{0, 0, 0, 0}}, from_u64((uint64_t)(-k)));
}
inline void alt_to_standard(uint32_t d, const F* fin, const F& cd, F* out) {
    static const long long NODES[5] = {0, 1, -1, 2, -2};
    static std::vector<F> cache[6];  // [d]: (d+1) x d basis values L_f(t), then (d+1) powers t^d
    static std::mutex mu;
    {
        std::lock_guard<std::mutex> lk(mu);
        std::vector<F>& m = cache[d];
        if (m.empty()) {
            std::vector<F> fresh((size_t)(d + 1) * d + (d + 1));
            for (uint32_t t = 0; t <= d; t++) {
                for (uint32_t f = 0; f < d; f++) {
                    F num = ONE, den = ONE;
                    for (uint32_t g = 0; g < d; g++) {
                        if (g == f) continue;
                        num = mul(num, from_i64((long long)t - NODES[g]));
                        den = mul(den, from_i64(NODES[f] - NODES[g]));
                    }
                    fresh[(size_t)t * d + f] = mul(num, inverse(den));
                }
                F pw = ONE;
                for (uint32_t e = 0; e < d; e++) pw = mul(pw, from_u64(t));
                fresh[(size_t)(d + 1) * d + t] = pw;
            }
            m.swap(fresh);
        }
    }
    const std::vector<F>& m = cache[d];
    F r[5];
    for (uint32_t f = 0; f < d; f++) {  // R(f) = P(f) - c_d f^d
        F pw = ONE;
        const F node = from_i64(NODES[f]);
        for (uint32_t e = 0; e < d; e++) pw = mul(pw, node);
        r[f] = sub(fin[f], mul(cd, pw));
    }
    for (uint32_t t = 0; t <= d; t++) {
        F acc = mul(cd, m[(size_t)(d + 1) * d + t]);
        for (uint32_t f = 0; f < d; f++) acc = add(acc, mul(m[(size_t)t * d + f], r[f]));
        out[t] = acc;
    }
}

// The value at r of the degree-d polynomial through (j, evals[j]), j = 0..d (interpolate_uni_poly, verifier.rs:139).
inline F interpolate(const F* evals, uint32_t d, const F& r) {
    const std::vector<F>& w = lagrange_weights(d);
    F diff[33], pre[34], suf[34];
    for (uint32_t k = 0; k <= d; k++) diff[k] = sub(r, from_u64(k));
    pre[0] = ONE;
    for (uint32_t k = 0; k <= d; k++) pre[k + 1] = mul(pre[k], diff[k]);
    suf[d + 1] = ONE;
    for (uint32_t k = d + 1; k-- > 0;) suf[k] = mul(suf[k + 1], diff[k]);
    F acc = {{0, 0, 0, 0}};
    for (uint32_t j = 0; j <= d; j++) acc = add(acc, mul(mul(evals[j], w[j]), mul(pre[j], suf[j + 1])));
    return acc;
}

// ---- tensor-core contraction rounds (gemm_sum.cuh) ------------------------------------------------------------------------------
// V mod p (a RAW residue, not Montgomery) of an n-limb (32-bit) integer: Horner over 256-bit chunks, acc * 2^256 = mul(acc, R^2).
inline F reduce_limbs(const uint32_t* v, uint32_t n) {
    F acc = {{0, 0, 0, 0}};
    for (int c = (int)((n + 7) / 8) - 1; c >= 0; c--) {
        F chunk = {{0, 0, 0, 0}};
        for (uint32_t i = 0; i < 8 && (uint32_t)c * 8 + i < n; i++) chunk.l[i >> 1] |= (uint64_t)v[(uint32_t)c * 8 + i] << (32 * (i & 1));
        for (int k = 0; k < 2; k++)  // 2^256 < 3p
            if (geq_p(chunk.l)) sub_p(chunk.l);
        acc = add(mul(acc, R2), chunk);
    }
    return acc;
}

// One product of m multiplicands split into an X side of kx and a Y side of ky = m - kx multiplicands (kx, ky in {1, 2}).  A side
// of ONE table contributes the values (a, b) = (table[2b], table[2b+1]): its line is (1-t) a + t b.  A side of TWO tables
// contributes (q0, q1, qs) = (a a', b b', (a+b)(a'+b')): the product of its two lines is (1-t)^2 q0 + t(1-t)(qs - q0 - q1) + t^2 q1.
// z[i * ny + j] = limbs of sum_b X_i * Y_j (n_limbs each, plain integers of Montgomery-form operands, i.e. value * R^m).
// out[t] = P(t) for t = 0..d in Montgomery form (unscaled by any deferred coefficient).
inline void side_weights(uint32_t k, long long t, F* w) {
    if (k == 1) {
        w[0] = from_i64(1 - t);
        w[1] = from_i64(t);
    } else {
        w[0] = from_i64((1 - t) * (1 - 2 * t));
        w[1] = from_i64(t * (2 * t - 1));
        w[2] = from_i64(t * (1 - t));
    }
}
// W[t][i * ny + j] = wX_i(t) * wY_j(t) * R^(2-m), so that mul(W, z) = weight * value * R for a raw residue
// z = value * R^m: the weights, the division by R^(m-1) and the conversion cost ONE multiplication per (t, i, j).  Cached per
// shape (this routine sits between two rounds of every proof: ~110 Montgomery multiplications before, ~50 now).
struct GemmWeights {
    F w[5][9];
    uint32_t d = 0;
};
inline const GemmWeights& gemm_weights(uint32_t kx, uint32_t ky) {
    static GemmWeights cache[3][3];
    static bool ready[3][3] = {};
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    GemmWeights& g = cache[kx][ky];
    if (!ready[kx][ky]) {
        const uint32_t nx = kx == 1 ? 2 : 3, ny = ky == 1 ? 2 : 3, m = kx + ky;
        const F one_int = {{1, 0, 0, 0}};
        g.d = m;
        for (uint32_t t = 0; t <= m; t++) {
            F wx[3], wy[3];
            side_weights(kx, (long long)t, wx);
            side_weights(ky, (long long)t, wy);
            for (uint32_t i = 0; i < nx; i++)
                for (uint32_t j = 0; j < ny; j++) {
                    F w = mul(wx[i], wy[j]);                                  // weight * R
                    for (uint32_t k = 1; k < m; k++) w = mul(w, one_int);     // weight * R^(2-m): mul(w, z) = w z / R = weight * value * R
                    g.w[t][i * ny + j] = w;
                }
        }
        ready[kx][ky] = true;
    }
    return g;
}
inline void gemm_finish(const uint32_t* z, uint32_t n_limbs, uint32_t kx, uint32_t ky, uint32_t d, F* out) {
    const uint32_t nx = kx == 1 ? 2 : 3, ny = ky == 1 ? 2 : 3;
    const GemmWeights& g = gemm_weights(kx, ky);
    F Z[9];
    for (uint32_t i = 0; i < nx * ny; i++) Z[i] = reduce_limbs(z + (size_t)i * n_limbs, n_limbs);  // value * R^m mod p (raw residue)
    for (uint32_t t = 0; t <= d; t++) {
        F acc = {{0, 0, 0, 0}};
        for (uint32_t i = 0; i < nx * ny; i++) acc = add(acc, mul(g.w[t][i], Z[i]));
        out[t] = acc;
    }
}
// a += b for n-limb integers (the pieces of a round summed by several launches)
inline void add_limbs(uint32_t* a, const uint32_t* b, uint32_t n) {
    uint64_t c = 0;
    for (uint32_t i = 0; i < n; i++) {
        c += (uint64_t)a[i] + b[i];
        a[i] = (uint32_t)c;
        c >>= 32;
    }
}

}  // namespace hfr
