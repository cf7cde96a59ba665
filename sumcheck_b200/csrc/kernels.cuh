// Round kernels of the sumcheck prover hot path.
//
// round_kernel<NPTS, FOLD> is ONE launch per protocol round and fuses the two steps of
// IPForMLSumcheck::prove_round (/root/reference/src/ml_sumcheck/protocol/prover.rs:74-153):
//   (a2) the fix_variables fold of every table on the previous challenge r (prover.rs:85-89 -> ark-poly
//        DenseMultilinearExtension::fix_variables: new[b] = old[2b] + r*(old[2b+1]-old[2b])), and
//   (a1) the (d+1)-point multiply-reduce over the folded tables (prover.rs:110-148).
// A thread that owns output pair-index b reads old[4b..4b+3] of each table (128 contiguous bytes), folds them to
// new[2b], new[2b+1] in registers, stores those 64 bytes, and feeds them straight into the product terms, so every
// table element is read from HBM exactly once per round and written once (SURVEY.md §8d "algorithmic bytes").
//
// Exactness: field arithmetic is exact and values are kept fully reduced, so factoring the coefficient out of a
// single product and summing in a different order give the same limbs the reference produces.
#pragma once
#include <cstdint>

#include "blake2b.cuh"
#include "fr.cuh"

namespace sck {

using fr::Fr;

constexpr int MAX_NPTS = 5;  // evaluation points handled by one launch; d+1 > 5 is split into several launches

struct RoundParams {
    const uint32_t* const* tab_in;  // [T] current tables (device pointers), length 4*n_pairs (FOLD) or 2*n_pairs
    uint32_t* const* tab_out;       // [T] folded tables (FOLD), length 2*n_pairs
    const uint32_t* prod_offsets;   // CSR of ProverState.list_of_products (prover.rs:24)
    const uint32_t* prod_indices;
    const uint8_t* prod_first;      // 1 where a CSR entry is the first use of its table (that use stores the fold)
    const uint32_t* coeffs;         // [n_products][8]
    uint32_t n_products;
    uint32_t n_tables;
    uint32_t defer_coeff;           // single product: multiply the sums by coeffs[0] once at the end
    uint32_t t0;                    // first evaluation point of this launch
    uint32_t write_fold;            // store folded tables (only the t0 == 0 launch of a round does)
    unsigned long long n_pairs;     // 2^(nv - round)
    uint32_t r[8];                  // challenge of the previous round (FOLD)
    uint32_t* partials;             // [gridDim.x][NPTS][8] scratch
    unsigned int* counter;          // last-block election
    uint32_t* evals_out;            // [(d+1)][8] Montgomery — ProverMsg.evaluations
    uint32_t* canon_out;            // [(d+1)][8] canonical integers (what ark-serialize writes to the transcript)
    // P(1) from the previous round's claim (rounds >= 2, single launch per round): the launch sums the points
    // t = 0, 2, 3, .., d only (skip1) and, when fix1 is set, the last block fills P(1) = P_prev(r) - P(0).
    uint32_t skip1, fix1, degree;
    // Result delivery without a copy: when set, the last block also writes the message (Montgomery then canonical,
    // (d+1)*8 words each) to mapped pinned host memory and then publishes `seq` in *host_flag; the host spins on it.
    uint32_t* host_out;
    volatile uint32_t* host_flag;
    uint32_t seq;
    // Multi-GPU: the last block exchanges this rank's NPTS partial sums with every peer through mailboxes in peer
    // device memory (NVLink stores + a flag), waits for all peers' partials and continues with the global sums —
    // the collective is fused into the round kernel, no NCCL call and no extra launch (capi_multi.inc).
    uint32_t* const* peer_mail;     // [n_ranks] mailbox base of every rank as mapped in THIS process; null = single GPU
    uint32_t n_ranks, rank, mail_slot, mail_seq;
    uint32_t* comm_error;           // mapped host word set to 1 when a peer does not answer in time
    long long mail_timeout;         // clock64 ticks to wait for a peer (SC_COMM_TIMEOUT_S, default 30 s)
    const uint32_t* prev_evals;     // [(d+1)][8] previous round's ProverMsg (may alias evals_out: read first)
    const uint32_t* lagrange;       // [2][(d+1)][8]: w_j = 1/prod_{k!=j}(j-k), then the field elements 0..d
    // TMA + tensor-core fold rounds (tc_round.cuh): [n_tables] CUtensorMap descriptors of tab_in (128-byte rows,
    // SWIZZLE_128B, 128-row boxes); null for the plain kernels
    const void* tmaps;
    // Raw delivery: the last block writes only the NPTS summed points (Montgomery, unscaled) to host_out and raises the
    // flag; the deferred coefficient, P(1) from the claim and the canonical forms are finished on the host (host_fr.h)
    uint32_t raw_out;
    // [n_products] 1 where the product's coefficient has already been multiplied into one of its tables (prover_init
    // pre-scales a table that only this product uses): the hot loop then skips the two coefficient multiplies per pair
    const uint8_t* prod_scaled;
    // round1_tma_kernel on a sub-range of the tables (pipelined upload, sc_prover_load_tables): first 64-row tile
    uint32_t tile_base;
};

// P_prev(r) by Lagrange interpolation through (j, prev[j]), j = 0..d — what the verifier computes at
// verifier.rs:114 (interpolate_uni_poly); exact field arithmetic, so any evaluation order gives the same element.
// Called by one whole warp; lanes j <= d each build one term; result valid in lane 0.  scratch: >= (d+1)*8 words.
__device__ __forceinline__ Fr claim_from_prev(const uint32_t* prev, const uint32_t* lagr, const Fr& r, uint32_t d, uint32_t* scratch) {
    const uint32_t lane = threadIdx.x & 31;
    if (lane <= d) {
        Fr term = fr::mul(fr::load(prev + lane * 8), fr::load(lagr + lane * 8));
        for (uint32_t k = 0; k <= d; k++) {
            if (k == lane) continue;
            term = fr::mul(term, fr::sub(r, fr::load(lagr + (size_t)(d + 1 + k) * 8)));
        }
#pragma unroll
        for (int i = 0; i < 8; i++) scratch[lane * 8 + i] = term.l[i];
    }
    __syncwarp();
    Fr acc = fr::zero();
    if (lane == 0) {
        for (uint32_t j = 0; j <= d; j++) {
            Fr t;
#pragma unroll
            for (int i = 0; i < 8; i++) t.l[i] = scratch[j * 8 + i];
            acc = fr::add(acc, t);
        }
    }
    __syncwarp();
    return acc;
}

// ---- block-wide sum of NPTS field elements per thread; result valid in thread 0 ------------------------------
__device__ __forceinline__ Fr shfl_down(const Fr& v, int delta) {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], delta);
    return r;
}

template <int NPTS>
__device__ __forceinline__ void block_reduce(Fr (&acc)[NPTS], uint32_t* smem /* [32][NPTS][8] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int t = 0; t < NPTS; t++) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) acc[t] = fr::add(acc[t], shfl_down(acc[t], d));
    }
    if (lane == 0) {
#pragma unroll
        for (int t = 0; t < NPTS; t++)
#pragma unroll
            for (int i = 0; i < 8; i++) smem[(warp * NPTS + t) * 8 + i] = acc[t].l[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int t = 0; t < NPTS; t++) {
            Fr v = fr::zero();
            if (lane < nwarps) {
#pragma unroll
                for (int i = 0; i < 8; i++) v.l[i] = smem[(lane * NPTS + t) * 8 + i];
            }
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) v = fr::add(v, shfl_down(v, d));
            acc[t] = v;
        }
    }
    __syncthreads();
}

// canonical integer of a Montgomery element: a * 1 * R^-1
__device__ __forceinline__ Fr to_canonical(const Fr& a) {
    Fr one_int = fr::zero();
    one_int.l[0] = 1;
    return fr::mul(a, one_int);
}

// coherent 256-bit load (L2 only): for buffers rewritten by this very kernel (tail_kernel ping-pong)
__device__ __forceinline__ Fr load_cg(const uint32_t* p) {
    Fr r;
    asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
                 : "l"(p)
                 : "memory");
    return r;
}

// The hot loop: for every output pair index b = b0, b0+stride, .. < p.n_pairs, (fold and) accumulate the product terms
// of all evaluation points into the thread's unreduced accumulators.  STREAM selects the read-only streaming loads
// (tables written by an EARLIER launch) or coherent L2 loads (tables written earlier in the SAME launch).
// Expand the round's challenge into the fold constants (fr.cuh fold_const) in shared memory: threads 0..7, one each.
__device__ __forceinline__ void prepare_fold_consts(const Fr& r, uint32_t* foldC /* [8][8] */) {
    if (threadIdx.x < 8) {
        Fr c = fr::fold_const(r, (int)threadIdx.x);
#pragma unroll
        for (int i = 0; i < 8; i++) foldC[threadIdx.x * 8 + i] = c.l[i];
    }
    __syncthreads();
}

// One multiplicand's pair (v0, v1) = (table[2b], table[2b+1]) of product k enters the running product terms of all
// evaluation points (prover.rs:116-128).  first/last: position of the multiplicand inside its product; kdeg: how many
// multiplicands the running product holds once this one is in.
// Where the lazily reduced sums of a thread live: in registers (17 limbs per evaluation point) ...
template <int NPTS>
struct RegAccs {
    fr::WideAcc (&a)[NPTS];
    __device__ __forceinline__ void mac(int t, const Fr& x, const Fr& y) { fr::wide_mac(a[t], x, y); }
    __device__ __forceinline__ void add_shifted(int t, const Fr& x) { fr::wide_add_shifted(a[t], x); }
    __device__ __forceinline__ void begin() {}
};
// ... or parked in tensor memory (tma_round1.cuh TmemAccs): same interface.

// SKIP: -1 = p.skip1 decides at run time, 0 / 1 = known when the kernel is built (round 1 never skips P(1), the fold rounds of
// a single launch always do).  SIMPLE: ONE product whose coefficient is deferred to the end of the round and a single launch per
// round (t0 == 0) — then nothing about the coefficient or the first evaluation point is left to decide per pair.
//
// ALT (needs SKIP known): the sums are taken at the points  0, 1, inf, -1, 2, -2  (round 1: the first d+1 of them; rounds that
// skip P(1): the same list without the 1) instead of 0, 1, 2, .., d.  "inf" is the coefficient of t^d, i.e. the product of the
// STEPS table[2b+1] - table[2b].  On these points a multiplicand's values cost at most ONE modular add each —
//     v(0) = table[2b]   v(1) = table[2b+1]   v(inf) = step   v(-1) = v(0) - step   v(2) = v(1) + step   v(-2) = v(-1) - step
// — where the consecutive points cost one add per point including t = 1 (4 instead of 2 add/sub per table at degree 3).  The host
// recovers P(0..d) exactly (host_fr.h alt_to_standard: interpolation of P(t) - c_d t^d through the d finite points), so the
// message is bit-identical.  `full`: the product has exactly `degree` multiplicands; a shorter one has no t^d term, so it is left
// out of the inf slot.
__device__ __forceinline__ int point_index(bool skip1, int s) { return skip1 ? (s == 0 ? 0 : s + 1) : s; }  // into 0, 1, inf, -1, 2, -2

template <int NPTS, bool ILP = false, int SKIP = -1, bool SIMPLE = false, bool ALT = false, class ACCS>
__device__ __forceinline__ void consume_pair_acc(const RoundParams& p, uint32_t k, bool first, bool last, uint32_t kdeg, bool full,
                                                 const Fr& v0, const Fr& v1, Fr (&prod)[NPTS], ACCS& accs) {
    static_assert(!ALT || SKIP >= 0, "the alternative points need a compile-time skip flag");
    const bool skip1 = SKIP < 0 ? (p.skip1 != 0) : (SKIP != 0);
    // prover.rs:119-124: start = table[2b], step = table[2b+1] - start; product[t] *= start; start += step
    Fr step = fr::sub(v1, v0);
    Fr cur = v0, one = v1;  // the multiplicand at t = 0 and t = 1 (ALT)
    if (!SIMPLE) {
        for (uint32_t s = 0; s < p.t0; s++) cur = fr::add(cur, step);  // only for d+1 > MAX_NPTS
        if (first && !p.defer_coeff && !(p.prod_scaled && p.prod_scaled[k])) {  // c_k * prod_j(...): scale the first multiplicand's line once
            Fr c = fr::load(p.coeffs + 8 * k);
            cur = fr::mul(cur, c);
            step = fr::mul(step, c);
            if (ALT) one = fr::add(cur, step);
        }
    }
    if (ALT) {
        // value of this multiplicand at the point of slot t (each visited once, in slot order: -1 before -2)
        Fr vm1 = cur;
#define SC_ALT_VALUE(t, out)                                                          \
        {                                                                             \
            const int pi_ = point_index(skip1, (t));                                  \
            if (pi_ == 0) out = cur;                                                  \
            else if (pi_ == 1) out = one;                                             \
            else if (pi_ == 2) out = step;                                            \
            else if (pi_ == 3) { vm1 = fr::sub(cur, step); out = vm1; }               \
            else if (pi_ == 4) out = fr::add(one, step);                              \
            else { vm1 = fr::sub(vm1, step); out = vm1; }                             \
        }
        if (first && last) {
            accs.begin();
#pragma unroll
            for (int t = 0; t < NPTS; t++) {
                Fr v;
                SC_ALT_VALUE(t, v)
                if (point_index(skip1, t) == 2 && !full) continue;
                accs.add_shifted(t, v);
            }
        } else if (first) {
#pragma unroll
            for (int t = 0; t < NPTS; t++) SC_ALT_VALUE(t, prod[t])
        } else if (last) {
            accs.begin();
#pragma unroll
            for (int t = 0; t < NPTS; t++) {
                Fr v;
                SC_ALT_VALUE(t, v)
                if (point_index(skip1, t) == 2 && !full) continue;
                accs.mac(t, prod[t], v);
            }
        } else if (ILP) {
            fr::FrN<NPTS> a, b;
#pragma unroll
            for (int t = 0; t < NPTS; t++) {
                a.v[t] = prod[t];
                SC_ALT_VALUE(t, b.v[t])
            }
            const fr::FrN<NPTS> r = fr::mul_lazy_n<NPTS>(a, b);
#pragma unroll
            for (int t = 0; t < NPTS; t++) prod[t] = r.v[t];
        } else {
            // After two multiplicands the running product q is quadratic in the evaluation point: from q(0), q(1), q(inf)
            //   q(-1) = 2 q(0) - q(1) + 2 q(inf)        q(2) = 2 q(1) - q(0) + 2 q(inf)
            // (exact), so round 1 multiplies three points of the second multiplicand and derives the others.
            const bool derive = !skip1 && kdeg == 2;
#pragma unroll
            for (int t = 0; t < NPTS; t++) {
                if (derive && t == 3) {
                    const Fr i2 = fr::add(prod[2], prod[2]);
                    prod[3] = fr::add(fr::sub(fr::add(prod[0], prod[0]), prod[1]), i2);
                } else if (derive && t == 4) {
                    const Fr i2 = fr::add(prod[2], prod[2]);
                    prod[4] = fr::add(fr::sub(fr::add(prod[1], prod[1]), prod[0]), i2);
                } else {
                    Fr v;
                    SC_ALT_VALUE(t, v)
                    // round 1 derives points from these products (canonical needed); the fold rounds only multiply them again
                    prod[t] = skip1 ? fr::mul_lazy(prod[t], v) : fr::mul(prod[t], v);
                }
            }
        }
#undef SC_ALT_VALUE
        return;
    }
    // slot s holds evaluation point t0+s, or with skip1 the points 0, 2, 3, ..: one extra step after slot 0
#define SC_NEXT_POINT(t)                                        \
    if ((t) + 1 < NPTS) {                                       \
        cur = fr::add(cur, step);                               \
        if ((t) == 0 && skip1) cur = fr::add(cur, step);        \
    }
    if (first && last) {  // single multiplicand: contributes its value itself
        accs.begin();
#pragma unroll
        for (int t = 0; t < NPTS; t++) {
            accs.add_shifted(t, cur);
            SC_NEXT_POINT(t)
        }
    } else if (first) {
#pragma unroll
        for (int t = 0; t < NPTS; t++) {
            prod[t] = cur;
            SC_NEXT_POINT(t)
        }
    } else if (last) {  // prover.rs:126-128 fused with the last multiply: products_sum[t] += product[t]*start
        accs.begin();
#pragma unroll
        for (int t = 0; t < NPTS; t++) {
            accs.mac(t, prod[t], cur);
            SC_NEXT_POINT(t)
        }
    } else if (ILP && skip1) {
        // latency-bound callers: the NPTS points of this multiplicand in ONE out-of-line call (interleaved carry chains)
        fr::FrN<NPTS> a, b;
#pragma unroll
        for (int t = 0; t < NPTS; t++) {
            a.v[t] = prod[t];
            b.v[t] = cur;
            SC_NEXT_POINT(t)
        }
        const fr::FrN<NPTS> r = fr::mul_lazy_n<NPTS>(a, b);
#pragma unroll
        for (int t = 0; t < NPTS; t++) prod[t] = r.v[t];
    } else {
        // After k multiplicands prod[.] is a degree-k polynomial in the evaluation point, so on consecutive
        // points only k+1 values need a multiply; the others follow by finite differences (integer
        // combinations, exact): k = 2: q(t) = 3(q(t-1) - q(t-2)) + q(t-3);  k = 3: q(t) = 4(q(t-1) + q(t-3)) -
        // 6 q(t-2) - q(t-4).  Saves one of the four multiplies of round 1 at degree 3.
        const bool consecutive = !skip1;
#pragma unroll
        for (int t = 0; t < NPTS; t++) {
            if (t >= 3 && consecutive && kdeg == 2) {
                Fr dd = fr::sub(prod[t - 1], prod[t - 2]);
                prod[t] = fr::add(fr::add(fr::add(dd, dd), dd), prod[t - 3]);
            } else if (t >= 4 && consecutive && kdeg == 3) {
                Fr s = fr::add(prod[t - 1], prod[t - 3]);
                s = fr::add(s, s);
                s = fr::add(s, s);
                Fr m2 = fr::add(prod[t - 2], prod[t - 2]);
                Fr m6 = fr::add(fr::add(m2, m2), m2);
                prod[t] = fr::sub(fr::sub(s, m6), prod[t - 4]);
            } else if (consecutive) {
                prod[t] = fr::mul(prod[t], cur);  // canonical: the finite differences above add and subtract these
                SC_NEXT_POINT(t)
            } else {
                // the running product is only ever multiplied again by a canonical operand or fed to the lazy
                // accumulator, so it may stay in [0, 2p): no conditional subtraction (fr::mul_lazy)
                prod[t] = fr::mul_lazy(prod[t], cur);
                SC_NEXT_POINT(t)
            }
        }
    }
#undef SC_NEXT_POINT
}

template <int NPTS, bool ILP = false, bool ALT = false>
__device__ __forceinline__ void consume_pair(const RoundParams& p, uint32_t k, bool first, bool last, uint32_t kdeg, bool full, const Fr& v0,
                                             const Fr& v1, Fr (&prod)[NPTS], fr::WideAcc (&accw)[NPTS]) {
    RegAccs<NPTS> accs{accw};
    // ALT is only used by the resident kernel, whose rounds always fold and skip P(1)
    consume_pair_acc<NPTS, ILP, ALT ? 1 : -1, false, ALT>(p, k, first, last, kdeg, full, v0, v1, prod, accs);
}

template <int NPTS, bool FOLD, bool STREAM, bool ILP = false, bool ALT = false>
__device__ __forceinline__ void accumulate_pairs(const RoundParams& p, const uint32_t* foldC, unsigned long long b0,
                                                 unsigned long long stride, fr::WideAcc (&accw)[NPTS]) {
    for (unsigned long long b = b0; b < p.n_pairs; b += stride) {
        for (uint32_t k = 0; k < p.n_products; k++) {
            Fr prod[NPTS];
            const uint32_t j0 = p.prod_offsets[k], j1 = p.prod_offsets[k + 1];
            for (uint32_t jj = j0; jj < j1; jj++) {
                const uint32_t idx = p.prod_indices[jj];
                const uint32_t* tin = p.tab_in[idx];
                Fr v0, v1;
                if (FOLD) {
                    const uint32_t* src = tin + b * 32;  // 4 elements x 8 words
                    Fr e0, e1, e2, e3;
                    if (STREAM) {
                        e0 = fr::load_stream(src); e1 = fr::load_stream(src + 8); e2 = fr::load_stream(src + 16); e3 = fr::load_stream(src + 24);
                    } else {
                        e0 = load_cg(src); e1 = load_cg(src + 8); e2 = load_cg(src + 16); e3 = load_cg(src + 24);
                    }
                    if (ILP) {
                        const fr::Fr2 f = fr::mul_round_const_x2(foldC, fr::sub(e1, e0), fr::sub(e3, e2));
                        v0 = fr::add(e0, f.a);
                        v1 = fr::add(e2, f.b);
                    } else {
                        v0 = fr::add(e0, FR_MUL_ROUND_CONST(foldC, fr::sub(e1, e0)));
                        v1 = fr::add(e2, FR_MUL_ROUND_CONST(foldC, fr::sub(e3, e2)));
                    }
                    if (p.write_fold && p.prod_first[jj]) {
                        uint32_t* dst = p.tab_out[idx] + b * 16;
                        fr::store(dst, v0);
                        fr::store(dst + 8, v1);
                    }
                } else {
                    const uint32_t* src = tin + b * 16;
                    v0 = fr::load_stream(src);
                    v1 = fr::load_stream(src + 8);
                }
                consume_pair<NPTS, ILP, ALT>(p, k, jj == j0, jj + 1 == j1, jj - j0 + 1, j1 - j0 == p.degree, v0, v1, prod, accw);
            }
        }
    }
}

// ---- fine-grained pairs (resident rounds with few pairs) ------------------------------------------------------------------------
// With one pair per thread a latency-bound round still walks ~3500 instructions per warp (6 folds, 6 multiplies, 3 lazy
// multiply-accumulates at degree 3) on one scheduler.  Here a pair is spread over LPP = 2^lpp_log2 lanes of a warp:
//   phase A  lane u < 2*nnz folds ONE element: CSR entry u/2 of the product list, half u%2 (new[2b] or new[2b+1]), stores it
//            (first use of the table) and parks it in the warp's shared-memory slots;
//   phase B  lane u < n_products*NPTS owns (product u/NPTS, evaluation point u%NPTS): it builds that point of every
//            multiplicand from the parked pair (prover.rs:119-124) and runs the product chain into ITS lazy accumulator.
// ~1000 instructions per lane instead of ~3500.  Needs 2*nnz <= 32 and n_products*NPTS <= 32; exact arithmetic, same limbs.
// slots: this warp's [32][8] words.  b = the pair this lane group works on (b >= p.n_pairs: idle).  On return `acc` holds the
// lane's contribution and `my_pt` its point slot (or -1).
template <int NPTS>
__device__ __forceinline__ void accumulate_fine(const RoundParams& p, const uint32_t* foldC, uint32_t* slots, unsigned long long b,
                                                uint32_t lpp_log2, fr::WideAcc& acc, int& my_pt) {
    const uint32_t lane = threadIdx.x & 31, lpp = 1u << lpp_log2, u = lane & (lpp - 1), base = lane - u;
    const uint32_t nnz = p.prod_offsets[p.n_products];
    const bool live = b < p.n_pairs;
    if (live && u < 2 * nnz) {  // phase A
        const uint32_t jj = u >> 1, h = u & 1u, idx = p.prod_indices[jj];
        const uint32_t* src = p.tab_in[idx] + b * 32 + h * 16;
        const Fr e0 = load_cg(src), e1 = load_cg(src + 8);
        const Fr v = fr::add(e0, FR_MUL_ROUND_CONST(foldC, fr::sub(e1, e0)));
        if (p.write_fold && p.prod_first[jj]) fr::store(p.tab_out[idx] + b * 16 + h * 8, v);
#pragma unroll
        for (int i = 0; i < 8; i++) slots[(base + u) * 8 + i] = v.l[i];
    }
    __syncwarp();
    my_pt = -1;
    fr::wide_zero(acc);
    if (live && u < p.n_products * NPTS) {  // phase B
        const uint32_t k = u / NPTS, s = u % NPTS;
        my_pt = (int)s;
        const int pi = point_index(true, (int)s);  // the resident rounds skip P(1): points 0, inf, -1, 2, -2 (consume_pair_acc ALT)
        const uint32_t j0 = p.prod_offsets[k], j1 = p.prod_offsets[k + 1];
        if (pi == 2 && j1 - j0 != p.degree) my_pt = -1;  // a product shorter than the degree has no t^d term
        Fr prod = fr::zero();
        for (uint32_t jj = j0; jj < j1 && my_pt >= 0; jj++) {
            Fr v0, v1;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                v0.l[i] = slots[(base + 2 * jj) * 8 + i];
                v1.l[i] = slots[(base + 2 * jj + 1) * 8 + i];
            }
            const Fr step = fr::sub(v1, v0);
            Fr cur;
            if (pi == 0) cur = v0;
            else if (pi == 2) cur = step;
            else if (pi == 3) cur = fr::sub(v0, step);
            else if (pi == 4) cur = fr::add(v1, step);
            else cur = fr::sub(fr::sub(v0, step), step);
            const bool first = jj == j0, last = jj + 1 == j1;
            if (first && !p.defer_coeff && !(p.prod_scaled && p.prod_scaled[k])) cur = fr::mul(cur, fr::load(p.coeffs + 8 * k));
            if (first && last) fr::wide_add_shifted(acc, cur);
            else if (first) prod = cur;
            else if (last) fr::wide_mac(acc, prod, cur);
            else prod = fr::mul_lazy(prod, cur);
        }
    }
}

// ---- integer block sum of lazily reduced accumulators (resident rounds) ------------------------------------------------------
// The per-thread sums of a round are UNREDUCED 17-limb integers (fr::WideAcc).  Instead of Montgomery-reducing every thread's
// sums and adding field elements through a shuffle tree (wide_reduce + block_reduce: ~5300 cycles of dependent latency for
// three points), the limbs are added as plain integers — 16-bit halves through redux.sync, so no carry crosses a lane — and the
// ONE resulting 17-limb integer per point goes to the host, which reduces it (host_fr.h; ~0.1 us of CPU).  A round has fewer
// than 2^34 products, so the total fits the 17 limbs (each product is below 2^510).
constexpr int WL = 17;

// per-limb 64-bit sums (tot[t * WL + i], shared or global memory) -> carry-propagated 17-limb integers in s_out; all threads call it
template <int NPTS, bool GLOBAL = false>
__device__ __forceinline__ void carry_wide(const unsigned long long* tot, uint32_t* s_out) {
    if (threadIdx.x < NPTS) {
        unsigned long long c = 0;
#pragma unroll
        for (int i = 0; i < WL; i++) {
            c += GLOBAL ? __ldcg(tot + threadIdx.x * WL + i) : tot[threadIdx.x * WL + i];
            s_out[threadIdx.x * WL + i] = (uint32_t)c;
            c >>= 32;
        }
    }
    __syncthreads();
}

// All threads call it.  s_part: [nwarps][NPTS * WL] u64, s_tot: [NPTS * WL] u64, s_out: [NPTS * WL] u32 (the result).
template <int NPTS>
__device__ __forceinline__ void block_sum_wide(const fr::WideAcc (&a)[NPTS], unsigned long long* s_part, unsigned long long* s_tot, uint32_t* s_out,
                                               bool carry = true) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int t = 0; t < NPTS; t++) {
#pragma unroll
        for (int i = 0; i < WL; i++) {
            const uint32_t lo = __reduce_add_sync(0xffffffffu, a[t].l[i] & 0xffffu);
            const uint32_t hi = __reduce_add_sync(0xffffffffu, a[t].l[i] >> 16);
            if (lane == 0) s_part[(size_t)warp * NPTS * WL + t * WL + i] = (unsigned long long)lo + ((unsigned long long)hi << 16);
        }
    }
    __syncthreads();
    if (threadIdx.x < NPTS * WL) {
        unsigned long long tot = 0;
        for (int w = 0; w < nwarps; w++) tot += s_part[(size_t)w * NPTS * WL + threadIdx.x];
        s_tot[threadIdx.x] = tot;
    }
    __syncthreads();
    if (!carry) return;  // the caller adds s_tot (per-limb sums, no carries yet) into the grid's totals
    carry_wide<NPTS>(s_tot, s_out);
}

// The same for ONE accumulator per thread that belongs to point `my_pt` (accumulate_fine): the other points see zeros.
template <int NPTS>
__device__ __forceinline__ void block_sum_wide_sel(const fr::WideAcc& a, int my_pt, unsigned long long* s_part, unsigned long long* s_tot,
                                                   uint32_t* s_out, bool carry = true) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int t = 0; t < NPTS; t++) {
        const uint32_t m = (my_pt == t) ? 0xffffffffu : 0u;
#pragma unroll
        for (int i = 0; i < WL; i++) {
            const uint32_t v = a.l[i] & m;
            const uint32_t lo = __reduce_add_sync(0xffffffffu, v & 0xffffu);
            const uint32_t hi = __reduce_add_sync(0xffffffffu, v >> 16);
            if (lane == 0) s_part[(size_t)warp * NPTS * WL + t * WL + i] = (unsigned long long)lo + ((unsigned long long)hi << 16);
        }
    }
    __syncthreads();
    if (threadIdx.x < NPTS * WL) {
        unsigned long long tot = 0;
        for (int w = 0; w < nwarps; w++) tot += s_part[(size_t)w * NPTS * WL + threadIdx.x];
        s_tot[threadIdx.x] = tot;
    }
    __syncthreads();
    if (!carry) return;  // the caller adds s_tot (per-limb sums, no carries yet) into the grid's totals
    carry_wide<NPTS>(s_tot, s_out);
}

// w += the 17-limb integer at src (last-arrival sum of the per-CTA partials; not a hot path)
__device__ __forceinline__ void wide_add_limbs(fr::WideAcc& w, const uint32_t* src) {
    unsigned long long c = 0;
#pragma unroll
    for (int i = 0; i < WL; i++) {
        c += (unsigned long long)w.l[i] + src[i];
        w.l[i] = (uint32_t)c;
        c >>= 32;
    }
}

// Executed by warp 0 (thread 0 holds the NPTS grand totals): deferred coefficient, P(1) from the claim, and the two
// output forms.  The d+1 output slots are spread over the lanes (each needs a coefficient multiply and a
// Montgomery->canonical multiply) so a latency-bound small round pays two multiplies, not 2(d+1).
// canon_smem (optional): (d+1)*8 words receiving the canonical limbs for an on-device transcript.
template <int NPTS>
__device__ __forceinline__ void publish_round(const RoundParams& p, const Fr (&acc)[NPTS], const Fr& r, uint32_t* scratch,
                                              uint32_t* canon_smem) {
    const uint32_t lane = threadIdx.x & 31;
    Fr claim = fr::zero();
    if (p.fix1) claim = claim_from_prev(p.prev_evals, p.lagrange, r, p.degree, scratch);  // reads prev before it is overwritten
    // hand the totals (and the claim) to the lanes through shared memory
    if (lane == 0) {
#pragma unroll
        for (int t = 0; t < NPTS; t++)
#pragma unroll
            for (int i = 0; i < 8; i++) scratch[t * 8 + i] = acc[t].l[i];
#pragma unroll
        for (int i = 0; i < 8; i++) scratch[NPTS * 8 + i] = claim.l[i];
    }
    __syncwarp();
    // lane s < NPTS owns summed point s; with skip1 lane NPTS owns P(1) = claim - P(0)
    const bool owner = lane < (uint32_t)NPTS, fixer = p.skip1 && lane == (uint32_t)NPTS;
    if (owner || fixer) {
        Fr v;
        const uint32_t src = owner ? lane : 0u;
#pragma unroll
        for (int i = 0; i < 8; i++) v.l[i] = scratch[src * 8 + i];
        if (p.defer_coeff) v = fr::mul(v, fr::load(p.coeffs));
        uint32_t slot = p.skip1 ? (lane == 0 ? 0u : lane + 1u) : p.t0 + lane;
        if (fixer) {
            slot = 1;
            if (p.fix1) {
                Fr cl;
#pragma unroll
                for (int i = 0; i < 8; i++) cl.l[i] = scratch[NPTS * 8 + i];
                v = fr::sub(cl, v);
            } else {
                v = fr::zero();  // sharded NCCL path: slot 1 stays zero for the all-gather
            }
        }
        const Fr cv = to_canonical(v);
        fr::store(p.evals_out + (size_t)slot * 8, v);
        if (p.canon_out) fr::store(p.canon_out + (size_t)slot * 8, cv);
        if (canon_smem) {
#pragma unroll
            for (int i = 0; i < 8; i++) canon_smem[slot * 8 + i] = cv.l[i];
        }
        if (p.host_out) {  // mapped pinned host memory: Montgomery words, then canonical words
            const uint32_t n = (p.degree + 1) * 8;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                p.host_out[slot * 8 + i] = v.l[i];
                p.host_out[n + slot * 8 + i] = cv.l[i];
            }
        }
    }
    if (p.host_flag) {
        __threadfence_system();
        __syncwarp();
    }
}

#ifndef SC_MIN_BLOCKS
#define SC_MIN_BLOCKS 3
#endif
#ifndef SC_THREADS
#define SC_THREADS 128
#endif
constexpr int ROUND_THREADS = SC_THREADS;  // CTA size of round_kernel in this translation unit
constexpr uint32_t MAIL_WORDS = 1024;  // per (slot, rank): up to 512 {value, sequence number} pairs — 6 points x 8 limbs, 5 x 17 limbs of unreduced sums, or
                                       // the 6 x 26 / 9 x 34 limbs of a contraction round (gemm_sum.cuh)
constexpr uint32_t MAIL_SLOTS = 64;

// One mailbox word = {limb, sequence number} packed into ONE 64-bit register and moved with a scalar 8-byte access.  The
// PTX memory model guarantees single-copy atomicity for naturally aligned scalar accesses up to 64 bits (a .v2.u32 access
// is two 32-bit accesses in unspecified order and gives no such guarantee), so a reader that sees the new sequence number
// sees the limb that was stored with it.  relaxed.sys: the word may live in a peer GPU's memory or in mapped host memory.
__device__ __forceinline__ void mail_store(uint32_t* dst, uint32_t value, uint32_t seq) {
    const unsigned long long w = (unsigned long long)value | ((unsigned long long)seq << 32);
    asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(dst), "l"(w) : "memory");
}
__device__ __forceinline__ void mail_load(const uint32_t* src, uint32_t& value, uint32_t& seq) {
    unsigned long long w;
    asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
    value = (uint32_t)w;
    seq = (uint32_t)(w >> 32);
}

// Executed by warp 0 of the last block; thread 0 holds this rank's NPTS partial sums.  All-to-all over NVLink peer
// memory with a flag-in-data protocol: every limb travels as ONE 8-byte store {limb, sequence number} into the
// receiver's mailbox, so the receiver just polls each word until its sequence number matches — one NVLink one-way
// latency, no system-scope fence, no separate flag write behind it (mail_store/mail_load above: the 8 bytes are
// single-copy atomic).  (The first version stored the limbs, fenced, then
// raised a flag: two dependent round trips per round.)  Mailbox slots rotate (MAIL_SLOTS) so a fast rank's next round
// never lands on words a slow rank has yet to read.  On return thread 0 holds the sums over all ranks, added in rank
// order (the same on every rank).
template <int NPTS>
__device__ __forceinline__ void exchange_partials(const RoundParams& p, Fr (&acc)[NPTS], uint32_t* scratch) {
    const uint32_t lane = threadIdx.x & 31, G = p.n_ranks;
    constexpr uint32_t NW = NPTS * 8;
    if (lane == 0) {
#pragma unroll
        for (int t = 0; t < NPTS; t++)
#pragma unroll
            for (int i = 0; i < 8; i++) scratch[t * 8 + i] = acc[t].l[i];
    }
    __syncwarp();
    for (uint32_t w = lane; w < NW; w += 32) {
        const uint32_t v = scratch[w];
        for (uint32_t g = 0; g < G; g++) {
            uint32_t* dst = p.peer_mail[g] + ((size_t)p.mail_slot * G + p.rank) * MAIL_WORDS + 2 * w;
            mail_store(dst, v, p.mail_seq);
        }
    }
    __syncwarp();
    const uint32_t* mine = p.peer_mail[p.rank] + (size_t)p.mail_slot * G * MAIL_WORDS;
    for (uint32_t w = lane; w < NW; w += 32) {
        for (uint32_t g = 0; g < G; g++) {
            const uint32_t* src = mine + (size_t)g * MAIL_WORDS + 2 * w;
            uint32_t d, f;
            const long long t0 = clock64();
            for (;;) {
                mail_load(src, d, f);
                if (f == p.mail_seq) break;
                if (clock64() - t0 > p.mail_timeout) {  // a peer died (default ~30 s): report instead of hanging the GPU
                    *p.comm_error = 1;
                    d = 0;
                    break;
                }
            }
            scratch[(g + 1) * NW + w] = d;  // rows 1..G: the ranks' partials (row 0 still holds this rank's own)
        }
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int t = 0; t < NPTS; t++) {
            Fr s = fr::zero();
            for (uint32_t g = 0; g < G; g++) {
                Fr x;
#pragma unroll
                for (int i = 0; i < 8; i++) x.l[i] = scratch[(g + 1) * NW + t * 8 + i];
                s = fr::add(s, x);
            }
            acc[t] = s;
        }
    }
    __syncwarp();
}

// The same exchange for UNREDUCED sums (resident rounds): s_vals holds this rank's NPTS * WL limbs; on return it holds the
// integer sums over all ranks (they fit the 17 limbs, see block_sum_wide).  Executed by warp 0.  s_rows: [G][NPTS * WL].
template <int NPTS>
__device__ __forceinline__ void exchange_wide(const RoundParams& p, uint32_t* s_vals, uint32_t* s_rows) {
    const uint32_t lane = threadIdx.x & 31, G = p.n_ranks;
    constexpr uint32_t NW = NPTS * WL;
    for (uint32_t w = lane; w < NW; w += 32) {
        const uint32_t v = s_vals[w];
        for (uint32_t g = 0; g < G; g++)
            mail_store(p.peer_mail[g] + ((size_t)p.mail_slot * G + p.rank) * MAIL_WORDS + 2 * w, v, p.mail_seq);
    }
    const uint32_t* mine = p.peer_mail[p.rank] + (size_t)p.mail_slot * G * MAIL_WORDS;
    for (uint32_t w = lane; w < NW; w += 32) {
        for (uint32_t g = 0; g < G; g++) {
            uint32_t d, f;
            const long long t0 = clock64();
            for (;;) {
                mail_load(mine + (size_t)g * MAIL_WORDS + 2 * w, d, f);
                if (f == p.mail_seq) break;
                if (clock64() - t0 > p.mail_timeout) {
                    *p.comm_error = 1;
                    d = 0;
                    break;
                }
            }
            s_rows[g * NW + w] = d;
        }
    }
    __syncwarp();
    if (lane < (uint32_t)NPTS) {
        unsigned long long c = 0;
        for (int i = 0; i < WL; i++) {
            for (uint32_t g = 0; g < G; g++) c += s_rows[g * NW + lane * WL + i];
            s_vals[lane * WL + i] = (uint32_t)c;
            c >>= 32;
        }
    }
    __syncwarp();
}

// Everything after the hot loop: per-thread lazy sums -> field elements -> block sum -> (last block) grid sum ->
// (sharded) exchange with the peer GPUs -> deferred coefficient, P(1) from the claim, publication.
template <int NPTS>
__device__ __forceinline__ void finish_round_reduced(const RoundParams& p, Fr (&acc)[NPTS], const Fr& r, uint32_t* s_red, bool* s_last_p);

template <int NPTS>
__device__ __forceinline__ void finish_round(const RoundParams& p, fr::WideAcc (&accw)[NPTS], const Fr& r, uint32_t* s_red, bool* s_last_p) {
    Fr acc[NPTS];
#pragma unroll
    for (int t = 0; t < NPTS; t++) acc[t] = fr::wide_reduce(accw[t]);
    finish_round_reduced<NPTS>(p, acc, r, s_red, s_last_p);
}

// acc[t]: this thread's sums as field elements
template <int NPTS>
__device__ __forceinline__ void finish_round_reduced(const RoundParams& p, Fr (&acc)[NPTS], const Fr& r, uint32_t* s_red, bool* s_last_p) {
    bool& s_last = *s_last_p;
    block_reduce<NPTS>(acc, s_red);
    if (gridDim.x > 1) {  // a single-CTA launch (small rounds) already holds the grand totals in thread 0
        if (threadIdx.x == 0) {
#pragma unroll
            for (int t = 0; t < NPTS; t++) fr::store(p.partials + ((size_t)blockIdx.x * NPTS + t) * 8, acc[t]);
            __threadfence();
            unsigned int ticket = atomicAdd(p.counter, 1u);
            s_last = (ticket == gridDim.x - 1);
        }
        __syncthreads();
        if (!s_last) return;
        // Last block to finish: sum the per-block partials (prover.rs:138-148, the rayon reduce)
        __threadfence();
#pragma unroll
        for (int t = 0; t < NPTS; t++) acc[t] = fr::zero();
        for (uint32_t g = threadIdx.x; g < gridDim.x; g += blockDim.x) {
#pragma unroll
            for (int t = 0; t < NPTS; t++) acc[t] = fr::add(acc[t], fr::load(p.partials + ((size_t)g * NPTS + t) * 8));
        }
        block_reduce<NPTS>(acc, s_red);
        if (threadIdx.x == 0) *p.counter = 0;
    }
    if (threadIdx.x >= 32) return;
    if (p.peer_mail) exchange_partials<NPTS>(p, acc, s_red);
    if (p.raw_out) {  // thread 0 holds the totals: one coalesced store of NPTS*8 words to mapped host memory, then the flag
        if (threadIdx.x == 0) {
#pragma unroll
            for (int t = 0; t < NPTS; t++)
#pragma unroll
                for (int i = 0; i < 8; i++) s_red[t * 8 + i] = acc[t].l[i];
        }
        __syncwarp();
        for (uint32_t w = threadIdx.x; w < (uint32_t)NPTS * 8; w += 32) p.host_out[w] = s_red[w];
        __threadfence_system();
        __syncwarp();
        if (threadIdx.x == 0) *p.host_flag = p.seq;
        return;
    }
    // deferred coefficient, P(1) from the claim, both output forms; with a flag the message also goes to mapped host memory
    publish_round<NPTS>(p, acc, r, s_red, nullptr);
    if (threadIdx.x == 0 && p.host_flag) *p.host_flag = p.seq;
}

template <int NPTS, bool FOLD>
__global__ void __launch_bounds__(SC_THREADS, SC_MIN_BLOCKS) round_kernel(const RoundParams p) {
    __shared__ uint32_t s_red[32 * NPTS * 8];
    __shared__ bool s_last;

    // Per-thread sums are kept UNREDUCED (fr::WideAcc): the last multiply of every product term is a plain 256x256-bit
    // integer product added into 17 limbs; one Montgomery reduction per evaluation point per thread at the end.
    fr::WideAcc accw[NPTS];
#pragma unroll
    for (int t = 0; t < NPTS; t++) fr::wide_zero(accw[t]);
    __shared__ __align__(16) uint32_t s_foldC[64];
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = p.r[i];
    if (FOLD) prepare_fold_consts(r, s_foldC);

    accumulate_pairs<NPTS, FOLD, true>(p, s_foldC, (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x,
                                       (unsigned long long)gridDim.x * blockDim.x, accw);
    finish_round<NPTS>(p, accw, r, s_red, &s_last);
}

}  // namespace sck
