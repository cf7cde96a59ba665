// GKR round-function initialisers (/root/reference/src/gkr_round_sumcheck/mod.rs:22-63) and small helpers.
//
// The reference walks BTreeMap/HashMap entries and scatter-accumulates into dense vectors.  On the GPU the
// scatter `dst[key] += value (mod p)` has no native atomic, so each destination is kept as 8 x u64 lanes holding
// plain integer sums of the 32-bit limbs (native 64-bit atomicAdd; up to 2^32 addends cannot overflow a lane) and
// a second pass carry-propagates and reduces mod p.  Integer sums commute, so the result is order-independent and
// equals the reference's value exactly.
#pragma once
#include <cstdint>

#include "fr.cuh"

namespace sck {

using fr::Fr;

struct Fr8 {  // a field element passed by value as a kernel argument
    uint32_t l[8];
};

__device__ __forceinline__ Fr fr_R2() {  // R^2 mod p: mul(x, R2) = x*R mod p
    Fr r = {{0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u}};
    return r;
}

// eq[b] = prod_j (b_j ? g_j : 1 - g_j)  — ark-poly precompute_eq (external), bit j of b <-> g[j].
// Two steps instead of dim multiplies per entry (the first version: O(N*dim)): the factors of the low
// `lo` variables and of the high dim-lo variables are tabulated separately (2^lo + 2^(dim-lo) short products), then
// eq[b] = eq_lo[b & (2^lo - 1)] * eq_hi[b >> lo] — one multiply per entry.  Same factors, exact arithmetic: same element.
__device__ __forceinline__ Fr eq_partial(const uint32_t* g, uint32_t first, uint32_t count, unsigned long long b) {
    Fr acc = fr::one();
    for (uint32_t j = 0; j < count; j++) {
        Fr gj = fr::load(g + 8 * (first + j));
        Fr f = ((b >> j) & 1) ? gj : fr::sub(fr::one(), gj);
        acc = (j == 0) ? f : fr::mul(acc, f);
    }
    return acc;
}
// halves[0 .. 2^lo) = eq over g[0..lo), halves[2^lo .. 2^lo + 2^(dim-lo)) = eq over g[lo..dim)
__global__ void __launch_bounds__(128) eq_halves_kernel(const uint32_t* g, uint32_t dim, uint32_t lo, uint32_t* halves) {
    const unsigned long long n_lo = 1ull << lo, n_hi = 1ull << (dim - lo);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_lo + n_hi;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const Fr v = i < n_lo ? eq_partial(g, 0, lo, i) : eq_partial(g, lo, dim - lo, i - n_lo);
        fr::store(halves + i * 8, v);
    }
}
__global__ void __launch_bounds__(128) eq_outer_kernel(const uint32_t* halves, uint32_t dim, uint32_t lo, uint32_t* out) {
    const unsigned long long n = 1ull << dim, n_lo = 1ull << lo;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < n;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        const Fr a = fr::load(halves + (b & (n_lo - 1)) * 8);
        if (lo == dim) { fr::store(out + b * 8, a); continue; }
        fr::store(out + b * 8, fr::mul(a, fr::load(halves + (n_lo + (b >> lo)) * 8)));
    }
}

__device__ __forceinline__ void lanes_add(unsigned long long* lanes, const Fr& v) {
#pragma unroll
    for (int i = 0; i < 8; i++) atomicAdd(lanes + i, (unsigned long long)v.l[i]);
}

// initialize_phase_one (mod.rs:30-38), per nonzero of f1 with idx = z | x << dim | y << 2dim:
//   w = eq_g[z] * v            (f1.fix_variables(g): f1_g[idx >> dim] += eq_g[idx & mask] * v)
//   hg_lanes[x] += w * f3[y]   (a_hg[x] += f1_g[xy] * f3[y])
// w is kept (keyed by xy = idx >> dim) for phase two.
__global__ void __launch_bounds__(128) gkr_phase1_scatter_kernel(const unsigned long long* f1_idx, const uint32_t* f1_val,
                                                                unsigned long long nnz, uint32_t dim, const uint32_t* eq_g,
                                                                const uint32_t* f3, uint32_t* w_out,
                                                                unsigned long long* hg_lanes) {
    const unsigned long long mask = (1ull << dim) - 1;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long idx = f1_idx[i];
        const unsigned long long z = idx & mask, x = (idx >> dim) & mask, y = idx >> (2 * dim);
        Fr w = fr::mul(fr::load(eq_g + z * 8), fr::load(f1_val + i * 8));
        fr::store(w_out + i * 8, w);
        if (hg_lanes) lanes_add(hg_lanes + x * 8, fr::mul(w, fr::load(f3 + y * 8)));
    }
}

// initialize_phase_two (mod.rs:57-63): f1_gu[y] += eq_u[x] * f1_g[x | y << dim]; keys are xy (already >> dim when
// `shift` = 0, or the original f1 index with shift = dim).
__global__ void __launch_bounds__(128) gkr_phase2_scatter_kernel(const unsigned long long* keys, uint32_t shift,
                                                                const uint32_t* w, unsigned long long nnz, uint32_t dim,
                                                                const uint32_t* eq_u, unsigned long long* out_lanes) {
    const unsigned long long mask = (1ull << dim) - 1;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long xy = keys[i] >> shift;
        const unsigned long long x = xy & mask, y = xy >> dim;
        lanes_add(out_lanes + y * 8, fr::mul(fr::load(eq_u + x * 8), fr::load(w + i * 8)));
    }
}

// lanes (8 x u64 integer sums of limbs) -> fully reduced field element
__device__ __forceinline__ Fr lanes_to_fr(const unsigned long long* lanes) {
    Fr lo;
    unsigned long long c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        unsigned long long s = lanes[i];
        unsigned long long lo32 = (s & 0xffffffffull) + (c & 0xffffffffull);
        lo.l[i] = (uint32_t)lo32;
        c = (s >> 32) + (c >> 32) + (lo32 >> 32);
    }
    // value = lo + c * 2^256 with lo < 2^256 < 3p: two conditional subtractions, then add c*R mod p
    lo = fr::reduce_once(fr::reduce_once(lo));
    if (c != 0) {
        Fr craw = fr::zero();
        craw.l[0] = (uint32_t)c;
        craw.l[1] = (uint32_t)(c >> 32);
        lo = fr::add(lo, fr::mul(craw, fr_R2()));
    }
    return lo;
}

__global__ void __launch_bounds__(128) lanes_normalise_kernel(const unsigned long long* lanes, unsigned long long n,
                                                             uint32_t* out) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        fr::store(out + i * 8, lanes_to_fr(lanes + i * 8));
}

// out[i] = s * in[i]   — `zero += (f2_u, f3)` at mod.rs:71-75
__global__ void __launch_bounds__(128) scale_kernel(const uint32_t* in, const uint32_t* s, unsigned long long n, uint32_t* out) {
    Fr sc = fr::load(s);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        fr::store(out + i * 8, fr::mul(sc, fr::load(in + i * 8)));
}

// out = t[0] + r * (t[1] - t[0]): last fold of a 2-entry table = DenseMLE::evaluate's final step (mod.rs:122)
__global__ void fold_pair_kernel(const uint32_t* t, const uint32_t* r, uint32_t* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        Fr a = fr::load(t), b = fr::load(t + 8), rr = fr::load(r);
        fr::store(out, fr::add(a, fr::mul(rr, fr::sub(b, a))));
    }
}

// Montgomery -> canonical integers (what ark-serialize emits), n elements
__global__ void __launch_bounds__(128) to_canonical_kernel(const uint32_t* in, unsigned long long n, uint32_t* out) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        Fr one_int = fr::zero();
        one_int.l[0] = 1;
        fr::store(out + i * 8, fr::mul(fr::load(in + i * 8), one_int));
    }
}

// ---- multi-GPU helpers (capi_multi.inc)
// Switch to replicated rounds: dst[j][g * len + e] = src[g * n_tables + j][e] — every rank PULLS the current shard of every
// table from every rank (its own included) over NVLink peer memory, one launch instead of an NCCL all-gather per proof.
// The peers' stores were fenced at GPU scope before their last block reported the previous round (finish_round), and this
// rank has seen that round's global sums, so the data is complete; the loads bypass L1 (a stale line of an earlier proof).
// blockIdx.y = g * n_tables + j.
__global__ void __launch_bounds__(128) gather_tables_kernel(const uint32_t* const* src, uint32_t n_tables, unsigned long long len,
                                                           uint32_t* dst, unsigned long long dst_table_stride) {
    const uint32_t g = blockIdx.y / n_tables, j = blockIdx.y % n_tables;
    const uint32_t* from = src[blockIdx.y];
    uint32_t* to = dst + (size_t)j * dst_table_stride * 8 + (size_t)g * len * 8;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < len; e += (unsigned long long)gridDim.x * blockDim.x) {
        Fr v;
        asm volatile("ld.relaxed.sys.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v.l[0]), "=r"(v.l[1]), "=r"(v.l[2]), "=r"(v.l[3]), "=r"(v.l[4]), "=r"(v.l[5]), "=r"(v.l[6]), "=r"(v.l[7])
                     : "l"(from + e * 8)
                     : "memory");
        fr::store(to + e * 8, v);
    }
}

// evals[t] = sum over ranks of gathered[g][t]  (+ canonical form for the transcript); npts <= 32, one warp.
// fix1: the ranks summed only t = 0, 2, .., d (slot 1 is zero); P(1) = P_prev(r) - P(0) with P_prev = evals_out's old content.
__global__ void sum_partials_kernel(const uint32_t* gathered, uint32_t n_ranks, uint32_t npts, uint32_t* evals_out, uint32_t* canon_out,
                                    uint32_t fix1, Fr8 r8, const uint32_t* lagrange, uint32_t* host_out,
                                    volatile uint32_t* host_flag, uint32_t seq) {
    __shared__ uint32_t scratch[32 * 8];
    const uint32_t t = threadIdx.x;
    Fr r, claim = fr::zero();
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = r8.l[i];
    if (fix1) claim = claim_from_prev(evals_out, lagrange, r, npts - 1, scratch);
    Fr acc = fr::zero();
    if (t < npts)
        for (uint32_t g = 0; g < n_ranks; g++) acc = fr::add(acc, fr::load(gathered + ((size_t)g * npts + t) * 8));
    if (fix1) {  // lane 1 needs lane 0's P(0) and claim
        Fr p0, cl;
#pragma unroll
        for (int i = 0; i < 8; i++) { p0.l[i] = __shfl_sync(0xffffffffu, acc.l[i], 0); cl.l[i] = __shfl_sync(0xffffffffu, claim.l[i], 0); }
        if (t == 1) acc = fr::sub(cl, p0);
    }
    if (t < npts) {
        fr::store(evals_out + (size_t)t * 8, acc);
        Fr one_int = fr::zero();
        one_int.l[0] = 1;
        Fr cv = fr::mul(acc, one_int);
        fr::store(canon_out + (size_t)t * 8, cv);
        if (host_out) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                host_out[t * 8 + i] = acc.l[i];
                host_out[npts * 8 + t * 8 + i] = cv.l[i];
            }
        }
    }
    if (host_flag) {  // publish after every lane's message words are on their way (mapped pinned memory)
        __threadfence_system();
        __syncwarp();
        if (t == 0) *host_flag = seq;
    }
}
// ---- verifier-side helpers: ListOfProductsOfPolynomials::evaluate and check_and_generate_subclaim ----
// DenseMultilinearExtension::fix_variables(&[r]) alone: out[b] = in[2b] + r (in[2b+1] - in[2b]); blockIdx.y = table
__global__ void __launch_bounds__(128) fold_only_kernel(const uint32_t* const* tab_in, uint32_t* const* tab_out, unsigned long long n_out, Fr8 r8) {
    __shared__ __align__(16) uint32_t s_foldC[64];
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = r8.l[i];
    prepare_fold_consts(r, s_foldC);
    const uint32_t* in = tab_in[blockIdx.y];
    uint32_t* out = tab_out[blockIdx.y];
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < n_out;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        Fr e0 = fr::load_stream(in + b * 16), e1 = fr::load_stream(in + b * 16 + 8);
        fr::store(out + b * 8, fr::add(e0, fr::mul_round_const(s_foldC, fr::sub(e1, e0))));
    }
}
// sum_k c_k * prod_j value[idx] over the (fully folded, 1-element) tables  — data_structures.rs:100-108
__global__ void poly_combine_kernel(const uint32_t* const* tabs, const uint32_t* offsets, const uint32_t* indices, const uint32_t* coeffs,
                                    uint32_t n_products, uint32_t* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fr sum = fr::zero();
    for (uint32_t k = 0; k < n_products; k++) {
        Fr pr = fr::load(coeffs + 8 * k);
        for (uint32_t j = offsets[k]; j < offsets[k + 1]; j++) pr = fr::mul(pr, fr::load(tabs[indices[j]]));
        sum = fr::add(sum, pr);
    }
    fr::store(out, sum);
}
// check_and_generate_subclaim (verifier.rs:90-121): for every round P(0)+P(1) == expected, expected = P(r_i) by
// interpolation.  One warp.  result[0] = 0 accept / (1 + round) reject; expected_out = final expected evaluation.
__global__ void verify_kernel(uint32_t nv, uint32_t d, const uint32_t* claimed, const uint32_t* evals, const uint32_t* rand,
                              const uint32_t* lagrange, uint32_t* result, uint32_t* expected_out) {
    __shared__ uint32_t scratch[32 * 8];
    Fr expected = fr::load(claimed);
    uint32_t bad = 0;
    for (uint32_t i = 0; i < nv && !bad; i++) {
        const uint32_t* ev = evals + (size_t)i * (d + 1) * 8;
        Fr s = fr::add(fr::load(ev), fr::load(ev + 8));
        if (!fr::eq(s, expected)) bad = 1 + i;   // identical on every lane
        Fr nxt = claim_from_prev(ev, lagrange, fr::load(rand + (size_t)i * 8), d, scratch);
#pragma unroll
        for (int w = 0; w < 8; w++) expected.l[w] = __shfl_sync(0xffffffffu, nxt.l[w], 0);
    }
    if (threadIdx.x == 0) {
        result[0] = bad;
        fr::store(expected_out, expected);
    }
}

// ---- merged sparse output for the stand-alone initialize_phase_one API: after sorting (key, position) by key,
// flag segment heads, and let one thread per head sum its segment.
__global__ void __launch_bounds__(128) seg_heads_kernel(const unsigned long long* keys, unsigned long long n, uint32_t* head) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x)
        head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}
__global__ void __launch_bounds__(128) seg_sum_kernel(const unsigned long long* keys, const uint32_t* pos, const uint32_t* head,
                                                     const uint32_t* head_scan /* exclusive */, unsigned long long n,
                                                     const uint32_t* w, unsigned long long* keys_out, uint32_t* vals_out) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        if (!head[i]) continue;
        Fr acc = fr::load(w + (unsigned long long)pos[i] * 8);
        for (unsigned long long j = i + 1; j < n && !head[j]; j++) acc = fr::add(acc, fr::load(w + (unsigned long long)pos[j] * 8));
        keys_out[head_scan[i]] = keys[i];
        fr::store(vals_out + (unsigned long long)head_scan[i] * 8, acc);
    }
}

}  // namespace sck
