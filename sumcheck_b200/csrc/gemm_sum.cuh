// The multiply-reduce of prove_round as a tensor-core contraction (sm_100a, tcgen05.mma.kind::i8).
//
// prover.rs:110-148 sums, over all pairs b of the hypercube, the product of the multiplicands' evaluation lines.  For a product
// of three tables that is  P(t) = sum_b q_b(t) * A2_b(t),  q_b(t) = A0_b(t) * A1_b(t).  Only q_b needs per-pair multiplications;
// the OUTER sum is a contraction over b, and over the BYTES of the two factors
//     sum_b x_b * y_b = sum_{u,v} 2^(8(u+v)) * sum_b x_b[u] * y_b[v]
// it is a u8 x u8 -> s32 matrix product with the pair index as the K dimension:  D = X^T Y,  X = [pairs][bytes of x],
// Y = [pairs][bytes of y] — exactly what tcgen05.mma computes (both operands MN-major: the bytes of one pair are contiguous).
// So a thread only has to produce, per pair, three PLAIN 256 x 256-bit products of table values
//     q0 = a0*a1,   q1 = b0*b1,   qs = (a0+b0)*(a1+b1)             (a_j = table_j[2b], b_j = table_j[2b+1]; Karatsuba: qs-q0-q1
// is the middle coefficient of q_b(t)) as 64-byte integers in shared memory — no Montgomery reduction, no modular add, no lazy
// accumulator — and the tensor core multiplies them with the 64 bytes [a2 | b2] of the third table and sums over the pairs.
// The six sums  Z[i][j] = sum_b q_i * y_j  (i in {q0, q1, qs}, y in {a2, b2}) determine P completely:
//     P(t) = sum_ij wX_i(t) wY_j(t) Z[i][j],   wX = ((1-t)(1-2t), t(2t-1), t(1-t)),   wY = (1-t, t)
// which the host evaluates at t = 0..d after reducing the six big integers mod p (host_fr.h gemm_finish) — all exact, so the
// message is bit-identical to the reference's.  Per pair at degree 3: 192 IMAD.WIDE instead of 589 (round 1) / 573 (fold rounds).
//
// Products of FOUR tables use the same idea with both operands computed (X = three plain products of tables 0 and 1, Y = three of
// tables 2 and 3: a 192 x 192 contraction, nine sums); products of TWO tables need no per-pair multiplication at all (X and Y are the
// tables' own pairs).  See Shape<MM> below.
//
// Kernel shape: ONE persistent CTA per SM = G compute groups of 128 threads (a group owns one 128-pair tile at a time) + producer warps,
// each a single thread: W_TMA stages table tiles, W_FOLD (fold rounds) issues the fix_variables MMAs (tc_fold.cuh) into per-group
// accumulators in tensor memory, W_SUM issues the contraction MMAs of every group into ONE accumulator set (a single issuing thread
// keeps the accumulating MMAs ordered).  After its last tile a CTA adds the anti-diagonals of D (sum over u+v = k) into 64-bit totals
// in global memory; the last CTA carries them into the integers, exchanges them with the peer GPUs when the polynomial is sharded,
// and publishes them to mapped host memory.  Inside whole-proof calls a fold round may be launched AHEAD of its challenge (PRE).
#pragma once
#include "kernels.cuh"
#include "tc_fold.cuh"

namespace gsum {

using fr::Fr;

constexpr uint32_t TILE = 128;               // pairs per work item = threads per compute group
constexpr uint32_t OUT_SLOT_WORDS = 512;     // mapped result slot: NB * OUT_LIMBS words, sequence flag in the last word
constexpr uint32_t MAX_ITEMS_PER_CTA = 256;  // 4 K-steps of 32 pairs each: every s32 accumulator stays below 2^31
constexpr uint32_t TOT_STRIDE = 128;         // u64 totals per block pair (>= number of anti-diagonals)
constexpr uint32_t MAX_PRODUCTS = 32;        // the CSR of the product list is cached in shared memory (3 or 4 entries per product)

// The product list as the kernels use it, in shared memory: a dependent chain of global loads (offsets -> indices -> table
// pointer) per work item costs the single producer thread ~1700 cycles per table tile (measured) and made it the bottleneck.
struct CsrCache {
    uint32_t idx[4 * MAX_PRODUCTS];          // table index of CSR entry MM k + j
    uint32_t first[4 * MAX_PRODUCTS];        // 1 where the entry is the first use of its table (that use stores the fold)
    const uint32_t* in[4 * MAX_PRODUCTS];    // tab_in of the entry
    uint32_t* out[4 * MAX_PRODUCTS];         // tab_out of the entry (fold rounds)
};
template <int MM>
__device__ __forceinline__ void load_csr(CsrCache& c, const sck::RoundParams& p, bool fold) {
    for (uint32_t e = threadIdx.x; e < MM * p.n_products; e += blockDim.x) {
        const uint32_t jj = p.prod_offsets[e / MM] + e % MM, idx = p.prod_indices[jj];
        c.idx[e] = idx;
        c.first[e] = p.prod_first[jj];
        c.in[e] = p.tab_in[idx];
        c.out[e] = fold ? p.tab_out[idx] : nullptr;
    }
}

struct Params {
    sck::RoundParams rp;          // tables, CSR, tmaps (fold rounds: 128-byte rows), n_pairs, tile_base, host_out/host_flag/seq, counter, r
    const void* ymaps;            // round 1: [n_tables] CUtensorMap with 64-byte rows (one pair), SWIZZLE_64B, 128-row boxes
    unsigned long long* totals;   // [NB][TOT_STRIDE], zero before the launch; the publishing launch leaves it zero again
    uint32_t publish;             // 0: only add into totals (a round split over several launches), 1: the last CTA publishes
    uint32_t items;               // work items (tile, product) of this launch
    long long* prof;              // SC_GEMM_PROF=1: [16] cycle counters summed over the CTAs (waits per role), else null
    // A fold round launched AHEAD of its challenge (the host launches round i+1 right behind round i and hashes round i's message
    // while this kernel's prologue and first table tiles are already under way): the eight limbs of r arrive through mapped host
    // memory as {limb, sequence number} words, as in the resident kernel.  Null: r is in rp.r.
    const unsigned long long* r_mail;
    unsigned long long* r_bcast;  // device memory: CTA 0 alone polls the host (1184 threads polling over PCIe delay the very write they
                                  // wait for by ~100 us: measured) and re-publishes the words here for the other CTAs
    uint32_t r_seq;
    long long r_timeout;          // clock64 ticks before the kernel gives up (continues with r = 0 and raises *r_error)
    uint32_t* r_error;            // mapped host word
};

// phase markers (SC_GEMM_PROF): max over the CTAs of the cycles since the CTA started, slots 12..
__device__ __forceinline__ void prof_mark(long long* prof, int slot, long long t_start) {
    if (prof && threadIdx.x == 0) atomicMax((unsigned long long*)prof + slot, (unsigned long long)(clock64() - t_start));
}

struct ProfTimer {  // accumulates clock64 intervals of one thread; flushed with one atomicAdd per counter at the end
    long long acc = 0, t0 = 0;
    __device__ __forceinline__ void start(bool on) { if (on) t0 = clock64(); }
    __device__ __forceinline__ void stop(bool on) { if (on) acc += clock64() - t0; }
    __device__ __forceinline__ void flush(long long* prof, int slot) { if (prof) atomicAdd((unsigned long long*)prof + slot, (unsigned long long)acc); }
};

// ---- MN-major operand descriptors (checked by tools/microbench/gemmsum.cu) ------------------------------------------------------
// [k = pair][bytes] with the bytes of a pair contiguous: 128-byte rows + SWIZZLE_128B (layout type 2, 8-row groups 1024 B apart)
// or 64-byte rows + SWIZZLE_64B (layout type 4, 8-row groups 512 B apart).  K = 32 pairs per instruction.
// lbo_bytes: distance between swizzle atoms along the byte (MN) direction, for operands wider than one atom.
__device__ __forceinline__ uint64_t mn_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout_type, uint32_t lbo_bytes = 16) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)layout_type << 61);
}

// ---- the two shapes --------------------------------------------------------------------------------------------------------------
// MM = multiplicands per product.  The X side is always the first TWO tables of a product: three plain products q0, q1, qs of 64 bytes
// (X = 192 bytes per pair: XA = [q0 | q1] with SWIZZLE_128B, XQ = qs with SWIZZLE_64B).  The Y side is
//   MM = 3: the third table's pair [a2 | b2], two blocks of 32 bytes (one SWIZZLE_64B array, N = 64);
//   MM = 4: the three plain products s0, s1, ss of tables 2 and 3, three blocks of 64 bytes (three SWIZZLE_64B arrays 8 KiB apart,
//           N = 192 through the descriptor's leading byte offset) — again no modular arithmetic anywhere in the hot loop.
// D = X^T Y has 3 x NBY block pairs; block pair (i, j) yields sum_b X_i * Y_j as an integer of OUT_LIMBS limbs.
//   MM = 2 (GKR phases, degree-2 lists): X = the first table's pair [a0 | b0], Y = the second table's pair — NO per-pair
//           multiplication at all: in round 1 both operands are the tiles as TMA lands them, a fold round only reads the folded pairs
//           out of tensor memory and writes them back as operands.  One M = 64 accumulator (XQ's place), two blocks of 32 bytes a side.
template <int MM>
struct Shape;
template <>
struct Shape<2> {
    static constexpr uint32_t NBX = 2, BX = 32, NBY = 2, BY = 32, N = 64, NB = 4, DIAG = 63, ES = 64, OUT_LIMBS = 18, Y_BYTES = 8192;
};
template <>
struct Shape<3> {
    static constexpr uint32_t NBX = 3, BX = 64, NBY = 2, BY = 32, N = 64, NB = 6, DIAG = 95, ES = 96, OUT_LIMBS = 26, Y_BYTES = 8192;
};
template <>
struct Shape<4> {
    static constexpr uint32_t NBX = 3, BX = 64, NBY = 3, BY = 64, N = 192, NB = 9, DIAG = 127, ES = 128, OUT_LIMBS = 34, Y_BYTES = 24576;
};
// c_format S32 @4, a/b U8, a_major = b_major = MN @15/@16, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t idesc_u8_mn(uint32_t M, uint32_t N) {
    return (2u << 4) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- per-pair arithmetic ------------------------------------------------------------------------------------------------------
// o = a * b as a plain 512-bit integer (64 IMAD.WIDE + the merge of the even/odd accumulators)
__device__ __forceinline__ void mul_plain(const Fr& a, const Fr& b, uint32_t (&o)[16]) {
    uint32_t ev[16], od[16], c0;
    fr::mul_wide_eo(ev, od, a.l, b.l);
    o[0] = ev[0];
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;\n\t"
        : "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]), "=r"(c0)
        : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]), "r"(ev[8]), "r"(od[0]), "r"(od[1]), "r"(od[2]),
          "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]));
    asm("{ .reg .u32 t_; add.cc.u32 t_, %21, 0xffffffff; }\n\t"
        "addc.cc.u32 %0, %7, %14;\n\t"
        "addc.cc.u32 %1, %8, %15;\n\t"
        "addc.cc.u32 %2, %9, %16;\n\t"
        "addc.cc.u32 %3, %10, %17;\n\t"
        "addc.cc.u32 %4, %11, %18;\n\t"
        "addc.cc.u32 %5, %12, %19;\n\t"
        "addc.u32 %6, %13, %20;\n\t"
        : "=r"(o[9]), "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15])
        : "r"(ev[9]), "r"(ev[10]), "r"(ev[11]), "r"(ev[12]), "r"(ev[13]), "r"(ev[14]), "r"(ev[15]), "r"(od[8]), "r"(od[9]), "r"(od[10]),
          "r"(od[11]), "r"(od[12]), "r"(od[13]), "r"(od[14]), "r"(c0));
}
// ---- one-level Karatsuba: 3 x 16 = 48 IMAD.WIDE instead of 64, the glue (~65 adds) on the ALU pipe, which idles while the multiply
// pipe is the binding unit of round 1 (ncu: 67 % against 17 %).  MEASURED SLOWER and therefore off: bit-exact, but round 1 0.343 ->
// 0.354 ms, round 2 0.394 -> 0.414 ms, config 4 2.45 -> 2.72 ms — the kernels are short of issue slots and latency hiding (12 warps
// per SM), not of multiplier throughput alone, and ptxas leaves 32 of the 144 products unfused (IMAD + IMAD.HI).  Kept as a switch.
#ifndef SC_GEMM_KARATSUBA
#define SC_GEMM_KARATSUBA 0
#endif
// z = a * b for 4-limb operands (8 limbs)
__device__ __forceinline__ void mul4(const uint32_t (&a)[4], const uint32_t (&b)[4], uint32_t (&z)[8]) {
    uint32_t ev[8], od[8];
    fr::mul_wide_eo4(ev, od, a, b);
    z[0] = ev[0];
    asm("add.cc.u32 %0, %7, %14;\n\t"
        "addc.cc.u32 %1, %8, %15;\n\t"
        "addc.cc.u32 %2, %9, %16;\n\t"
        "addc.cc.u32 %3, %10, %17;\n\t"
        "addc.cc.u32 %4, %11, %18;\n\t"
        "addc.cc.u32 %5, %12, %19;\n\t"
        "addc.u32 %6, %13, %20;\n\t"
        : "=r"(z[1]), "=r"(z[2]), "=r"(z[3]), "=r"(z[4]), "=r"(z[5]), "=r"(z[6]), "=r"(z[7])
        : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]), "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]),
          "r"(od[4]), "r"(od[5]), "r"(od[6]));
}
// r = x + y (4 limbs), returns the carry
__device__ __forceinline__ uint32_t add4(const uint32_t* x, const uint32_t* y, uint32_t (&r)[4]) {
    uint32_t c;
    asm("add.cc.u32 %0, %5, %9;\n\t"
        "addc.cc.u32 %1, %6, %10;\n\t"
        "addc.cc.u32 %2, %7, %11;\n\t"
        "addc.cc.u32 %3, %8, %12;\n\t"
        "addc.u32 %4, 0, 0;\n\t"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(c)
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]));
    return c;
}
// o = a * b as a plain 512-bit integer: z0 = a_lo b_lo, z2 = a_hi b_hi, z1 = (a_lo + a_hi)(b_lo + b_hi) - z0 - z2
__device__ __forceinline__ void mul_plain_karatsuba(const Fr& a, const Fr& b, uint32_t (&o)[16]) {
    uint32_t alo[4] = {a.l[0], a.l[1], a.l[2], a.l[3]}, ahi[4] = {a.l[4], a.l[5], a.l[6], a.l[7]};
    uint32_t blo[4] = {b.l[0], b.l[1], b.l[2], b.l[3]}, bhi[4] = {b.l[4], b.l[5], b.l[6], b.l[7]};
    uint32_t z0[8], z2[8], zm[9], sa[4], sb[4];
    mul4(alo, blo, z0);
    mul4(ahi, bhi, z2);
    const uint32_t ca = add4(alo, ahi, sa), cb = add4(blo, bhi, sb);
    {
        uint32_t t[8];
        mul4(sa, sb, t);
#pragma unroll
        for (int i = 0; i < 8; i++) zm[i] = t[i];
    }
    // (sa + ca 2^128)(sb + cb 2^128) = sa sb + (ca sb + cb sa) 2^128 + ca cb 2^256
    const uint32_t ma = 0u - ca, mb = 0u - cb;
    asm("add.cc.u32 %0, %0, %5;\n\t"
        "addc.cc.u32 %1, %1, %6;\n\t"
        "addc.cc.u32 %2, %2, %7;\n\t"
        "addc.cc.u32 %3, %3, %8;\n\t"
        "addc.u32 %4, %9, 0;\n\t"
        : "+r"(zm[4]), "+r"(zm[5]), "+r"(zm[6]), "+r"(zm[7]), "=r"(zm[8])
        : "r"(sb[0] & ma), "r"(sb[1] & ma), "r"(sb[2] & ma), "r"(sb[3] & ma), "r"(ca & cb));
    asm("add.cc.u32 %0, %0, %5;\n\t"
        "addc.cc.u32 %1, %1, %6;\n\t"
        "addc.cc.u32 %2, %2, %7;\n\t"
        "addc.cc.u32 %3, %3, %8;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        : "+r"(zm[4]), "+r"(zm[5]), "+r"(zm[6]), "+r"(zm[7]), "+r"(zm[8])
        : "r"(sa[0] & mb), "r"(sa[1] & mb), "r"(sa[2] & mb), "r"(sa[3] & mb));
    // z1 = zm - z0 - z2 (non-negative, below 2^258: nine limbs)
#define SC_SUB9(z)                                                                                                              \
    asm("sub.cc.u32 %0, %0, %9;\n\t"                                                                                             \
        "subc.cc.u32 %1, %1, %10;\n\t"                                                                                           \
        "subc.cc.u32 %2, %2, %11;\n\t"                                                                                           \
        "subc.cc.u32 %3, %3, %12;\n\t"                                                                                           \
        "subc.cc.u32 %4, %4, %13;\n\t"                                                                                           \
        "subc.cc.u32 %5, %5, %14;\n\t"                                                                                           \
        "subc.cc.u32 %6, %6, %15;\n\t"                                                                                           \
        "subc.cc.u32 %7, %7, %16;\n\t"                                                                                           \
        "subc.u32 %8, %8, 0;\n\t"                                                                                                \
        : "+r"(zm[0]), "+r"(zm[1]), "+r"(zm[2]), "+r"(zm[3]), "+r"(zm[4]), "+r"(zm[5]), "+r"(zm[6]), "+r"(zm[7]), "+r"(zm[8])      \
        : "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]))
    SC_SUB9(z0);
    SC_SUB9(z2);
#undef SC_SUB9
    // o = z0 + z1 2^128 + z2 2^256
    o[0] = z0[0]; o[1] = z0[1]; o[2] = z0[2]; o[3] = z0[3];
    uint32_t c0;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;\n\t"
        : "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]), "=r"(o[9]), "=r"(o[10]), "=r"(o[11]), "=r"(c0)
        : "r"(z0[4]), "r"(z0[5]), "r"(z0[6]), "r"(z0[7]), "r"(z2[0]), "r"(z2[1]), "r"(z2[2]), "r"(z2[3]), "r"(zm[0]), "r"(zm[1]), "r"(zm[2]),
          "r"(zm[3]), "r"(zm[4]), "r"(zm[5]), "r"(zm[6]), "r"(zm[7]));
    asm("{ .reg .u32 t_; add.cc.u32 t_, %9, 0xffffffff; }\n\t"
        "addc.cc.u32 %0, %4, %8;\n\t"
        "addc.cc.u32 %1, %5, 0;\n\t"
        "addc.cc.u32 %2, %6, 0;\n\t"
        "addc.u32 %3, %7, 0;\n\t"
        : "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15])
        : "r"(z2[4]), "r"(z2[5]), "r"(z2[6]), "r"(z2[7]), "r"(zm[8]), "r"(c0));
}

// a + b as a plain integer (both canonical: the sum is below 2p < 2^256)
__device__ __forceinline__ Fr add_plain(const Fr& a, const Fr& b) {
    Fr r;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;\n\t"
        : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]), "r"(b.l[0]), "r"(b.l[1]),
          "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
    return r;
}

// 16 limbs (64 bytes) of pair `row` into a [pairs][128 B] SWIZZLE_128B operand at 16-byte chunks c0..c0+3 of the row
__device__ __forceinline__ void sts_sw128(uint8_t* base, uint32_t row, uint32_t c0, const uint32_t* v) {
#pragma unroll
    for (uint32_t c = 0; c < 4; c++)
        *reinterpret_cast<uint4*>(base + row * 128u + (((c0 + c) ^ (row & 7u)) << 4)) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
// ... into a [pairs][64 B] SWIZZLE_64B operand
__device__ __forceinline__ void sts_sw64(uint8_t* base, uint32_t row, const uint32_t* v) {
#pragma unroll
    for (uint32_t c = 0; c < 4; c++)
        *reinterpret_cast<uint4*>(base + row * 64u + ((c ^ ((row >> 1) & 3u)) << 4)) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}

__device__ __forceinline__ void mul_plain_sel(const Fr& a, const Fr& b, uint32_t (&o)[16]) {
#if SC_GEMM_KARATSUBA
    mul_plain_karatsuba(a, b, o);
#else
    mul_plain(a, b, o);
#endif
}
// The three products of one pair into the group's X operand: XA row = [q0 | q1] (SWIZZLE_128B), XQ row = qs (SWIZZLE_64B)
__device__ __forceinline__ void products_to_smem(const Fr& a0, const Fr& b0, const Fr& a1, const Fr& b1, uint8_t* XA, uint8_t* XQ, uint32_t row) {
    uint32_t o[16];
    mul_plain_sel(a0, a1, o);
    sts_sw128(XA, row, 0, o);
    mul_plain_sel(b0, b1, o);
    sts_sw128(XA, row, 4, o);
    mul_plain_sel(add_plain(a0, b0), add_plain(a1, b1), o);
    sts_sw64(XQ, row, o);
}

// The three products of the Y side (MM = 4) into the group's three Y arrays
__device__ __forceinline__ void y_products_to_smem(const Fr& a2, const Fr& b2, const Fr& a3, const Fr& b3, uint8_t* Y, uint32_t row) {
    uint32_t o[16];
    mul_plain_sel(a2, a3, o);
    sts_sw64(Y, row, o);
    mul_plain_sel(b2, b3, o);
    sts_sw64(Y + 8192, row, o);
    mul_plain_sel(add_plain(a2, b2), add_plain(a3, b3), o);
    sts_sw64(Y + 16384, row, o);
}

// ---- contraction of one 128-pair tile: D1 (128 x N) += XA^T Y, D2 (64 x N) += XQ^T Y; issued by ONE thread -----------------------
template <int MM>
__device__ __forceinline__ void issue_contraction(uint32_t xa_smem, uint32_t xq_smem, uint32_t y_smem, uint32_t tmem_d, uint32_t accumulate) {
    constexpr uint32_t N = Shape<MM>::N, I128 = idesc_u8_mn(128, N), I64 = idesc_u8_mn(64, N);
#pragma unroll
    for (uint32_t ks = 0; ks < 4; ks++) {
        const uint64_t b = mn_desc(y_smem + ks * 2048u, 512, 4, 8192);
        if (MM == 2) {  // one 64-byte X operand (the first table's pair), one accumulator
            umma(tmem_d, mn_desc(xq_smem + ks * 2048u, 512, 4), b, I64, accumulate | ks);
        } else {
            umma(tmem_d, mn_desc(xa_smem + ks * 4096u, 1024, 2), b, I128, accumulate | ks);
            umma(tmem_d + N, mn_desc(xq_smem + ks * 2048u, 512, 4), b, I64, accumulate | ks);
        }
    }
}

// ---- epilogue: D -> anti-diagonal sums -> global totals -> (last CTA) the NB integers --------------------------------------------
// D1 lanes 0..127 = byte u of q0 (lanes 0..63) / q1 (64..127), columns 0..N-1; D2 rows 0..63 = byte u of qs in lanes
// (u % 16) + 32 (u / 16), columns N..2N-1; column c = byte c % BY of Y block c / BY.  Block pair bp = NBY i + j, diagonal k = u + v.
constexpr uint32_t NB3 = Shape<3>::NB, OUT_LIMBS3 = Shape<3>::OUT_LIMBS;

// s_E: [2][NB * ES] (low / high 16-bit halves of the accumulators, summed as u32).  Called by the whole CTA; warps 0..3 read TMEM.
// scratch: shared memory that is idle by now (the operand ring), >= (1 + n_ranks) * NB * OUT_LIMBS words.
template <int MM>
__device__ __forceinline__ void epilogue(const Params& P, uint32_t tmem_d, bool has_work, uint32_t* s_E, bool* s_last, uint32_t* scratch,
                                         long long t_start = 0) {
    using S_ = Shape<MM>;
    constexpr uint32_t NB = S_::NB, ES = S_::ES, DIAG = S_::DIAG, OUT_LIMBS = S_::OUT_LIMBS, N = S_::N, BY = S_::BY, NBY = S_::NBY;
    constexpr uint32_t PARTS = (MM == 2 ? 1 : 2) * N / 32;  // MM = 2 has the M = 64 accumulator only (at column 0)
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    prof_mark(P.prof, 13, t_start);  // main loop done
    for (uint32_t i = tid; i < 2 * NB * ES; i += blockDim.x) s_E[i] = 0;
    __syncthreads();
    tcf::tc_fence_after();
    if (warp < 4 && has_work) {  // (a CTA without work never wrote its accumulators)
        const uint32_t lane_addr = tmem_d + ((warp * 32u) << 16);
        uint32_t S[32];
#pragma unroll 1
        for (uint32_t part = 0; part < PARTS; part++) {  // 32 columns at a time: D1 first, then D2
            const bool second = MM == 2 || part >= N / 32;
            tcf::tmem_ld32(lane_addr + part * 32u, S);  // (whole warp: .sync.aligned)
            tcf::tmem_ld_wait();
            if (second && lane >= 16) continue;  // an M = 64 accumulator lives in lanes 0..15 of every 32-lane quadrant
            const uint32_t row = warp * 16u + lane;  // (of an M = 64 accumulator)
            const uint32_t i = MM == 2 ? (row >> 5) : (second ? 2u : (tid >> 6)), u = MM == 2 ? (row & 31u) : (second ? row : (tid & 63u));
            const uint32_t c0 = ((second && MM != 2) ? part - N / 32 : part) * 32u, j = c0 / BY, v0 = c0 % BY;
            uint32_t* lo = s_E + (NBY * i + j) * ES + u + v0;
            uint32_t* hi = lo + NB * ES;
#pragma unroll
            for (int v = 0; v < 32; v++) {
                atomicAdd(lo + v, S[v] & 0xffffu);
                atomicAdd(hi + v, S[v] >> 16);
            }
        }
    }
    tcf::tc_fence_before();
    __syncthreads();
    for (uint32_t i = tid; i < NB * ES; i += blockDim.x) {
        const uint32_t bp = i / ES, k = i % ES;
        if (k >= DIAG) continue;
        const unsigned long long v = (unsigned long long)s_E[i] + ((unsigned long long)s_E[NB * ES + i] << 16);
        if (v) atomicAdd(P.totals + bp * TOT_STRIDE + k, v);
    }
    __threadfence();
    __syncthreads();
    prof_mark(P.prof, 14, t_start);  // totals added
    if (tid == 0) {
        const unsigned int ticket = atomicAdd(P.rp.counter, 1u);
        *s_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!*s_last) return;
    __threadfence();
    if (tid == 0) *P.rp.counter = 0;
    if (!P.publish) return;
    // the grid's totals into shared memory with independent loads (a dependent chain of L2 round trips per integer cost
    // ~25 us per launch), leaving the global copy zero for the next launch
    unsigned long long* s_tot = reinterpret_cast<unsigned long long*>(s_E);  // [NB][ES]
    __syncthreads();
    for (uint32_t i = tid; i < NB * ES; i += blockDim.x) {
        const uint32_t bp = i / ES, k = i % ES;
        unsigned long long v = 0;
        if (k < DIAG) {
            v = __ldcg(P.totals + bp * TOT_STRIDE + k);
            P.totals[bp * TOT_STRIDE + k] = 0;
        }
        s_tot[i] = v;
    }
    __syncthreads();
    if (tid < NB) {  // Z = sum_k 2^(8k) totals[k]: byte-serial carry into OUT_LIMBS limbs
        unsigned long long acc = 0;
        uint32_t limb = 0;
#pragma unroll 8
        for (uint32_t k = 0; k < OUT_LIMBS * 4; k++) {
            if (k < DIAG) acc += s_tot[tid * ES + k];
            limb |= (uint32_t)(acc & 0xffu) << (8 * (k & 3u));
            acc >>= 8;
            if ((k & 3u) == 3u) {
                scratch[tid * OUT_LIMBS + (k >> 2)] = limb;
                limb = 0;
            }
        }
    }
    __syncthreads();
    if (tid >= 32) return;
    constexpr uint32_t NW = NB * OUT_LIMBS;
    static_assert(2 * NW <= sck::MAIL_WORDS, "a rank's integers must fit its mailbox slot ({limb, sequence number} pairs)");
    static_assert(NW < OUT_SLOT_WORDS - 40, "the integers must leave room for the challenge words and the flag in the result slot");
    if (P.rp.peer_mail) {
        // sharded polynomial: all-to-all of the integers over NVLink peer memory — every limb ONE 8-byte {limb, sequence
        // number} store into the receiver's mailbox (kernels.cuh mail_store: single-copy atomic), then the integer sums over
        // the ranks (identical on every rank; they fit the limbs: 8 ranks add 3 bits)
        const sck::RoundParams& p = P.rp;
        const uint32_t G = p.n_ranks;
        uint32_t* rows = scratch + NW;
        for (uint32_t w = lane; w < NW; w += 32) {
            const uint32_t v = scratch[w];
            for (uint32_t g = 0; g < G; g++) sck::mail_store(p.peer_mail[g] + ((size_t)p.mail_slot * G + p.rank) * sck::MAIL_WORDS + 2 * w, v, p.mail_seq);
        }
        const uint32_t* mine = p.peer_mail[p.rank] + (size_t)p.mail_slot * G * sck::MAIL_WORDS;
        for (uint32_t w = lane; w < NW; w += 32)
            for (uint32_t g = 0; g < G; g++) {
                uint32_t d, f;
                const long long t0 = clock64();
                for (;;) {
                    sck::mail_load(mine + (size_t)g * sck::MAIL_WORDS + 2 * w, d, f);
                    if (f == p.mail_seq) break;
                    if (clock64() - t0 > p.mail_timeout) {  // a peer died: report instead of hanging the GPU
                        *p.comm_error = 1;
                        d = 0;
                        break;
                    }
                }
                rows[g * NW + w] = d;
            }
        __syncwarp();
        if (lane < NB) {
            unsigned long long c = 0;
            for (uint32_t i = 0; i < OUT_LIMBS; i++) {
                for (uint32_t g = 0; g < G; g++) c += rows[g * NW + lane * OUT_LIMBS + i];
                scratch[lane * OUT_LIMBS + i] = (uint32_t)c;
                c >>= 32;
            }
        }
        __syncwarp();
    }
    for (uint32_t w = tid; w < NW; w += 32) P.rp.host_out[w] = scratch[w];
    __threadfence_system();
    __syncwarp();
    if (tid == 0 && P.rp.host_flag) *P.rp.host_flag = P.rp.seq;
    prof_mark(P.prof, 15, t_start);  // published (last CTA)
}

// Work split of one CTA: group g owns the items w = blockIdx.x * G + g + n * stride, n < n_items(g).  n_items is non-increasing
// in g and differs by at most one between groups, so all G groups take part in every step n < n_min and the first `rem` groups
// in the (possibly) partial last step n = n_min.  Unit (n, j, g) is the u-th of the CTA with u = MM G n + j * groups(n) + g (MM table tiles per item).
template <int G, int MM = 3>
struct Split {
    uint32_t n_min, rem, stride, first;
    __device__ __forceinline__ Split(uint32_t items) {
        stride = gridDim.x * G;
        first = blockIdx.x * G;
        const uint32_t last_g = first + G - 1;
        n_min = last_g < items ? (items - last_g + stride - 1) / stride : 0u;
        rem = 0;
#pragma unroll
        for (int g = 0; g < G - 1; g++) {
            const uint32_t f = first + g;
            const uint32_t ni = f < items ? (items - f + stride - 1) / stride : 0u;
            rem += ni > n_min ? 1u : 0u;
        }
    }
    __device__ __forceinline__ uint32_t n_items(uint32_t g) const { return n_min + (g < rem ? 1u : 0u); }
    __device__ __forceinline__ uint32_t steps() const { return n_min + (rem ? 1u : 0u); }
    __device__ __forceinline__ uint32_t groups(uint32_t n) const { return n < n_min ? (uint32_t)G : rem; }
    __device__ __forceinline__ uint32_t unit(uint32_t n, uint32_t j, uint32_t g) const { return (uint32_t)MM * G * n + j * groups(n) + g; }
    __device__ __forceinline__ uint32_t item(uint32_t n, uint32_t g) const { return first + g + n * stride; }
};

// ================================================================================================ round 1 (no fold)
// MM = 3: tables 0 and 1 of the product are read by the threads (two LDG.256 per table, one item ahead), table 2 goes HBM -> shared
// memory by TMA with 64-byte rows and is the Y operand as it lands.  MM = 4: all four tables are read by the threads (both
// operands are products).
template <int G, int MM>
struct R1Smem {
    static constexpr uint32_t XA = 0, XQ = 16384, Y0 = 24576, Y1 = Y0 + Shape<MM>::Y_BYTES;  // MM = 3: two Y slots (TMA, double-buffered)
    static constexpr uint32_t GROUP = MM == 3 ? Y1 + 8192 : Y1;
    static constexpr size_t BYTES = (size_t)G * GROUP;
};

template <int G, int MM>
__global__ void __launch_bounds__(G * 128 + 64, 1) gemm_round1_kernel(const Params P) {
    using L = R1Smem<G, MM>;
    using S_ = Shape<MM>;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t y_full[G][2], y_empty[G][2], x_full[G], x_empty[G];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint32_t s_E[2 * S_::NB * S_::ES];
    __shared__ bool s_last;
    __shared__ CsrCache s_csr;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const sck::RoundParams& p = P.rp;
    load_csr<MM>(s_csr, p, false);
    if (tid == 0) {
        if (tcf::smem_u32(smem) & 1023u) __trap();
        for (int g = 0; g < G; g++) {
            tcf::mbar_init(&y_full[g][0], 1);
            tcf::mbar_init(&y_full[g][1], 1);
            tcf::mbar_init(&y_empty[g][0], 1);
            tcf::mbar_init(&y_empty[g][1], 1);
            tcf::mbar_init(&x_full[g], 4);
            tcf::mbar_init(&x_empty[g], 1);
        }
        tcf::fence_mbar_init();
    }
    constexpr uint32_t TMEM_COLS = MM == 3 ? 128 : 512;  // 2 N accumulator columns, a power of two
    if (warp == 0) tcf::tmem_alloc(&s_tmem, TMEM_COLS);
    tcf::tc_fence_before();
    __syncthreads();
    tcf::tc_fence_after();
    const uint32_t tmem = s_tmem;
    const Split<G> sp(P.items);
    const bool one_product = p.n_products == 1;
    auto locate = [&](uint32_t n, uint32_t g, uint32_t& k, uint32_t& tile) {
        const uint32_t w = sp.item(n, g);
        k = one_product ? 0u : w % p.n_products;
        tile = p.tile_base + (one_product ? w : w / p.n_products);
    };
    if (warp < G * 4) {
        // ---------------------------------------------------------------------------------------------------- compute group
        const uint32_t g = warp >> 2, t = tid & 127u;
        uint8_t* const base = smem + (size_t)g * L::GROUP;
        const uint32_t n_items = sp.n_items(g);
        // this thread's pairs of two tables of the product for item n
        auto fetch = [&](uint32_t n, uint32_t j, Fr& x0, Fr& y0, Fr& x1, Fr& y1) {
            uint32_t k, tile;
            locate(n, g, k, tile);
            const unsigned long long b = (unsigned long long)tile * TILE + t;
            const uint32_t* s0 = s_csr.in[MM * k + j] + b * 16;
            const uint32_t* s1 = s_csr.in[MM * k + j + 1] + b * 16;
            x0 = fr::load_stream(s0);
            y0 = fr::load_stream(s0 + 8);
            x1 = fr::load_stream(s1);
            y1 = fr::load_stream(s1 + 8);
        };
        Fr a0, b0, a1, b1, na0, nb0, na1, nb1;
        if (n_items) fetch(0, 0, a0, b0, a1, b1);
        for (uint32_t n = 0; n < n_items; n++) {
            Fr a2, b2, a3, b3;
            if (MM == 4) fetch(n, 2, a2, b2, a3, b3);
            // the next item's pairs are requested before this item's products are computed: a group always has loads in flight
            if (n + 1 < n_items) fetch(n + 1, 0, na0, nb0, na1, nb1);
            if (n > 0) tcf::mbar_wait(&x_empty[g], (n - 1) & 1u);  // the contraction of item n-1 has read the operands
            products_to_smem(a0, b0, a1, b1, base + L::XA, base + L::XQ, t);
            if (MM == 4) y_products_to_smem(a2, b2, a3, b3, base + L::Y0, t);
            a0 = na0; b0 = nb0; a1 = na1; b1 = nb1;
            tcf::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tcf::mbar_arrive(&x_full[g]);
        }
    } else if (warp == G * 4) {
        // ---------------------------------------------------------------------------------------------------- TMA of the Y tiles (MM = 3)
        if (MM == 3 && lane == 0) {
            const uint32_t steps = sp.steps();
            for (uint32_t n = 0; n < steps; n++) {
                const uint32_t ng = sp.groups(n);
#pragma unroll
                for (uint32_t g = 0; g < (uint32_t)G; g++) {
                    if (g >= ng) continue;
                    const uint32_t s = n & 1u;
                    if (n >= 2) tcf::mbar_wait(&y_empty[g][s], ((n >> 1) - 1u) & 1u);
                    uint32_t k, tile;
                    locate(n, g, k, tile);
                    tcf::mbar_expect_tx(&y_full[g][s], 8192);
                    tcf::tma_load_tile(smem + (size_t)g * L::GROUP + (s ? L::Y1 : L::Y0), (const uint8_t*)P.ymaps + (size_t)s_csr.idx[MM * k + 2] * 128,
                                       &y_full[g][s], tile * TILE);
                }
            }
        }
    } else {
        // ---------------------------------------------------------------------------------------------------- contraction MMAs
        if (lane == 0) {
            uint32_t acc = 0;
            const uint32_t steps = sp.steps();
            for (uint32_t n = 0; n < steps; n++) {
                const uint32_t ng = sp.groups(n);
#pragma unroll
                for (uint32_t g = 0; g < (uint32_t)G; g++) {
                    if (g >= ng) continue;
                    const uint32_t s = MM == 3 ? (n & 1u) : 0u, gb = tcf::smem_u32(smem + (size_t)g * L::GROUP);
                    if (MM == 3) tcf::mbar_wait(&y_full[g][s], (n >> 1) & 1u);
                    tcf::mbar_wait(&x_full[g], n & 1u);
                    tcf::tc_fence_after();
                    issue_contraction<MM>(gb + L::XA, gb + L::XQ, gb + (s ? L::Y1 : L::Y0), tmem, acc);
                    acc = 1;
                    tcf::umma_commit(&x_empty[g]);
                    if (MM == 3) tcf::umma_commit(&y_empty[g][s]);
                }
            }
            // every MMA has completed once the last commit of every group has arrived
#pragma unroll
            for (uint32_t g = 0; g < (uint32_t)G; g++)
                if (sp.n_items(g)) tcf::mbar_wait(&x_empty[g], (sp.n_items(g) - 1) & 1u);
        }
    }
    __syncwarp();
    tcf::tc_fence_before();
    __syncthreads();
    epilogue<MM>(P, tmem, sp.n_items(0) > 0, s_E, &s_last, reinterpret_cast<uint32_t*>(smem));
    __syncthreads();
    if (warp == 0) tcf::tmem_dealloc(tmem, TMEM_COLS);
}

// ================================================================================================ round 1 of two-table products
// Both operands are the tables' tiles as TMA lands them (64-byte rows = one pair, SWIZZLE_64B): the kernel is a TMA warp, an MMA warp
// and four warps that only run the epilogue — not one multiplication on the CUDA cores.  Items (tile, product) are dealt to the CTAs
// round-robin; a ring of RAW_STAGES (X, Y) tile pairs.
constexpr uint32_t RAW_STAGES = 6, RAW_STAGE_BYTES = 16384, RAW_THREADS = 192;  // 96 KiB + 64 tensor-memory columns: two CTAs per SM
constexpr size_t RAW_SMEM = (size_t)RAW_STAGES * RAW_STAGE_BYTES;

template <int MM = 2>  // (a template only so that the header can be included by several translation units)
__global__ void __launch_bounds__(RAW_THREADS, 2) gemm_round1_raw_kernel(const Params P) {
    static_assert(MM == 2, "two-table products");
    using S_ = Shape<2>;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[RAW_STAGES], empty[RAW_STAGES], done;
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint32_t s_E[2 * S_::NB * S_::ES];
    __shared__ bool s_last;
    __shared__ CsrCache s_csr;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const sck::RoundParams& p = P.rp;
    load_csr<2>(s_csr, p, false);
    if (tid == 0) {
        if (tcf::smem_u32(smem) & 1023u) __trap();
        for (uint32_t s = 0; s < RAW_STAGES; s++) {
            tcf::mbar_init(&full[s], 1);
            tcf::mbar_init(&empty[s], 1);
        }
        tcf::mbar_init(&done, 1);
        tcf::fence_mbar_init();
    }
    if (warp == 0) tcf::tmem_alloc(&s_tmem, 64);
    tcf::tc_fence_before();
    __syncthreads();
    tcf::tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t n_items = blockIdx.x < P.items ? (P.items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    const bool one_product = p.n_products == 1;
    if (warp == 4) {
        if (lane == 0)
            for (uint32_t n = 0; n < n_items; n++) {
                const uint32_t slot = n % RAW_STAGES, w = blockIdx.x + n * gridDim.x;
                const uint32_t k = one_product ? 0u : w % p.n_products, tile = p.tile_base + (one_product ? w : w / p.n_products);
                if (n >= RAW_STAGES) tcf::mbar_wait(&empty[slot], ((n / RAW_STAGES) - 1u) & 1u);
                uint8_t* dst = smem + (size_t)slot * RAW_STAGE_BYTES;
                tcf::mbar_expect_tx(&full[slot], RAW_STAGE_BYTES);
                tcf::tma_load_tile(dst, (const uint8_t*)P.ymaps + (size_t)s_csr.idx[2 * k] * 128, &full[slot], tile * TILE);
                tcf::tma_load_tile(dst + 8192, (const uint8_t*)P.ymaps + (size_t)s_csr.idx[2 * k + 1] * 128, &full[slot], tile * TILE);
            }
    } else if (warp == 5) {
        if (lane == 0 && n_items) {
            for (uint32_t n = 0; n < n_items; n++) {
                const uint32_t slot = n % RAW_STAGES, sb = tcf::smem_u32(smem + (size_t)slot * RAW_STAGE_BYTES);
                tcf::mbar_wait(&full[slot], (n / RAW_STAGES) & 1u);
                tcf::tc_fence_after();
                issue_contraction<2>(0, sb, sb + 8192, tmem, n ? 1u : 0u);
                tcf::umma_commit(&empty[slot]);
            }
            tcf::umma_commit(&done);
            tcf::mbar_wait(&done, 0);
        }
    }
    __syncwarp();
    tcf::tc_fence_before();
    __syncthreads();
    epilogue<2>(P, tmem, n_items > 0, s_E, &s_last, reinterpret_cast<uint32_t*>(smem));
    __syncthreads();
    if (warp == 0) tcf::tmem_dealloc(tmem, 64);
}

// ================================================================================================ fold rounds, degree 3
// Every table tile (128 rows x 128 bytes = old[4b..4b+3]) goes HBM -> shared memory by TMA into a ring shared by the groups
// (warp W_TMA); warp W_FOLD issues the fix_variables MMAs (tc_fold.cuh) into the owning group's accumulator (two per group,
// alternating); the group's thread t reads out new[2b], new[2b+1] of pair b = tile * 128 + t, stores them (the folded table)
// and, once it holds the pairs of the product's first two tables, writes the three plain products; the folded pair of the
// third table is the Y operand; warp W_SUM issues the contraction MMAs.  Each producer warp walks the same sequence of units
// (item n, multiplicand j, group g) — g innermost, so consecutive units belong to different groups — and never computes
// anything but a ring slot and a barrier phase (a producer that also chased the product list through global memory cost
// ~1700 cycles per unit and starved every group: measured).
template <int G, int MM>
struct FoldSmem {
    static constexpr uint32_t RING_SLOTS = 6;
    static constexpr uint32_t GROUPS = RING_SLOTS * tcf::TILE_BYTES;
    static constexpr uint32_t XA = 0, XQ = 16384, Y = 24576, GROUP = Y + Shape<MM>::Y_BYTES;
    static constexpr uint32_t BMAT = GROUPS + G * GROUP;
    static constexpr size_t BYTES = (size_t)BMAT + tcf::BMAT_BYTES;
};

// Tensor memory: 2 N columns of contraction accumulators, then NACC fix_variables accumulators of 64 columns per group — MM = 3: three
// groups x two (alternating), MM = 4: two groups x one (the 384 contraction columns leave room for two; a group's compute per table
// tile is several times the latency of its MMAs, so one accumulator per group does not starve it).  512 columns either way.
// PRE: the build for rounds launched ahead of their challenge.  A separate instantiation, because the mere presence of the polling loop
// in the prologue made ptxas schedule the main loop ~3 % slower (measured A/B on one box: round 2 of nv = 24 0.390 vs 0.403 ms) — the
// ordinary build must not carry it.
template <int G, int MM, bool PRE = false>
__global__ void __launch_bounds__(G * 128 + 96, 1) gemm_fold_kernel(const Params P) {
    using L = FoldSmem<G, MM>;
    using S_ = Shape<MM>;
    constexpr uint32_t R = L::RING_SLOTS;
    constexpr uint32_t W_TMA = G * 4, W_FOLD = G * 4 + 1, W_SUM = G * 4 + 2;
    constexpr uint32_t NACC = MM == 4 ? 1 : 2, ACC0 = 2 * S_::N;
    static_assert(ACC0 + G * NACC * 64 <= 512, "tensor memory");
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t slot_full[R], slot_empty[R], acc_full[G][2], acc_empty[G][2], x_full[G], x_empty[G];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint32_t s_E[2 * S_::NB * S_::ES];
    __shared__ bool s_last;
    __shared__ CsrCache s_csr;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const sck::RoundParams& p = P.rp;
    const long long t_start = P.prof ? clock64() : 0;
    load_csr<MM>(s_csr, p, true);
    uint8_t* const bmat = smem + L::BMAT;
    if (tid == 0) {
        if (tcf::smem_u32(smem) & 1023u) __trap();
        for (uint32_t s = 0; s < R; s++) {
            tcf::mbar_init(&slot_full[s], 1);
            tcf::mbar_init(&slot_empty[s], 1);
        }
        for (int g = 0; g < G; g++) {
            tcf::mbar_init(&acc_full[g][0], 1);
            tcf::mbar_init(&acc_full[g][1], 1);
            tcf::mbar_init(&acc_empty[g][0], 4);
            tcf::mbar_init(&acc_empty[g][1], 4);
            tcf::mbar_init(&x_full[g], 4);
            tcf::mbar_init(&x_empty[g], 1);
        }
        tcf::fence_mbar_init();
    }
    if (warp == 0) tcf::tmem_alloc(&s_tmem, 512);
    tcf::tc_fence_before();
    __syncthreads();  // barriers, tensor memory and the product list are ready: the TMA warp starts staging tiles right away ...
    tcf::tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (warp < 2) {   // ... while warps 0 and 1 expand the challenge into the constants matrix, which only the first
        Fr r;         // fix_variables MMA waits for (named barrier 1: these two warps arrive, the MMA-issuing warp syncs)
        if (PRE) {  // launched ahead of the challenge: wait for the host to send it
            __shared__ uint32_t s_r[8];
            if (tid < 8) {
                unsigned long long w;
                const long long t0 = clock64();
                const unsigned long long* src = blockIdx.x == 0 ? P.r_mail + tid : P.r_bcast + tid;
                for (;;) {
                    if (blockIdx.x == 0) asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
                    else asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
                    if ((uint32_t)(w >> 32) == P.r_seq) break;
                    if (clock64() - t0 > P.r_timeout) {
                        *P.r_error = 1;
                        w = (unsigned long long)P.r_seq << 32;  // (release the other CTAs too)
                        break;
                    }
                    if (blockIdx.x != 0) __nanosleep(200);
                }
                if (blockIdx.x == 0) asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(P.r_bcast + tid), "l"(w) : "memory");
                s_r[tid] = (uint32_t)w;
            }
            asm volatile("bar.sync 2, 64;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 8; i++) r.l[i] = s_r[i];
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) r.l[i] = p.r[i];
        }
        tcf::build_bmat(r, bmat);
        tcf::fence_proxy_async_smem();
        asm volatile("bar.arrive 1, 96;" ::: "memory");
    }
    const Split<G, MM> sp(P.items);
    const bool one_product = p.n_products == 1;
    const bool pf = P.prof != nullptr && lane == 0;
    prof_mark(P.prof, 12, t_start);  // prologue done
    if (warp < G * 4) {
        // ---------------------------------------------------------------------------------------------------- compute group
        const uint32_t g = warp >> 2, t = tid & 127u;
        uint8_t* const base = smem + L::GROUPS + (size_t)g * L::GROUP;
        const uint32_t lane_taddr = tmem + ACC0 + g * NACC * 64u + (((warp & 3u) * 32u) << 16);
        const uint32_t n_items = sp.n_items(g);
        ProfTimer t_acc, t_x, t_all;
        t_all.start(pf);
        for (uint32_t n = 0; n < n_items; n++) {
            const uint32_t w = sp.item(n, g);
            const uint32_t k = one_product ? 0u : w % p.n_products, tile = p.tile_base + (one_product ? w : w / p.n_products);
            const unsigned long long b = (unsigned long long)tile * TILE + t;
            Fr e0, o0;
#pragma unroll
            for (uint32_t j = 0; j < (uint32_t)MM; j++) {
                const uint32_t q = MM * n + j, a = q % NACC;
                t_acc.start(pf);
                tcf::mbar_wait(&acc_full[g][a], (q / NACC) & 1u);
                t_acc.stop(pf);
                // the fold MMAs of this unit have completed: its ring slot is free again
                if ((warp & 3u) == 0 && lane == 0) tcf::mbar_arrive(&slot_empty[sp.unit(n, j, g) % R]);
                tcf::tc_fence_after();
                uint32_t S[32];
                tcf::tmem_ld32(lane_taddr + a * 64u, S);
                tcf::tmem_ld_wait();
                const Fr v0 = tcf::columns_to_fr(S);
                tcf::tmem_ld32(lane_taddr + a * 64u + 32u, S);
                tcf::tmem_ld_wait();
                tcf::tc_fence_before();
                __syncwarp();
                if (lane == 0) tcf::mbar_arrive(&acc_empty[g][a]);
                const Fr v1 = tcf::columns_to_fr(S);
                if (s_csr.first[MM * k + j]) {
                    uint32_t* dst = s_csr.out[MM * k + j] + b * 16;
                    fr::store(dst, v0);
                    fr::store(dst + 8, v1);
                }
                if (MM == 2) {  // no products: the folded pair of table 0 is the X operand, that of table 1 the Y operand
                    uint32_t y[16];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        y[i] = v0.l[i];
                        y[8 + i] = v1.l[i];
                    }
                    if (j == 0) {
                        t_x.start(pf);
                        if (n > 0) tcf::mbar_wait(&x_empty[g], (n - 1) & 1u);
                        t_x.stop(pf);
                        sts_sw64(base + L::XQ, t, y);
                    } else {
                        sts_sw64(base + L::Y, t, y);
                    }
                } else if (j == 0 || (MM == 4 && j == 2)) {
                    e0 = v0;
                    o0 = v1;
                } else if (j == 1) {
                    t_x.start(pf);
                    if (n > 0) tcf::mbar_wait(&x_empty[g], (n - 1) & 1u);  // the contraction of item n-1 has read X and Y
                    t_x.stop(pf);
                    products_to_smem(e0, o0, v0, v1, base + L::XA, base + L::XQ, t);
                } else if (MM == 4) {
                    y_products_to_smem(e0, o0, v0, v1, base + L::Y, t);
                } else {
                    uint32_t y[16];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        y[i] = v0.l[i];
                        y[8 + i] = v1.l[i];
                    }
                    sts_sw64(base + L::Y, t, y);
                }
            }
            tcf::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) tcf::mbar_arrive(&x_full[g]);
        }
        t_all.stop(pf);
        if (pf) { t_acc.flush(P.prof, 0); t_x.flush(P.prof, 1); t_all.flush(P.prof, 2); }
    } else if (warp == W_TMA) {
        // ---------------------------------------------------------------------------------------------------- TMA: table tiles into the ring
        if (lane == 0) {
            ProfTimer t_se, t_all;
            t_all.start(pf);
            uint32_t u = 0;
            const uint32_t steps = sp.steps();
            for (uint32_t n = 0; n < steps; n++) {
                const uint32_t ng = sp.groups(n);
#pragma unroll
                for (uint32_t j = 0; j < (uint32_t)MM; j++)
#pragma unroll
                    for (uint32_t g = 0; g < (uint32_t)G; g++) {
                        if (g >= ng) continue;
                        const uint32_t slot = u % R;
                        t_se.start(pf);
                        if (u >= R) tcf::mbar_wait(&slot_empty[slot], ((u / R) - 1u) & 1u);
                        t_se.stop(pf);
                        const uint32_t w = sp.item(n, g);
                        const uint32_t k = one_product ? 0u : w % p.n_products, tile = p.tile_base + (one_product ? w : w / p.n_products);
                        tcf::mbar_expect_tx(&slot_full[slot], tcf::TILE_BYTES);
                        tcf::tma_load_tile(smem + (size_t)slot * tcf::TILE_BYTES, (const uint8_t*)p.tmaps + (size_t)s_csr.idx[MM * k + j] * 128, &slot_full[slot],
                                           tile * TILE);
                        u++;
                    }
            }
            t_all.stop(pf);
            if (pf) { t_se.flush(P.prof, 5); t_all.flush(P.prof, 11); }
        }
    } else if (warp == W_FOLD) {
        // ---------------------------------------------------------------------------------------------------- fix_variables MMAs
        asm volatile("bar.sync 1, 96;" ::: "memory");  // the constants matrix is in shared memory (written through the generic proxy, fenced)
        tcf::tc_fence_after();
        if (lane == 0) {
            ProfTimer t_sf, t_ae, t_all, t_mma;
            t_all.start(pf);
            const uint32_t ring = tcf::smem_u32(smem), bmat_smem = tcf::smem_u32(bmat);
            uint32_t u = 0;
            const uint32_t steps = sp.steps();
            for (uint32_t n = 0; n < steps; n++) {
                const uint32_t ng = sp.groups(n);
#pragma unroll
                for (uint32_t j = 0; j < (uint32_t)MM; j++)
#pragma unroll
                    for (uint32_t g = 0; g < (uint32_t)G; g++) {
                        if (g >= ng) continue;
                        const uint32_t slot = u % R, q = MM * n + j, a = q % NACC;
                        t_sf.start(pf);
                        tcf::mbar_wait(&slot_full[slot], (u / R) & 1u);
                        t_sf.stop(pf);
                        t_ae.start(pf);
                        if (q >= NACC) tcf::mbar_wait(&acc_empty[g][a], ((q / NACC) - 1u) & 1u);
                        t_ae.stop(pf);
                        tcf::tc_fence_after();
                        t_mma.start(pf);
                        tcf::issue_fold_mma(ring + slot * tcf::TILE_BYTES, bmat_smem, tmem + ACC0 + (g * NACC + a) * 64u);
                        tcf::umma_commit(&acc_full[g][a]);
                        t_mma.stop(pf);
                        u++;
                    }
            }
            t_all.stop(pf);
            if (pf) { t_sf.flush(P.prof, 3); t_ae.flush(P.prof, 4); t_all.flush(P.prof, 6); t_mma.flush(P.prof, 8); }
        }
    } else {
        // ---------------------------------------------------------------------------------------------------- contraction MMAs
        if (lane == 0) {
            ProfTimer t_xf, t_iss;
            uint32_t acc = 0;
            const uint32_t steps = sp.steps();
            for (uint32_t n = 0; n < steps; n++) {
                const uint32_t ng = sp.groups(n);
#pragma unroll
                for (uint32_t g = 0; g < (uint32_t)G; g++) {
                    if (g >= ng) continue;
                    const uint32_t gb = tcf::smem_u32(smem + L::GROUPS + (size_t)g * L::GROUP);
                    t_xf.start(pf);
                    tcf::mbar_wait(&x_full[g], n & 1u);
                    t_xf.stop(pf);
                    tcf::tc_fence_after();
                    t_iss.start(pf);
                    issue_contraction<MM>(gb + L::XA, gb + L::XQ, gb + L::Y, tmem, acc);
                    acc = 1;
                    tcf::umma_commit(&x_empty[g]);
                    t_iss.stop(pf);
                }
            }
            // every MMA has completed once the last commit of every group has arrived
#pragma unroll
            for (uint32_t g = 0; g < (uint32_t)G; g++)
                if (sp.n_items(g)) tcf::mbar_wait(&x_empty[g], (sp.n_items(g) - 1) & 1u);
            if (pf) { t_xf.flush(P.prof, 7); t_iss.flush(P.prof, 10); }
        }
    }
    __syncwarp();
    tcf::tc_fence_before();
    __syncthreads();
    epilogue<MM>(P, tmem, sp.n_items(0) > 0, s_E, &s_last, reinterpret_cast<uint32_t*>(smem), t_start);
    __syncthreads();
    if (warp == 0) tcf::tmem_dealloc(tmem, 512);
}

// ---- launchers (gemm.cu) ------------------------------------------------------------------------------------------------------------
cudaError_t init_constants();
// mm = multiplicands per product (3 or 4)
unsigned long long max_items_round1(int sms, int mm);  // work items one launch may carry (s32 accumulator head-room)
unsigned long long max_items_fold(int sms, int mm);
cudaError_t launch_round1(const Params& P, int mm, int sms, cudaStream_t stream);
cudaError_t launch_fold(const Params& P, int mm, int sms, cudaStream_t stream);

}  // namespace gsum
