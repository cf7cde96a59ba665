// fix_variables on the 5th-generation tensor cores (tcgen05, sm_100a).
//
// DenseMultilinearExtension::fix_variables(&[r]) (what prover.rs:85-89 calls once per round) is
//     new[b] = old[2b] + r*(old[2b+1] - old[2b]) = (1-r)*old[2b] + r*old[2b+1]          (mod p)
// with ONE r for the whole round, i.e. a linear map applied to every 64-byte pair of the table.  Written over the 64
// little-endian BYTES of the pair (x_0..x_31 = old[2b], x_32..x_63 = old[2b+1], Montgomery form as stored)
//     new[b] == sum_k x_k * C_k  (mod p),   C_k = (1-r)*2^(8k) mod p  (k < 32),   C_k = r*2^(8(k-32)) mod p  (k >= 32)
// and, over the 32 bytes of every constant,  sum_k x_k*C_k = sum_j 2^(8j) * S_j  with  S_j = sum_k x_k * C_k[j] < 2^22.
// S = X * Cmat is a u8 x u8 -> s32 matrix product with a (64 x 32) right-hand side shared by every row of the round:
// exactly what tcgen05.mma.kind::i8 computes.  So the table tile that TMA lands in shared memory (128-byte rows =
// old[4b..4b+3], SWIZZLE_128B) is used AS IS as the A operand: two K=64 products per 128-row tile give the column
// sums of new[2b] and new[2b+1] in tensor memory; a thread then carries its row's 32 columns into limbs and reduces
// mod p (columns_to_fr: ~90 ALU instructions + 8 IMAD.WIDE) instead of running 2 x 76 IMAD.WIDE for the two folds.
// Field arithmetic is exact and the result is fully reduced, so the limbs are the ones the reference produces.
#pragma once
#include <cstdint>

#include "fr.cuh"

namespace tcf {

constexpr uint32_t TILE_ROWS = 128;              // output pairs per tile = UMMA M = TMEM lanes = threads per CTA
constexpr uint32_t ROW_BYTES = 128;              // old[4b..4b+3]
constexpr uint32_t TILE_BYTES = TILE_ROWS * ROW_BYTES;  // 16 KiB per table per tile
constexpr uint32_t BMAT_BYTES = 32 * 128;        // constants matrix: N = 32 rows x 128-byte pitch (K = 64 bytes used)
constexpr uint32_t ACC_COLS = 64;                // TMEM columns per table tile: new[2b] (32) | new[2b+1] (32)
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 (2) @4, a/b format U8 (0) @7/@10, K-major A and B,
// n_dim = N>>3 = 4 @17, m_dim = M>>4 = 8 @24
constexpr uint32_t IDESC_U8_M128_N32 = (2u << 4) | (4u << 17) | (8u << 24);
constexpr uint32_t MU_270 = 0x8d54u;             // floor(2^270 / p)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a lost arrival (a bug, never expected) traps the kernel instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s
    }
}

// ---- TMA: one 128-row x 128-byte box of a table into shared memory (SWIZZLE_128B), completion on an mbarrier -------
__device__ __forceinline__ void tma_load_tile(void* smem_dst, const void* tmap, uint64_t* bar, uint32_t row0) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(0), "r"(row0)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// generic-proxy writes to shared memory (the constants matrix) must be fenced before the tensor core reads them
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tensor memory ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp (the allocating one)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 @0, leading
// byte offset (unused for swizzled K-major, 1) @16, stride byte offset = 1024 B between 8-row groups @32, version 1
// @46, layout type 2 (SWIZZLE_128B) @61.  The tile base must be 1024-byte aligned; a K step of 32 bytes inside the
// 128-byte swizzle atom is taken by adding (32 >> 4) to the start-address field.
__device__ __forceinline__ uint64_t sw128_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

// D[tmem] (+)= A[smem] * B[smem]^T, u8 x u8 -> s32, M = 128, N = 32, K = 32; issued by ONE thread
__device__ __forceinline__ void umma_u8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC_U8_M128_N32), "r"(accumulate)
        : "memory");
}
// arrive on `bar` when every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// One table tile: new[2b] columns <- bytes 0..63 of every row, new[2b+1] columns <- bytes 64..127; same constants.
__device__ __forceinline__ void issue_fold_mma(uint32_t tile_smem, uint32_t bmat_smem, uint32_t tmem_acc) {
    const uint64_t a = sw128_desc(tile_smem), b = sw128_desc(bmat_smem);
    umma_u8(tmem_acc, a, b, 0);                    // e0 bytes  x C[0..31]
    umma_u8(tmem_acc, a + 2, b + 2, 1);            // e1 bytes  x C[32..63]
    umma_u8(tmem_acc + 32, a + 4, b, 0);           // e2 bytes  x C[0..31]
    umma_u8(tmem_acc + 32, a + 6, b + 2, 1);       // e3 bytes  x C[32..63]
}

// this thread's TMEM lane (warp w of the CTA owns lanes 32w..32w+31), 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// this thread's TMEM lane, 16 consecutive 32-bit columns (whole warp, convergent)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                 "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
}

// ---- the constants matrix -----------------------------------------------------------------------------------------
// Byte (n, k) of the B operand (N = 32 rows, K-major, SWIZZLE_128B, 128-byte pitch): B[n][k] = byte n of C_k.
__device__ __forceinline__ uint32_t bmat_offset(uint32_t n, uint32_t k) {
    return n * 128u + ((((k >> 4) ^ (n & 7u)) << 4) | (k & 15u));
}

// 2^(8k) mod p as a raw integer, k = 0..31 (all below 2^248 < p): mul(x_mont, pow256(k)) = x * 2^(8k) mod p, canonical.
__device__ __forceinline__ fr::Fr pow256(uint32_t k) {
    fr::Fr x;
#pragma unroll
    for (uint32_t i = 0; i < 8; i++) x.l[i] = (i == (k >> 2)) ? (1u << (8 * (k & 3))) : 0u;
    return x;
}

// Threads 0..63 of the CTA: thread k writes the 32 bytes of C_k into the swizzled matrix.  r_mont = the challenge.
__device__ __forceinline__ void build_bmat(const fr::Fr& r_mont, uint8_t* bmat) {
    const uint32_t k = threadIdx.x;
    if (k < 64) {
        const fr::Fr base = (k < 32) ? fr::sub(fr::one(), r_mont) : r_mont;
        const fr::Fr c = fr::mul(base, pow256(k & 31));
#pragma unroll
        for (uint32_t n = 0; n < 32; n++) bmat[bmat_offset(n, k)] = (uint8_t)(c.l[n >> 2] >> (8 * (n & 3)));
    }
}

// ---- column sums -> field element ----------------------------------------------------------------------------------
// V = sum_j S[j] * 2^(8j), S[j] < 2^22, V < 64*255*p < 2^269.  Returns V mod p, fully reduced.
// Barrett with a 16-bit reciprocal: q = floor(floor(V / 2^238) * floor(2^270 / p) / 2^32) satisfies
// floor(V/p) - 1 <= q <= floor(V/p) (x*mu/2^32 > V/p - V/2^270 - 2^238/p > V/p - 0.46), so V - q*p < 2p: one
// conditional subtraction finishes.
__device__ __forceinline__ fr::Fr columns_to_fr(const uint32_t (&S)[32]) {
    uint32_t l[9];
    uint32_t hi = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t t0 = S[4 * i] + (S[4 * i + 1] << 8);      // < 2^31
        const uint32_t t1 = S[4 * i + 2] + (S[4 * i + 3] << 8);  // < 2^31
        // (measured: forcing this step onto the ALU pipe with PRMT byte shifts + carry-chained IADD3 removes 16 IMAD, 8
        // IMAD.WIDE, 8 IMAD.X and ~20 IMAD.MOV per call but adds as many ALU instructions — 2 % slower overall)
        const uint64_t w = (uint64_t)t0 + ((uint64_t)t1 << 16) + hi;
        l[i] = (uint32_t)w;
        hi = (uint32_t)(w >> 32);
    }
    l[8] = hi;  // < 2^14
    const uint32_t x = (l[7] >> 14) | (l[8] << 18);  // floor(V / 2^238) < 2^32
    const uint32_t q = __umulhi(x, MU_270);
    // t = q * p (q < 2^15: 8 limbs and a small top), V -= t
    uint32_t t[9];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)q * fr::FR_MODULUS[i];
        t[i] = (uint32_t)c;
        c >>= 32;
    }
    t[8] = (uint32_t)c;
    fr::Fr v;
    asm("sub.cc.u32 %0, %8, %17;\n\t"
        "subc.cc.u32 %1, %9, %18;\n\t"
        "subc.cc.u32 %2, %10, %19;\n\t"
        "subc.cc.u32 %3, %11, %20;\n\t"
        "subc.cc.u32 %4, %12, %21;\n\t"
        "subc.cc.u32 %5, %13, %22;\n\t"
        "subc.cc.u32 %6, %14, %23;\n\t"
        "subc.u32 %7, %15, %24;\n\t"
        : "=r"(v.l[0]), "=r"(v.l[1]), "=r"(v.l[2]), "=r"(v.l[3]), "=r"(v.l[4]), "=r"(v.l[5]), "=r"(v.l[6]), "=r"(v.l[7])
        : "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]), "r"(l[8]), "r"(t[0]), "r"(t[1]),
          "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]));
    // V - q*p < 2p < 2^256: limb 8 of the difference is zero, so l[8] and t[8] need not be subtracted
    (void)t[8];
    return fr::reduce_once(v);
}

}  // namespace tcf
