// BLS12-381 scalar field Fr on sm_100a: 8 x u32 limbs, Montgomery form R = 2^256, values always fully reduced —
// bit-identical to ark-ff's Fp<MontBackend<FrConfig,4>,4> in memory (4 x u64 LE), which is what the reference's
// tables hold (ml_sumcheck/protocol/prover.rs:26 `flattened_ml_extensions`).  Field arithmetic is exact and the
// representation canonical, so any evaluation order reproduces the reference's limbs.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fr_asm.cuh"

namespace fr {

struct Fr {
    uint32_t l[8];
};

// p, little-endian 32-bit limbs.  p[0] = 1 and -p^-1 mod 2^32 = 0xffffffff.
#define FR_P0 0x00000001u
#define FR_P1 0xffffffffu
#define FR_P2 0xfffe5bfeu
#define FR_P3 0x53bda402u
#define FR_P4 0x09a1d805u
#define FR_P5 0x3339d808u
#define FR_P6 0x299d7d48u
#define FR_P7 0x73eda753u

__device__ __forceinline__ Fr zero() {
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = 0;
    return r;
}
// R mod p  (Montgomery form of 1)
__device__ __forceinline__ Fr one() {
    Fr r = {{0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u}};
    return r;
}

// 256-bit vector accesses (LDG.E.256 / STG.E.256 on sm_100): one instruction per element.
__device__ __forceinline__ Fr load(const uint32_t* p) {
    Fr r;
    asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
                 : "l"(p));
    return r;
}
// streaming read: read-only path, do not keep in L1 (every table element is consumed exactly once per round)
__device__ __forceinline__ Fr load_stream(const uint32_t* p) {
    Fr r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void store(uint32_t* p, const Fr& v) {
    asm volatile("st.global.v8.u32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"r"(v.l[0]), "r"(v.l[1]), "r"(v.l[2]), "r"(v.l[3]),
                 "r"(v.l[4]), "r"(v.l[5]), "r"(v.l[6]), "r"(v.l[7]), "l"(p)
                 : "memory");
}

__device__ __forceinline__ bool is_zero(const Fr& a) {
    return (a.l[0] | a.l[1] | a.l[2] | a.l[3] | a.l[4] | a.l[5] | a.l[6] | a.l[7]) == 0;
}
__device__ __forceinline__ bool eq(const Fr& a, const Fr& b) {
    uint32_t d = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) d |= a.l[i] ^ b.l[i];
    return d == 0;
}

// r = a - p if a >= p else a      (a < 2p)
__device__ __forceinline__ Fr reduce_once(const Fr& a) {
    Fr s;
    uint32_t borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;\n\t"
        : "=r"(s.l[0]), "=r"(s.l[1]), "=r"(s.l[2]), "=r"(s.l[3]), "=r"(s.l[4]), "=r"(s.l[5]), "=r"(s.l[6]), "=r"(s.l[7]),
          "=r"(borrow)
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]), "r"(FR_P0),
          "r"(FR_P1), "r"(FR_P2), "r"(FR_P3), "r"(FR_P4), "r"(FR_P5), "r"(FR_P6), "r"(FR_P7));
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = borrow ? a.l[i] : s.l[i];
    return r;
}

__device__ __forceinline__ Fr add(const Fr& a, const Fr& b) {
    Fr t;  // a + b < 2p < 2^256: no carry out
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;\n\t"
        : "=r"(t.l[0]), "=r"(t.l[1]), "=r"(t.l[2]), "=r"(t.l[3]), "=r"(t.l[4]), "=r"(t.l[5]), "=r"(t.l[6]), "=r"(t.l[7])
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]), "r"(b.l[0]),
          "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
    return reduce_once(t);
}

__device__ __forceinline__ Fr sub(const Fr& a, const Fr& b) {
    Fr t;
    uint32_t borrow;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;\n\t"
        : "=r"(t.l[0]), "=r"(t.l[1]), "=r"(t.l[2]), "=r"(t.l[3]), "=r"(t.l[4]), "=r"(t.l[5]), "=r"(t.l[6]), "=r"(t.l[7]),
          "=r"(borrow)
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]), "r"(b.l[0]),
          "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
    // borrow = 0xffffffff when a < b: add p back (masked)
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;\n\t"
        : "+r"(t.l[0]), "+r"(t.l[1]), "+r"(t.l[2]), "+r"(t.l[3]), "+r"(t.l[4]), "+r"(t.l[5]), "+r"(t.l[6]), "+r"(t.l[7])
        : "r"(FR_P0 & borrow), "r"(FR_P1 & borrow), "r"(FR_P2 & borrow), "r"(FR_P3 & borrow), "r"(FR_P4 & borrow),
          "r"(FR_P5 & borrow), "r"(FR_P6 & borrow), "r"(FR_P7 & borrow));
    return t;
}

// Montgomery product a*b*2^-256 mod p WITHOUT the final conditional subtraction: 63 IMAD.WIDE for the product + 48 for the
// reduction.  For a < 2p and b < p the result is < p*(2p/2^256 + 1) < 1.91p < 2p: a value that is only multiplied again
// (by a canonical operand) or fed to the lazy accumulator may stay in [0, 2p).
__device__ __forceinline__ Fr mul_lazy_impl(const Fr& a, const Fr& b) {
    uint32_t ev[16], od[16];
    mul_wide_eo(ev, od, a.l, b.l);
    uint32_t c = redc_eo(ev, od);
    // quotient = sum_{k=8..15} (ev[k] + od[k-1]) 2^(32(k-8)) + c   (< 2p, fits 8 limbs)
    Fr t;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;\n\t"
        : "=r"(t.l[0]), "=r"(t.l[1]), "=r"(t.l[2]), "=r"(t.l[3]), "=r"(t.l[4]), "=r"(t.l[5]), "=r"(t.l[6]), "=r"(t.l[7])
        : "r"(ev[8]), "r"(ev[9]), "r"(ev[10]), "r"(ev[11]), "r"(ev[12]), "r"(ev[13]), "r"(ev[14]), "r"(ev[15]), "r"(od[7]),
          "r"(od[8]), "r"(od[9]), "r"(od[10]), "r"(od[11]), "r"(od[12]), "r"(od[13]), "r"(od[14]));
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.u32 %7, %7, 0;\n\t"
        : "+r"(t.l[0]), "+r"(t.l[1]), "+r"(t.l[2]), "+r"(t.l[3]), "+r"(t.l[4]), "+r"(t.l[5]), "+r"(t.l[6]), "+r"(t.l[7])
        : "r"(c));
    return t;
}
__device__ __forceinline__ Fr mul_impl(const Fr& a, const Fr& b) { return reduce_once(mul_lazy_impl(a, b)); }


// FR_COMPACT (latency-bound kernels such as the fused tail): ONE out-of-line copy of the multiplier, arguments by value
// (registers), so the whole round loop stays resident in the instruction cache instead of streaming ~100 KB of inlined
// straight-line code per round.  Throughput kernels keep the inlined form.
#ifdef FR_COMPACT
static __device__ __noinline__ Fr mul_outlined(Fr a, Fr b) { return mul_impl(a, b); }
__device__ __forceinline__ Fr mul(const Fr& a, const Fr& b) { return mul_outlined(a, b); }
static __device__ __noinline__ Fr mul_lazy_outlined(Fr a, Fr b) { return mul_lazy_impl(a, b); }
__device__ __forceinline__ Fr mul_lazy(const Fr& a, const Fr& b) { return mul_lazy_outlined(a, b); }
#else
__device__ __forceinline__ Fr mul(const Fr& a, const Fr& b) { return mul_impl(a, b); }
__device__ __forceinline__ Fr mul_lazy(const Fr& a, const Fr& b) { return mul_lazy_impl(a, b); }
#endif

// ---- multiplication by the round's challenge (the fix_variables fold) --------------------------------------------
// Every fold of a round multiplies by the SAME r, so r is expanded once per round into eight plain-integer constants
// C[k] = r * 2^(32k+64) mod p (k = 0..7).  For a Montgomery-form d,  V = sum_k d.l[k] * C[k]  is 64 wide MACs whose rows
// all land on limb 0, V == r*d*2^64 (mod p), V < 2^290; a 2-digit Montgomery reduction (12 wide MACs) then gives
// r*d mod p < 2p.  76 IMAD.WIDE instead of 111 for the general product — and r no longer occupies 8 registers.
__device__ __forceinline__ Fr fold_const(const Fr& r_mont, int k) {
    // X_k = 2^(32k+64) mod p as a raw integer; mul(r_mont, X_k) = r * X_k mod p (r = the challenge's canonical value)
    Fr x = zero();
    if (k < 6) {
        x.l[k + 2] = 1;
    } else if (k == 6) {
        x = one();  // 2^256 mod p
    } else {
        Fr x7 = {{0xcaaf6b13u, 0x355094eau, 0x69a568efu, 0xf6b10cb3u, 0x40cc3869u, 0xe2c926a6u, 0xed269aadu, 0x736a6d3bu}};  // 2^288 mod p
        x = x7;
    }
    return mul(r_mont, x);
}

// r * d (Montgomery form in, Montgomery form out), C = the 8 x 8 limbs written by fold_const (shared memory)
__device__ __forceinline__ Fr mul_round_const(const uint32_t* C, const Fr& d) {
    uint32_t ev[10], od[10];
    mulc_rows_eo(ev, od, C, d.l);
    uint32_t c = redc2_eo10(ev, od);
    // quotient limbs j = 0..7: ev[2+j] + od[1+j] (+ c at j = 0); it is < 2p, so od[9] and the carry out are zero
    Fr t;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;\n\t"
        : "=r"(t.l[0]), "=r"(t.l[1]), "=r"(t.l[2]), "=r"(t.l[3]), "=r"(t.l[4]), "=r"(t.l[5]), "=r"(t.l[6]), "=r"(t.l[7])
        : "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]), "r"(ev[8]), "r"(ev[9]), "r"(od[1]),
          "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]), "r"(od[7]), "r"(od[8]));
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.u32 %7, %7, 0;\n\t"
        : "+r"(t.l[0]), "+r"(t.l[1]), "+r"(t.l[2]), "+r"(t.l[3]), "+r"(t.l[4]), "+r"(t.l[5]), "+r"(t.l[6]), "+r"(t.l[7])
        : "r"(c));
    return reduce_once(t);
}
#ifdef FR_COMPACT
static __device__ __noinline__ Fr mul_round_const_outlined(const uint32_t* C, Fr d) { return mul_round_const(C, d); }
#define FR_MUL_ROUND_CONST(C, d) fr::mul_round_const_outlined(C, d)
#else
#define FR_MUL_ROUND_CONST(C, d) fr::mul_round_const(C, d)
#endif

// ILP forms for latency-bound kernels (resident rounds): ONE out-of-line routine works on several independent operands, so
// ptxas interleaves their carry chains (a lone warp on a scheduler otherwise waits ~4 cycles between dependent IMAD.WIDE).
// fold2: the two folds of one table row; mul_lazy_n: the N evaluation points of one multiplicand.
struct Fr2 {
    Fr a, b;
};
#ifdef FR_COMPACT
static __device__ __noinline__ Fr2 mul_round_const_x2(const uint32_t* C, Fr d0, Fr d1) {
    Fr2 r;
    r.a = mul_round_const(C, d0);
    r.b = mul_round_const(C, d1);
    return r;
}
template <int N>
struct FrN {
    Fr v[N];
};
template <int N>
static __device__ __noinline__ FrN<N> mul_lazy_n(FrN<N> a, FrN<N> b) {
    FrN<N> r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = mul_lazy_impl(a.v[i], b.v[i]);
    return r;
}
#else
__device__ __forceinline__ Fr2 mul_round_const_x2(const uint32_t* C, const Fr& d0, const Fr& d1) {
    Fr2 r;
    r.a = mul_round_const(C, d0);
    r.b = mul_round_const(C, d1);
    return r;
}
template <int N>
struct FrN {
    Fr v[N];
};
template <int N>
__device__ __forceinline__ FrN<N> mul_lazy_n(const FrN<N>& a, const FrN<N>& b) {
    FrN<N> r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = mul_lazy_impl(a.v[i], b.v[i]);
    return r;
}
#endif

// ---- lazily reduced inner products -----------------------------------------------------------------------------
// A WideAcc holds an UNREDUCED integer sum of Montgomery products x*y (each < p^2 < 2^510) in 17 limbs, so 2^34
// products can be accumulated before overflow.  One Montgomery reduction at the end turns the whole sum into a field
// element: sum_b x_b*y_b*R^-1 mod p — the same value as reducing every product first, because the field arithmetic is
// exact.  This halves the IMAD.WIDE cost of the last multiply of every product term (64 instead of 112).
struct WideAcc {
    uint32_t l[17];
};

__device__ __forceinline__ void wide_zero(WideAcc& w) {
#pragma unroll
    for (int i = 0; i < 17; i++) w.l[i] = 0;
}

// w += a * b   (integer product, 16 limbs): 64 IMAD.WIDE + 33 carry-chained adds
#if defined(FR_COMPACT) && !defined(FR_INLINE_WIDE_MAC)
static __device__ __noinline__ WideAcc wide_mac_outlined(WideAcc w, Fr a, Fr b) {
    wide_mac_limbs(w.l, a.l, b.l);
    return w;
}
__device__ __forceinline__ void wide_mac(WideAcc& w, const Fr& a, const Fr& b) { w = wide_mac_outlined(w, a, b); }
#else
__device__ __forceinline__ void wide_mac(WideAcc& w, const Fr& a, const Fr& b) { wide_mac_limbs(w.l, a.l, b.l); }
#endif

// w += a * 2^256: after the final reduction this contributes exactly `a` (used for single-multiplicand products)
__device__ __forceinline__ void wide_add_shifted(WideAcc& w, const Fr& a) {
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %12;\n\t"
        "addc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\t"
        "addc.cc.u32 %6, %6, %15;\n\t"
        "addc.cc.u32 %7, %7, %16;\n\t"
        "addc.u32 %8, %8, 0;\n\t"
        : "+r"(w.l[8]), "+r"(w.l[9]), "+r"(w.l[10]), "+r"(w.l[11]), "+r"(w.l[12]), "+r"(w.l[13]), "+r"(w.l[14]),
          "+r"(w.l[15]), "+r"(w.l[16])
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]));
}

__device__ __forceinline__ Fr R2() {  // R^2 mod p: mul(x, R2()) = x * R mod p for any x < 2^256
    Fr r = {{0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u}};
    return r;
}

// Montgomery-reduce a WideAcc (value V < 2^544) to the canonical field element V * R^-1 mod p.
#ifdef FR_COMPACT
static __device__ __noinline__ Fr wide_reduce(WideAcc w) {
#else
__device__ __forceinline__ Fr wide_reduce(const WideAcc& w) {
#endif
    uint32_t ev[17], od[17];
#pragma unroll
    for (int i = 0; i < 17; i++) { ev[i] = w.l[i]; od[i] = 0; }
    uint32_t c = redc_eo17(ev, od);
    // quotient q = sum_{k=8..16} (ev[k] + od[k-1]) 2^(32(k-8)) + od[16] 2^288 + c  < 2^289: 10 limbs
    uint32_t q[10];
    asm("add.cc.u32 %0, %10, %19;\n\t"
        "addc.cc.u32 %1, %11, %20;\n\t"
        "addc.cc.u32 %2, %12, %21;\n\t"
        "addc.cc.u32 %3, %13, %22;\n\t"
        "addc.cc.u32 %4, %14, %23;\n\t"
        "addc.cc.u32 %5, %15, %24;\n\t"
        "addc.cc.u32 %6, %16, %25;\n\t"
        "addc.cc.u32 %7, %17, %26;\n\t"
        "addc.cc.u32 %8, %18, %27;\n\t"
        "addc.u32 %9, %28, 0;\n\t"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9])
        : "r"(ev[8]), "r"(ev[9]), "r"(ev[10]), "r"(ev[11]), "r"(ev[12]), "r"(ev[13]), "r"(ev[14]), "r"(ev[15]), "r"(ev[16]),
          "r"(od[7]), "r"(od[8]), "r"(od[9]), "r"(od[10]), "r"(od[11]), "r"(od[12]), "r"(od[13]), "r"(od[14]), "r"(od[15]),
          "r"(od[16]));
    asm("add.cc.u32 %0, %0, %10;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.cc.u32 %8, %8, 0;\n\t"
        "addc.u32 %9, %9, 0;\n\t"
        : "+r"(q[0]), "+r"(q[1]), "+r"(q[2]), "+r"(q[3]), "+r"(q[4]), "+r"(q[5]), "+r"(q[6]), "+r"(q[7]), "+r"(q[8]), "+r"(q[9])
        : "r"(c));
    // q = lo + hi * 2^256 with lo < 2^256 < 3p and hi < 2^64:  q mod p = (lo mod p) + hi*R mod p
    Fr lo, hi = zero();
#pragma unroll
    for (int i = 0; i < 8; i++) lo.l[i] = q[i];
    hi.l[0] = q[8];
    hi.l[1] = q[9];
    lo = reduce_once(reduce_once(lo));
    // the two top limbs are zero unless several products were accumulated (each is < 2^510): small rounds skip a multiply
    if ((q[8] | q[9]) == 0) return lo;
    return add(lo, mul(hi, R2()));
}

// Reference formulation in plain C (32-bit limbs, 64-bit accumulators); used by the micro-benchmark as a baseline.
__device__ __forceinline__ Fr mul_c64(const Fr& a, const Fr& b) {
    const uint32_t P[8] = {FR_P0, FR_P1, FR_P2, FR_P3, FR_P4, FR_P5, FR_P6, FR_P7};
    uint32_t t[10];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)a.l[j] * b.l[i] + t[j];
            t[j] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[8] = (uint32_t)c;
        t[9] = (uint32_t)(c >> 32);
        uint32_t m = 0u - t[0];
        c = (uint64_t)m * P[0] + t[0];
        c >>= 32;
#pragma unroll
        for (int j = 1; j < 8; j++) {
            c += (uint64_t)m * P[j] + t[j];
            t[j - 1] = (uint32_t)c;
            c >>= 32;
        }
        c += t[8];
        t[7] = (uint32_t)c;
        t[8] = t[9] + (uint32_t)(c >> 32);
    }
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = t[i];
    return reduce_once(r);
}

}  // namespace fr
