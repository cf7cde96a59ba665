// Feasibility check for the NEXT step named in DESIGN.md §10: the Montgomery REDUCTION on the tensor cores.
//   a*b*R^-1 mod p == sum_k T_k * K_k (mod p),  T = a*b (512 bits, bytes T_0..T_63),
//   K_k = 2^(8k) * R^-1 mod p for k < 32 and K_k = 2^(8(k-32)) for k >= 32 (the high half of T needs no constants).
// Each thread forms the plain 512-bit product (64 IMAD.WIDE), writes its 64 bytes into a SWIZZLE_128B shared-memory
// tile with generic stores, one thread issues two K = 32 tcgen05.mma.kind::i8 against the constants matrix, and every
// thread carries its 32 column sums out with the same tcf::columns_to_fr the fold uses (V < 64*255*p as well).
// Compared against fr::mul on the same inputs.  Not used by the library; round-2 material.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I sumcheck_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc_fold.cuh"

using fr::Fr;

__global__ void __launch_bounds__(128) redc_tc_kernel(const uint32_t* a_in, const uint32_t* b_in, uint32_t* out_tc, uint32_t* out_ref,
                                                      long long* clk) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* tile = smem;                    // 128 rows x 128 B (bytes 0..63 of a row = this thread's T)
    uint8_t* kmat = smem + tcf::TILE_BYTES;  // N = 32 rows x 128-byte pitch, K = 64 bytes used
    __shared__ __align__(8) uint64_t bar_mma;
    __shared__ uint32_t tmem_slot;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        tcf::mbar_init(&bar_mma, 1);
        tcf::fence_mbar_init();
    }
    if (warp == 0) tcf::tmem_alloc(&tmem_slot, 32);
    // constants: thread k < 64 writes the 32 bytes of K_k into the swizzled B layout
    if (tid < 64) {
        Fr c;
        if (tid < 32) {
            Fr one_int = fr::zero();
            one_int.l[0] = 1;
            c = fr::mul(tcf::pow256(tid), one_int);  // 2^(8k) * R^-1 mod p
        } else {
            c = tcf::pow256(tid - 32);
        }
#pragma unroll
        for (uint32_t n = 0; n < 32; n++) kmat[tcf::bmat_offset(n, tid)] = (uint8_t)(c.l[n >> 2] >> (8 * (n & 3)));
    }
    const size_t row = (size_t)blockIdx.x * 128 + tid;
    const Fr a = fr::load(a_in + row * 8), b = fr::load(b_in + row * 8);
    const long long c0 = clock64();
    // T = a*b as 16 limbs: even/odd accumulators merged
    uint32_t ev[16], od[16], T[16];
    fr::mul_wide_eo(ev, od, a.l, b.l);
    {
        uint64_t carry = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const uint64_t s = (uint64_t)ev[i] + (i > 0 ? od[i - 1] : 0u) + carry;
            T[i] = (uint32_t)s;
            carry = s >> 32;
        }
    }
    // this thread's row: logical 16-byte chunk c at (c ^ (row % 8))
#pragma unroll
    for (uint32_t c = 0; c < 4; c++)
        *reinterpret_cast<uint4*>(tile + tid * 128 + ((c ^ (tid & 7)) << 4)) = make_uint4(T[4 * c], T[4 * c + 1], T[4 * c + 2], T[4 * c + 3]);
    tcf::fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core
    tcf::tc_fence_before();
    __syncthreads();
    tcf::tc_fence_after();
    const uint32_t taddr = tmem_slot;
    long long c1 = 0;
    if (tid == 0) {
        const uint64_t ad = tcf::sw128_desc(tcf::smem_u32(tile)), bd = tcf::sw128_desc(tcf::smem_u32(kmat));
        tcf::umma_u8(taddr, ad, bd, 0);
        tcf::umma_u8(taddr, ad + 2, bd + 2, 1);
        tcf::umma_commit(&bar_mma);
        c1 = clock64();
    }
    tcf::mbar_wait(&bar_mma, 0);
    tcf::tc_fence_after();
    uint32_t S[32];
    tcf::tmem_ld32(taddr + ((warp * 32u) << 16), S);
    tcf::tmem_ld_wait();
    const Fr v = tcf::columns_to_fr(S);
    const long long c2 = clock64();
    fr::store(out_tc + row * 8, v);
    fr::store(out_ref + row * 8, fr::mul(a, b));
    if (blockIdx.x == 0 && tid == 0) { clk[0] = c1 - c0; clk[1] = c2 - c0; }
    tcf::tc_fence_before();
    __syncthreads();
    if (warp == 0) tcf::tmem_dealloc(taddr, 32);
}

static uint64_t rng_state = 0x1234567ULL;
static uint64_t rnd() { uint64_t z = (rng_state += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
static const uint64_t P[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static void rand_fr(uint64_t* x) {
    for (;;) {
        x[0] = rnd(); x[1] = rnd(); x[2] = rnd(); x[3] = rnd() & 0x7fffffffffffffffULL;
        for (int i = 3; i >= 0; i--) { if (x[i] < P[i]) return; if (x[i] > P[i]) break; }
    }
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

int main(int argc, char** argv) {
    const int tiles = argc > 1 ? atoi(argv[1]) : 64;
    const size_t rows = (size_t)tiles * 128;
    CK(fr::fr_init_constants());
    std::vector<uint64_t> a(rows * 4), b(rows * 4);
    for (size_t i = 0; i < rows; i++) { rand_fr(&a[i * 4]); rand_fr(&b[i * 4]); }
    for (int i = 0; i < 4; i++) { a[i] = 0; b[4 + i] = 0; a[8 + i] = P[i] - (i == 0); b[8 + i] = P[i] - (i == 0); }  // 0*x, x*0, (p-1)^2
    uint32_t *da, *db, *dtc, *dref; long long* dclk;
    CK(cudaMalloc(&da, rows * 32)); CK(cudaMalloc(&db, rows * 32)); CK(cudaMalloc(&dtc, rows * 32)); CK(cudaMalloc(&dref, rows * 32));
    CK(cudaMalloc(&dclk, 64)); CK(cudaMemset(dclk, 0, 64)); CK(cudaMemset(dtc, 0xff, rows * 32));
    CK(cudaMemcpy(da, a.data(), rows * 32, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, b.data(), rows * 32, cudaMemcpyHostToDevice));
    const size_t smem = tcf::TILE_BYTES + tcf::BMAT_BYTES + 1024;
    CK(cudaFuncSetAttribute(redc_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    redc_tc_kernel<<<tiles, 128, smem>>>(da, db, dtc, dref, dclk);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> tc(rows * 8), ref(rows * 8);
    long long clk[8];
    CK(cudaMemcpy(tc.data(), dtc, rows * 32, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ref.data(), dref, rows * 32, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(clk, dclk, 64, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < rows; i++)
        if (memcmp(&tc[i * 8], &ref[i * 8], 32)) { if (bad++ < 4) printf("  row %zu: tensor-core reduction %08x.. vs fr::mul %08x..\n", i, tc[i * 8], ref[i * 8]); }
    printf("rows=%zu mismatches=%zu | cycles: product+stage+issue %lld, whole round trip (product .. reduced element) %lld\n", rows, bad, clk[0], clk[1]);
    printf(bad ? "REDC_TC FAIL\n" : "REDC_TC OK\n");
    return bad ? 1 : 0;
}
