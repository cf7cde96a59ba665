// Stand-alone check of the tensor-core fold (sumcheck_b200/csrc/tc_fold.cuh) on one B200:
//   TMA (SWIZZLE_128B tensor map) -> shared memory -> tcgen05.mma.kind::i8 -> tensor memory -> columns_to_fr
// against (a) the IMAD fold the round kernel used so far, computed by the same threads from plain global loads, and
// (b) a host big-integer model.  Prints the first mismatches of every stage so a wrong descriptor / swizzle / layout
// can be told apart in ONE run.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I sumcheck_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc_fold.cuh"
#include "tmap_host.h"

using fr::Fr;

__global__ void __launch_bounds__(128) tcfold_test_kernel(const CUtensorMap* tmap, const uint32_t* table, const uint32_t* r_mont,
                                                          uint32_t* raw_cols, uint32_t* fr_tc, uint32_t* fr_imad, uint32_t* lds_rows,
                                                          uint8_t* bmat_out, long long* clk) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* tile = smem;
    uint8_t* bmat = smem + tcf::TILE_BYTES;
    __shared__ __align__(8) uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_slot;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = r_mont[i];
    if (tid == 0) {
        tcf::mbar_init(&bar_full, 1);
        tcf::mbar_init(&bar_mma, 1);
        tcf::fence_mbar_init();
        tcf::prefetch_tmap(tmap);
    }
    if (warp == 0) tcf::tmem_alloc(&tmem_slot, 64);
    for (uint32_t i = tid; i < tcf::BMAT_BYTES; i += blockDim.x) bmat[i] = 0;
    __syncthreads();
    tcf::build_bmat(r, bmat);
    tcf::fence_proxy_async_smem();
    tcf::tc_fence_before();
    __syncthreads();
    tcf::tc_fence_after();
    const uint32_t taddr = tmem_slot;
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if (tid == 0) {
        c0 = clock64();
        tcf::mbar_expect_tx(&bar_full, tcf::TILE_BYTES);
        tcf::tma_load_tile(tile, tmap, &bar_full, blockIdx.x * tcf::TILE_ROWS);
    }
    tcf::mbar_wait(&bar_full, 0);
    if (tid == 0) {
        c1 = clock64();
        tcf::tc_fence_after();
        tcf::issue_fold_mma(tcf::smem_u32(tile), tcf::smem_u32(bmat), taddr);
        tcf::umma_commit(&bar_mma);
        c2 = clock64();
        tcf::mbar_wait(&bar_mma, 0);
        c3 = clock64();
        // a second, warm round of MMAs on the same tile: issue -> completion latency without cold descriptors
        tcf::issue_fold_mma(tcf::smem_u32(tile), tcf::smem_u32(bmat), taddr);
        tcf::umma_commit(&bar_full);  // reuse: phase 1 of bar_full
        const long long c4 = clock64();
        tcf::mbar_wait(&bar_full, 1);
        const long long c5 = clock64();
        if (blockIdx.x == 0) { clk[0] = c1 - c0; clk[1] = c2 - c1; clk[2] = c3 - c2; clk[3] = c5 - c4; }
    }
    // stage 1 check: this thread's row as TMA + swizzle left it
    const size_t row = (size_t)blockIdx.x * tcf::TILE_ROWS + tid;
#pragma unroll
    for (uint32_t c = 0; c < 8; c++) {
        const uint4 v = *reinterpret_cast<const uint4*>(tile + tid * 128 + ((c ^ (tid & 7)) << 4));
        *reinterpret_cast<uint4*>(lds_rows + row * 32 + c * 4) = v;
    }
    if (blockIdx.x == 0)
        for (uint32_t i = tid; i < tcf::BMAT_BYTES; i += blockDim.x) bmat_out[i] = bmat[i];
    tcf::mbar_wait(&bar_mma, 0);
    tcf::tc_fence_after();
    uint32_t S0[32], S1[32];
    const uint32_t lane_addr = taddr + ((warp * 32u) << 16);
    const long long l0 = clock64();
    tcf::tmem_ld32(lane_addr, S0);
    tcf::tmem_ld32(lane_addr + 32, S1);
    tcf::tmem_ld_wait();
    const long long l1 = clock64();
    if (blockIdx.x == 0 && tid == 0) clk[4] = l1 - l0;
#pragma unroll
    for (int j = 0; j < 32; j++) {
        raw_cols[row * 64 + j] = S0[j];
        raw_cols[row * 64 + 32 + j] = S1[j];
    }
    fr::store(fr_tc + row * 16, tcf::columns_to_fr(S0));
    fr::store(fr_tc + row * 16 + 8, tcf::columns_to_fr(S1));
    // the fold as the round kernel computed it so far
    const uint32_t* src = table + row * 32;
    const Fr e0 = fr::load(src), e1 = fr::load(src + 8), e2 = fr::load(src + 16), e3 = fr::load(src + 24);
    fr::store(fr_imad + row * 16, fr::add(e0, fr::mul(r, fr::sub(e1, e0))));
    fr::store(fr_imad + row * 16 + 8, fr::add(e2, fr::mul(r, fr::sub(e3, e2))));
    tcf::tc_fence_before();
    __syncthreads();
    if (warp == 0) tcf::tmem_dealloc(taddr, 64);
}

// ---- host big integers (4 x u64), only what the model needs ---------------------------------------------------------
struct U256 { uint64_t w[4]; };
static const U256 P = {{0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL}};
static int cmp(const U256& a, const U256& b) {
    for (int i = 3; i >= 0; i--) { if (a.w[i] < b.w[i]) return -1; if (a.w[i] > b.w[i]) return 1; }
    return 0;
}
static U256 addmod(const U256& a, const U256& b) {
    U256 r; unsigned __int128 c = 0;
    for (int i = 0; i < 4; i++) { c += (unsigned __int128)a.w[i] + b.w[i]; r.w[i] = (uint64_t)c; c >>= 64; }
    if (c || cmp(r, P) >= 0) { unsigned __int128 bw = 0; for (int i = 0; i < 4; i++) { unsigned __int128 d = (unsigned __int128)r.w[i] - P.w[i] - (uint64_t)bw; r.w[i] = (uint64_t)d; bw = (d >> 64) & 1; } }
    return r;
}
static U256 submod(const U256& a, const U256& b) {
    U256 r; unsigned __int128 bw = 0;
    for (int i = 0; i < 4; i++) { unsigned __int128 d = (unsigned __int128)a.w[i] - b.w[i] - (uint64_t)bw; r.w[i] = (uint64_t)d; bw = (d >> 64) & 1; }
    if (bw) { unsigned __int128 c = 0; for (int i = 0; i < 4; i++) { c += (unsigned __int128)r.w[i] + P.w[i]; r.w[i] = (uint64_t)c; c >>= 64; } }
    return r;
}
static U256 mulmod(const U256& a, const U256& b) {  // double-and-add
    U256 acc = {{0, 0, 0, 0}};
    for (int i = 255; i >= 0; i--) {
        acc = addmod(acc, acc);
        if ((b.w[i >> 6] >> (i & 63)) & 1) acc = addmod(acc, a);
    }
    return acc;
}
static uint64_t rng_state = 0x9e3779b97f4a7c15ULL;
static uint64_t rnd() { uint64_t z = (rng_state += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
static U256 rand_fr() { for (;;) { U256 x = {{rnd(), rnd(), rnd(), rnd() & 0x7fffffffffffffffULL}}; if (cmp(x, P) < 0) return x; } }

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

int main(int argc, char** argv) {
    const int tiles = argc > 1 ? atoi(argv[1]) : 4;
    const size_t rows = (size_t)tiles * 128, elems = rows * 4;
    CK(fr::fr_init_constants());
    std::vector<U256> tab(elems);
    for (auto& x : tab) x = rand_fr();
    // edge values in the first rows: 0, p-1, all-ones-ish bytes
    memset(&tab[0], 0, sizeof(U256));
    tab[1] = submod(tab[0], U256{{1, 0, 0, 0}});
    tab[2] = tab[1]; tab[3] = tab[1];
    const U256 r = rand_fr();  // canonical challenge
    const U256 Rm = {{0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL}};
    const U256 r_mont = mulmod(r, Rm), one = {{1, 0, 0, 0}}, omr = submod(one, r);
    // host model
    std::vector<U256> want(rows * 2);
    for (size_t b = 0; b < rows * 2; b++) want[b] = addmod(mulmod(tab[2 * b], omr), mulmod(tab[2 * b + 1], r));
    std::vector<U256> C(64);
    for (int k = 0; k < 64; k++) { U256 c = k < 32 ? omr : r; for (int s = 0; s < 8 * (k & 31); s++) c = addmod(c, c); C[k] = c; }

    uint32_t *d_tab, *d_r, *d_raw, *d_tc, *d_imad, *d_lds; uint8_t* d_bmat; CUtensorMap* d_map; long long* d_clk; CK(cudaMalloc(&d_clk, 64)); CK(cudaMemset(d_clk, 0, 64));
    CK(cudaMalloc(&d_tab, elems * 32)); CK(cudaMalloc(&d_r, 32)); CK(cudaMalloc(&d_raw, rows * 64 * 4));
    CK(cudaMalloc(&d_tc, rows * 64)); CK(cudaMalloc(&d_imad, rows * 64)); CK(cudaMalloc(&d_lds, rows * 128));
    CK(cudaMalloc(&d_bmat, tcf::BMAT_BYTES)); CK(cudaMalloc(&d_map, sizeof(CUtensorMap)));
    CK(cudaMemcpy(d_tab, tab.data(), elems * 32, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_r, &r_mont, 32, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_raw, 0xff, rows * 64 * 4)); CK(cudaMemset(d_tc, 0xff, rows * 64)); CK(cudaMemset(d_lds, 0xff, rows * 128));
    CUtensorMap hmap;
    if (!tmaph::make_table_map(&hmap, d_tab, rows, 128)) { printf("FAIL: cuTensorMapEncodeTiled\n"); return 1; }
    CK(cudaMemcpy(d_map, &hmap, sizeof(hmap), cudaMemcpyHostToDevice));
    const size_t smem = tcf::TILE_BYTES + tcf::BMAT_BYTES + 1024;
    CK(cudaFuncSetAttribute(tcfold_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tcfold_test_kernel<<<tiles, 128, smem>>>(d_map, d_tab, d_r, d_raw, d_tc, d_imad, d_lds, d_bmat, d_clk);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> raw(rows * 64), tc(rows * 16), imad(rows * 16), lds(rows * 32);
    std::vector<uint8_t> bm(tcf::BMAT_BYTES);
    CK(cudaMemcpy(raw.data(), d_raw, rows * 256, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(tc.data(), d_tc, rows * 64, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(imad.data(), d_imad, rows * 64, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(lds.data(), d_lds, rows * 128, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(bm.data(), d_bmat, tcf::BMAT_BYTES, cudaMemcpyDeviceToHost));

    long long hclk[8];
    CK(cudaMemcpy(hclk, d_clk, 64, cudaMemcpyDeviceToHost));
    printf("cycles: TMA 16 KiB issue->landed %lld | 4 UMMA issue+commit %lld | commit->mbarrier (cold) %lld | issue->mbarrier (warm) %lld | 2 x LDTM.x32 + wait %lld\n",
           hclk[0], hclk[1], hclk[2], hclk[3], hclk[4]);
    int bad_lds = 0, bad_bm = 0, bad_raw = 0, bad_tc = 0, bad_imad = 0;
    const uint8_t* tb = (const uint8_t*)tab.data();
    for (size_t i = 0; i < rows; i++)
        if (memcmp(&lds[i * 32], tb + i * 128, 128)) { if (bad_lds++ < 4) printf("  swizzled row %zu differs from the table (first words %08x vs %08x)\n", i, lds[i * 32], *(const uint32_t*)(tb + i * 128)); }
    for (uint32_t n = 0; n < 32; n++)
        for (uint32_t k = 0; k < 64; k++) {
            const uint8_t w = ((const uint8_t*)&C[k])[n];
            const uint8_t g = bm[n * 128 + ((((k >> 4) ^ (n & 7)) << 4) | (k & 15))];
            if (w != g && bad_bm++ < 4) printf("  constants matrix (n=%u,k=%u): got %02x want %02x\n", n, k, g, w);
        }
    for (size_t i = 0; i < rows; i++)
        for (int h = 0; h < 2; h++)
            for (int j = 0; j < 32; j++) {
                uint32_t s = 0;
                for (int k = 0; k < 64; k++) s += (uint32_t)tb[i * 128 + h * 64 + k] * ((const uint8_t*)&C[k])[j];
                const uint32_t g = raw[i * 64 + h * 32 + j];
                if (s != g && bad_raw++ < 8) printf("  column sum row %zu half %d col %d: got %u want %u\n", i, h, j, g, s);
            }
    for (size_t b = 0; b < rows * 2; b++) {
        if (memcmp(&tc[b * 8], &want[b], 32) && bad_tc++ < 4) printf("  tensor-core fold %zu: got %08x.. want %08x..\n", b, tc[b * 8], (uint32_t)want[b].w[0]);
        if (memcmp(&imad[b * 8], &want[b], 32) && bad_imad++ < 4) printf("  IMAD fold %zu: got %08x.. want %08x..\n", b, imad[b * 8], (uint32_t)want[b].w[0]);
    }
    printf("rows=%zu  swizzle(TMA->LDS) bad=%d  constants bad=%d  column sums bad=%d  tc fold bad=%d  imad fold bad=%d\n", rows, bad_lds, bad_bm,
           bad_raw, bad_tc, bad_imad);
    const bool ok = !(bad_lds | bad_bm | bad_raw | bad_tc | bad_imad);
    printf(ok ? "TCFOLD OK\n" : "TCFOLD FAIL\n");
    return ok ? 0 : 1;
}
