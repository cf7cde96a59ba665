// Stand-alone check of the operand layouts the "contraction on the tensor cores" kernels rely on (csrc/gemm_sum.cuh):
//   sum_b X_b (x) Y_b  over the pairs b of a tile, X_b / Y_b = the BYTES of big integers each thread / TMA left in shared memory,
// as ONE u8 x u8 -> s32 tcgen05.mma with the PAIR index as the K dimension.  Both operands are therefore MN-major (the bytes
// of one pair are contiguous, K = the row index).  Checked here against a host model:
//   D1  A = XA [128 pairs][128 B] SWIZZLE_128B MN-major, M = 128;  B = Y [128 pairs][64 B] SWIZZLE_64B MN-major as TMA writes it
//   D2  A = XB [128 pairs][ 64 B] SWIZZLE_64B  MN-major, M = 64 (where do the 64 rows land in tensor memory?)
//   D3  A = XC [128 pairs][128 B] SWIZZLE_128B, M = 128, only the first 64 bytes of a row meaningful
//   D4  as D1 with Y written by the threads (the same swizzle by hand)
// and times a stream of MMAs of each shape.
// build: nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a -I sumcheck_b200/csrc tools/microbench/gemmsum.cu -o tools/microbench/gemmsum
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc_fold.cuh"
#include "tmap_host.h"

constexpr uint32_t PAIRS = 128;

__device__ __forceinline__ uint64_t mn_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout_type, uint32_t lbo_bytes = 16) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ uint32_t idesc_u8_mn(uint32_t M, uint32_t N) {
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

__global__ void __launch_bounds__(128) gemmsum_test_kernel(const CUtensorMap* ymap, const uint8_t* xbytes /* [128][192] */,
                                                            const uint8_t* ybytes /* [128][64] */, uint32_t* dump /* [128 lanes][256 cols] */,
                                                            long long* clk) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* XA = smem;                  // 16 KB
    uint8_t* XC = smem + 16384;          // 16 KB
    uint8_t* XB = smem + 32768;          // 8 KB
    uint8_t* Y = smem + 40960;           // 8 KB (TMA)
    uint8_t* Y2 = smem + 49152;          // 8 KB (by hand)
    __shared__ __align__(8) uint64_t bar_tma, bar_mma;
    __shared__ uint32_t s_tmem;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        if (tcf::smem_u32(smem) & 1023u) __trap();
        tcf::mbar_init(&bar_tma, 1);
        tcf::mbar_init(&bar_mma, 1);
        tcf::fence_mbar_init();
    }
    if (warp == 0) tcf::tmem_alloc(&s_tmem, 256);
    // thread k = pair k writes its bytes
    const uint8_t* xr = xbytes + (size_t)tid * 192;
#pragma unroll
    for (uint32_t c = 0; c < 8; c++) {
        *reinterpret_cast<uint4*>(XA + tid * 128 + ((c ^ (tid & 7)) << 4)) = *reinterpret_cast<const uint4*>(xr + c * 16);
        uint4 junk = make_uint4(0x01010101u * (tid + c), 0xffffffffu, 0x7f7f7f7fu, tid);
        *reinterpret_cast<uint4*>(XC + tid * 128 + ((c ^ (tid & 7)) << 4)) = c < 4 ? *reinterpret_cast<const uint4*>(xr + 128 + c * 16) : junk;
    }
#pragma unroll
    for (uint32_t c = 0; c < 4; c++) {
        *reinterpret_cast<uint4*>(XB + tid * 64 + ((c ^ ((tid >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(xr + 128 + c * 16);
        *reinterpret_cast<uint4*>(Y2 + tid * 64 + ((c ^ ((tid >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(ybytes + (size_t)tid * 64 + c * 16);
    }
    tcf::fence_proxy_async_smem();
    tcf::tc_fence_before();
    __syncthreads();
    tcf::tc_fence_after();
    const uint32_t taddr = s_tmem;
    if (tid == 0) {
        tcf::mbar_expect_tx(&bar_tma, 8192);
        tcf::tma_load_tile(Y, ymap, &bar_tma, 0);
        tcf::mbar_wait(&bar_tma, 0);
        tcf::tc_fence_after();
        const uint32_t i128 = idesc_u8_mn(128, 64), i64 = idesc_u8_mn(64, 64);
        for (uint32_t ks = 0; ks < 4; ks++) {  // K = 32 pairs per instruction
            const uint64_t a1 = mn_desc(tcf::smem_u32(XA) + ks * 4096, 1024, 2), a3 = mn_desc(tcf::smem_u32(XC) + ks * 4096, 1024, 2);
            const uint64_t a2 = mn_desc(tcf::smem_u32(XB) + ks * 2048, 512, 4);
            const uint64_t b = mn_desc(tcf::smem_u32(Y) + ks * 2048, 512, 4), b2 = mn_desc(tcf::smem_u32(Y2) + ks * 2048, 512, 4);
            umma(taddr + 0, a1, b, i128, ks);
            umma(taddr + 64, a2, b, i64, ks);
            umma(taddr + 128, a3, b, i128, ks);
            umma(taddr + 192, a1, b2, i128, ks);
        }
        tcf::umma_commit(&bar_mma);
    }
    tcf::mbar_wait(&bar_mma, 0);
    tcf::tc_fence_after();
    const uint32_t lane_addr = taddr + ((warp * 32u) << 16);
    for (uint32_t c0 = 0; c0 < 256; c0 += 32) {
        uint32_t S[32];
        tcf::tmem_ld32(lane_addr + c0, S);
        tcf::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j++) dump[(size_t)tid * 256 + c0 + j] = S[j];
    }
    tcf::tc_fence_before();
    __syncthreads();
    // ---- timing: streams of 256 MMAs of one shape
    if (tid == 0) {
        tcf::tc_fence_after();
        const uint64_t a1 = mn_desc(tcf::smem_u32(XA), 1024, 2), a2 = mn_desc(tcf::smem_u32(XB), 512, 4), b = mn_desc(tcf::smem_u32(Y), 512, 4);
        for (int shape = 0; shape < 2; shape++) {
            const uint32_t id = shape == 0 ? idesc_u8_mn(128, 64) : idesc_u8_mn(64, 64);
            const long long t0 = clock64();
            for (uint32_t i = 0; i < 256; i++) umma(taddr + (shape == 0 ? 0 : 64), shape == 0 ? a1 : a2, b, id, 1);
            tcf::umma_commit(&bar_mma);
            const long long t1 = clock64();
            tcf::mbar_wait(&bar_mma, (shape + 1) & 1);
            const long long t2 = clock64();
            clk[shape * 2] = t1 - t0;
            clk[shape * 2 + 1] = t2 - t0;
        }
    }
    tcf::tc_fence_before();
    __syncthreads();
    if (warp == 0) tcf::tmem_dealloc(taddr, 256);
}

//   D5  A = XA (M = 128), B = three [128 pairs][64 B] SWIZZLE_64B arrays 8 KiB apart = N 192 through the leading byte offset
//   D6  A = XB (M = 64),  same B
__global__ void __launch_bounds__(128) gemmsum_wide_kernel(const uint8_t* xbytes /* [128][192] */, const uint8_t* ybytes /* [128][192] */,
                                                            uint32_t* dump /* [128 lanes][384 cols] */) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* XA = smem;             // 16 KB
    uint8_t* XB = smem + 16384;     // 8 KB
    uint8_t* Y3 = smem + 24576;     // 3 x 8 KB
    __shared__ __align__(8) uint64_t bar_mma;
    __shared__ uint32_t s_tmem;
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        tcf::mbar_init(&bar_mma, 1);
        tcf::fence_mbar_init();
    }
    if (warp == 0) tcf::tmem_alloc(&s_tmem, 512);
    const uint8_t* xr = xbytes + (size_t)tid * 192;
    const uint8_t* yr = ybytes + (size_t)tid * 192;
#pragma unroll
    for (uint32_t c = 0; c < 8; c++) *reinterpret_cast<uint4*>(XA + tid * 128 + ((c ^ (tid & 7)) << 4)) = *reinterpret_cast<const uint4*>(xr + c * 16);
#pragma unroll
    for (uint32_t c = 0; c < 4; c++) {
        *reinterpret_cast<uint4*>(XB + tid * 64 + ((c ^ ((tid >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(xr + 128 + c * 16);
        for (uint32_t a = 0; a < 3; a++)
            *reinterpret_cast<uint4*>(Y3 + a * 8192 + tid * 64 + ((c ^ ((tid >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(yr + a * 64 + c * 16);
    }
    tcf::fence_proxy_async_smem();
    tcf::tc_fence_before();
    __syncthreads();
    tcf::tc_fence_after();
    const uint32_t taddr = s_tmem;
    if (tid == 0) {
        const uint32_t i128 = idesc_u8_mn(128, 192), i64 = idesc_u8_mn(64, 192);
        for (uint32_t ks = 0; ks < 4; ks++) {
            const uint64_t b = mn_desc(tcf::smem_u32(Y3) + ks * 2048, 512, 4, 8192);
            umma(taddr + 0, mn_desc(tcf::smem_u32(XA) + ks * 4096, 1024, 2), b, i128, ks);
            umma(taddr + 192, mn_desc(tcf::smem_u32(XB) + ks * 2048, 512, 4), b, i64, ks);
        }
        tcf::umma_commit(&bar_mma);
    }
    tcf::mbar_wait(&bar_mma, 0);
    tcf::tc_fence_after();
    const uint32_t lane_addr = taddr + ((warp * 32u) << 16);
    for (uint32_t c0 = 0; c0 < 384; c0 += 32) {
        uint32_t S[32];
        tcf::tmem_ld32(lane_addr + c0, S);
        tcf::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j++) dump[(size_t)tid * 384 + c0 + j] = S[j];
    }
    tcf::tc_fence_before();
    __syncthreads();
    if (warp == 0) tcf::tmem_dealloc(taddr, 512);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

int main() {
    std::vector<uint8_t> x(PAIRS * 192), y(PAIRS * 64);
    uint64_t s = 0x9e3779b97f4a7c15ULL;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint8_t)(s >> 24); };
    for (auto& v : x) v = rnd();
    for (auto& v : y) v = rnd();
    for (int i = 0; i < 192; i++) x[i] = 0xff;  // extreme rows
    for (int i = 0; i < 64; i++) y[i] = 0xff;
    uint8_t *dx, *dy; uint32_t* dd; long long* dclk; CUtensorMap* dmap;
    CK(cudaMalloc(&dx, x.size())); CK(cudaMalloc(&dy, y.size())); CK(cudaMalloc(&dd, 128 * 256 * 4)); CK(cudaMalloc(&dclk, 64)); CK(cudaMalloc(&dmap, sizeof(CUtensorMap)));
    CK(cudaMemcpy(dx, x.data(), x.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dy, y.data(), y.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dd, 0xee, 128 * 256 * 4));
    CUtensorMap hmap;
    {
        tmaph::EncodeTiledFn enc = tmaph::encode_tiled_fn();
        if (!enc) { printf("FAIL: no cuTensorMapEncodeTiled\n"); return 1; }
        const cuuint64_t dims[2] = {16, PAIRS}; const cuuint64_t strides[1] = {64}; const cuuint32_t box[2] = {16, 128}; const cuuint32_t estr[2] = {1, 1};
        if (enc(&hmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, dy, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("FAIL: encode\n"); return 1; }
    }
    CK(cudaMemcpy(dmap, &hmap, sizeof(hmap), cudaMemcpyHostToDevice));
    const size_t smem = 57344 + 1024;
    CK(cudaFuncSetAttribute(gemmsum_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemmsum_test_kernel<<<1, 128, smem>>>(dmap, dx, dy, dd, dclk);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> d(128 * 256);
    long long clk[8];
    CK(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(clk, dclk, 64, cudaMemcpyDeviceToHost));
    // host model: W[r][c] = sum_k x[k][r] * y[k][c], r < 192, c < 64
    std::vector<uint32_t> W(192 * 64, 0);
    for (uint32_t k = 0; k < PAIRS; k++)
        for (uint32_t r = 0; r < 192; r++)
            for (uint32_t c = 0; c < 64; c++) W[r * 64 + c] += (uint32_t)x[k * 192 + r] * y[k * 64 + c];
    int bad1 = 0, bad3 = 0, bad4 = 0;
    for (uint32_t r = 0; r < 128; r++)
        for (uint32_t c = 0; c < 64; c++) {
            if (d[r * 256 + c] != W[r * 64 + c] && bad1++ < 4) printf("  D1[%u][%u] got %u want %u\n", r, c, d[r * 256 + c], W[r * 64 + c]);
            if (d[r * 256 + 192 + c] != W[r * 64 + c] && bad4++ < 4) printf("  D4[%u][%u] got %u want %u\n", r, c, d[r * 256 + 192 + c], W[r * 64 + c]);
            if (r < 64 && d[r * 256 + 128 + c] != W[(128 + r) * 64 + c] && bad3++ < 4) printf("  D3[%u][%u] got %u want %u\n", r, c, d[r * 256 + 128 + c], W[(128 + r) * 64 + c]);
        }
    // D2 (M = 64): find the lane every row landed in
    int bad2 = 0, lane_of[64];
    for (uint32_t r = 0; r < 64; r++) {
        lane_of[r] = -1;
        for (uint32_t l = 0; l < 128; l++) {
            bool ok = true;
            for (uint32_t c = 0; c < 64 && ok; c++) ok = d[l * 256 + 64 + c] == W[(128 + r) * 64 + c];
            if (ok) { lane_of[r] = (int)l; break; }
        }
        if (lane_of[r] < 0) bad2++;
    }
    printf("M=64 rows -> TMEM lanes:");
    for (int r = 0; r < 64; r++) printf(" %d", lane_of[r]);
    printf("\n");
    printf("256 MMAs M=128 N=64 K=32: issue %lld cycles, issue->done %lld (%.1f per MMA);  M=64: issue %lld, done %lld (%.1f per MMA)\n", clk[0], clk[1],
           clk[1] / 256.0, clk[2], clk[3], clk[3] / 256.0);
    printf("D1 (SW128 A, TMA SW64 B) bad=%d   D2 (M=64, SW64 A) rows not found=%d   D3 (M=128, half rows) bad=%d   D4 (hand-swizzled B) bad=%d\n", bad1, bad2, bad3, bad4);
    // ---- N = 192 through the leading byte offset
    int bad5 = 0, bad6 = 0;
    {
        std::vector<uint8_t> y3(PAIRS * 192);
        for (auto& v : y3) v = rnd();
        uint8_t* dy3; uint32_t* dd3;
        CK(cudaMalloc(&dy3, y3.size())); CK(cudaMalloc(&dd3, 128 * 384 * 4));
        CK(cudaMemcpy(dy3, y3.data(), y3.size(), cudaMemcpyHostToDevice));
        const size_t smem3 = 49152 + 1024;
        CK(cudaFuncSetAttribute(gemmsum_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
        gemmsum_wide_kernel<<<1, 128, smem3>>>(dx, dy3, dd3);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> d3(128 * 384);
        CK(cudaMemcpy(d3.data(), dd3, d3.size() * 4, cudaMemcpyDeviceToHost));
        std::vector<uint32_t> W3(192 * 192, 0);
        for (uint32_t k = 0; k < PAIRS; k++)
            for (uint32_t r = 0; r < 192; r++)
                for (uint32_t c = 0; c < 192; c++) W3[r * 192 + c] += (uint32_t)x[k * 192 + r] * y3[k * 192 + c];
        for (uint32_t r = 0; r < 128; r++)
            for (uint32_t c = 0; c < 192; c++)
                if (d3[r * 384 + c] != W3[r * 192 + c] && bad5++ < 4) printf("  D5[%u][%u] got %u want %u\n", r, c, d3[r * 384 + c], W3[r * 192 + c]);
        for (uint32_t r = 0; r < 64; r++) {
            const uint32_t l = (r % 16) + 32 * (r / 16);
            for (uint32_t c = 0; c < 192; c++)
                if (d3[l * 384 + 192 + c] != W3[(128 + r) * 192 + c] && bad6++ < 4) printf("  D6[%u][%u] got %u want %u\n", r, c, d3[l * 384 + 192 + c], W3[(128 + r) * 192 + c]);
        }
        printf("D5 (M=128, N=192 as three SWIZZLE_64B atoms, LBO 8192) bad=%d   D6 (M=64, N=192) bad=%d\n", bad5, bad6);
    }
    const bool ok = !(bad1 | bad2 | bad3 | bad4 | bad5 | bad6);
    printf(ok ? "GEMMSUM OK\n" : "GEMMSUM FAIL\n");
    return ok ? 0 : 1;
}
