// Translation unit of the tensor-core contraction kernels (gemm_sum.cuh): launchers only.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "gemm_sum.cuh"

namespace gsum {

cudaError_t init_constants() { return fr::fr_init_constants(); }  // this TU's copy of the __constant__ modulus

#ifndef SC_GEMM_G1
#define SC_GEMM_G1 3  // compute groups per CTA, round 1 (three: 144 registers per thread leave room for the next item's pairs)
#endif
#ifndef SC_GEMM_GF
#define SC_GEMM_GF 3  // compute groups per CTA, fold rounds of three-table products (shared memory and tensor memory allow three)
#endif
#ifndef SC_GEMM_GF4
#define SC_GEMM_GF4 2  // ... of four-table products: the 192 x 192 contraction takes 384 of the 512 tensor-memory columns
#endif
constexpr int G1 = SC_GEMM_G1, GF = SC_GEMM_GF, GF4 = SC_GEMM_GF4;

// TAG: one instantiation (and one set of per-device flags) per kernel — the kernels share a function-pointer type
template <int TAG, class K>
static cudaError_t prepare(K kernel, size_t smem) {
    static bool ready_dev[64] = {};
    static std::mutex mu;  // handles may live on different host threads (one per rank of an in-process group)
    std::lock_guard<std::mutex> lk(mu);
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (ready_dev[dev]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    if (getenv("SC_DEBUG")) {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess)
            fprintf(stderr, "gemm kernel %d: %d registers, %zu B static + %zu B dynamic shared memory\n", TAG, fa.numRegs, fa.sharedSizeBytes, smem);
    }
    ready_dev[dev] = true;
    return cudaSuccess;
}

// grid: one CTA per SM, never more CTAs than there are items for their first groups
static int grid_for(uint32_t items, int G, int sms) {
    const uint32_t need = (items + G - 1) / G;
    return (int)(need < (uint32_t)sms ? need : (uint32_t)sms);
}

// items a single launch may carry (every CTA at most MAX_ITEMS_PER_CTA per group)
unsigned long long max_items_round1(int sms, int mm) { return (unsigned long long)sms * (mm == 2 ? 2 : G1) * MAX_ITEMS_PER_CTA; }  // (mm = 2: two CTAs per SM)
unsigned long long max_items_fold(int sms, int mm) { return (unsigned long long)sms * (mm == 4 ? GF4 : GF) * MAX_ITEMS_PER_CTA; }

template <int TAG, int G, int MM>
static cudaError_t launch_r1(const Params& P, int sms, cudaStream_t stream) {
    const size_t smem = R1Smem<G, MM>::BYTES;
    cudaError_t e = prepare<TAG>(gemm_round1_kernel<G, MM>, smem);
    if (e != cudaSuccess) return e;
    gemm_round1_kernel<G, MM><<<grid_for(P.items, G, sms), G * 128 + 64, smem, stream>>>(P);
    return cudaGetLastError();
}
template <int TAG, int G, int MM, bool PRE>
static cudaError_t launch_f(const Params& P, int sms, cudaStream_t stream) {
    const size_t smem = FoldSmem<G, MM>::BYTES;
    cudaError_t e = prepare<TAG>(gemm_fold_kernel<G, MM, PRE>, smem);
    if (e != cudaSuccess) return e;
    gemm_fold_kernel<G, MM, PRE><<<grid_for(P.items, G, sms), G * 128 + 96, smem, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_round1(const Params& P, int mm, int sms, cudaStream_t stream) {
    if (mm == 2) {  // two-table products: TMA + MMA only
        cudaError_t e = prepare<7>(gemm_round1_raw_kernel<2>, RAW_SMEM);
        if (e != cudaSuccess) return e;
        gemm_round1_raw_kernel<2><<<grid_for(P.items, 1, 2 * sms), RAW_THREADS, RAW_SMEM, stream>>>(P);
        return cudaGetLastError();
    }
    return mm == 4 ? launch_r1<3, G1, 4>(P, sms, stream) : launch_r1<1, G1, 3>(P, sms, stream);
}
cudaError_t launch_fold(const Params& P, int mm, int sms, cudaStream_t stream) {
    if (mm == 2) return P.r_mail ? launch_f<9, GF, 2, true>(P, sms, stream) : launch_f<8, GF, 2, false>(P, sms, stream);
    if (P.r_mail) return mm == 4 ? launch_f<6, GF4, 4, true>(P, sms, stream) : launch_f<5, GF, 3, true>(P, sms, stream);
    return mm == 4 ? launch_f<4, GF4, 4, false>(P, sms, stream) : launch_f<2, GF, 3, false>(P, sms, stream);
}

}  // namespace gsum
