// Translation unit of the fused tail kernel (see tail_kernel.cuh).  Built with -DFR_COMPACT.
#if !defined(FR_COMPACT) && !defined(FR_COMPACT_OFF)
#error "tail.cu must be compiled with -DFR_COMPACT"
#endif
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include "resident_kernel.cuh"
#include "tail_kernel.cuh"
#include "tc_round.cuh"

namespace sck {

cudaError_t tail_init_constants() { return fr::fr_init_constants(); }  // this TU has its own copy of the __constant__ modulus

// Fold rounds (FOLD = true) are also launched from this compact translation unit: measured on B200 the out-of-line
// multiplier makes rounds >= 2 3-15 % faster (smaller code, no instruction-fetch stalls), while round 1 (no fold, pure
// streaming multiply-reduce) is 10 % faster fully inlined and stays in sumcheck.cu.
int fold_round_threads() { return ROUND_THREADS; }

int fold_round_occupancy(uint32_t npts) {
    int nb = 0;
    switch (npts) {
        case 1: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, round_kernel<1, true>, ROUND_THREADS, 0); break;
        case 2: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, round_kernel<2, true>, ROUND_THREADS, 0); break;
        case 3: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, round_kernel<3, true>, ROUND_THREADS, 0); break;
        case 4: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, round_kernel<4, true>, ROUND_THREADS, 0); break;
        default: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, round_kernel<5, true>, ROUND_THREADS, 0); break;
    }
    return nb < 1 ? 1 : nb;
}

cudaError_t launch_fold_round(uint32_t npts, int grid, const RoundParams& rp, cudaStream_t stream) {
    switch (npts) {
        case 1: round_kernel<1, true><<<grid, ROUND_THREADS, 0, stream>>>(rp); break;
        case 2: round_kernel<2, true><<<grid, ROUND_THREADS, 0, stream>>>(rp); break;
        case 3: round_kernel<3, true><<<grid, ROUND_THREADS, 0, stream>>>(rp); break;
        case 4: round_kernel<4, true><<<grid, ROUND_THREADS, 0, stream>>>(rp); break;
        case 5: round_kernel<5, true><<<grid, ROUND_THREADS, 0, stream>>>(rp); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// TMA + tensor-core fold rounds (tc_round.cuh): 128 threads, TC_DYN_SMEM bytes of dynamic shared memory
unsigned long long tc_min_pairs() { return TC_MIN_PAIRS; }

template <int NPTS, int M = 0>
static cudaError_t tc_prepare(int* occ) {
    static bool ready_dev[64] = {};
    static int blocks_dev[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    bool& ready = ready_dev[dev];  // function attributes are per device
    int& blocks = blocks_dev[dev];
    if (!ready) {
        cudaError_t e = cudaFuncSetAttribute(round_tc_kernel<NPTS, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_DYN_SMEM);
        if (e != cudaSuccess) return e;
        // three CTAs of ~58 KB each per SM: ask for the largest shared-memory carve-out (the default heuristic sizes it
        // for one CTA and would leave the SM with a single resident block)
        e = cudaFuncSetAttribute(round_tc_kernel<NPTS, M>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        // Resident CTAs per SM, by hand: cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for this kernel on
        // B200 whatever the shared-memory size (measured; three CTAs do co-reside and run 2x faster than one), so the
        // limits are taken from the function and device attributes: registers, shared memory, tensor-memory columns.
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, round_tc_kernel<NPTS, M>);
        if (e != cudaSuccess) return e;
        int regs_sm = 0, smem_sm = 0;
        cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
        cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        const int regs_cta = ((fa.numRegs + 7) / 8 * 8) * (int)TC_THREADS;
        const int smem_cta = (int)fa.sharedSizeBytes + (int)TC_DYN_SMEM + 1024;  // + the per-CTA system reservation
        blocks = regs_sm / regs_cta;
        if (smem_sm / smem_cta < blocks) blocks = smem_sm / smem_cta;
        if ((int)(512 / TC_TMEM_COLS) < blocks) blocks = 512 / TC_TMEM_COLS;
        if (blocks < 1) blocks = 1;
        if (getenv("SC_DEBUG"))
            fprintf(stderr, "round_tc_kernel<%d,%d>: %d CTAs/SM (regs %d, static smem %zu, dynamic smem %zu)\n", NPTS, M, blocks, fa.numRegs,
                    fa.sharedSizeBytes, (size_t)TC_DYN_SMEM);
        if (const char* f = getenv("SC_TC_OCC")) blocks = atoi(f);
        ready = true;
    }
    *occ = blocks;
    return cudaSuccess;
}

cudaError_t launch_fold_round_tc(uint32_t npts, uint32_t m, int sms, int max_grid, const RoundParams& rp, cudaStream_t stream) {
    int occ = 1;
    cudaError_t e;
    const bool s2 = m == 2 && npts == 2, s3 = m == 3 && npts == 3;  // the single-product builds
    switch (npts) {
        case 1: e = tc_prepare<1>(&occ); break;
        case 2: e = s2 ? tc_prepare<2, 2>(&occ) : tc_prepare<2>(&occ); break;
        case 3: e = s3 ? tc_prepare<3, 3>(&occ) : tc_prepare<3>(&occ); break;
        case 4: e = tc_prepare<4>(&occ); break;
        case 5: e = tc_prepare<5>(&occ); break;
        default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    const unsigned long long n_tiles = rp.n_pairs / tcf::TILE_ROWS;
    unsigned long long cap = (unsigned long long)sms * occ;
    if (cap > (unsigned long long)max_grid) cap = max_grid;
    const int grid = (int)(n_tiles < cap ? n_tiles : cap);
    switch (npts) {
        case 1: round_tc_kernel<1><<<grid, TC_THREADS, TC_DYN_SMEM, stream>>>(rp); break;
        case 2:
            if (s2) round_tc_kernel<2, 2><<<grid, TC_THREADS, TC_DYN_SMEM, stream>>>(rp);
            else round_tc_kernel<2><<<grid, TC_THREADS, TC_DYN_SMEM, stream>>>(rp);
            break;
        case 3:
            if (s3) round_tc_kernel<3, 3><<<grid, TC_THREADS, TC_DYN_SMEM, stream>>>(rp);
            else round_tc_kernel<3><<<grid, TC_THREADS, TC_DYN_SMEM, stream>>>(rp);
            break;
        case 4: round_tc_kernel<4><<<grid, TC_THREADS, TC_DYN_SMEM, stream>>>(rp); break;
        default: round_tc_kernel<5><<<grid, TC_THREADS, TC_DYN_SMEM, stream>>>(rp); break;
    }
    return cudaGetLastError();
}

// ---- resident rounds (resident_kernel.cuh): cooperative launch, so that every CTA is co-resident by construction
template <int NPTS>
static int resident_blocks_per_sm() {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, resident_kernel<NPTS>, RES_THREADS, 0);
    return nb < 1 ? 1 : (nb > 2 ? 2 : nb);
}
int resident_max_grid(uint32_t npts, int device, int sms) {
    static int cached[64][MAX_NPTS + 1] = {};
    int& v = cached[device & 63][npts <= (uint32_t)MAX_NPTS ? npts : 0];
    if (!v) {
        switch (npts) {
            case 1: v = resident_blocks_per_sm<1>(); break;
            case 2: v = resident_blocks_per_sm<2>(); break;
            case 3: v = resident_blocks_per_sm<3>(); break;
            case 4: v = resident_blocks_per_sm<4>(); break;
            default: v = resident_blocks_per_sm<5>(); break;
        }
    }
    return v * sms;
}

cudaError_t launch_resident(uint32_t npts, int grid, const ResidentParams& rp, cudaStream_t stream, bool cooperative) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(RES_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (cooperative && !getenv("SC_RES_NO_COOP")) ? 1 : 0;  // measured: no difference in launch cost; cooperative guarantees co-residency
    switch (npts) {
        case 1: return cudaLaunchKernelEx(&cfg, resident_kernel<1>, rp);
        case 2: return cudaLaunchKernelEx(&cfg, resident_kernel<2>, rp);
        case 3: return cudaLaunchKernelEx(&cfg, resident_kernel<3>, rp);
        case 4: return cudaLaunchKernelEx(&cfg, resident_kernel<4>, rp);
        case 5: return cudaLaunchKernelEx(&cfg, resident_kernel<5>, rp);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_tail(uint32_t degree, const TailParams& tp, cudaStream_t stream) {
    switch (degree) {
        case 1: tail_kernel<1><<<1, TAIL_THREADS, 0, stream>>>(tp); break;
        case 2: tail_kernel<2><<<1, TAIL_THREADS, 0, stream>>>(tp); break;
        case 3: tail_kernel<3><<<1, TAIL_THREADS, 0, stream>>>(tp); break;
        case 4: tail_kernel<4><<<1, TAIL_THREADS, 0, stream>>>(tp); break;
        case 5: tail_kernel<5><<<1, TAIL_THREADS, 0, stream>>>(tp); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace sck
