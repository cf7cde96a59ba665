// Translation unit of the fused tail kernel (see tail_kernel.cuh).  Built with -DFR_COMPACT.
#ifndef FR_COMPACT
#error "tail.cu must be compiled with -DFR_COMPACT"
#endif
#include <cuda_runtime.h>

#include "tail_kernel.cuh"

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
