// Parameters of the fused tail kernel, shared by the launcher TU (sumcheck.cu) and the kernel TU (tail.cu).
#pragma once
#include "blake2b.cuh"
#include "kernels.cuh"

namespace sck {

constexpr unsigned long long TAIL_PAIRS = 1024;
constexpr int TAIL_THREADS = 256;

struct TailParams {
    RoundParams rp;                 // products, coefficients, lagrange, degree, r of the first tail round, prev_evals
    uint32_t* const* ptrs[3];       // device arrays of table pointers: pristine, ping, pong
    int cur;                        // which of them holds the tables entering the first tail round
    uint32_t n_rounds;              // rounds to run
    unsigned long long n_pairs_first;
    const b2::State* st_in;         // transcript after the previous round's challenge was drawn (buflen % 8 == 0)
    b2::State* st_out;
    uint32_t* evals_all;            // [n_rounds][d+1][8] Montgomery: the ProverMsgs of the tail rounds
    uint32_t* chal_all;             // [n_rounds][8] challenge drawn after each tail round
    long long* prof;                // optional [n_rounds][4] cycle counts: accumulate, reduce, publish, transcript (SC_TAIL_PROF)
};

cudaError_t tail_init_constants();
int fold_round_occupancy(uint32_t npts);
int fold_round_threads();
cudaError_t launch_fold_round(uint32_t npts, int grid, const RoundParams& rp, cudaStream_t stream);
unsigned long long tc_min_pairs();
cudaError_t launch_fold_round_tc(uint32_t npts, int sms, int max_grid, const RoundParams& rp, cudaStream_t stream);
cudaError_t launch_tail(uint32_t degree, const TailParams& tp, cudaStream_t stream);

}  // namespace sck
