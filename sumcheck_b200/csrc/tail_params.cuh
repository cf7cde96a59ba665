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

constexpr int RES_THREADS = 128;       // one warp per scheduler: these rounds are latency-bound
constexpr uint32_t RES_CONST_WORDS = 64;

struct ResidentParams {
    RoundParams rp;                    // products, coefficients, peer mailboxes; skip1 = 1, write_fold = 1, t0 = 0
    uint32_t* const* ptrs[3];          // device arrays of table pointers: pristine, ping, pong
    int cur;                           // which of them holds the tables entering the first resident round
    uint32_t n_rounds;
    unsigned long long n_pairs_first;  // output pairs of the first resident round
    uint32_t seq0;                     // sequence number of the first resident round; +1 per round
    uint32_t mail_seq0;                // sharded: mailbox sequence number of the first resident round
    const uint32_t* h_consts;          // mapped host memory: [64] {limb, seq}
    uint32_t* h_sums;                  // mapped host memory: [NPTS * 17] {limb, seq}: unreduced sums
    uint32_t* h_error;                 // mapped host word: set to 1 when the host (or a peer) did not answer in time
    const uint32_t* h_abort;           // mapped host word: the host sets it to seq0 when it abandons the proof (error paths)
    uint32_t* d_abort;                 // device word through which CTA 0 passes the abort on
    uint32_t* d_bcast;                 // device memory: [64] {limb, seq}, CTA 0 -> the other CTAs
    unsigned long long* totals;        // device memory: [2][NPTS][17] per-limb sums over the CTAs of a round, zeroed before the launch
    unsigned int* counters;            // device memory: [n_rounds], zeroed before the launch
    // fine-grained rounds (kernels.cuh accumulate_fine): rounds with at most fine_max_pairs pairs spread every pair over
    // 2^lpp_log2 lanes; 0 = never
    unsigned long long fine_max_pairs;
    uint32_t lpp_log2;
    long long timeout;                 // clock64 ticks to wait for the host
    long long* prof;                   // optional [n_rounds][4] cycle counts of CTA 0: wait, accumulate, reduce, publish (SC_RES_PROF)
};

cudaError_t tail_init_constants();
int resident_max_grid(uint32_t npts, int device, int sms);  // co-resident CTAs of resident_kernel<npts> (cooperative launch)
// cooperative = false for ranks that share a device: two cooperative grids that wait for each other's partial sums must not be
// serialised by the launch mechanism; their co-residency then follows from the grid cap alone (resident_launch)
cudaError_t launch_resident(uint32_t npts, int grid, const ResidentParams& rp, cudaStream_t stream, bool cooperative);
int fold_round_occupancy(uint32_t npts);
int fold_round_threads();
cudaError_t launch_fold_round(uint32_t npts, int grid, const RoundParams& rp, cudaStream_t stream);
unsigned long long tc_min_pairs();
// m > 0: the single-product build (one product of m == npts multiplicands, deferred coefficient); 0: the CSR-driven build
cudaError_t launch_fold_round_tc(uint32_t npts, uint32_t m, int sms, int max_grid, const RoundParams& rp, cudaStream_t stream);
cudaError_t launch_tail(uint32_t degree, const TailParams& tp, cudaStream_t stream);

}  // namespace sck
