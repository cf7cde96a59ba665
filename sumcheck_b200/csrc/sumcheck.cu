// libsumcheck_b200.so — C ABI (include/sumcheck_b200.h) over the sm_100a kernels in kernels.cuh.
// Host logic here is only what the reference's L3 drivers do around the hot path: the prove_round state machine
// (prover.rs:74-98), the round loop + Fiat-Shamir transcript (ml_sumcheck/mod.rs:50-70), buffer rotation.  All field
// arithmetic runs on the device; there is no CPU fallback — without a CUDA device every entry point fails.
#include <cuda_runtime.h>

#include <sched.h>
#include <time.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#endif

#include "../../include/sumcheck_b200.h"
#include "blake2b.cuh"
#include "kernels.cuh"
#include "gkr_kernels.cuh"
#include "tail_params.cuh"
#include "tmap_host.h"
#include "host_fr.h"
#include "host_copy.h"
#include "tma_round1.cuh"
#include "gemm_sum.cuh"

static_assert(sizeof(sc_blake2b512_rng) == sizeof(b2::State), "sc_blake2b512_rng must be layout-identical to b2::State");

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) return fail(SC_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

// Host-side wait for a word the GPU writes into mapped pinned memory.  The expected wait is a few microseconds (one
// kernel), so the first polls only PAUSE (frees the sibling hyper-thread, keeps the store visible at once); a long wait
// (a multi-millisecond round of a big proof, or a peer rank still uploading) backs off to sched_yield so that a caller
// which also runs a rayon/OpenMP pool on the same cores is not starved (SURVEY §8b "threading").
struct SpinWait {
    unsigned long long n = 0;
    inline void pause() {
        ++n;
        if (n < 2048) {
#if defined(__x86_64__) || defined(__i386__)
            _mm_pause();
#elif defined(__aarch64__)
            asm volatile("yield" ::: "memory");
#endif
        } else if (n < 65536) {
#if defined(__x86_64__) || defined(__i386__)
            for (int i = 0; i < 16; i++) _mm_pause();
#endif
        } else {
            sched_yield();
        }
    }
    inline bool check_now() const { return (n & 0xffff) == 0; }  // time to ask the driver whether the kernel died
};

// Nsight Compute makes every launch synchronous and replays kernels: a kernel that waits for words the host writes AFTER the launch call
// returns (the resident rounds, a fold round launched ahead of its challenge) can never finish under it.  ncu marks the profiled process
// with these variables; with them set the library runs one self-contained launch per round, so `ncu python bench.py` just works.
bool profiler_attached() {
    static const bool on = getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") != nullptr || getenv("NV_NSIGHT_INJECTION_TRANSPORT_TYPE") != nullptr ||
                           (getenv("CUDA_LAUNCH_BLOCKING") && atoi(getenv("CUDA_LAUNCH_BLOCKING")) != 0);  // (the same holds for blocking launches)
    return on;
}

struct DeviceInfo {
    bool ready = false;
    int sms = 0;
    int khz = 0;  // SM clock: cudaDeviceGetAttribute(cudaDevAttrClockRate) takes 1-4 ms per call (measured), so it is read once
};
DeviceInfo g_dev[64];
std::mutex g_dev_mu;  // one-time per-device initialisation (constants, attributes); handles may live on different threads

int ensure_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(SC_ERR_NO_DEVICE, "no CUDA device (%s); sumcheck_b200 has no CPU path", cudaGetErrorString(e));
    if (device < 0 || device >= n || device >= 64) return fail(SC_ERR_BAD_INPUT, "device %d out of range (%d visible)", device, n);
    CUDA_TRY(cudaSetDevice(device));
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (!g_dev[device].ready) {
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) return fail(SC_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        CUDA_TRY(fr::fr_init_constants());
        CUDA_TRY(sck::tail_init_constants());
        CUDA_TRY(gsum::init_constants());
        g_dev[device].sms = prop.multiProcessorCount;
        cudaDeviceGetAttribute(&g_dev[device].khz, cudaDevAttrClockRate, device);
        g_dev[device].ready = true;
    }
    return SC_OK;
}

// ---- allocation cache ---------------------------------------------------------------------------------------------
// cudaMalloc / cudaHostAlloc / cudaStreamCreate cost 0.1-10 ms each and vary wildly from call to call (measured: a
// dim-18 GKR proof, which creates two provers, took 8.6 to 75 ms end to end on the same box).  Handles therefore return
// their device slab (up to 256 MiB), their pinned+mapped result block and their stream to a small per-device cache that
// the next handle of a similar size reuses.  sc_release_cached_memory() empties it.
struct CachedBlock {
    void* p;
    size_t bytes;
};
struct AllocCache {
    std::vector<CachedBlock> dev, host;
    std::vector<cudaStream_t> streams;
};
AllocCache g_cache[64];
std::mutex g_cache_mu;
// resident rounds: host <-> device words (see resident_kernel.cuh); byte offsets inside the mapped block
constexpr size_t RES_OFF_CONSTS = 0, RES_OFF_SUMS = 512, RES_OFF_ERROR = 1536, RES_OFF_ABORT = 1600, RES_HOST_BYTES = 1792;
constexpr size_t RES_BCAST_BYTES = 64 * 8 + 64;
constexpr unsigned long long RES_MAX_PAIRS_DEFAULT = 1ull << 16;
constexpr uint32_t EAGER_CHUNKS = 8;
constexpr uint32_t EAGER_SLOT_WORDS = sck::MAX_NPTS * 8 + 8;  // raw sums, then the flag word
constexpr size_t CACHE_MAX_ENTRIES = 96;  // a 16-layer GKR batch holds 16 instance slabs, 32 prover slabs, 32 pinned blocks and 32 streams at once
// Largest block the cache keeps and the most it holds per device.  The nv = 24 one-shot proof (MLSumcheck::prove: create +
// prove + destroy) allocates a 1.5 GiB and a 1.1 GiB slab: with the first-round limit of 256 MiB both went through
// cudaMalloc/cudaFree on every call.  180 GB of HBM make 8 GiB of parked slabs cheap;
// SC_CACHE_MAX_MB=<n> changes the budget, sc_release_cached_memory() returns everything.
size_t cache_budget() {
    static const size_t v = [] {
        const char* e = getenv("SC_CACHE_MAX_MB");
        return (size_t)(e ? strtoull(e, nullptr, 10) : 8192) << 20;
    }();
    return v;
}

bool cache_enabled() {
    static const bool on = !getenv("SC_NO_ALLOC_CACHE");
    return on;
}

bool cache_take(std::vector<CachedBlock>& v, size_t bytes, void** out, size_t* got) {
    size_t best = v.size();
    for (size_t i = 0; i < v.size(); i++)
        if (v[i].bytes >= bytes && v[i].bytes <= bytes + bytes / 2 + 4096 && (best == v.size() || v[i].bytes < v[best].bytes)) best = i;
    if (best == v.size()) return false;
    *out = v[best].p;
    *got = v[best].bytes;
    v.erase(v.begin() + best);
    return true;
}
cudaError_t device_alloc(void** out, size_t bytes, size_t* got, int device) {
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        if (cache_take(g_cache[device & 63].dev, bytes, out, got)) return cudaSuccess;
    }
    *got = bytes;
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e == cudaErrorMemoryAllocation) {  // parked slabs may be what is in the way: give them back and retry once
        cudaGetLastError();
        sc_release_cached_memory();
        e = cudaMalloc(out, bytes ? bytes : 1);
    }
    return e;
}
void device_free(void* p, size_t bytes, int device) {
    if (!p) return;
    if (cache_enabled()) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto& v = g_cache[device & 63].dev;
        size_t held = 0;
        for (auto& b : v) held += b.bytes;
        if (v.size() < CACHE_MAX_ENTRIES && held + bytes <= cache_budget()) { v.push_back({p, bytes}); return; }
    }
    cudaFree(p);
}
cudaError_t host_mapped_alloc(void** out, size_t bytes, size_t* got, int device) {
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        if (cache_take(g_cache[device & 63].host, bytes, out, got)) return cudaSuccess;
    }
    *got = bytes;
    return cudaHostAlloc(out, bytes, cudaHostAllocMapped);
}
void host_mapped_free(void* p, size_t bytes, int device) {
    if (!p) return;
    if (cache_enabled()) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto& v = g_cache[device & 63].host;
        if (v.size() < CACHE_MAX_ENTRIES) { v.push_back({p, bytes}); return; }
    }
    cudaFreeHost(p);
}
cudaError_t stream_acquire(cudaStream_t* out, int device) {
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto& v = g_cache[device & 63].streams;
        if (!v.empty()) { *out = v.back(); v.pop_back(); return cudaSuccess; }
    }
    return cudaStreamCreateWithFlags(out, cudaStreamNonBlocking);
}
void stream_release(cudaStream_t s, int device) {  // the caller has synchronised it
    if (!s) return;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto& v = g_cache[device & 63].streams;
        if (v.size() < CACHE_MAX_ENTRIES && cache_enabled()) { v.push_back(s); return; }
    }
    cudaStreamDestroy(s);
}

}  // namespace

// ================================================================================================ prover handle
struct sc_prover {
    int device = 0;
    uint32_t nv = 0, T = 0, n_products = 0, d = 0, round = 0;
    uint64_t N = 0;
    bool owns_tab0 = true;
    bool is_shard = false;  // one rank's shard of a sharded polynomial (capi_multi.inc); set before the first upload
    std::vector<uint64_t> randomness;  // ProverState.randomness, 4 u64 each
    std::vector<uint32_t*> tab0, bufA, bufB;  // per-table device pointers
    uint32_t *slab0 = nullptr, *slabA = nullptr, *slabB = nullptr;
    size_t slab0_bytes = 0, slabA_bytes = 0, h_result_bytes = 0;  // as handed out by the allocation cache
    uint32_t** d_ptr0 = nullptr;  // device arrays of table pointers
    uint32_t** d_ptrA = nullptr;
    uint32_t** d_ptrB = nullptr;
    uint32_t *d_offsets = nullptr, *d_indices = nullptr, *d_coeffs = nullptr;
    uint8_t* d_first = nullptr;
    uint32_t *d_partials = nullptr, *d_evals = nullptr, *d_canon = nullptr, *d_lagrange = nullptr;
    unsigned int* d_counter = nullptr;
    uint32_t *h_evals = nullptr, *h_canon = nullptr;  // pinned + mapped: [evals | canon | flag]
    uint32_t* h_result = nullptr;   // one allocation: (d+1)*8 evals, (d+1)*8 canon, then the flag word
    uint32_t* d_result = nullptr;   // device alias of h_result
    uint32_t seq = 0;               // flag value of the last issued round
    bool want_timing = false;       // record per-round CUDA events (sc_prover_set_timing)
    // fused tail (device transcript): [nv][d+1][8] evals, [nv][8] challenges, transcript state in/out
    uint32_t *d_tail_evals = nullptr, *d_tail_chal = nullptr, *h_tail = nullptr;
    b2::State *d_st = nullptr, *h_st = nullptr;
    int max_grid = 0;
    // TMA descriptors of every table in each of the three buffers (tab0, A, B): [3][T] CUtensorMap in device memory;
    // tc_buf_ok[c]: rounds reading buffer c may use the TMA + tensor-core fold kernel (tc_round.cuh)
    uint8_t* d_maps = nullptr;
    std::vector<CUtensorMap> h_maps;  // staging copy (kept alive: uploaded asynchronously)
    bool tc_buf_ok[3] = {false, false, false};
    bool r1_ok = false;  // [3][T..2T): the pristine tables again with 64-row boxes, for round1_tma_kernel
    unsigned long long tc_min_pairs = 0;
    int cur = 0;  // which buffer holds the current tables: 0 = tab0, 1 = A, 2 = B
    cudaStream_t stream = nullptr;      // the stream work is issued on
    cudaStream_t own_stream = nullptr;  // created by the handle
    std::vector<cudaEvent_t> ev;  // 2 per round
    std::vector<float> round_ms;
    bool timing = false;
    bool used_skip1 = false;  // last device round summed t = 0, 2, .., d only
    // host_post: rounds deliver only their raw sums; coefficient, claim and canonical forms are finished on the host
    bool host_post = false, raw_active = false;
    bool alt_active = false;  // the raw sums just delivered are at the alternative points 0, 1, inf, -1, 2, -2 (kernels.cuh ALT)
    // Pipelined upload (sc_prover_load_tables): the tables arrive in EAGER_CHUNKS pieces on a copy stream while round 1 —
    // which needs no challenge — is summed piece by piece behind them; the first prove_round then only adds the pieces up.
    bool eager_valid = false;
    bool eager_gemm = false;  // the chunks were summed by the contraction kernel: six integers per chunk in h_gemm slots 1..
    uint32_t eager_epoch = 0;
    uint32_t* h_eager = nullptr;   // pinned+mapped: [EAGER_CHUNKS][(MAX_NPTS * 8) sums + 8 (flag word first)]
    uint32_t* d_eager = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t eager_ev[8] = {};
    uint32_t raw_npts = 0;
    std::vector<uint64_t> h_prev;  // the previous round's ProverMsg (d+1 elements), for the claim
    bool direct_results = true;  // rounds deliver their message through mapped host memory + flag
    bool direct_active = false;  // ... and the round just issued did so
    bool exchange = false;       // sharded round: fuse the partial-sum exchange into the round kernel
    uint64_t launches = 0, tc_rounds = 0, res_rounds = 0, gemm_rounds = 0;
    // Tensor-core contraction rounds (gemm_sum.cuh): products of three tables; the sums arrive as big integers in h_gemm
    bool gemm_shape = false;     // the list of products has the shape those kernels serve
    uint32_t gemm_m = 0;         // ... every product has gemm_m multiplicands (3 or 4)
    bool gemm_r1_ok = false;     // ... and the 64-byte-row descriptors of the pristine tables exist (round 1)
    bool gemm_active = false;    // the round just issued delivers through h_gemm
    uint32_t gemm_seq = 0, gemm_want = 0;  // flag values of the contraction launches: last handed out / the one the current round publishes
    // a fold round launched ahead of its challenge (run_rounds only): its 1-based round number (0 = none) and its flag value
    uint32_t gemm_pre_round = 0, gemm_pre_want = 0;
    bool prelaunch_ok = false;   // inside a whole-proof call, where the next round always follows
    uint8_t* d_ymaps = nullptr;  // [T] CUtensorMap: tab0 as rows of one pair (64 bytes), SWIZZLE_64B, 128-row boxes
    std::vector<CUtensorMap> h_ymaps;
    unsigned long long* d_gemm_totals = nullptr;
    uint32_t *h_gemm = nullptr, *d_gemm = nullptr;  // mapped: [1 + EAGER_CHUNKS] slots of gsum::OUT_SLOT_WORDS
    // where the d+1 results of the last round live on the device (local buffers, the summed copies, or the sub-prover's)
    uint32_t *out_evals = nullptr, *out_canon = nullptr;
    // ---- multi-GPU (capi_multi.inc): nv is GLOBAL, nv_local = nv - log2(ranks) is what this rank's shard spans
    uint32_t nv_local = 0;
    struct sc_comm* comm = nullptr;
    sc_prover* sub = nullptr;  // replicated prover for the last rounds
    sc_prover* wait_on = nullptr;  // whose mapped result block the round just issued will signal (this or sub)
    uint32_t switch_round = 0;     // first global round run replicated
    bool switched = false;         // ... and this proof has gathered the tables for it
    bool collect_pending = false;  // prove_round_issue launched a round whose message prove_round_collect has yet to fetch
    bool host_done = false;        // the round just run left its message in h_evals / h_canon (host-side exchange)
    uint32_t** d_peer_tabs = nullptr;  // [3][n_ranks][T] every rank's table pointers in each buffer, as seen from this device
    std::vector<void*> ipc_opened;     // peer slabs mapped through CUDA IPC (multi-process)
    // ---- single-process multi-GPU facade (sc_prover_create_multi): the shards, one comm and one worker thread per rank
    std::vector<sc_prover*> group;
    std::vector<struct sc_comm*> group_comms;
    struct MultiWorkers* workers = nullptr;
    uint32_t *d_gather = nullptr, *d_evals_g = nullptr, *d_canon_g = nullptr, *d_sub_tabs = nullptr;
    std::vector<uint64_t> h_coeffs;
    // Pre-applied coefficients: h_scaled[k] = 1 when product k's coefficient lives in table scaled_table[k] (a table only
    // that product uses, scaled once at prover_init / load_tables); sc_prover_table divides it out again on export.
    std::vector<uint8_t> h_scaled;
    std::vector<int> scaled_table;
    uint8_t* d_scaled = nullptr;
    std::vector<uint32_t> h_offsets, h_indices;
    size_t h_nnz = 0;                  // offsets[n_products]
    std::vector<uint64_t> h_lagrange;  // staging copy of d_lagrange (uploaded asynchronously)
    // Resident rounds (resident_kernel.cuh): set by run_rounds for the proof in flight
    uint32_t res_first = 0;            // first (1-based) round served by the resident kernel; 0 = none
    bool res_running = false;          // the kernel is on the GPU, waiting for fold constants
    int res_share = 1;                 // independent provers whose resident kernels are on this device at the same time (batches)
    uint32_t res_seq0 = 0, res_last_seq = 0;  // sequence numbers of its first and last round
    unsigned long long res_max_pairs = 0;
    uint32_t *h_res = nullptr, *d_res = nullptr;  // mapped block: [64 x {limb,seq}] constants | [40 x {limb,seq}] sums | error | abort
    uint32_t* d_res_bcast = nullptr;   // device: [64 x {limb,seq}] + abort word
    unsigned int* d_res_counters = nullptr;
    long long* d_res_prof = nullptr;   // SC_RES_PROF=1: per-round cycle counts of CTA 0
    double res_host_us[64][3] = {};
    double res_t0 = 0, res_t1 = 0;     // (profiling) time stamps between resident_post and resident_collect    // ... and host-side microseconds per resident round: constants out, wait, finish
    void* adopted = nullptr;           // a device block whose ownership was handed to this handle (freed on destroy)
    size_t adopted_bytes = 0;
};

namespace {

// [w_0..w_d | 0..d] as Montgomery elements: w_j = 1 / prod_{k != j} (j - k) (kernels.cuh claim_from_prev).  The first
// version ran a 256-step square-and-multiply per lane on one warp at every prover_init (195 us).
std::vector<uint64_t> lagrange_block(uint32_t d) {
    const std::vector<hfr::F>& w = hfr::lagrange_weights(d);
    std::vector<uint64_t> out((size_t)2 * (d + 1) * 4);
    for (uint32_t j = 0; j <= d; j++) {
        memcpy(&out[(size_t)j * 4], &w[j], 32);
        const hfr::F fj = hfr::from_u64(j);
        memcpy(&out[(size_t)(d + 1 + j) * 4], &fj, 32);
    }
    return out;
}

template <int NPTS>
int occupancy_blocks_nofold() {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sck::round_kernel<NPTS, false>, sck::ROUND_THREADS, 0);
    return nb < 1 ? 1 : nb;
}

template <int NPTS>
cudaError_t launch_round(sc_prover* p, bool fold, const sck::RoundParams& rp) {
    const int threads = fold ? sck::fold_round_threads() : sck::ROUND_THREADS;
    unsigned long long need = (rp.n_pairs + threads - 1) / threads;
    int occ = fold ? sck::fold_round_occupancy(NPTS) : occupancy_blocks_nofold<NPTS>();
    unsigned long long cap = (unsigned long long)g_dev[p->device].sms * occ;
    if (cap > (unsigned long long)p->max_grid) cap = p->max_grid;
    int grid = (int)(need < cap ? need : cap);
    if (grid < 1) grid = 1;
    p->launches++;
    if (fold) return sck::launch_fold_round(NPTS, grid, rp, p->stream);  // compact translation unit (tail.cu)
    sck::round_kernel<NPTS, false><<<grid, threads, 0, p->stream>>>(rp);
    return cudaGetLastError();
}

// Round 1 on the TMA-staged kernel (tma_round1.cuh).  Resident CTAs per SM from the function attributes (registers,
// shared memory): see tail.cu tc_prepare for why the occupancy API is not used.
template <int NPTS, int M = 0>
cudaError_t launch_round1_tma_m(sc_prover* p, const sck::RoundParams& rp) {
    static bool ready_dev[64] = {};
    static int blocks_dev[64] = {};
    const int dev = p->device & 63;
    if (!ready_dev[dev]) {
        cudaError_t e = cudaFuncSetAttribute(sck::round1_tma_kernel<NPTS, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sck::R1_DYN_SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(sck::round1_tma_kernel<NPTS, M>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        cudaFuncAttributes fa;
        e = cudaFuncGetAttributes(&fa, sck::round1_tma_kernel<NPTS, M>);
        if (e != cudaSuccess) return e;
        int regs_sm = 0, smem_sm = 0;
        cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, p->device);
        cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, p->device);
        const int regs_cta = ((fa.numRegs + 7) / 8 * 8) * (int)sck::R1_THREADS;
        const int smem_cta = (int)fa.sharedSizeBytes + (int)sck::R1_DYN_SMEM + 1024;
        int blocks = regs_sm / regs_cta;
        if (smem_sm / smem_cta < blocks) blocks = smem_sm / smem_cta;
        if (blocks < 1) blocks = 1;
        if (getenv("SC_DEBUG")) fprintf(stderr, "round1_tma_kernel<%d,%d>: %d CTAs/SM (regs %d)\n", NPTS, M, blocks, fa.numRegs);
        blocks_dev[dev] = blocks;
        ready_dev[dev] = true;
    }
    const unsigned long long n_tiles = rp.n_pairs / sck::R1_THREADS;
    unsigned long long cap = (unsigned long long)g_dev[p->device].sms * blocks_dev[dev];
    if (cap > (unsigned long long)p->max_grid) cap = p->max_grid;
    const int grid = (int)(n_tiles < cap ? n_tiles : cap);
    p->launches++;
    sck::round1_tma_kernel<NPTS, M><<<grid, sck::R1_THREADS, sck::R1_DYN_SMEM, p->stream>>>(rp);
    return cudaGetLastError();
}

// One product of M = NPTS - 1 multiplicands with a deferred coefficient (the shape of BASELINE configs 2, 3 and of both GKR
// phases): the build with the loops over products and multiplicands unrolled.  Anything else: the CSR-driven kernel.
bool single_product_shape(const sc_prover* p, const sck::RoundParams& rp, uint32_t m) {
    static const bool off = getenv("SC_NO_SPECIALISE") != nullptr;
    return !off && p->n_products == 1 && rp.defer_coeff && !rp.prod_scaled && rp.t0 == 0 && p->h_nnz == m && p->d == m;
}
// (Measured at nv = 24, degree 3: the unrolled build helps the tensor-core fold kernel by 1-3 % — round 2 0.697 -> 0.678 ms — but
// makes round 1 MUCH slower, 1.09 -> 1.68 ms: its fully inlined body, three times as long once unrolled, no longer fits the
// instruction cache.  Round 1 therefore always runs the CSR-driven build.)
template <int NPTS>
cudaError_t launch_round1_tma(sc_prover* p, const sck::RoundParams& rp) {
    return launch_round1_tma_m<NPTS, 0>(p, rp);
}

void set_exchange_params(sc_prover* p, sck::RoundParams& rp);  // capi_multi.inc
bool comm_fused_exchange_ok(const sc_prover* p);  // the sharded rounds of this handle run the fused peer-memory exchange
int comm_device_share(const sc_prover* p);        // ranks of the group that share this rank's device
bool comm_in_process(const sc_prover* p);         // the ranks are host threads of ONE process (sc_prover_create_multi)
uint32_t set_exchange_params_resident(sc_prover* p, sck::RoundParams& rp, uint32_t n_rounds);
bool comm_failed(const sc_prover* p);
void comm_clear_error(sc_prover* p);
// single-process multi-GPU facade (capi_multi.inc)
void multi_destroy(sc_prover* P);
int multi_reset(sc_prover* P);
int multi_load_tables(sc_prover* P, const uint64_t* const* tables);
int multi_prove_round(sc_prover* P, const uint64_t* r_or_null, uint64_t* evals_out);
int multi_ml_prove(sc_prover* P, sc_blake2b512_rng* rng, uint64_t* evals_out, uint64_t* randomness_out);
int multi_table(const sc_prover* P, uint32_t j, uint64_t* out, uint64_t cap_elems, uint64_t* len_out);
inline const sc_prover* lead(const sc_prover* p) { return (p && !p->group.empty()) ? p->group[0] : p; }

// ---- tensor-core contraction rounds (gemm_sum.cuh) --------------------------------------------------------------------------------
inline uint32_t gemm_limbs(const sc_prover* p) {
    return p->gemm_m == 4 ? gsum::Shape<4>::OUT_LIMBS : (p->gemm_m == 2 ? gsum::Shape<2>::OUT_LIMBS : gsum::Shape<3>::OUT_LIMBS);
}
inline uint32_t gemm_nb(const sc_prover* p) { return p->gemm_m == 4 ? gsum::Shape<4>::NB : (p->gemm_m == 2 ? gsum::Shape<2>::NB : gsum::Shape<3>::NB); }

// May this round (n_pairs output pairs; fold = rounds >= 2) run on the contraction kernels?
bool gemm_round_ok(const sc_prover* p, unsigned long long n_pairs, bool fold) {
    if (!p->gemm_shape || !p->host_post || !p->direct_results || p->n_products > gsum::MAX_PRODUCTS) return false;
    // a shard of a sharded polynomial: only when this round runs the fused peer-memory exchange (the last CTA then exchanges
    // the six integers with the other ranks before it publishes)
    if ((p->comm || p->is_shard) && !(p->comm && p->exchange)) return false;
    if (p->n_products > 1)
        for (uint32_t k = 0; k < p->n_products; k++)
            if (!p->h_scaled[k]) return false;  // a coefficient that is not inside a table would have to be applied per product
    if (n_pairs < p->tc_min_pairs || n_pairs % gsum::TILE) return false;
    return fold ? p->tc_buf_ok[p->cur] : (p->gemm_r1_ok && p->cur == 0);
}

// Launch the round over the tiles [tile0, tile0 + n_tiles) (several launches when one could overflow the s32 accumulators); the
// last launch publishes the six integers into mapped slot `slot` of h_gemm and then `seq` into the slot's last word.
constexpr uint32_t GEMM_RMAIL_WORD = 480, GEMM_RERROR_WORD = 496;  // inside slot 0 of h_gemm (results use < 320 words, the flag the last one)

int launch_gemm_round(sc_prover* p, const sck::RoundParams& rp, bool fold, unsigned long long tile0, unsigned long long n_tiles, uint32_t slot, uint32_t seq,
                      bool wait_r = false) {
    gsum::Params G;
    memset(&G, 0, sizeof(G));
    G.rp = rp;
    if (wait_r) {
        G.r_mail = (const unsigned long long*)(p->d_gemm + GEMM_RMAIL_WORD);
        G.r_bcast = p->d_gemm_totals + (size_t)9 * gsum::TOT_STRIDE;
        G.r_seq = seq;
        G.r_error = p->d_gemm + GEMM_RERROR_WORD;
        const int khz = g_dev[p->device].khz;
        static const double secs = getenv("SC_RES_TIMEOUT_S") ? atof(getenv("SC_RES_TIMEOUT_S")) : 10.0;
        G.r_timeout = (long long)(secs * (khz > 0 ? khz : 1965000) * 1000.0);
    }
    G.ymaps = p->d_ymaps;
    G.totals = p->d_gemm_totals;
    G.rp.host_out = p->d_gemm + (size_t)slot * gsum::OUT_SLOT_WORDS;
    G.rp.host_flag = G.rp.host_out + gsum::OUT_SLOT_WORDS - 1;
    G.rp.seq = seq;
    const int sms = g_dev[p->device].sms;
    const int mm = (int)p->gemm_m;
    unsigned long long cap = (fold ? gsum::max_items_fold(sms, mm) : gsum::max_items_round1(sms, mm)) / p->n_products;  // tiles per launch
    if (const char* env = getenv("SC_GEMM_MAX_TILES")) {  // tests: force the several-launches-per-round path at small sizes
        const unsigned long long v = strtoull(env, nullptr, 10);
        if (v >= 1 && v < cap) cap = v;
    }
    if (cap < 1) cap = 1;
    for (unsigned long long t = 0; t < n_tiles; t += cap) {
        const unsigned long long take = n_tiles - t < cap ? n_tiles - t : cap;
        G.rp.tile_base = (uint32_t)(tile0 + t);
        G.items = (uint32_t)(take * p->n_products);
        G.publish = (t + take == n_tiles) ? 1u : 0u;
        p->launches++;
        static const bool prof = getenv("SC_GEMM_PROF") != nullptr;  // debugging aid: wait cycles per role, printed per launch
        long long* d_prof = nullptr;
        if (prof) {
            cudaMalloc(&d_prof, 16 * sizeof(long long));
            cudaMemsetAsync(d_prof, 0, 16 * sizeof(long long), p->stream);
            G.prof = d_prof;
        }
        cudaError_t e = fold ? gsum::launch_fold(G, mm, sms, p->stream) : gsum::launch_round1(G, mm, sms, p->stream);
        if (e != cudaSuccess) return fail(SC_ERR_CUDA, "contraction kernel launch: %s", cudaGetErrorString(e));
        if (prof) {
            long long h[16];
            cudaStreamSynchronize(p->stream);
            cudaMemcpy(h, d_prof, sizeof(h), cudaMemcpyDeviceToHost);
            cudaFree(d_prof);
            const double ctas = (double)((G.items + 2) / 3 < (uint32_t)sms ? (G.items + 2) / 3 : (uint32_t)sms), cw = ctas * 12;
            fprintf(stderr, "gemm %s items %u: per compute warp: total %.0f, wait acc_full %.0f, wait x_empty %.0f | fold warp per CTA: total %.0f, wait slot_full %.0f, acc_empty %.0f, mma issue + commit %.0f | TMA warp: total %.0f, wait slot_empty %.0f | sum warp: wait x_full %.0f, issue %.0f | cycles since CTA start (max): prologue %lld, loop %lld, totals %lld, published %lld\n",
                    fold ? "fold" : "round1", G.items, h[2] / cw, h[0] / cw, h[1] / cw, h[6] / ctas, h[3] / ctas, h[4] / ctas, h[8] / ctas, h[11] / ctas, h[5] / ctas, h[7] / ctas, h[10] / ctas, h[12], h[13], h[14], h[15]);
        }
    }
    return SC_OK;
}

// The challenge of a fold round that was launched ahead: eight {limb, sequence number} words, each ONE aligned 8-byte store
void gemm_send_challenge(sc_prover* p, const uint64_t* r /* null: zeros (abandoned proof) */, uint32_t seq) {
    uint64_t* m = (uint64_t*)(p->h_gemm + GEMM_RMAIL_WORD);
    for (int k = 0; k < 4; k++) {
        const uint64_t limb = r ? r[k] : 0;
        __atomic_store_n(m + 2 * k, (uint64_t)(uint32_t)limb | ((uint64_t)seq << 32), __ATOMIC_RELAXED);
        __atomic_store_n(m + 2 * k + 1, (limb >> 32) | ((uint64_t)seq << 32), __ATOMIC_RELAXED);
    }
    __sync_synchronize();
}

// An abandoned proof (error paths, reset, destroy): release a kernel that still waits for its challenge and let it drain
void gemm_abort_prelaunch(sc_prover* p) {
    if (!p->gemm_pre_round) return;
    gemm_send_challenge(p, nullptr, p->gemm_pre_want);
    cudaSetDevice(p->device);
    cudaStreamSynchronize(p->stream);
    p->gemm_pre_round = 0;
}

// Right after round p->round has been issued (p->cur names the buffer it writes): if the NEXT round is a contraction fold round
// with a launch of its own, launch it now — it starts when this round's kernel ends, runs its prologue and stages its first table
// tiles while the host is still finishing and hashing this round's message, and picks the challenge up from mapped memory.
// Saves the launch latency, the prologue and the first TMA round trip of every large fold round (~8 us each).
int gemm_prelaunch(sc_prover* p) {
    static const bool off = getenv("SC_NO_PRELAUNCH") != nullptr || profiler_attached();
    if (off || !p->prelaunch_ok || (p->timing && p->want_timing) || p->gemm_pre_round) return SC_OK;
    const uint32_t i2 = p->round + 1;
    if (i2 < 2 || i2 > p->nv_local || (p->res_first && i2 >= p->res_first)) return SC_OK;
    const unsigned long long n_pairs = (unsigned long long)1 << (p->nv_local - i2);
    {   // the build that can wait for its challenge runs its main loop ~3 % slower (gemm_sum.cuh): only worth it where that is less than
        // the ~8 us a launch ahead saves, i.e. for rounds of up to 2^21 pairs x 3 tables
        static const unsigned long long cap = getenv("SC_PRELAUNCH_MAX_PAIRS") ? strtoull(getenv("SC_PRELAUNCH_MAX_PAIRS"), nullptr, 10) : (3ull << 21);
        if (n_pairs * p->h_nnz > cap) return SC_OK;
    }
    if (p->comm) {
        // a shard: only sharded rounds with the fused exchange, and only when this rank has its GPU to itself — a kernel that
        // waits for its challenge holds every SM, and a rank sharing the device could then never finish the round it waits for
        // (and only with one process per GPU: with the ranks as threads of one process a proof now and then ran into the exchange
        // time-out — seen once in ~10 runs on 2 GPUs, never with separate processes; CUDA calls of different threads share locks, and a
        // kernel that waits for its host thread while that thread's next call waits behind another rank's is one deadlock too many)
        if (!comm_fused_exchange_ok(p) || comm_device_share(p) > 1 || comm_in_process(p) || i2 >= p->switch_round) return SC_OK;
        p->exchange = true;
        const bool ok = gemm_round_ok(p, n_pairs, true);
        p->exchange = false;
        if (!ok) return SC_OK;
    } else if (p->is_shard || !gemm_round_ok(p, n_pairs, true)) {
        return SC_OK;
    }
    sck::RoundParams rp;
    memset(&rp, 0, sizeof(rp));
    uint32_t** in = p->cur == 0 ? p->d_ptr0 : (p->cur == 1 ? p->d_ptrA : p->d_ptrB);
    uint32_t** out = p->cur == 1 ? p->d_ptrB : p->d_ptrA;
    rp.tab_in = (const uint32_t* const*)in;
    rp.tab_out = out;
    rp.prod_offsets = p->d_offsets; rp.prod_indices = p->d_indices; rp.prod_first = p->d_first; rp.coeffs = p->d_coeffs;
    rp.prod_scaled = p->d_scaled;
    rp.n_products = p->n_products; rp.n_tables = p->T; rp.defer_coeff = (p->n_products == 1) ? 1u : 0u;
    rp.n_pairs = n_pairs;
    rp.counter = p->d_counter;
    rp.degree = p->d;
    rp.write_fold = 1;
    rp.tmaps = p->d_maps + (size_t)p->cur * p->T * sizeof(CUtensorMap);
    if (p->comm) set_exchange_params(p, rp);  // mailbox slot and sequence number of that round (handed out in launch order on every rank)
    const uint32_t want = ++p->gemm_seq;
    int rc = launch_gemm_round(p, rp, true, 0, n_pairs / gsum::TILE, 0, want, true);
    if (rc) return rc;
    p->gemm_pre_round = i2;
    p->gemm_pre_want = want;
    return SC_OK;
}

// One protocol round on the device: (fold on r) + sums for all d+1 points.  Results land in d_evals / d_canon.
int run_round_device(sc_prover* p, const uint64_t* r /* null in round 1 */) {
    const uint32_t i = p->round;  // already incremented: 1-based round being computed
    const bool fold = (i >= 2);
    sck::RoundParams rp;
    memset(&rp, 0, sizeof(rp));
    uint32_t** in = p->cur == 0 ? p->d_ptr0 : (p->cur == 1 ? p->d_ptrA : p->d_ptrB);
    int next = p->cur;
    if (fold) next = (p->cur == 1) ? 2 : 1;
    uint32_t** out = next == 1 ? p->d_ptrA : p->d_ptrB;
    if (p->gemm_pre_round) {
        if (p->gemm_pre_round != i || !fold) {  // (cannot happen inside run_rounds; a stale kernel must not outlive its proof)
            gemm_abort_prelaunch(p);
        } else {
            // this round's kernel is already on the GPU (gemm_prelaunch): all that is left to do is to send the challenge
            gemm_send_challenge(p, r, p->gemm_pre_want);
            p->gemm_want = p->gemm_pre_want;
            p->gemm_pre_round = 0;
            p->direct_active = true;
            p->raw_active = true;
            p->alt_active = false;
            p->gemm_active = true;
            p->used_skip1 = false;
            p->gemm_rounds++;
            p->tc_rounds++;
            p->cur = next;
            return SC_OK;
        }
    }
    rp.tab_in = (const uint32_t* const*)in;
    rp.tab_out = out;
    rp.prod_offsets = p->d_offsets;
    rp.prod_indices = p->d_indices;
    rp.prod_first = p->d_first;
    rp.coeffs = p->d_coeffs;
    rp.prod_scaled = p->d_scaled;
    rp.n_products = p->n_products;
    rp.n_tables = p->T;
    rp.defer_coeff = (p->n_products == 1) ? 1u : 0u;
    rp.n_pairs = (unsigned long long)1 << (p->nv_local - i);
    if (fold) memcpy(rp.r, r, 32);
    rp.partials = p->d_partials;
    rp.counter = p->d_counter;
    rp.evals_out = p->d_evals;
    rp.canon_out = p->d_canon;
    rp.degree = p->d;
    rp.prev_evals = p->d_evals;
    rp.lagrange = p->d_lagrange;
    const bool direct = (!p->comm || p->exchange) && p->direct_results;  // results straight into mapped host memory
    p->direct_active = direct;
    if (p->exchange) {  // sharded + fused exchange: the kernel's outputs are the GLOBAL message; keep it where the NCCL path does
        set_exchange_params(p, rp);
        rp.evals_out = p->d_evals_g;
        rp.canon_out = p->d_canon_g;
        rp.prev_evals = p->d_evals_g;
    }
    if (direct) {
        rp.host_out = p->d_result;
        rp.seq = ++p->seq;
    }
    p->raw_active = false;
    p->alt_active = false;
    p->gemm_active = false;
    if (direct && gemm_round_ok(p, rp.n_pairs, fold)) {
        // products of three tables: the per-pair work shrinks to three plain products, the sum runs on the tensor cores and the
        // host turns the six integers it receives into P(0..d) (host_finish_round)
        rp.write_fold = 1;
        rp.tmaps = fold ? p->d_maps + (size_t)p->cur * p->T * sizeof(CUtensorMap) : nullptr;
        p->gemm_want = ++p->gemm_seq;
        int rc = launch_gemm_round(p, rp, fold, 0, rp.n_pairs / gsum::TILE, 0, p->gemm_want);
        if (rc) return rc;
        p->gemm_active = true;
        p->raw_active = true;
        p->used_skip1 = false;
        p->gemm_rounds++;
        if (fold) p->tc_rounds++;
        p->cur = next;
        return SC_OK;
    }
    if (fold && p->d <= (uint32_t)sck::MAX_NPTS && p->d_lagrange) {
        // rounds >= 2: P(0) + P(1) = P_prev(r) (the verifier's check, verifier.rs:109), so only t = 0, 2, .., d are summed
        rp.skip1 = 1;
        rp.fix1 = (p->comm && !p->exchange) ? 0u : 1u;  // NCCL path: the fix needs the GLOBAL P(0) -> after the all-gather
        rp.t0 = 0;
        rp.write_fold = 1;
        if (direct) rp.host_flag = p->d_result + (size_t)(p->d + 1) * 16;
        if (direct && p->host_post) {
            rp.raw_out = 1;
            p->raw_active = true;
            p->raw_npts = p->d;
        }
        cudaError_t e;
        if (p->tc_buf_ok[p->cur] && rp.n_pairs >= p->tc_min_pairs && rp.raw_out) {
            // large fold round: tables staged by TMA, fix_variables on the tensor cores (tc_round.cuh).  These kernels sum
            // at the alternative points, which only the host finishing converts back (raw delivery)
            p->alt_active = true;
            rp.tmaps = p->d_maps + (size_t)p->cur * p->T * sizeof(CUtensorMap);
            p->launches++;
            p->tc_rounds++;
            e = sck::launch_fold_round_tc(p->d, (single_product_shape(p, rp, 2) || single_product_shape(p, rp, 3)) ? p->d : 0, g_dev[p->device].sms,
                                          p->max_grid, rp, p->stream);
        } else
        switch (p->d) {
            case 1: e = launch_round<1>(p, true, rp); break;
            case 2: e = launch_round<2>(p, true, rp); break;
            case 3: e = launch_round<3>(p, true, rp); break;
            case 4: e = launch_round<4>(p, true, rp); break;
            default: e = launch_round<5>(p, true, rp); break;
        }
        if (e != cudaSuccess) return fail(SC_ERR_CUDA, "round kernel launch: %s", cudaGetErrorString(e));
        p->cur = next;
        p->used_skip1 = true;
        return SC_OK;
    }
    p->used_skip1 = false;
    uint32_t remaining = p->d + 1, t0 = 0;
    while (remaining > 0) {
        uint32_t take;
        if (remaining <= (uint32_t)sck::MAX_NPTS) take = remaining;
        else if (remaining == (uint32_t)sck::MAX_NPTS + 1) take = 3;
        else take = sck::MAX_NPTS;
        rp.t0 = t0;
        rp.write_fold = (t0 == 0) ? 1u : 0u;
        rp.host_flag = (direct && take == remaining) ? p->d_result + (size_t)(p->d + 1) * 16 : nullptr;
        rp.host_out = direct ? p->d_result : nullptr;
        if (direct && p->host_post && take == p->d + 1) {  // the whole message in one launch: raw delivery
            rp.raw_out = 1;
            p->raw_active = true;
            p->raw_npts = take;
        }
        cudaError_t e;
        if (!fold && p->r1_ok && p->cur == 0 && take == p->d + 1 && rp.n_pairs >= p->tc_min_pairs && rp.raw_out) {
            p->alt_active = true;  // round1_tma_kernel sums at the alternative points (raw delivery only)
            rp.tmaps = p->d_maps + (size_t)3 * p->T * sizeof(CUtensorMap);  // round 1, one launch: TMA-staged kernel
            switch (take) {
                case 1: e = launch_round1_tma<1>(p, rp); break;
                case 2: e = launch_round1_tma<2>(p, rp); break;
                case 3: e = launch_round1_tma<3>(p, rp); break;
                case 4: e = launch_round1_tma<4>(p, rp); break;
                default: e = launch_round1_tma<5>(p, rp); break;
            }
        } else
        switch (take) {
            case 1: e = launch_round<1>(p, fold, rp); break;
            case 2: e = launch_round<2>(p, fold, rp); break;
            case 3: e = launch_round<3>(p, fold, rp); break;
            case 4: e = launch_round<4>(p, fold, rp); break;
            default: e = launch_round<5>(p, fold, rp); break;
        }
        if (e != cudaSuccess) return fail(SC_ERR_CUDA, "round kernel launch: %s", cudaGetErrorString(e));
        t0 += take;
        remaining -= take;
    }
    p->cur = next;
    return SC_OK;
}

int validate_products(uint32_t n_tables, uint32_t n_products, const uint32_t* offsets, const uint32_t* indices, uint32_t* d_out) {
    if (n_products == 0 || n_tables == 0) return fail(SC_ERR_BAD_INPUT, "empty polynomial");
    if (!offsets || !indices) return fail(SC_ERR_BAD_INPUT, "null product list");
    if (offsets[0] != 0) return fail(SC_ERR_BAD_INPUT, "offsets[0] must be 0 (got %u)", offsets[0]);
    uint32_t d = 0;
    for (uint32_t k = 0; k < n_products; k++) {
        if (offsets[k + 1] <= offsets[k]) return fail(SC_ERR_BAD_INPUT, "product %u is empty (data_structures.rs:78)", k);
        uint32_t m = offsets[k + 1] - offsets[k];
        if (m > d) d = m;
        for (uint32_t j = offsets[k]; j < offsets[k + 1]; j++)
            if (indices[j] >= n_tables) return fail(SC_ERR_BAD_INPUT, "table index %u out of range", indices[j]);
    }
    *d_out = d;
    return SC_OK;
}

int prescale_tables(sc_prover* p);
int upload_tables(sc_prover* p, const uint64_t* const* tables);

int create_common(sc_prover** out, uint32_t nv, uint32_t T, const uint64_t* const* tables, bool tables_on_device,
                  uint32_t n_products, const uint64_t* coeffs, const uint32_t* offsets, const uint32_t* indices, int device,
                  const uint8_t* inherit_scaled = nullptr, bool shard = false) {
    if (!out) return fail(SC_ERR_BAD_INPUT, "null output handle");
    *out = nullptr;
    if (nv == 0) return fail(SC_ERR_PANIC_CONSTANT, "Attempt to prove a constant.");
    if (!tables || !coeffs) return fail(SC_ERR_BAD_INPUT, "null tables or coefficients");
    if (nv > 40) return fail(SC_ERR_BAD_INPUT, "nv = %u too large", nv);
    uint32_t d = 0;
    int rc = validate_products(T, n_products, offsets, indices, &d);
    if (rc) return rc;
    rc = ensure_device(device);
    if (rc) return rc;
    sc_prover* p = new sc_prover();
    p->is_shard = shard;
    p->device = device; p->nv = nv; p->nv_local = nv; p->T = T; p->n_products = n_products; p->d = d; p->N = (uint64_t)1 << nv;
    auto bail = [&](int code) { sc_prover_destroy(p); return code; };
#define TRY_P(expr)                                                                                        \
    do {                                                                                                   \
        cudaError_t e__ = (expr);                                                                          \
        if (e__ != cudaSuccess) return bail(fail(SC_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__))); \
    } while (0)
    TRY_P(stream_acquire(&p->own_stream, device));
    p->stream = p->own_stream;
    const size_t elem = 32, N = p->N;
    const size_t nA = N / 2 ? N / 2 : 1, nB = N / 4 ? N / 4 : 1;
    p->tab0.resize(T); p->bufA.resize(T); p->bufB.resize(T);
    if (tables_on_device) {
        p->owns_tab0 = false;
        for (uint32_t j = 0; j < T; j++) p->tab0[j] = (uint32_t*)tables[j];
    } else {
        TRY_P(device_alloc((void**)&p->slab0, (size_t)T * N * elem, &p->slab0_bytes, device));
        for (uint32_t j = 0; j < T; j++) p->tab0[j] = p->slab0 + (size_t)j * N * 8;  // filled by upload_tables below
    }
    // Two allocations for everything else (creation cost matters for one-shot proofs and the two GKR phases):
    // one device slab = ping-pong tables + all small arrays, one pinned+mapped host block.
    const uint32_t nnz = offsets[n_products];
    p->h_nnz = nnz;
    p->max_grid = g_dev[device].sms * 32;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += up(bytes ? bytes : 1); return o; };
    const size_t oA = take((size_t)T * nA * elem), oB = take((size_t)T * nB * elem);
    const size_t oP0 = take(T * sizeof(uint32_t*)), oPA = take(T * sizeof(uint32_t*)), oPB = take(T * sizeof(uint32_t*));
    const size_t oOff = take((n_products + 1) * sizeof(uint32_t)), oIdx = take(nnz * sizeof(uint32_t)), oFirst = take(nnz);
    const size_t oCoef = take((size_t)n_products * 32), oPart = take((size_t)p->max_grid * sck::MAX_NPTS * 32), oCnt = take(4);
    const size_t oEv = take((size_t)(d + 1) * 32), oCa = take((size_t)(d + 1) * 32), oLag = take((size_t)2 * (d + 1) * 32);
    const size_t oTe = take((size_t)nv * (d + 1) * 32), oTc = take((size_t)nv * 32), oSt = take(2 * sizeof(b2::State));
    const size_t oMaps = take((size_t)4 * T * sizeof(CUtensorMap));
    const size_t oScaled = take(n_products);
    const size_t oResB = take(RES_BCAST_BYTES), oResC = take(64 * sizeof(unsigned int));
    const size_t oYmaps = take((size_t)T * sizeof(CUtensorMap)), oGemmTot = take((size_t)9 * gsum::TOT_STRIDE * sizeof(unsigned long long) + 64);
    TRY_P(device_alloc((void**)&p->slabA, off, &p->slabA_bytes, device));
    uint8_t* base = (uint8_t*)p->slabA;
    for (uint32_t j = 0; j < T; j++) {
        p->bufA[j] = (uint32_t*)(base + oA) + (size_t)j * nA * 8;
        p->bufB[j] = (uint32_t*)(base + oB) + (size_t)j * nB * 8;
    }
    p->d_ptr0 = (uint32_t**)(base + oP0); p->d_ptrA = (uint32_t**)(base + oPA); p->d_ptrB = (uint32_t**)(base + oPB);
    p->d_offsets = (uint32_t*)(base + oOff); p->d_indices = (uint32_t*)(base + oIdx); p->d_first = base + oFirst;
    p->d_coeffs = (uint32_t*)(base + oCoef); p->d_partials = (uint32_t*)(base + oPart); p->d_counter = (unsigned int*)(base + oCnt);
    p->d_evals = (uint32_t*)(base + oEv); p->d_canon = (uint32_t*)(base + oCa);
    p->d_res_bcast = (uint32_t*)(base + oResB); p->d_res_counters = (unsigned int*)(base + oResC);
    TRY_P(cudaMemsetAsync(p->d_res_bcast, 0, RES_BCAST_BYTES, p->stream));  // a recycled slab may hold another handle's sequence numbers
    p->d_ymaps = base + oYmaps;
    p->d_gemm_totals = (unsigned long long*)(base + oGemmTot);
    TRY_P(cudaMemsetAsync(p->d_gemm_totals, 0, (size_t)9 * gsum::TOT_STRIDE * sizeof(unsigned long long) + 64, p->stream));  // + the 8 challenge words
    p->d_tail_evals = (uint32_t*)(base + oTe); p->d_tail_chal = (uint32_t*)(base + oTc); p->d_st = (b2::State*)(base + oSt);
    {
        // TMA descriptors (fold rounds with >= tc_min_pairs output pairs run on the TMA + tensor-core kernel)
        static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
        p->d_maps = base + oMaps;
        const char* env = getenv("SC_TC_MIN_PAIRS");
        p->tc_min_pairs = env ? strtoull(env, nullptr, 10) : sck::tc_min_pairs();
        if (p->tc_min_pairs < 128) p->tc_min_pairs = 128;
        std::vector<CUtensorMap>& maps = p->h_maps;
        maps.resize((size_t)4 * T);
        const size_t len[3] = {N, nA, nB};
        for (int c = 0; c < 3 && !getenv("SC_NO_TC"); c++) {
            const uint64_t rows = len[c] / 4;
            bool ok = rows >= 128;
            for (uint32_t j = 0; j < T && ok; j++) {
                const uint32_t* tb = c == 0 ? p->tab0[j] : (c == 1 ? p->bufA[j] : p->bufB[j]);
                ok = tmaph::make_table_map(&maps[(size_t)c * T + j], tb, rows, 128);
            }
            p->tc_buf_ok[c] = ok;
        }
        if (!getenv("SC_NO_TC") && !getenv("SC_NO_TMA_R1")) {  // round 1 reads 64-byte pairs: 64-row boxes = 128 pairs per tile
            bool ok = N / 4 >= sck::R1_TILE_ROWS;
            for (uint32_t j = 0; j < T && ok; j++) ok = tmaph::make_table_map(&maps[(size_t)3 * T + j], p->tab0[j], N / 4, sck::R1_TILE_ROWS);
            p->r1_ok = ok;
        }
        TRY_P(cudaMemcpyAsync(p->d_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, p->stream));
    }
    TRY_P(cudaMemcpyAsync(p->d_ptr0, p->tab0.data(), T * sizeof(uint32_t*), cudaMemcpyHostToDevice, p->stream));
    TRY_P(cudaMemcpyAsync(p->d_ptrA, p->bufA.data(), T * sizeof(uint32_t*), cudaMemcpyHostToDevice, p->stream));
    TRY_P(cudaMemcpyAsync(p->d_ptrB, p->bufB.data(), T * sizeof(uint32_t*), cudaMemcpyHostToDevice, p->stream));
    std::vector<uint8_t> first(nnz, 0);
    {
        std::vector<uint8_t> seen(T, 0);
        for (uint32_t j = 0; j < nnz; j++)
            if (!seen[indices[j]]) { seen[indices[j]] = 1; first[j] = 1; }
    }
    TRY_P(cudaMemcpyAsync(p->d_offsets, offsets, (n_products + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, p->stream));
    TRY_P(cudaMemcpyAsync(p->d_indices, indices, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, p->stream));
    TRY_P(cudaMemcpyAsync(p->d_first, first.data(), nnz, cudaMemcpyHostToDevice, p->stream));
    TRY_P(cudaMemcpyAsync(p->d_coeffs, coeffs, (size_t)n_products * 32, cudaMemcpyHostToDevice, p->stream));
    TRY_P(cudaMemsetAsync(p->d_counter, 0, sizeof(unsigned int), p->stream));
    {
        // Coefficient pre-scaling (several products only; a single product's coefficient is applied to the sums): pick for
        // every product a table that nothing else uses and multiply OUR copy of it by the coefficient once.
        p->h_scaled.assign(n_products, 0);
        p->scaled_table.assign(n_products, -1);
        if (p->h_coeffs.empty()) p->h_coeffs.assign(coeffs, coeffs + (size_t)n_products * 4);
        if (inherit_scaled) {  // replicated sub-prover of a sharded proof: its tables were gathered from scaled shards
            p->h_scaled.assign(inherit_scaled, inherit_scaled + n_products);
        } else if (!tables_on_device && n_products > 1 && !getenv("SC_NO_PRESCALE")) {
            std::vector<uint32_t> uses(T, 0);
            for (uint32_t j = 0; j < nnz; j++) uses[indices[j]]++;
            for (uint32_t k = 0; k < n_products; k++) {
                const uint64_t* c = coeffs + (size_t)k * 4;
                if ((c[0] | c[1] | c[2] | c[3]) == 0) continue;  // zero has no inverse to export the table with
                for (uint32_t j = offsets[k]; j < offsets[k + 1]; j++)
                    if (uses[indices[j]] == 1) { p->scaled_table[k] = (int)indices[j]; p->h_scaled[k] = 1; break; }
            }
        }
        bool any = false;
        for (uint8_t f : p->h_scaled) any = any || f;
        if (any) {
            p->d_scaled = base + oScaled;
            TRY_P(cudaMemcpyAsync(p->d_scaled, p->h_scaled.data(), n_products, cudaMemcpyHostToDevice, p->stream));
        }
    }
    {
        // Tensor-core contraction rounds (gemm_sum.cuh): every product has exactly two (d = 2), three (d = 3) or four (d = 4)
        // multiplicands; with several products every coefficient must already live in a table (checked per round: gemm_round_ok)
        bool shape = (d >= 2 && d <= 4) && !getenv("SC_NO_GEMM") && !getenv("SC_NO_TC") && !(d == 4 && getenv("SC_NO_GEMM4")) && !(d == 2 && getenv("SC_NO_GEMM2"));
        for (uint32_t k = 0; k < n_products && shape; k++) shape = offsets[k + 1] - offsets[k] == d;
        p->gemm_shape = shape;
        p->gemm_m = shape ? d : 0;
        if (shape && N / 2 >= gsum::TILE) {
            p->h_ymaps.resize(T);
            bool ok = true;  // (only three-table products read a table through these descriptors: the Y operand of round 1)
            for (uint32_t j = 0; j < T && ok; j++) ok = tmaph::make_pair_map(&p->h_ymaps[j], p->tab0[j], N / 2, gsum::TILE);
            p->gemm_r1_ok = ok;
            if (ok) TRY_P(cudaMemcpyAsync(p->d_ymaps, p->h_ymaps.data(), (size_t)T * sizeof(CUtensorMap), cudaMemcpyHostToDevice, p->stream));
        }
    }
    if (d + 1 <= 32) {  // Lagrange weights + the nodes 0..d for the P(1)-from-claim shortcut: host table (cached per degree)
        p->d_lagrange = (uint32_t*)(base + oLag);
        p->h_lagrange = lagrange_block(d);
        TRY_P(cudaMemcpyAsync(p->d_lagrange, p->h_lagrange.data(), (size_t)2 * (d + 1) * 32, cudaMemcpyHostToDevice, p->stream));
    }
    const size_t hRes = up((size_t)(d + 1) * 64 + 64), hTail = up((size_t)nv * (d + 2) * 32), hSt = up(2 * sizeof(b2::State));
    const size_t hEager = up((size_t)EAGER_CHUNKS * EAGER_SLOT_WORDS * 4);
    const size_t hGemm = up((size_t)(1 + EAGER_CHUNKS) * gsum::OUT_SLOT_WORDS * 4);
    TRY_P(host_mapped_alloc((void**)&p->h_result, hRes + hTail + hSt + hEager + up(RES_HOST_BYTES) + hGemm, &p->h_result_bytes, device));
    memset(p->h_result, 0, hRes);
    TRY_P(cudaHostGetDevicePointer((void**)&p->d_result, p->h_result, 0));
    p->h_evals = p->h_result;
    p->h_canon = p->h_result + (size_t)(d + 1) * 8;
    p->h_tail = (uint32_t*)((uint8_t*)p->h_result + hRes);
    p->h_st = (b2::State*)((uint8_t*)p->h_result + hRes + hTail);
    p->h_eager = (uint32_t*)((uint8_t*)p->h_result + hRes + hTail + hSt);
    p->d_eager = (uint32_t*)((uint8_t*)p->d_result + hRes + hTail + hSt);
    memset(p->h_eager, 0, hEager);
    p->h_res = (uint32_t*)((uint8_t*)p->h_result + hRes + hTail + hSt + hEager);
    p->d_res = (uint32_t*)((uint8_t*)p->d_result + hRes + hTail + hSt + hEager);
    memset(p->h_res, 0, RES_HOST_BYTES);  // recycled pinned blocks hold another handle's sequence numbers
    p->h_gemm = (uint32_t*)((uint8_t*)p->h_res + up(RES_HOST_BYTES));
    p->d_gemm = (uint32_t*)((uint8_t*)p->d_res + up(RES_HOST_BYTES));
    memset(p->h_gemm, 0, hGemm);
    {
        const char* env = getenv("SC_RES_MAX_PAIRS");
        p->res_max_pairs = (getenv("SC_NO_RESIDENT") || profiler_attached()) ? 0 : (env ? strtoull(env, nullptr, 10) : RES_MAX_PAIRS_DEFAULT);
    }
    p->host_post = !getenv("SC_TAIL") && !getenv("SC_NO_HOST_POST");
    p->h_prev.assign((size_t)(d + 1) * 4, 0);
    p->ev.assign(2 * (size_t)nv, nullptr);  // CUDA events are created on demand (sc_prover_set_timing)
    p->round_ms.assign(nv, 0.f);
    p->randomness.reserve((size_t)nv * 4);
    TRY_P(cudaStreamSynchronize(p->stream));
#undef TRY_P
    if (!tables_on_device) {
        // the deep copy of the caller's tables (prover.rs:55-59), pipelined with round 1 where the shape allows; returns when
        // the caller may free / modify its buffers
        int rc2 = upload_tables(p, tables);
        if (rc2) return bail(rc2);
    }
    *out = p;
    return SC_OK;
}

// tab0[scaled_table[k]] *= coeffs[k] for every pre-scaled product (after an upload of the pristine tables)
int prescale_tables(sc_prover* p) {
    if (!p->d_scaled || !p->owns_tab0) return SC_OK;
    for (uint32_t k = 0; k < p->n_products; k++) {
        const int j = p->scaled_table[k];
        if (!p->h_scaled[k] || j < 0) continue;
        unsigned long long need = (p->N + 127) / 128, cap = (unsigned long long)g_dev[p->device].sms * 16;
        const int grid = (int)(need < cap ? need : cap);
        sck::scale_kernel<<<grid < 1 ? 1 : grid, 128, 0, p->stream>>>(p->tab0[j], p->d_coeffs + (size_t)k * 8, p->N, p->tab0[j]);
        CUDA_TRY(cudaGetLastError());
        p->launches++;
    }
    return SC_OK;
}

// ---- table upload --------------------------------------------------------------------------------------------------
// Pinned bounce slots for pageable sources (host_copy.h), shared by all handles of a device; an upload holds the mutex.
constexpr int BOUNCE_MAX_SLOTS = 16;
size_t bounce_slot_bytes() {
    static const size_t v = [] {
        const char* e = getenv("SC_BOUNCE_SLOT_KB");
        size_t kb = e ? strtoull(e, nullptr, 10) : 16384;
        if (kb < 64) kb = 64;
        if (kb > (1u << 20)) kb = 1u << 20;
        return kb << 10;
    }();
    return v;
}
int bounce_slots() {
    static const int v = [] {
        const char* e = getenv("SC_BOUNCE_SLOTS");
        int n = e ? atoi(e) : 4;
        return n < 2 ? 2 : (n > BOUNCE_MAX_SLOTS ? BOUNCE_MAX_SLOTS : n);
    }();
    return v;
}
#define BOUNCE_SLOT_BYTES bounce_slot_bytes()
#define BOUNCE_SLOTS bounce_slots()
struct Bounce {
    std::mutex mu;
    uint8_t* slot[BOUNCE_MAX_SLOTS] = {};
    cudaEvent_t ev[BOUNCE_MAX_SLOTS] = {};
    bool busy[BOUNCE_MAX_SLOTS] = {};
    size_t k = 0;
};
Bounce g_bounce[64];

bool host_pointer_is_pageable(const void* ptr) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// dst (device) <- src (host) on stream s.  A pageable source is staged: the copy pool fills a pinned slot with
// non-temporal stores while the DMA engine drains the previous one.
cudaError_t h2d(Bounce* B, void* dst, const void* src, size_t bytes, cudaStream_t s) {
    if (!B) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
    hcopy::Pool& pool = hcopy::Pool::get();
    for (size_t off = 0; off < bytes; off += BOUNCE_SLOT_BYTES) {
        const size_t len = bytes - off < BOUNCE_SLOT_BYTES ? bytes - off : BOUNCE_SLOT_BYTES;
        const int k = (int)(B->k++ % BOUNCE_SLOTS);
        cudaError_t e;
        if (B->busy[k] && (e = cudaEventSynchronize(B->ev[k])) != cudaSuccess) return e;
        pool.copy(B->slot[k], (const uint8_t*)src + off, len);
        if ((e = cudaMemcpyAsync((uint8_t*)dst + off, B->slot[k], len, cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
        if ((e = cudaEventRecord(B->ev[k], s)) != cudaSuccess) return e;
        B->busy[k] = true;
    }
    return cudaSuccess;
}

// Locks and (first time) allocates the device's bounce slots when any source table is pageable and large enough to matter.
struct BounceLease {
    Bounce* B = nullptr;
    ~BounceLease() { release(); }
    void release() {
        if (B) B->mu.unlock();
        B = nullptr;
    }
    cudaError_t acquire(const sc_prover* p, const uint64_t* const* tables) {
        if (p->N * 32 < ((size_t)4 << 20)) return cudaSuccess;
        return acquire_raw(p->device, (const void* const*)tables, p->T);
    }
    cudaError_t acquire_raw(int device, const void* const* ptrs, uint32_t n) {
        if (getenv("SC_NO_BOUNCE")) return cudaSuccess;
        bool pageable = false;
        for (uint32_t j = 0; j < n && !pageable; j++) pageable = host_pointer_is_pageable(ptrs[j]);
        if (!pageable) return cudaSuccess;
        Bounce& b = g_bounce[device & 63];
        b.mu.lock();
        B = &b;
        for (int k = 0; k < BOUNCE_SLOTS; k++) {
            if (!b.slot[k]) {
                cudaError_t e = cudaHostAlloc((void**)&b.slot[k], BOUNCE_SLOT_BYTES, cudaHostAllocDefault);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b.ev[k], cudaEventDisableTiming);
                if (e != cudaSuccess) { release(); return e; }
            }
            b.busy[k] = false;  // the previous upload synchronised its stream before it let go of the slots
        }
        return cudaSuccess;
    }
};

// H2D of all tables into the pristine copies, then rewind to round 0 (prover_init's deep copy and sc_prover_load_tables).
int upload_tables(sc_prover* p, const uint64_t* const* tables) {
    CUDA_TRY(cudaSetDevice(p->device));
    BounceLease lease;
    CUDA_TRY(lease.acquire(p, tables));
    // Pipelined path: round 1 needs no challenge, so it is summed chunk by chunk behind the copies (the upload of 1.5 GiB
    // takes ~29 ms over PCIe; the 1.1 ms of round 1 disappear behind it).  One product with a deferred coefficient only
    // (pre-scaled tables would have to be scaled before they are summed), single GPU, d + 1 <= 5 points, host finishing.
    const unsigned long long tiles = p->N / 256;  // 64-row tiles of round 1 (128 pairs each)
    const bool eager = p->r1_ok && p->host_post && !p->comm && !p->is_shard && !p->d_scaled && p->n_products == 1 && p->d + 1 <= (uint32_t)sck::MAX_NPTS &&
                       tiles % EAGER_CHUNKS == 0 && tiles / EAGER_CHUNKS >= 1024 && !getenv("SC_NO_EAGER_R1");
    if (!eager) {
        for (uint32_t j = 0; j < p->T; j++) CUDA_TRY(h2d(lease.B, p->tab0[j], tables[j], p->N * 32, p->stream));
        int rc = prescale_tables(p);
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(p->stream));
        return sc_prover_reset(p);
    }
    if (!p->copy_stream) CUDA_TRY(stream_acquire(&p->copy_stream, p->device));
    for (auto& e : p->eager_ev)
        if (!e) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // the copies may only start once earlier work on the proving stream (a previous proof reading tab0) has finished
    CUDA_TRY(cudaEventRecord(p->eager_ev[0], p->stream));
    CUDA_TRY(cudaStreamWaitEvent(p->copy_stream, p->eager_ev[0], 0));
    sc_prover_reset(p);
    const uint32_t epoch = ++p->eager_epoch;
    const size_t chunk_elems = p->N / EAGER_CHUNKS;
    sck::RoundParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.tab_in = (const uint32_t* const*)p->d_ptr0;
    rp.prod_offsets = p->d_offsets; rp.prod_indices = p->d_indices; rp.prod_first = p->d_first; rp.coeffs = p->d_coeffs;
    rp.n_products = p->n_products; rp.n_tables = p->T; rp.defer_coeff = 1;
    rp.n_pairs = chunk_elems / 2;
    rp.partials = p->d_partials; rp.counter = p->d_counter;
    rp.evals_out = p->d_evals; rp.canon_out = p->d_canon; rp.degree = p->d;
    rp.tmaps = p->d_maps + (size_t)3 * p->T * sizeof(CUtensorMap);
    rp.raw_out = 1;
    rp.seq = epoch;
    const bool gemm = gemm_round_ok(p, rp.n_pairs, false);
    p->eager_gemm = gemm;
    for (uint32_t c = 0; c < EAGER_CHUNKS; c++) {
        for (uint32_t j = 0; j < p->T; j++)
            CUDA_TRY(h2d(lease.B, p->tab0[j] + c * chunk_elems * 8, tables[j] + c * chunk_elems * 4, chunk_elems * 32, p->copy_stream));
        CUDA_TRY(cudaEventRecord(p->eager_ev[c], p->copy_stream));
        CUDA_TRY(cudaStreamWaitEvent(p->stream, p->eager_ev[c], 0));
        rp.tile_base = (uint32_t)(c * (tiles / EAGER_CHUNKS));
        rp.host_out = p->d_eager + (size_t)c * EAGER_SLOT_WORDS;
        rp.host_flag = p->d_eager + (size_t)c * EAGER_SLOT_WORDS + sck::MAX_NPTS * 8;
        if (gemm) {
            int rc = launch_gemm_round(p, rp, false, rp.tile_base, tiles / EAGER_CHUNKS, 1 + c, epoch);
            if (rc) return rc;
            continue;
        }
        cudaError_t e;
        switch (p->d + 1) {
            case 1: e = launch_round1_tma<1>(p, rp); break;
            case 2: e = launch_round1_tma<2>(p, rp); break;
            case 3: e = launch_round1_tma<3>(p, rp); break;
            case 4: e = launch_round1_tma<4>(p, rp); break;
            default: e = launch_round1_tma<5>(p, rp); break;
        }
        if (e != cudaSuccess) return fail(SC_ERR_CUDA, "round-1 chunk launch: %s", cudaGetErrorString(e));
    }
    CUDA_TRY(cudaStreamSynchronize(p->copy_stream));  // the caller's buffers are free again; the last chunk's sum may still run
    p->eager_valid = true;
    return SC_OK;
}


int sharded_round(sc_prover* p, const uint64_t* r);  // capi_multi.inc

// Finish a round whose kernel delivered raw sums (RoundParams::raw_out) into w->h_result: scale by the deferred
// coefficient of a single product (c * sum == sum of c * terms, exact), fill P(1) = P_prev(r) - P(0) when the launch
// summed t = 0, 2, .., d only (the identity the verifier checks, verifier.rs:109), and produce the canonical integers
// ark-serialize feeds the transcript.  Writes the message into p->h_evals / p->h_canon.
void host_finish_round(sc_prover* p, sc_prover* w, const uint64_t* r) {
    const uint32_t d = p->d, ns = w->raw_npts;
    hfr::F sums[8], out[8], coeff;
    if (w->gemm_active) {  // the big integers from the contraction kernels (six at degree 3, nine at degree 4) -> P(0..d)
        hfr::gemm_finish(w->h_gemm, gemm_limbs(p), p->gemm_m == 2 ? 1 : 2, p->gemm_m == 2 ? 1 : p->gemm_m - 2, d, out);
        if (p->n_products == 1) {
            memcpy(&coeff, p->h_coeffs.data(), 32);
            for (uint32_t t = 0; t <= d; t++) out[t] = hfr::mul(out[t], coeff);
        }
        for (uint32_t t = 0; t <= d; t++) {
            const hfr::F c = hfr::to_canonical(out[t]);
            memcpy(p->h_evals + (size_t)t * 8, &out[t], 32);
            memcpy(p->h_canon + (size_t)t * 8, &c, 32);
        }
        return;
    }
    memcpy(sums, w->h_result, (size_t)ns * 32);
    const bool scale = p->n_products == 1;
    if (scale) memcpy(&coeff, p->h_coeffs.data(), 32);
    auto fin = [&](const hfr::F& x) { return scale ? hfr::mul(x, coeff) : x; };
    if (w->alt_active && d >= 2) {
        // sums at the alternative points: round 1 slots = (0, 1, inf, -1, 2)[0..d]; a round that skipped P(1): (0, inf, -1, 2, -2)[0..d-1]
        hfr::F at[5], cd;  // at[i] = P at the i-th of (0, 1, -1, 2, -2)
        if (w->used_skip1) {
            hfr::F rr, prev[8];
            memcpy(&rr, r, 32);
            memcpy(prev, p->h_prev.data(), (size_t)(d + 1) * 32);
            at[0] = fin(sums[0]);
            at[1] = hfr::sub(hfr::interpolate(prev, d, rr), at[0]);  // P(1) = P_prev(r) - P(0) (verifier.rs:109)
            cd = fin(sums[1]);
            for (uint32_t i = 2; i < d; i++) at[i] = fin(sums[i]);
        } else {
            at[0] = fin(sums[0]);
            at[1] = fin(sums[1]);
            cd = fin(sums[2]);
            for (uint32_t i = 2; i < d; i++) at[i] = fin(sums[i + 1]);
        }
        hfr::alt_to_standard(d, at, cd, out);
    } else if (w->used_skip1) {  // sums hold P(0), P(2), .., P(d)
        hfr::F rr, prev[8];
        memcpy(&rr, r, 32);
        memcpy(prev, p->h_prev.data(), (size_t)(d + 1) * 32);
        const hfr::F claim = hfr::interpolate(prev, d, rr);
        out[0] = fin(sums[0]);
        out[1] = hfr::sub(claim, out[0]);
        for (uint32_t t = 2; t <= d; t++) out[t] = fin(sums[t - 1]);
    } else {
        for (uint32_t t = 0; t <= d; t++) out[t] = fin(sums[t]);
    }
    for (uint32_t t = 0; t <= d; t++) {
        const hfr::F c = hfr::to_canonical(out[t]);
        memcpy(p->h_evals + (size_t)t * 8, &out[t], 32);
        memcpy(p->h_canon + (size_t)t * 8, &c, 32);
    }
}

// ---- resident rounds (resident_kernel.cuh) ---------------------------------------------------------------------------
// First (1-based) round the resident kernel serves for a whole-proof call (run_rounds), or 0.
bool comm_fused_exchange_ok(const sc_prover* p);  // capi_multi.inc: the sharded rounds of this handle run the fused peer-memory exchange
int comm_device_share(const sc_prover* p);

uint32_t resident_first_round(const sc_prover* p) {
    if (p->res_max_pairs == 0 || !p->host_post || !p->direct_results || p->d < 1 || p->d > (uint32_t)sck::MAX_NPTS || !p->d_lagrange || getenv("SC_TAIL"))
        return 0;
    // sharded handle: the resident kernel serves the sharded rounds up to the switch (exchange of the partial sums inside the
    // kernel); the replicated rounds after the switch run on the sub-prover's own resident launch
    if (p->comm && !comm_fused_exchange_ok(p)) return 0;
    const uint32_t last = p->comm ? p->switch_round - 1 : p->nv_local;
    // The resident kernel multiplies on the CUDA cores, so a round costs it time in proportion to the entries of the product list;
    // where the contraction kernels are available (a launch of theirs costs ~25-40 us whatever the list) the hand-over moves to
    // smaller rounds for long lists: 2^16 pairs for one product of three tables, 2^13 for BASELINE config 4 (16 entries)
    unsigned long long max_pairs = p->res_max_pairs;
    if (p->gemm_shape && p->h_nnz > 3 && !getenv("SC_RES_MAX_PAIRS")) {
        max_pairs = max_pairs * 3 / p->h_nnz;
        if (max_pairs < p->tc_min_pairs / 2) max_pairs = p->tc_min_pairs / 2;  // (rounds below tc_min_pairs have no launch of their own to go to)
    }
    for (uint32_t i = 2; i <= last; i++)
        if (((unsigned long long)1 << (p->nv_local - i)) <= max_pairs) return i;
    return 0;
}

// Enqueue the resident kernel behind the launch of round res_first - 1 (p->cur already names the buffer that round writes).
int resident_launch(sc_prover* p) {
    const uint32_t first = p->res_first, nv = p->nv_local, d = p->d;
    const uint32_t last = p->comm ? p->switch_round - 1 : nv;  // sharded: up to the switch to replicated rounds
    sck::ResidentParams q;
    memset(&q, 0, sizeof(q));
    sck::RoundParams& rp = q.rp;
    rp.prod_offsets = p->d_offsets; rp.prod_indices = p->d_indices; rp.prod_first = p->d_first; rp.coeffs = p->d_coeffs;
    rp.prod_scaled = p->d_scaled;
    rp.n_products = p->n_products; rp.n_tables = p->T; rp.defer_coeff = (p->n_products == 1) ? 1u : 0u;
    rp.t0 = 0; rp.write_fold = 1; rp.skip1 = 1; rp.degree = d;
    q.ptrs[0] = p->d_ptr0; q.ptrs[1] = p->d_ptrA; q.ptrs[2] = p->d_ptrB;
    q.cur = p->cur;
    q.n_rounds = last - first + 1;
    q.n_pairs_first = (unsigned long long)1 << (nv - first);
    q.seq0 = p->seq + 1;
    if (p->comm) q.mail_seq0 = set_exchange_params_resident(p, rp, q.n_rounds);
    q.h_consts = (const uint32_t*)((uint8_t*)p->d_res + RES_OFF_CONSTS);
    q.h_sums = (uint32_t*)((uint8_t*)p->d_res + RES_OFF_SUMS);
    q.h_error = (uint32_t*)((uint8_t*)p->d_res + RES_OFF_ERROR);
    q.h_abort = (const uint32_t*)((uint8_t*)p->d_res + RES_OFF_ABORT);
    q.d_bcast = p->d_res_bcast;
    q.d_abort = p->d_res_bcast + 128;
    q.totals = (unsigned long long*)p->d_partials;  // [2][NPTS][17] u64
    q.counters = p->d_res_counters;
    CUDA_TRY(cudaMemsetAsync(p->d_partials, 0, 2 * sck::MAX_NPTS * 17 * sizeof(unsigned long long), p->stream));
    {   // fine-grained rounds: every pair spread over 2^lpp lanes (kernels.cuh accumulate_fine)
        const uint32_t nnz = (uint32_t)p->h_nnz, units = 2 * nnz > p->n_products * d ? 2 * nnz : p->n_products * d;
        if (units <= 32 && !getenv("SC_RES_NO_FINE")) {
            uint32_t lg = 0;
            while ((1u << lg) < units) lg++;
            q.lpp_log2 = lg;
            const char* e = getenv("SC_RES_FINE_MAX_PAIRS");
            q.fine_max_pairs = e ? strtoull(e, nullptr, 10) : 4096;  // measured: 512..8192 are within noise at nv = 20 and 24 (tools/tune_resident.sh)
        }
    }
    const int khz = g_dev[p->device].khz;
    static const double secs = getenv("SC_RES_TIMEOUT_S") ? atof(getenv("SC_RES_TIMEOUT_S")) : 10.0;
    q.timeout = (long long)(secs * (khz > 0 ? khz : 1965000) * 1000.0);
    *(volatile uint32_t*)((uint8_t*)p->h_res + RES_OFF_ERROR) = 0;
    if (getenv("SC_RES_PROF")) {
        if (!p->d_res_prof) CUDA_TRY(cudaMalloc(&p->d_res_prof, 64 * 4 * sizeof(long long)));
        CUDA_TRY(cudaMemsetAsync(p->d_res_prof, 0, 64 * 4 * sizeof(long long), p->stream));
        q.prof = p->d_res_prof;
    }
    CUDA_TRY(cudaMemsetAsync(p->d_res_counters, 0, 64 * sizeof(unsigned int), p->stream));
    unsigned long long need = (q.n_pairs_first + sck::RES_THREADS - 1) / sck::RES_THREADS;
    for (uint32_t rd = 0; rd < q.n_rounds && q.fine_max_pairs; rd++) {  // the fine-grained rounds use more CTAs per pair
        const unsigned long long n = q.n_pairs_first >> rd, per = sck::RES_THREADS >> q.lpp_log2;
        if (n <= q.fine_max_pairs && (n + per - 1) / per > need) need = (n + per - 1) / per;
    }
    unsigned long long cap = (unsigned long long)sck::resident_max_grid(d, p->device, g_dev[p->device].sms);
    // ranks sharing a device (sc_prover_create_multi with a repeated device id): their resident kernels wait for each other's
    // partial sums, so ALL of them must fit on the device at once — next to the launch-per-round kernels of a rank that is
    // still a round behind (half of the device is left to those)
    const int share = comm_device_share(p) > p->res_share ? comm_device_share(p) : p->res_share;
    if (comm_device_share(p) > 1) cap /= 2ull * (unsigned long long)comm_device_share(p);
    else if (share > 1) cap = cap / (unsigned long long)share;  // a batch: the kernels do not depend on each other, they only share the SMs
    if (cap < 1) cap = 1;
    const int grid = (int)(need < cap ? (need ? need : 1) : cap);
    cudaError_t e = sck::launch_resident(d, grid, q, p->stream, share == 1);
    if (e != cudaSuccess) return fail(SC_ERR_CUDA, "resident kernel launch: %s", cudaGetErrorString(e));
    p->launches++;
    p->res_running = true;
    p->res_seq0 = q.seq0;
    p->res_last_seq = q.seq0 + q.n_rounds - 1;
    return SC_OK;
}

// The host abandons a proof whose resident kernel is still waiting: tell it to leave (it would otherwise time out).
void resident_abort(sc_prover* p) {
    if (!p->res_running) return;
    *(volatile uint32_t*)((uint8_t*)p->h_res + RES_OFF_ABORT) = p->res_seq0;
    __sync_synchronize();
    cudaStreamSynchronize(p->stream);
    p->res_running = false;
}

// One round served by the resident kernel: prove_round's state machine, then fold constants out / raw sums in.
// `w` is the handle whose kernel is on the GPU: p itself, or the replicated sub-prover of a sharded p.
int resident_collect(sc_prover* p, sc_prover* w, const uint64_t* r);

int resident_post(sc_prover* p, sc_prover* w, const uint64_t* r) {
    p->randomness.insert(p->randomness.end(), r, r + 4);  // prover.rs:82
    p->round += 1;
    if (p->round > p->nv) return fail(SC_ERR_PANIC_NOT_ACTIVE, "Prover is not active");
    if (w != p) w->round += 1;
    const uint32_t seq = ++w->seq, d = p->d;
    const bool prof = w->d_res_prof != nullptr;
    auto now_us = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
    const double h0 = prof ? now_us() : 0;
    // C[k] = r * 2^(32k+64) mod p as plain integers (fr.cuh fold_const): Montgomery-multiply r by the raw 2^(32k+64) mod p
    static const hfr::F X7 = {{0x355094eacaaf6b13ULL, 0xf6b10cb369a568efULL, 0xe2c926a640cc3869ULL, 0x736a6d3bed269aadULL}};  // 2^288 mod p
    hfr::F rr;
    memcpy(&rr, r, 32);
    uint64_t* hc = (uint64_t*)((uint8_t*)w->h_res + RES_OFF_CONSTS);
    for (int k = 0; k < 8; k++) {
        hfr::F x = {{0, 0, 0, 0}};
        if (k < 6) x.l[(k + 2) / 2] = (uint64_t)1 << (32 * ((k + 2) & 1));
        else if (k == 6) x = hfr::ONE;
        else x = X7;
        const hfr::F c = hfr::mul(rr, x);
        for (int i = 0; i < 4; i++) {
            __atomic_store_n(hc + k * 8 + 2 * i, (uint64_t)(uint32_t)c.l[i] | ((uint64_t)seq << 32), __ATOMIC_RELAXED);
            __atomic_store_n(hc + k * 8 + 2 * i + 1, (uint64_t)(c.l[i] >> 32) | ((uint64_t)seq << 32), __ATOMIC_RELAXED);
        }
    }
    w->res_t0 = h0;
    w->res_t1 = prof ? now_us() : 0;
    return SC_OK;
}

int resident_collect(sc_prover* p, sc_prover* w, const uint64_t* r) {
    const uint32_t seq = w->seq, d = p->d;
    const bool prof = w->d_res_prof != nullptr;
    auto now_us = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; };
    const double h0 = w->res_t0, h1 = w->res_t1;
    // wait for the d sums: unreduced 17-limb integers, every limb one {limb, seq} word
    const uint64_t* hs = (const uint64_t*)((uint8_t*)w->h_res + RES_OFF_SUMS);
    const uint32_t n_words = d * 17;
    uint32_t wide[sck::MAX_NPTS * 17];
    SpinWait sw;
    for (uint32_t k = 0; k < n_words; k++) {
        uint64_t v;
        while (((v = __atomic_load_n(hs + k, __ATOMIC_RELAXED)) >> 32) != seq) {
            sw.pause();
            if (sw.check_now()) {
                if (*(volatile uint32_t*)((uint8_t*)w->h_res + RES_OFF_ERROR)) {
                    w->res_running = false;
                    return fail(p->comm ? SC_ERR_COMM : SC_ERR_CUDA, "resident kernel gave up waiting (host constants or a peer GPU did not arrive in time)");
                }
                cudaError_t q = cudaStreamQuery(w->stream);
                if (q != cudaSuccess && q != cudaErrorNotReady) { w->res_running = false; return fail(SC_ERR_CUDA, "resident kernel failed: %s", cudaGetErrorString(q)); }
                if (q == cudaSuccess && (__atomic_load_n(hs + k, __ATOMIC_RELAXED) >> 32) != seq) {
                    w->res_running = false;
                    return fail(SC_ERR_CUDA, "resident kernel ended without publishing round %u", p->round);
                }
            }
        }
        wide[k] = (uint32_t)v;
    }
    __sync_synchronize();
    const double h2 = prof ? now_us() : 0;
    for (uint32_t t = 0; t < d; t++) {  // V * R^-1 mod p: what fr::wide_reduce computes on the device
        const hfr::F s = hfr::reduce_wide17(wide + 17 * t);
        memcpy(w->h_result + (size_t)t * 8, &s, 32);
    }
    w->raw_npts = d;
    w->used_skip1 = true;
    w->alt_active = true;  // the resident kernel sums at the alternative points
    w->gemm_active = false;
    host_finish_round(p, w, r);
    memcpy(p->h_prev.data(), p->h_evals, (size_t)(d + 1) * 32);
    w->cur = (w->cur == 1) ? 2 : 1;
    if (p->comm && comm_failed(p)) { w->res_running = false; return fail(SC_ERR_COMM, "a peer GPU did not deliver its partial sums in time"); }
    if (prof) {
        double* h = w->res_host_us[(seq - w->res_seq0) & 63];
        h[0] = h1 - h0; h[1] = h2 - h1; h[2] = now_us() - h2;
        if (seq == w->res_last_seq) {
            cudaStreamSynchronize(w->stream);
            long long hp[64 * 4];
            cudaMemcpy(hp, w->d_res_prof, sizeof(hp), cudaMemcpyDeviceToHost);
            for (uint32_t k = 0; k <= w->res_last_seq - w->res_seq0 && k < 64; k++)
                fprintf(stderr, "resident round %u: device wait %lld accumulate %lld reduce %lld publish %lld cycles | host consts %.2f wait %.2f finish %.2f us\n",
                        p->round - (w->res_last_seq - w->res_seq0) + k, hp[k * 4], hp[k * 4 + 1], hp[k * 4 + 2], hp[k * 4 + 3], w->res_host_us[k][0], w->res_host_us[k][1], w->res_host_us[k][2]);
        }
    }
    p->res_rounds++;
    if (seq == w->res_last_seq) w->res_running = false;  // the kernel has published its last round and leaves by itself
    p->out_evals = nullptr;  // the message of a resident round lives on the host only
    p->out_canon = nullptr;
    return SC_OK;
}

int resident_round(sc_prover* p, sc_prover* w, const uint64_t* r) {
    int rc = resident_post(p, w, r);
    if (rc) return rc;
    return resident_collect(p, w, r);
}

// prove_round state machine (prover.rs:78-98) + device round + D2H of the d+1 results into the pinned buffers.
// Split in two so that a batch driver (capi_gkr.inc sc_gkr_prove_batch) can issue the round of many independent provers before
// it collects the first result: prove_round_issue launches, prove_round_collect waits for the message and finishes it.
int prove_round_issue(sc_prover* p, const uint64_t* r_or_null) {
    p->collect_pending = false;
    if (r_or_null) {
        if (p->round == 0) return fail(SC_ERR_PANIC_FIRST_ROUND_MSG, "first round should be prover first.");
        p->randomness.insert(p->randomness.end(), r_or_null, r_or_null + 4);
    } else if (p->round > 0) {
        return fail(SC_ERR_PANIC_MISSING_MSG, "verifier message is empty");
    }
    p->round += 1;
    if (p->round > p->nv) return fail(SC_ERR_PANIC_NOT_ACTIVE, "Prover is not active");
    CUDA_TRY(cudaSetDevice(p->device));
    const bool timed = p->timing && p->want_timing;
    if (timed) CUDA_TRY(cudaEventRecord(p->ev[2 * (p->round - 1)], p->stream));
    // prover.rs:85-86: r = randomness[round-1] — the challenge just pushed
    int rc;
    if (p->eager_valid && p->round == 1 && !p->comm) {
        // round 1 was summed in EAGER_CHUNKS pieces behind the upload: wait for the last piece, add the pieces, finish
        p->eager_valid = false;
        volatile uint32_t* flag = p->eager_gemm ? p->h_gemm + (size_t)(EAGER_CHUNKS + 1) * gsum::OUT_SLOT_WORDS - 1
                                                : p->h_eager + (size_t)(EAGER_CHUNKS - 1) * EAGER_SLOT_WORDS + sck::MAX_NPTS * 8;
        SpinWait sw;
        while (*flag != p->eager_epoch) {
            sw.pause();
            if (sw.check_now()) {
                cudaError_t q = cudaStreamQuery(p->stream);
                if (q != cudaSuccess && q != cudaErrorNotReady) return fail(SC_ERR_CUDA, "round-1 chunk kernel failed: %s", cudaGetErrorString(q));
                if (q == cudaSuccess && *flag != p->eager_epoch) return fail(SC_ERR_CUDA, "round-1 chunks finished without publishing");
            }
        }
        __sync_synchronize();
        const uint32_t npts = p->d + 1;
        if (p->eager_gemm) {  // add the chunks' integers up in slot 0
            const uint32_t nb = gemm_nb(p), nl = gemm_limbs(p);
            memset(p->h_gemm, 0, (size_t)nb * nl * 4);
            for (uint32_t c = 0; c < EAGER_CHUNKS; c++)
                for (uint32_t i = 0; i < nb; i++)
                    hfr::add_limbs(p->h_gemm + (size_t)i * nl, p->h_gemm + (size_t)(1 + c) * gsum::OUT_SLOT_WORDS + (size_t)i * nl, nl);
            p->gemm_active = true;
            p->gemm_rounds++;
        } else {
            hfr::F tot[8];
            memset(tot, 0, sizeof(tot));
            for (uint32_t c = 0; c < EAGER_CHUNKS; c++)
                for (uint32_t t = 0; t < npts; t++) {
                    hfr::F x;
                    memcpy(&x, p->h_eager + (size_t)c * EAGER_SLOT_WORDS + t * 8, 32);
                    tot[t] = hfr::add(tot[t], x);
                }
            memcpy(p->h_result, tot, (size_t)npts * 32);
            p->gemm_active = false;
            p->alt_active = true;  // the chunks were summed by round1_tma_kernel
        }
        p->raw_npts = npts;
        p->used_skip1 = false;
        host_finish_round(p, p, nullptr);
        memcpy(p->h_prev.data(), p->h_evals, (size_t)npts * 32);
        p->out_evals = p->d_evals;
        p->out_canon = p->d_canon;
        p->direct_active = true;
        if (timed) CUDA_TRY(cudaEventRecord(p->ev[2 * (p->round - 1) + 1], p->stream));  // the round's work ran during the upload
        if (p->res_first && p->round + 1 == p->res_first) return resident_launch(p);
        return gemm_prelaunch(p);  // nothing left to collect; round 2 can start its prologue while the caller hashes round 1
    }
    if (p->comm) {
        rc = sharded_round(p, r_or_null);
        if (!rc) rc = gemm_prelaunch(p);
    } else {
        rc = run_round_device(p, r_or_null);
        p->out_evals = p->d_evals;
        p->out_canon = p->d_canon;
        if (!rc) rc = gemm_prelaunch(p);  // the next large fold round right behind this one (whole-proof calls only)
    }
    if (rc) return rc;
    if (timed) CUDA_TRY(cudaEventRecord(p->ev[2 * (p->round - 1) + 1], p->stream));
    if (p->host_done) {  // the message is already complete on the host
        p->host_done = false;
        return SC_OK;
    }
    p->collect_pending = true;
    if (p->res_first && p->round + 1 == p->res_first) {
        // the next round is the first resident one: queue the kernel now, so that it is already polling when this
        // round's challenge has been drawn
        if (timed) CUDA_TRY(cudaEventRecord(p->ev[2 * p->round], p->stream));
        rc = resident_launch(p);
        if (rc) return rc;
        if (timed) {  // one interval for the whole resident launch; the later rounds have no launch of their own
            CUDA_TRY(cudaEventRecord(p->ev[2 * p->round + 1], p->stream));
            for (uint32_t i = p->round + 1; i < p->nv; i++) {
                CUDA_TRY(cudaEventRecord(p->ev[2 * i], p->stream));
                CUDA_TRY(cudaEventRecord(p->ev[2 * i + 1], p->stream));
            }
        }
    }
    return SC_OK;
}

int prove_round_collect(sc_prover* p, const uint64_t* r_or_null, bool sync_out) {
    if (!p->collect_pending) return SC_OK;
    p->collect_pending = false;
    if (p->direct_active) {
        // the last block wrote the message into mapped pinned memory and then the flag: spin instead of copy + sync
        sc_prover* w = (p->comm && p->wait_on) ? p->wait_on : p;
        volatile uint32_t* flag = w->gemm_active ? w->h_gemm + gsum::OUT_SLOT_WORDS - 1 : w->h_result + (size_t)(p->d + 1) * 16;
        const uint32_t want = w->gemm_active ? w->gemm_want : w->seq;
        SpinWait sw;
        while (*flag != want) {
            sw.pause();
            if (sw.check_now()) {  // every 64K polls make sure the kernel has not died
                cudaError_t q = cudaStreamQuery(p->stream);
                if (q != cudaSuccess && q != cudaErrorNotReady) return fail(SC_ERR_CUDA, "round kernel failed: %s", cudaGetErrorString(q));
                if (q == cudaSuccess && *flag != want) return fail(SC_ERR_CUDA, "round finished without publishing its result");
            }
        }
        __sync_synchronize();
        if (w->gemm_active && w->h_gemm[GEMM_RERROR_WORD]) {
            w->h_gemm[GEMM_RERROR_WORD] = 0;
            return fail(SC_ERR_CUDA, "a fold round launched ahead of its challenge gave up waiting for it");
        }
        if (w->raw_active) {
            host_finish_round(p, w, r_or_null);
        } else if (w != p) {
            memcpy(p->h_result, w->h_result, (size_t)(p->d + 1) * 64);
        }
        memcpy(p->h_prev.data(), p->h_evals, (size_t)(p->d + 1) * 32);
        if (p->comm && comm_failed(p)) return fail(SC_ERR_COMM, "a peer GPU did not deliver its partial sums in time");
        return SC_OK;
    }
    const size_t bytes = (size_t)(p->d + 1) * 32;
    CUDA_TRY(cudaMemcpyAsync(p->h_evals, p->out_evals, bytes, cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->h_canon, p->out_canon, bytes, cudaMemcpyDeviceToHost, p->stream));
    if (sync_out) {
        CUDA_TRY(cudaStreamSynchronize(p->stream));
        memcpy(p->h_prev.data(), p->h_evals, (size_t)(p->d + 1) * 32);
    }
    return SC_OK;
}

int prove_round_impl(sc_prover* p, const uint64_t* r_or_null, bool sync_out) {
    int rc = prove_round_issue(p, r_or_null);
    if (rc) return rc;
    return prove_round_collect(p, r_or_null, sync_out);
}

// First round (1-based) handled by the fused tail kernel, or nv+1 when the tail is not used.
uint32_t tail_first_round(const sc_prover* p) {
    // Opt-in (SC_TAIL=1): measured on B200 the single-CTA tail with the device-side Blake2b transcript costs ~60 us per
    // round against ~36 us for a host-driven round now that results arrive through mapped memory (profiles/README.md).
    if (p->comm || p->d < 1 || p->d > (uint32_t)sck::MAX_NPTS || !p->d_lagrange || !getenv("SC_TAIL")) return p->nv + 1;
    for (uint32_t i = 2; i <= p->nv; i++)
        if (((unsigned long long)1 << (p->nv - i)) <= sck::TAIL_PAIRS) return i;
    return p->nv + 1;
}

// Rounds first..nv in ONE launch with the transcript on the device (kernels.cuh tail_kernel).  `r` = challenge drawn
// after round first-1, `st` = transcript at that point.  Fills evals_out / challenges_out rows first-1 .. nv-1.
int run_tail(sc_prover* p, b2::State* st, uint32_t first, const uint64_t* r, uint64_t* evals_out, uint64_t* challenges_out) {
    const uint32_t nv = p->nv, d = p->d, n_rounds = nv - first + 1;
    CUDA_TRY(cudaSetDevice(p->device));
    sck::TailParams tp;
    memset(&tp, 0, sizeof(tp));
    sck::RoundParams& rp = tp.rp;
    rp.prod_offsets = p->d_offsets; rp.prod_indices = p->d_indices; rp.prod_first = p->d_first; rp.coeffs = p->d_coeffs;
    rp.prod_scaled = p->d_scaled;
    rp.n_products = p->n_products; rp.n_tables = p->T; rp.defer_coeff = (p->n_products == 1) ? 1u : 0u;
    rp.t0 = 0; rp.write_fold = 1; rp.skip1 = 1; rp.fix1 = 1; rp.degree = d;
    rp.prev_evals = p->d_evals;  // ProverMsg of round first-1 (host-driven)
    rp.lagrange = p->d_lagrange;
    memcpy(rp.r, r, 32);
    tp.ptrs[0] = p->d_ptr0; tp.ptrs[1] = p->d_ptrA; tp.ptrs[2] = p->d_ptrB;
    tp.cur = p->cur;
    tp.n_rounds = n_rounds;
    tp.n_pairs_first = (unsigned long long)1 << (nv - first);
    tp.st_in = p->d_st; tp.st_out = p->d_st + 1;
    tp.evals_all = p->d_tail_evals; tp.chal_all = p->d_tail_chal;
    long long* d_prof = nullptr;
    if (getenv("SC_TAIL_PROF")) { CUDA_TRY(cudaMalloc(&d_prof, (size_t)n_rounds * 4 * sizeof(long long))); tp.prof = d_prof; }
    memcpy(p->h_st, st, sizeof(b2::State));
    CUDA_TRY(cudaMemcpyAsync(p->d_st, p->h_st, sizeof(b2::State), cudaMemcpyHostToDevice, p->stream));
    const bool timed = p->timing && p->want_timing;
    if (timed) CUDA_TRY(cudaEventRecord(p->ev[2 * (first - 1)], p->stream));
    CUDA_TRY(sck::launch_tail(d, tp, p->stream));
    p->launches++;
    if (timed) {
        CUDA_TRY(cudaEventRecord(p->ev[2 * (first - 1) + 1], p->stream));
        for (uint32_t i = first; i < nv; i++) {  // later tail rounds have no launch of their own: zero-length intervals
            CUDA_TRY(cudaEventRecord(p->ev[2 * i], p->stream));
            CUDA_TRY(cudaEventRecord(p->ev[2 * i + 1], p->stream));
        }
    }
    const size_t eb = (size_t)n_rounds * (d + 1) * 32, cb = (size_t)n_rounds * 32;
    CUDA_TRY(cudaMemcpyAsync(p->h_tail, p->d_tail_evals, eb, cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaMemcpyAsync((uint8_t*)p->h_tail + eb, p->d_tail_chal, cb, cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->h_st + 1, p->d_st + 1, sizeof(b2::State), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    if (d_prof) {
        std::vector<long long> hp((size_t)n_rounds * 4);
        cudaMemcpy(hp.data(), d_prof, hp.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        for (uint32_t k = 0; k < n_rounds; k++)
            fprintf(stderr, "tail round %u: accumulate %lld reduce %lld publish %lld transcript %lld cycles\n", first + k, hp[k * 4], hp[k * 4 + 1], hp[k * 4 + 2], hp[k * 4 + 3]);
        cudaFree(d_prof);
    }
    memcpy(evals_out + (size_t)(first - 1) * (d + 1) * 4, p->h_tail, eb);
    memcpy(challenges_out + (size_t)(first - 1) * 4, (uint8_t*)p->h_tail + eb, cb);
    memcpy(st, p->h_st + 1, sizeof(b2::State));
    // ProverState bookkeeping, as if prove_round had been called for every tail round (prover.rs:82,94)
    p->randomness.insert(p->randomness.end(), r, r + 4);
    const uint64_t* ch = challenges_out + (size_t)(first - 1) * 4;
    p->randomness.insert(p->randomness.end(), ch, ch + (size_t)(n_rounds - 1) * 4);
    for (uint32_t k = 0; k < n_rounds; k++) p->cur = (p->cur == 1) ? 2 : 1;
    p->round = nv;
    // the last tail round's message stays readable through the usual buffers
    CUDA_TRY(cudaMemcpyAsync(p->d_evals, p->d_tail_evals + (size_t)(n_rounds - 1) * (d + 1) * 8, (size_t)(d + 1) * 32,
                             cudaMemcpyDeviceToDevice, p->stream));
    return SC_OK;
}

int switch_to_replicated(sc_prover* p);  // capi_multi.inc

// One round of a whole-proof call: the resident kernel when it is serving this round, a launch otherwise.
int round_dispatch(sc_prover* p, const uint64_t* r) {
    const uint32_t i = p->round + 1;  // the (global) round about to run
    if (!p->comm || i < p->switch_round)
        return (p->res_first && i >= p->res_first) ? resident_round(p, p, r) : prove_round_impl(p, r, true);
    // sharded handle, replicated rounds: the sub-prover continues the same proof on the gathered (small) tables
    if (i == p->switch_round) {
        int rc = switch_to_replicated(p);
        if (rc) return rc;
        sc_prover* w = p->sub;
        w->res_first = resident_first_round(w) == 2 ? 2 : 0;  // every replicated round, or none
        if (w->res_first) {
            rc = resident_launch(w);
            if (rc) return rc;
        }
    }
    if (p->sub->res_first) return resident_round(p, p->sub, r);
    return prove_round_impl(p, r, true);
}

// The round loop shared by MLSumcheck (mod.rs:59-64) and the two GKR phases (gkr mod.rs:111-119, 126-133):
// prove_round -> rng.feed(&prover_msg) -> sample_round.  challenges_out: nv*4 u64.
int run_rounds(sc_prover* p, b2::State* st, uint64_t* evals_out, uint64_t* challenges_out) {
    const uint32_t nv = p->nv, d = p->d;
    std::vector<uint8_t> msg(8 + 32 * (size_t)(d + 1));
    b2::put_u64(msg.data(), d + 1);
    uint64_t r[4];
    bool have_r = false;
    const uint32_t tail_first = tail_first_round(p);
    p->timing = true;
    p->res_first = resident_first_round(p);
    p->prelaunch_ok = true;
    auto bail = [&](int rc) {
        p->timing = false;
        p->prelaunch_ok = false;
        gemm_abort_prelaunch(p);
        resident_abort(p);
        if (p->sub) { resident_abort(p->sub); p->sub->res_first = 0; }
        p->res_first = 0;
        return rc;
    };
    for (uint32_t i = 0; i < nv; i++) {
        if (i + 1 == tail_first && have_r && st->buflen % 8 == 0) {
            int rc = run_tail(p, st, tail_first, r, evals_out, challenges_out);
            if (rc) return bail(rc);
            break;
        }
        int rc = round_dispatch(p, have_r ? r : nullptr);
        if (rc) return bail(rc);
        memcpy(evals_out + (size_t)i * (d + 1) * 4, p->h_evals, (size_t)(d + 1) * 32);
        memcpy(msg.data() + 8, p->h_canon, (size_t)(d + 1) * 32);
        b2::update(st, msg.data(), msg.size());
        b2::sample_fr(st, r);
        have_r = true;
        memcpy(challenges_out + (size_t)i * 4, r, 32);
    }
    p->res_first = 0;
    if (p->sub) p->sub->res_first = 0;
    p->timing = false;
    p->prelaunch_ok = false;
    gemm_abort_prelaunch(p);  // (none is pending after a complete proof)
    if (p->want_timing) {
        cudaStreamSynchronize(p->stream);
        for (uint32_t i = 0; i < nv; i++) cudaEventElapsedTime(&p->round_ms[i], p->ev[2 * i], p->ev[2 * i + 1]);
    }
    return SC_OK;
}

// The same loop for SEVERAL independent provers of one shape (the layers of sc_gkr_prove_batch), each with its own transcript:
// every round is issued for all provers before the first result is collected, so the ~8 us a latency-bound round costs is
// shared by the whole batch instead of paid once per prover.
int run_rounds_batch(std::vector<sc_prover*>& ps, std::vector<b2::State*>& sts, std::vector<uint64_t*>& evals_out,
                     std::vector<uint64_t*>& challenges_out) {
    const size_t L = ps.size();
    if (L == 1) return run_rounds(ps[0], sts[0], evals_out[0], challenges_out[0]);
    const uint32_t nv = ps[0]->nv, d = ps[0]->d;
    std::vector<uint8_t> msg(8 + 32 * (size_t)(d + 1));
    b2::put_u64(msg.data(), d + 1);
    std::vector<uint64_t> r(L * 4, 0);
    for (sc_prover* p : ps) p->res_first = resident_first_round(p);
    auto bail = [&](int rc) {
        for (sc_prover* p : ps) {
            resident_abort(p);
            p->res_first = 0;
        }
        return rc;
    };
    for (uint32_t i = 0; i < nv; i++) {
        for (size_t l = 0; l < L; l++) {
            sc_prover* p = ps[l];
            const uint64_t* rl = i ? &r[l * 4] : nullptr;
            int rc = (p->res_first && i + 1 >= p->res_first) ? resident_post(p, p, rl) : prove_round_issue(p, rl);
            if (rc) return bail(rc);
        }
        for (size_t l = 0; l < L; l++) {
            sc_prover* p = ps[l];
            const uint64_t* rl = i ? &r[l * 4] : nullptr;
            int rc = (p->res_first && i + 1 >= p->res_first) ? resident_collect(p, p, rl) : prove_round_collect(p, rl, true);
            if (rc) return bail(rc);
            memcpy(evals_out[l] + (size_t)i * (d + 1) * 4, p->h_evals, (size_t)(d + 1) * 32);
            memcpy(msg.data() + 8, p->h_canon, (size_t)(d + 1) * 32);
            b2::update(sts[l], msg.data(), msg.size());
            b2::sample_fr(sts[l], &r[l * 4]);
            memcpy(challenges_out[l] + (size_t)i * 4, &r[l * 4], 32);
        }
    }
    for (sc_prover* p : ps) p->res_first = 0;
    return SC_OK;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

const char* sc_last_error(void) { return g_err.c_str(); }

int sc_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(SC_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

void sc_rng_setup(sc_blake2b512_rng* rng) { b2::init((b2::State*)rng); }
void sc_rng_feed_bytes(sc_blake2b512_rng* rng, const uint8_t* b, size_t n) { b2::update((b2::State*)rng, b, n); }
void sc_rng_fill_bytes(sc_blake2b512_rng* rng, uint8_t* dest, size_t n) { b2::fill_bytes((b2::State*)rng, dest, n); }
uint64_t sc_rng_next_u64(sc_blake2b512_rng* rng) { return b2::next_u64((b2::State*)rng); }
void sc_rng_sample_fr(sc_blake2b512_rng* rng, uint64_t out[4]) { b2::sample_fr((b2::State*)rng, out); }

int sc_prover_create(sc_prover** out, uint32_t nv, uint32_t n_tables, const uint64_t* const* tables, uint32_t n_products,
                     const uint64_t* coeffs, const uint32_t* offsets, const uint32_t* indices, int device) {
    return create_common(out, nv, n_tables, tables, false, n_products, coeffs, offsets, indices, device);
}
int sc_prover_create_device(sc_prover** out, uint32_t nv, uint32_t n_tables, const uint64_t* const* d_tables,
                            uint32_t n_products, const uint64_t* coeffs, const uint32_t* offsets, const uint32_t* indices,
                            int device) {
    return create_common(out, nv, n_tables, d_tables, true, n_products, coeffs, offsets, indices, device);
}

void sc_prover_destroy(sc_prover* p) {
    if (!p) return;
    if (!p->group.empty() || p->workers) { multi_destroy(p); return; }
    cudaSetDevice(p->device);
    gemm_abort_prelaunch(p);
    resident_abort(p);
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->sub) { p->sub->stream = p->sub->own_stream; sc_prover_destroy(p->sub); cudaSetDevice(p->device); }
    for (void* q : p->ipc_opened) cudaIpcCloseMemHandle(q);
    cudaFree(p->d_peer_tabs);
    cudaFree(p->d_gather); cudaFree(p->d_evals_g); cudaFree(p->d_canon_g); cudaFree(p->d_sub_tabs);
    if (p->owns_tab0) device_free(p->slab0, p->slab0_bytes, p->device);
    cudaFree(p->d_res_prof);
    device_free(p->adopted, p->adopted_bytes, p->device);
    device_free(p->slabA, p->slabA_bytes, p->device);  // one slab: ping-pong tables and every small device array
    host_mapped_free(p->h_result, p->h_result_bytes, p->device);  // one pinned block: results, tail read-back, transcript state
    for (auto e : p->ev) if (e) cudaEventDestroy(e);
    if (p->copy_stream) { cudaStreamSynchronize(p->copy_stream); stream_release(p->copy_stream, p->device); }
    for (auto& e : p->eager_ev) if (e) cudaEventDestroy(e);
    stream_release(p->own_stream, p->device);  // synchronised above
    delete p;
}

#define NEED_HANDLE(p) \
    if (!(p)) return fail(SC_ERR_BAD_INPUT, "null prover handle")

int sc_prover_reset(sc_prover* p) {
    NEED_HANDLE(p);
    if (!p->group.empty()) return multi_reset(p);
    gemm_abort_prelaunch(p);
    p->round = 0;
    p->cur = 0;
    p->randomness.clear();
    p->launches = 0;
    p->tc_rounds = 0;
    p->res_rounds = 0;
    p->gemm_rounds = 0;
    p->eager_valid = false;  // a pre-computed first round belongs to the proof that follows its upload only
    p->switched = false;
    if (p->comm) comm_clear_error(p);  // a timed-out exchange invalidated the previous proof, not the communicator
    return SC_OK;
}

int sc_prover_load_tables(sc_prover* p, const uint64_t* const* tables) {
    NEED_HANDLE(p);
    if (!tables) return fail(SC_ERR_BAD_INPUT, "null table list");
    if (!p->group.empty()) return multi_load_tables(p, tables);
    if (!p->owns_tab0) return fail(SC_ERR_BAD_INPUT, "handle was created over caller-owned device tables");
    return upload_tables(p, tables);
}

int sc_prover_set_stream(sc_prover* p, void* cuda_stream) {
    NEED_HANDLE(p);
    if (!p->group.empty()) return fail(SC_ERR_BAD_INPUT, "a multi-device handle runs on its own per-device streams");
    CUDA_TRY(cudaSetDevice(p->device));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    p->stream = cuda_stream ? (cudaStream_t)cuda_stream : p->own_stream;
    return SC_OK;
}

int sc_prove_round(sc_prover* p, const uint64_t* r_or_null, uint64_t* evals_out) {
    NEED_HANDLE(p);
    if (!evals_out) return fail(SC_ERR_BAD_INPUT, "null output buffer");
    if (!p->group.empty()) return multi_prove_round(p, r_or_null, evals_out);
    int rc = prove_round_impl(p, r_or_null, true);
    if (rc) return rc;
    memcpy(evals_out, p->h_evals, (size_t)(p->d + 1) * 32);
    return SC_OK;
}

uint32_t sc_prover_max_multiplicands(const sc_prover* p) { return p ? p->d : 0; }
uint32_t sc_prover_num_vars(const sc_prover* p) { return p ? p->nv : 0; }
uint32_t sc_prover_round(const sc_prover* p) { return p ? lead(p)->round : 0; }
uint32_t sc_prover_randomness(const sc_prover* p, uint64_t* out, uint32_t cap) {
    if (!p) return 0;
    p = lead(p);
    uint32_t n = (uint32_t)(p->randomness.size() / 4);
    uint32_t c = n < cap ? n : cap;
    if (out && c) memcpy(out, p->randomness.data(), (size_t)c * 32);
    return n;
}

int sc_prover_push_randomness(sc_prover* p, const uint64_t r[4]) {
    NEED_HANDLE(p);
    if (!r) return fail(SC_ERR_BAD_INPUT, "null challenge");
    for (sc_prover* q : p->group) q->randomness.insert(q->randomness.end(), r, r + 4);
    p->randomness.insert(p->randomness.end(), r, r + 4);
    return SC_OK;
}

int sc_prover_table(const sc_prover* p, uint32_t j, uint64_t* out, uint64_t cap_elems, uint64_t* len_out) {
    NEED_HANDLE(p);
    if (j >= p->T) return fail(SC_ERR_BAD_INPUT, "table %u out of range", j);
    if (!p->group.empty()) return multi_table(p, j, out, cap_elems, len_out);
    // after round i >= 2 the tables have been folded i-1 times
    if (p->comm && p->round >= p->switch_round) {  // replicated rounds: the full (small) tables live on the sub-prover
        if (!p->sub) return fail(SC_ERR_BAD_INPUT, "sharded prover: no replicated state yet");
        return sc_prover_table(p->sub, j, out, cap_elems, len_out);
    }
    uint64_t len = p->round <= 1 ? p->N : (p->N >> (p->round - 1));  // sharded: this rank's shard
    if (len_out) *len_out = len;
    if (!out) return SC_OK;
    if (cap_elems < len) return fail(SC_ERR_BAD_INPUT, "buffer too small: %llu < %llu", (unsigned long long)cap_elems, (unsigned long long)len);
    const uint32_t* src = p->cur == 0 ? p->tab0[j] : (p->cur == 1 ? p->bufA[j] : p->bufB[j]);
    CUDA_TRY(cudaSetDevice(p->device));
    int scaled_by = -1;
    for (uint32_t k = 0; k < p->n_products && k < p->scaled_table.size(); k++)
        if (p->h_scaled[k] && p->scaled_table[k] == (int)j) scaled_by = (int)k;
    if (scaled_by >= 0) {
        // this table carries its product's coefficient (prescale_tables): the reference's table is ours / c — exact
        hfr::F c, cinv;
        memcpy(&c, p->h_coeffs.data() + (size_t)scaled_by * 4, 32);
        cinv = hfr::inverse(c);
        uint32_t* tmp = nullptr;
        size_t got = 0;
        CUDA_TRY(device_alloc((void**)&tmp, len * 32 + 32, &got, p->device));
        CUDA_TRY(cudaMemcpyAsync(tmp + len * 8, &cinv, 32, cudaMemcpyHostToDevice, p->stream));
        unsigned long long need = (len + 127) / 128, cap = (unsigned long long)g_dev[p->device].sms * 16;
        const int grid = (int)(need < cap ? need : cap);
        sck::scale_kernel<<<grid < 1 ? 1 : grid, 128, 0, p->stream>>>(src, tmp + len * 8, len, tmp);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(out, tmp, len * 32, cudaMemcpyDeviceToHost, p->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
        device_free(tmp, got, p->device);
        if (e != cudaSuccess) return fail(SC_ERR_CUDA, "table export: %s", cudaGetErrorString(e));
        return SC_OK;
    }
    CUDA_TRY(cudaMemcpyAsync(out, src, len * 32, cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return SC_OK;
}

int sc_ml_prove(sc_prover* p, sc_blake2b512_rng* rng, uint64_t* evals_out, uint64_t* randomness_out) {
    NEED_HANDLE(p);
    if (!rng || !evals_out) return fail(SC_ERR_BAD_INPUT, "null rng or output buffer");
    if (!p->group.empty()) return multi_ml_prove(p, rng, evals_out, randomness_out);
    if (p->round != 0) return fail(SC_ERR_BAD_INPUT, "sc_ml_prove needs a prover at round 0 (got %u)", p->round);
    b2::State* st = (b2::State*)rng;
    uint8_t info[16];
    b2::put_u64(info, p->d);       // PolynomialInfo.max_multiplicands  (data_structures.rs:50-55)
    b2::put_u64(info + 8, p->nv);  // PolynomialInfo.num_variables
    b2::update(st, info, 16);      // mod.rs:54 fs_rng.feed(&polynomial.info())
    std::vector<uint64_t> ch((size_t)p->nv * 4);
    int rc = run_rounds(p, st, evals_out, ch.data());  // mod.rs:59-64
    if (rc) return rc;
    p->randomness.insert(p->randomness.end(), ch.end() - 4, ch.end());  // mod.rs:65-67
    if (randomness_out) memcpy(randomness_out, p->randomness.data(), (size_t)p->nv * 32);
    return SC_OK;
}

int sc_ml_prove_oneshot(uint32_t nv, uint32_t n_tables, const uint64_t* const* tables, uint32_t n_products,
                        const uint64_t* coeffs, const uint32_t* offsets, const uint32_t* indices, int device,
                        uint64_t* evals_out, uint64_t* randomness_out) {
    sc_prover* p = nullptr;
    int rc = sc_prover_create(&p, nv, n_tables, tables, n_products, coeffs, offsets, indices, device);
    if (rc) return rc;
    sc_blake2b512_rng rng;
    sc_rng_setup(&rng);  // mod.rs:43
    rc = sc_ml_prove(p, &rng, evals_out, randomness_out);
    sc_prover_destroy(p);
    return rc;
}

size_t sc_serialize_proof(const uint64_t* evals, uint32_t nv, uint32_t d, uint8_t* out);  // defined in gkr/serialize section

void sc_synth_table(uint64_t* out, uint64_t n_elems, uint64_t seed) { sc_synth_table_at(out, 0, n_elems, seed); }

void sc_synth_table_at(uint64_t* out, uint64_t first_elem, uint64_t n_elems, uint64_t seed) {
    const uint64_t GAMMA = 0x9e3779b97f4a7c15ULL;
    const uint64_t P[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
    auto mix = [](uint64_t z) {
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    };
    for (uint64_t e0 = 0; e0 < n_elems; e0++) {
        const uint64_t e = first_elem + e0;
        uint64_t t[4];
        for (uint64_t k = 0; k < 8; k++) {
            for (uint64_t i = 0; i < 4; i++) t[i] = mix(seed + GAMMA * (((e * 8 + k) * 4) + i + 1));
            t[3] &= 0x7fffffffffffffffULL;
            bool lt = false;
            for (int i = 3; i >= 0; i--) {
                if (t[i] < P[i]) { lt = true; break; }
                if (t[i] > P[i]) break;
            }
            if (lt) break;
            if (k == 7) t[3] &= 0x3fffffffffffffffULL;
        }
        memcpy(out + 4 * e0, t, 32);
    }
}

uint32_t sc_prover_round_times_ms(const sc_prover* p, float* out, uint32_t cap) {
    if (!p) return 0;
    p = lead(p);
    uint32_t c = p->nv < cap ? p->nv : cap;
    if (out && c) memcpy(out, p->round_ms.data(), c * sizeof(float));
    return p->nv;
}
uint64_t sc_prover_launch_count(const sc_prover* p) { return p ? lead(p)->launches : 0; }
uint64_t sc_prover_tc_round_count(const sc_prover* p) { return p ? lead(p)->tc_rounds : 0; }
uint64_t sc_prover_resident_round_count(const sc_prover* p) { return p ? lead(p)->res_rounds : 0; }
uint64_t sc_prover_gemm_round_count(const sc_prover* p) { return p ? lead(p)->gemm_rounds : 0; }

void sc_release_cached_memory(void) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (int d = 0; d < 64; d++) {
        AllocCache& c = g_cache[d];
        if (c.dev.empty() && c.host.empty() && c.streams.empty()) continue;
        cudaSetDevice(d);
        for (auto& b : c.dev) cudaFree(b.p);
        for (auto& b : c.host) cudaFreeHost(b.p);
        for (auto& st : c.streams) cudaStreamDestroy(st);
        c.dev.clear(); c.host.clear(); c.streams.clear();
    }
    cudaSetDevice(cur);
}

int sc_fr_interpolate(const uint64_t* evals, uint32_t n_evals, const uint64_t r[4], uint64_t out[4]) {
    if (n_evals == 0 || n_evals > 33) return fail(SC_ERR_BAD_INPUT, "n_evals = %u out of range (1..33)", n_evals);
    std::vector<hfr::F> ev(n_evals);
    memcpy(ev.data(), evals, (size_t)n_evals * 32);
    hfr::F rr;
    memcpy(&rr, r, 32);
    const hfr::F v = hfr::interpolate(ev.data(), n_evals - 1, rr);
    memcpy(out, &v, 32);
    return SC_OK;
}
int sc_fr_contraction_finish(const uint32_t* z, uint32_t n_limbs, uint32_t kx, uint32_t ky, uint64_t* out) {
    if (!z || !out || kx < 1 || kx > 2 || ky < 1 || ky > 2 || n_limbs < 1 || n_limbs > 64) return fail(SC_ERR_BAD_INPUT, "contraction_finish: bad shape");
    hfr::F ev[8];
    hfr::gemm_finish(z, n_limbs, kx, ky, kx + ky, ev);
    memcpy(out, ev, (size_t)(kx + ky + 1) * 32);
    return SC_OK;
}
int sc_prover_set_timing(sc_prover* p, int enabled) {
    NEED_HANDLE(p);
    if (!p->group.empty()) {
        for (sc_prover* q : p->group) {
            int rc = sc_prover_set_timing(q, enabled);
            if (rc) return rc;
        }
        return SC_OK;
    }
    p->want_timing = enabled != 0;
    if (p->want_timing) {
        CUDA_TRY(cudaSetDevice(p->device));
        for (auto& e : p->ev)
            if (!e) CUDA_TRY(cudaEventCreate(&e));
    }
    return SC_OK;
}

}  // extern "C"

#include "capi_gkr.inc"
#include "capi_multi.inc"
#include "capi_verify.inc"
