// Host side of the upload of caller tables that live in PAGEABLE memory (a Rust Vec<F>): prover_init's deep copy
// (prover.rs:55-59) becomes pageable -> pinned bounce slot -> HBM.  cudaMemcpyAsync from pageable memory runs at 11 GB/s on
// the B200 box (one driver thread staging through its own bounce buffer); the PCIe 5 x16 link carries 55 GB/s.  Measured
// there (tools/microbench/hostprobe.cu, 16 vCPUs): 12 threads filling pinned slots reach 65 GB/s alone and 47 GB/s while the
// DMA engine reads the previous slot (host DRAM bandwidth is shared by the copy's reads, its writes and the DMA's reads).
// This is a memcpy, not a CPU path of the protocol: no field arithmetic happens here.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace hcopy {

#if defined(__x86_64__)
// dst 32-byte aligned, n a multiple of 128: streaming stores, no read-for-ownership of the destination
__attribute__((target("avx2"))) inline void copy_nt_avx2(uint8_t* dst, const uint8_t* src, size_t n) {
    for (size_t i = 0; i < n; i += 128) {
        const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
        const __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64)), d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
        _mm256_stream_si256((__m256i*)(dst + i), a);
        _mm256_stream_si256((__m256i*)(dst + i + 32), b);
        _mm256_stream_si256((__m256i*)(dst + i + 64), c);
        _mm256_stream_si256((__m256i*)(dst + i + 96), d);
    }
    _mm_sfence();
}
inline bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}
#endif

// Streaming stores by default; SC_COPY_NT=0 selects cached stores.  Measured end to end on the B200 box (one-shot nv = 24 proof
// from pageable tables, tools/e2e_sweep.sh; PCIe alone 29.1 ms): 12 threads + streaming stores 34.8-35.3 ms in every run; 8
// threads + cached stores 33.2 ms in one process and 64 ms in another (the DMA engine sometimes takes the fresh lines from the
// last-level cache and sometimes fights the copy for them) — the stable setting is the default.
inline bool use_nt() {
    static const bool v = !(getenv("SC_COPY_NT") && atoi(getenv("SC_COPY_NT")) == 0);
    return v;
}

inline void copy_piece(uint8_t* dst, const uint8_t* src, size_t n) {
#if defined(__x86_64__)
    if (use_nt() && have_avx2() && ((uintptr_t)dst & 31) == 0) {
        const size_t body = n & ~(size_t)127;
        copy_nt_avx2(dst, src, body);
        if (n > body) memcpy(dst + body, src + body, n - body);
        return;
    }
#endif
    memcpy(dst, src, n);
}

// A small persistent pool: workers sleep on a condition variable between uploads, so an idle library holds no core.
class Pool {
public:
    static Pool& get() {
        static Pool* p = new Pool();  // leaked on purpose: worker threads must not be joined from a static destructor
        return *p;
    }
    int threads() const { return (int)workers_.size() + 1; }

    // dst[0..n) = src[0..n), split into pieces taken by the workers and the calling thread; returns when all are done
    void copy(uint8_t* dst, const uint8_t* src, size_t n) {
        size_t piece = (size_t)1 << 20;
        if (n < piece * (workers_.size() + 1)) piece = ((n / (workers_.size() + 1)) + 4095) & ~(size_t)4095;  // small slots: one piece per thread
        if (workers_.empty() || n <= (size_t)256 << 10) {
            copy_piece(dst, src, n);
            return;
        }
        std::lock_guard<std::mutex> serial(job_mu_);  // one job at a time (handles on different threads share the pool)
        Job job;
        {
            std::lock_guard<std::mutex> lk(mu_);
            job_.dst = dst; job_.src = src; job_.n = n; job_.piece = piece;
            job_.n_pieces = (n + piece - 1) / piece;
            job_.gen = ++generation_;
            // the claim counter carries the job's generation in its high half, so a worker that is late leaving the
            // previous job can neither claim nor skip a piece of this one
            next_.store((uint64_t)(job_.gen & 0xffffffffu) << 32, std::memory_order_relaxed);
            done_.store(0, std::memory_order_relaxed);
            job = job_;
            gen_hint_.store(job_.gen, std::memory_order_relaxed);
        }
        cv_.notify_all();
        work(job);
        // the caller finished its share: wait for the stragglers (each at most one piece)
        while (done_.load(std::memory_order_acquire) < job.n_pieces) {
#if defined(__x86_64__)
            _mm_pause();
#endif
        }
    }

private:
    Pool() {
        unsigned hw = std::thread::hardware_concurrency();
        int want = hw >= 16 ? 12 : (hw > 2 ? (int)(hw * 3 / 4) : 1);
        if (const char* e = getenv("SC_COPY_THREADS")) want = atoi(e);
        if (want < 1) want = 1;
        if (want > 64) want = 64;
        for (int i = 1; i < want; i++) workers_.emplace_back([this] { loop(); });
        for (auto& t : workers_) t.detach();
    }
    struct Job {
        uint8_t* dst = nullptr;
        const uint8_t* src = nullptr;
        size_t n = 0, piece = 0, n_pieces = 0;
        unsigned long long gen = 0;
    };
    void work(const Job& j) {
        const uint64_t tag = (uint64_t)(j.gen & 0xffffffffu) << 32;
        for (;;) {
            uint64_t v = next_.load(std::memory_order_relaxed);
            if ((v & 0xffffffff00000000ull) != tag) return;  // a newer job owns the counter
            const size_t k = (size_t)(v & 0xffffffffu);
            if (k >= j.n_pieces) return;
            if (!next_.compare_exchange_weak(v, v + 1, std::memory_order_relaxed)) continue;
            const size_t o = k * j.piece, len = (o + j.piece > j.n) ? j.n - o : j.piece;
            copy_piece(j.dst + o, j.src + o, len);
            done_.fetch_add(1, std::memory_order_release);
        }
    }
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            Job j;
            // the next slot of an upload arrives within a millisecond: poll for a moment before going to sleep
            for (int spin = 0; spin < 20000 && gen_hint_.load(std::memory_order_relaxed) == seen; spin++) {
#if defined(__x86_64__)
                _mm_pause();
#endif
            }
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                j = job_;
            }
            work(j);
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, job_mu_;
    std::condition_variable cv_;
    unsigned long long generation_ = 0;
    Job job_;
    std::atomic<uint64_t> next_{0};
    std::atomic<unsigned long long> gen_hint_{0};
    std::atomic<size_t> done_{0};
};

}  // namespace hcopy
