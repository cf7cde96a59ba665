// Host-side probe for the one-shot (pageable -> HBM) upload path: how fast can the box move 1.5 GiB of caller memory to the
// GPU?  Prints: core count, multi-thread memcpy rate into pinned bounce buffers, H2D rates (pinned, pageable,
// cudaHostRegister + copy), cudaMalloc/cudaFree/cudaHostAlloc costs.  Stand-alone: nvcc -O3 -o hostprobe hostprobe.cu -lpthread
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e__)); return 1; } } while (0)

#include <immintrin.h>
static void copy_movsb(char* dst, const char* src, size_t n) { asm volatile("rep movsb" : "+D"(dst), "+S"(src), "+c"(n)::"memory"); }
__attribute__((target("avx2"))) static void copy_nt(char* dst, const char* src, size_t n) {  // 32-byte aligned dst, n % 128 == 0
    for (size_t i = 0; i < n; i += 128) {
        __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
        __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64)), d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
        _mm256_stream_si256((__m256i*)(dst + i), a); _mm256_stream_si256((__m256i*)(dst + i + 32), b);
        _mm256_stream_si256((__m256i*)(dst + i + 64), c); _mm256_stream_si256((__m256i*)(dst + i + 96), d);
    }
    _mm_sfence();
}
static int g_mode = 0;  // 0 memcpy, 1 rep movsb, 2 avx2 non-temporal
static void copy_any(char* dst, const char* src, size_t n) {
    if (g_mode == 0) memcpy(dst, src, n);
    else if (g_mode == 1) copy_movsb(dst, src, n);
    else copy_nt(dst, src, n);
}
static void par_memcpy(char* dst, const char* src, size_t n, int threads) {
    std::vector<std::thread> th;
    size_t per = (n / threads + 4095) & ~(size_t)4095;
    for (int t = 0; t < threads; t++) {
        size_t o = (size_t)t * per;
        if (o >= n) break;
        size_t len = (o + per > n) ? n - o : per;
        th.emplace_back([=] { copy_any(dst + o, src + o, len); });
    }
    for (auto& x : th) x.join();
}

int main() {
    printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
    const size_t total = (size_t)1536 << 20, chunk = (size_t)64 << 20;
    char* page = (char*)malloc(total);
    for (size_t i = 0; i < total; i += 4096) page[i] = (char)i;  // touch
    memset(page, 1, total);
    char *pin = nullptr, *dev = nullptr;
    double t0 = now();
    CK(cudaFree(0));
    printf("context create %.1f ms\n", (now() - t0) * 1e3);
    t0 = now();
    CK(cudaHostAlloc(&pin, 4 * chunk, cudaHostAllocDefault));
    printf("cudaHostAlloc 256 MiB %.2f ms\n", (now() - t0) * 1e3);
    t0 = now();
    CK(cudaMalloc(&dev, total));
    printf("cudaMalloc 1.5 GiB %.2f ms\n", (now() - t0) * 1e3);
    t0 = now();
    CK(cudaFree(dev));
    printf("cudaFree 1.5 GiB %.2f ms\n", (now() - t0) * 1e3);
    t0 = now();
    CK(cudaMalloc(&dev, total));
    printf("cudaMalloc 1.5 GiB again %.2f ms\n", (now() - t0) * 1e3);
    // host memcpy rates into pinned memory
    for (g_mode = 0; g_mode < 3; g_mode++)
    for (int threads : {1, 4, 8, 12, 16}) {
        double best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            t0 = now();
            for (size_t o = 0; o + chunk <= total; o += chunk) par_memcpy(pin + (o / chunk % 4) * chunk, page + o, chunk, threads);
            double dt = now() - t0;
            if (dt < best) best = dt;
        }
        printf("mode %d (0 memcpy, 1 rep movsb, 2 avx2-nt) pageable->pinned %2d threads (spawned per 64 MiB chunk): %.1f GB/s (%.1f ms for 1.5 GiB)\n", threads, g_mode, threads, total / best / 1e9, best * 1e3);
    }
    // H2D pinned
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    for (int rep = 0; rep < 2; rep++) {
        t0 = now();
        for (size_t o = 0; o + chunk <= total; o += chunk) CK(cudaMemcpyAsync(dev + o, pin + (o / chunk % 4) * chunk, chunk, cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        double dt = now() - t0;
        printf("H2D pinned 1.5 GiB: %.1f ms = %.1f GB/s\n", dt * 1e3, total / dt / 1e9);
    }
    // H2D pageable
    for (int rep = 0; rep < 2; rep++) {
        t0 = now();
        CK(cudaMemcpyAsync(dev, page, total, cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        double dt = now() - t0;
        printf("H2D pageable 1.5 GiB (one cudaMemcpyAsync): %.1f ms = %.1f GB/s\n", dt * 1e3, total / dt / 1e9);
    }
    // pipelined bounce: T threads fill pinned slot k while slot k-1 is in flight
    for (g_mode = 0; g_mode < 3; g_mode++)
    for (int threads : {8, 12, 16}) {
        cudaEvent_t ev[4];
        for (auto& e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        const size_t ck = (size_t)32 << 20;
        t0 = now();
        size_t k = 0;
        for (size_t o = 0; o + ck <= total; o += ck, k++) {
            int slot = (int)(k % 4);
            if (k >= 4) CK(cudaEventSynchronize(ev[slot]));
            par_memcpy(pin + slot * chunk, page + o, ck, threads);
            CK(cudaMemcpyAsync(dev + o, pin + slot * chunk, ck, cudaMemcpyHostToDevice, s));
            CK(cudaEventRecord(ev[slot], s));
        }
        CK(cudaStreamSynchronize(s));
        double dt = now() - t0;
        printf("bounce pipeline mode %d (32 MiB chunks, 4 slots, %d copy threads): %.1f ms = %.1f GB/s\n", g_mode, threads, dt * 1e3, total / dt / 1e9);
    }
    // register in place
    t0 = now();
    cudaError_t e = cudaHostRegister(page, total, cudaHostRegisterDefault);
    double treg = now() - t0;
    if (e == cudaSuccess) {
        t0 = now();
        CK(cudaMemcpyAsync(dev, page, total, cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        double dt = now() - t0;
        t0 = now();
        cudaHostUnregister(page);
        printf("cudaHostRegister 1.5 GiB %.1f ms, copy %.1f ms, unregister %.1f ms\n", treg * 1e3, dt * 1e3, (now() - t0) * 1e3);
    } else {
        printf("cudaHostRegister failed: %s\n", cudaGetErrorString(e));
    }
    return 0;
}
