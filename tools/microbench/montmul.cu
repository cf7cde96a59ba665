// Montgomery-multiplier shoot-out for BLS12-381 Fr on sm_100a.  Each variant is checked against a host
// __int128 CIOS and then timed (modmul/s) so the product kernels can adopt the fastest formulation.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../sumcheck_b200/csrc -o montmul montmul.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "fr.cuh"

typedef unsigned __int128 u128;
static const uint64_t HP[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static void host_mul(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * 0xfffffffeffffffffULL;
        c = (u128)m * HP[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * HP[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    int ge = 1;
    for (int i = 3; i >= 0; i--) { if (t[i] > HP[i]) break; if (t[i] < HP[i]) { ge = 0; break; } }
    if (ge) { u128 bw = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)t[i] - HP[i] - bw; t[i] = (uint64_t)d; bw = (d >> 64) & 1; } }
    for (int i = 0; i < 4; i++) r[i] = t[i];
}

template <int V>
__device__ __forceinline__ fr::Fr mulv(const fr::Fr& a, const fr::Fr& b) {
    if (V == 0) return fr::mul(a, b);
    if (V == 1) return fr::mul_c64(a, b);
    return fr::mul(a, b);
}

template <int V>
__global__ void check_kernel(const uint32_t* a, const uint32_t* b, uint32_t* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr::Fr x = fr::load(a + 8 * i), y = fr::load(b + 8 * i);
    fr::Fr z = mulv<V>(x, y);
    fr::store(out + 8 * i, z);
}

template <int V, int ILP>
__global__ void __launch_bounds__(128) bench_kernel(const uint32_t* a, uint32_t* out, int iters) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    fr::Fr x[ILP], y = fr::load(a + 8 * (i % 1024));
#pragma unroll
    for (int k = 0; k < ILP; k++) x[k] = fr::load(a + 8 * ((i + k * 37) % 1024));
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) x[k] = mulv<V>(x[k], y);
    }
    fr::Fr s = x[0];
#pragma unroll
    for (int k = 1; k < ILP; k++) s = fr::add(s, x[k]);
    fr::store(out + 8 * i, s);
}

template <int V, int ILP>
void bench(const char* name, const uint32_t* da, uint32_t* dout, int sms, int warps_per_sm) {
    int threads = 128, blocks = sms * warps_per_sm / 4, iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    bench_kernel<V, ILP><<<blocks, threads>>>(da, dout, 100);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    bench_kernel<V, ILP><<<blocks, threads>>>(da, dout, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double muls = (double)blocks * threads * iters * ILP;
    printf("%-22s ILP=%d warps/SM=%2d  %7.3f ms  %7.2f G modmul/s  (%s)\n", name, ILP, warps_per_sm, ms, muls / ms * 1e-6,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int n = 1 << 14;
    uint64_t *ha = (uint64_t*)malloc(n * 32), *hb = (uint64_t*)malloc(n * 32), *hr = (uint64_t*)malloc(n * 32), *hd = (uint64_t*)malloc(n * 32);
    uint64_t s = 88172645463325252ULL;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 4; k++) { ha[4 * i + k] = rnd(); hb[4 * i + k] = rnd(); }
        ha[4 * i + 3] &= 0x3fffffffffffffffULL; hb[4 * i + 3] &= 0x3fffffffffffffffULL;  // < 2^254 < p
    }
    // edge values
    for (int k = 0; k < 4; k++) { ha[k] = 0; hb[k] = HP[k]; ha[4 + k] = HP[k]; hb[4 + k] = HP[k]; }
    ha[4] -= 1; hb[4] -= 1; hb[0] -= 1;  // p-1
    ha[8] = 1; ha[9] = ha[10] = ha[11] = 0;
    for (int i = 0; i < n; i++) host_mul(hr + 4 * i, ha + 4 * i, hb + 4 * i);
    fr::fr_init_constants();
    uint32_t *da, *db, *dout;
    cudaMalloc(&da, n * 32); cudaMalloc(&db, n * 32); cudaMalloc(&dout, 1 << 26);
    cudaMemcpy(da, ha, n * 32, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, n * 32, cudaMemcpyHostToDevice);
    for (int v = 0; v < 2; v++) {
        if (v == 0) check_kernel<0><<<n / 128, 128>>>(da, db, dout, n);
        if (v == 1) check_kernel<1><<<n / 128, 128>>>(da, db, dout, n);
        cudaMemcpy(hd, dout, n * 32, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < 4 * n; i++) if (hd[i] != hr[i]) { if (bad < 4) printf("  mismatch v%d at elem %d limb %d: %016llx vs %016llx\n", v, i / 4, i % 4, (unsigned long long)hd[i], (unsigned long long)hr[i]); bad++; }
        printf("variant %d check: %s (%d bad limbs) [%s]\n", v, bad ? "FAIL" : "ok", bad, cudaGetErrorString(cudaGetLastError()));
    }
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    for (int w : {4, 8, 16, 32}) {
        bench<0, 1>("asm even/odd", da, dout, sms, w);
        bench<0, 2>("asm even/odd", da, dout, sms, w);
        bench<0, 4>("asm even/odd", da, dout, sms, w);
        bench<1, 1>("C 64-bit acc", da, dout, sms, w);
        bench<1, 2>("C 64-bit acc", da, dout, sms, w);
        bench<1, 4>("C 64-bit acc", da, dout, sms, w);
    }
    return 0;
}
