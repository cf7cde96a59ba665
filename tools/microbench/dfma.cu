// Round-2 material: is a 256x256-bit integer product cheaper on the FP64 pipe than on the IMAD.WIDE pipe of a B200?
// tools/microbench/pipes.cu measured DFMA at 62.6 lane-ops/clk/SM against 30.6 for IMAD.WIDE (same pipe), so a product
// in 5 x 52-bit limbs — 25 limb products, each 2 DFMA + 1 DADD with the classic hi/lo trick
//     hi = fma_rz(a, b, 2^104);  lo = fma_rz(a, b, (2^104 + 2^52) - hi)     (both exact; the mantissas hold the halves)
// plus 64-bit integer column sums on the ALU pipe — occupies the pipe for 150 cycles per warp instead of 256 for the 64
// IMAD.WIDE of the 32-bit schoolbook product.  This measures both as dependent chains at full occupancy and checks
// the FP64 product bit for bit against the integer one.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I sumcheck_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fr.cuh"

constexpr int ITER = 256;
constexpr double C1 = 20282409603651670423947251286016.0;                    // 2^104
constexpr double C2 = 20282409603651670423947251286016.0 + 4503599627370496.0;  // 2^104 + 2^52
constexpr unsigned long long M52 = (1ull << 52) - 1;
constexpr unsigned long long EXP_LO = 0x433ull << 52;  // exponent field of 2^52
constexpr unsigned long long EXP_HI = 0x467ull << 52;  // exponent field of 2^104

// columns[k] = sum_{i+j=k} lo(a_i b_j) + sum_{i+j=k-1} hi(a_i b_j), k = 0..9 (each < 2^56)
__device__ __forceinline__ void mul52(const double (&a)[5], const double (&b)[5], unsigned long long (&col)[10]) {
#pragma unroll
    for (int k = 0; k < 10; k++) col[k] = 0;
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const double hi = __fma_rz(a[i], b[j], C1);
            const double lo = __fma_rz(a[i], b[j], C2 - hi);
            col[i + j] += (unsigned long long)__double_as_longlong(lo);
            col[i + j + 1] += (unsigned long long)__double_as_longlong(hi);
        }
    // remove the exponent fields: column k received n_lo(k) low halves and n_hi(k) high halves
#pragma unroll
    for (int k = 0; k < 10; k++) {
        const int n_lo = (k <= 4) ? k + 1 : ((k <= 8) ? 9 - k : 0);
        const int n_hi = (k >= 1 && k <= 5) ? k : ((k >= 6) ? 10 - k : 0);
        col[k] -= (unsigned long long)n_lo * EXP_LO + (unsigned long long)n_hi * EXP_HI;
    }
}

__device__ __forceinline__ double to_double52(unsigned long long x) {  // x < 2^52, exact
    return __longlong_as_double((long long)(x | EXP_LO)) - 4503599627370496.0;
}

__global__ void __launch_bounds__(256) dfma_chain(const unsigned long long* in, unsigned long long* out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double a[5], b[5];
#pragma unroll
    for (int i = 0; i < 5; i++) { a[i] = to_double52(in[t * 10 + i] & M52); b[i] = to_double52(in[t * 10 + 5 + i] & M52); }
    unsigned long long col[10];
    for (int it = 0; it < ITER; it++) {
        mul52(a, b, col);
#pragma unroll
        for (int i = 0; i < 5; i++) a[i] = to_double52((col[i] ^ col[i + 5]) & M52);  // keep the chain dependent
    }
#pragma unroll
    for (int k = 0; k < 10; k++) out[t * 10 + k] = col[k];
}

__global__ void __launch_bounds__(256) imad_chain(const uint32_t* in, uint32_t* out) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t a[8], b[8], ev[16], od[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = in[t * 16 + i]; b[i] = in[t * 16 + 8 + i]; }
    for (int it = 0; it < ITER; it++) {
        fr::mul_wide_eo(ev, od, a, b);
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = ev[i] ^ od[i] ^ ev[i + 8];
    }
#pragma unroll
    for (int i = 0; i < 16; i++) out[t * 16 + i] = ev[i] + od[i];
}

// one product each way on the same 260-bit inputs, for the host to compare
__global__ void check_kernel(const unsigned long long* in, unsigned long long* out, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double a[5], b[5];
    for (int i = 0; i < 5; i++) { a[i] = to_double52(in[t * 10 + i] & M52); b[i] = to_double52(in[t * 10 + 5 + i] & M52); }
    unsigned long long col[10];
    mul52(a, b, col);
    for (int k = 0; k < 10; k++) out[t * 10 + k] = col[k];
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)
static uint64_t rng_state = 42;
static uint64_t rnd() { uint64_t z = (rng_state += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }

int main() {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // ---- exactness: FP64 columns against __int128 arithmetic on the host
    const int NCHK = 4096;
    std::vector<unsigned long long> hin((size_t)NCHK * 10), hout((size_t)NCHK * 10);
    for (auto& v : hin) v = rnd() & M52;
    for (int i = 0; i < 10; i++) { hin[i] = M52; hin[10 + i] = 0; hin[20 + i] = (i < 5) ? M52 : 1; }  // extremes
    unsigned long long *din, *dout;
    CK(cudaMalloc(&din, hin.size() * 8)); CK(cudaMalloc(&dout, hout.size() * 8));
    CK(cudaMemcpy(din, hin.data(), hin.size() * 8, cudaMemcpyHostToDevice));
    check_kernel<<<(NCHK + 127) / 128, 128>>>(din, dout, NCHK);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hout.data(), dout, hout.size() * 8, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (int t = 0; t < NCHK; t++) {
        unsigned long long want[10] = {0};
        for (int i = 0; i < 5; i++)
            for (int j = 0; j < 5; j++) {
                const unsigned __int128 pr = (unsigned __int128)hin[t * 10 + i] * hin[t * 10 + 5 + j];
                want[i + j] += (unsigned long long)(pr & M52);
                want[i + j + 1] += (unsigned long long)(pr >> 52);
            }
        for (int k = 0; k < 10; k++)
            if (want[k] != hout[t * 10 + k] && bad++ < 4) printf("  product %d column %d: got %llx want %llx\n", t, k, hout[t * 10 + k], want[k]);
    }
    printf("exactness: %d products, %zu bad columns\n", NCHK, bad);
    // ---- throughput: dependent chains, 8 warps per scheduler
    const int blocks = sms * 8, threads = 256;
    const size_t n = (size_t)blocks * threads;
    std::vector<unsigned long long> big(n * 10);
    for (auto& v : big) v = rnd();
    unsigned long long *d1, *d2;
    uint32_t *d3, *d4;
    CK(cudaMalloc(&d1, n * 80)); CK(cudaMalloc(&d2, n * 80)); CK(cudaMalloc(&d3, n * 64)); CK(cudaMalloc(&d4, n * 64));
    CK(cudaMemcpy(d1, big.data(), n * 80, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d3, big.data(), n * 64, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_d = 0, ms_i = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); dfma_chain<<<blocks, threads>>>(d1, d2); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms_d, e0, e1);
        cudaEventRecord(e0); imad_chain<<<blocks, threads>>>(d3, d4); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms_i, e0, e1);
    }
    const double prods = (double)n * ITER;
    printf("256x256-bit products: FP64 (5x52-bit limbs, 75 FP64 ops + int64 columns) %.3f ms = %.1f G/s | IMAD.WIDE (64 wide MACs) %.3f ms = %.1f G/s | ratio %.2f\n",
           ms_d, prods / ms_d / 1e6, ms_i, prods / ms_i / 1e6, ms_i / ms_d);
    printf(bad ? "DFMA FAIL\n" : "DFMA OK\n");
    return bad ? 1 : 0;
}
