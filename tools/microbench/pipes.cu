// Pipe-throughput microbenchmark for sm_100a: which integer multiply form should the
// Montgomery multiplier be built from?  Prints warp-instructions / clk / SM for each variant.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define NACC 8

template <int V>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
    uint32_t a[NACC], b[NACC];
    uint64_t w[NACC];
    double d[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) {
        a[i] = seed * (threadIdx.x + 1) + i;
        b[i] = seed ^ (i * 0x9e3779b9u + threadIdx.x);
        w[i] = ((uint64_t)a[i] << 32) | b[i];
        d[i] = (double)a[i];
    }
    uint32_t x = seed | 1, y = seed * 3 + 7;
    double dx = 1.0000001, dy = 0.5;
    if (V == 10) {
        uint32_t c[32];
#pragma unroll
        for (int i = 0; i < 32; i++) c[i] = a[i % 8] + i;
        for (int it = 0; it < ITERS; it++) {
#pragma unroll
            for (int g = 0; g < 4; g++) {  // 4 independent chains of 4 wide-mads = 16 IMAD.WIDE per iteration
                asm volatile(
                    "mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
                    "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
                    "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
                    "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
                    "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
                    "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
                    "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
                    "madc.hi.u32 %7, %11, %12, %7;\n\t"
                    : "+r"(c[g * 8 + 0]), "+r"(c[g * 8 + 1]), "+r"(c[g * 8 + 2]), "+r"(c[g * 8 + 3]),
                      "+r"(c[g * 8 + 4]), "+r"(c[g * 8 + 5]), "+r"(c[g * 8 + 6]), "+r"(c[g * 8 + 7])
                    : "r"(a[g]), "r"(a[g + 1]), "r"(a[g + 2]), "r"(a[g + 3]), "r"(b[g]));
            }
        }
#pragma unroll
        for (int i = 0; i < 32; i++) a[i % 8] ^= c[i];
    } else
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (V == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(y), "r"(x));
            if (V == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(y), "r"(x));
            if (V == 2) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(y)); a[i] = (uint32_t)w[i]; }
            if (V == 3) {  // carry-chained lo/hi pair (2 IMAD-class instructions)
                asm volatile("{ .reg .u32 t; mad.lo.cc.u32 t, %0, %2, %0; madc.hi.cc.u32 %1, %0, %2, %1; mov.u32 %0, t; }"
                             : "+r"(a[i]), "+r"(b[i]) : "r"(y));
            }
            if (V == 4) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dx), "d"(dy));
            if (V == 5) {  // 1 IMAD + 1 IADD3: do the pipes co-issue?
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(y), "r"(x));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(x));
            }
            if (V == 6) asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(x));
            if (V == 7) {  // wide mad + carry add into a third limb: Comba step with IMAD.WIDE
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(y));
                asm volatile("add.cc.u32 %0, %0, %1; addc.u32 %1, %1, 0;" : "+r"(a[i]), "+r"(b[i]));
            }
            if (V == 8) {  // 1 IMAD.WIDE + 1 DFMA: separate pipes?
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(y)); a[i] = (uint32_t)w[i];
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(dx), "d"(dy));
            }
            if (V == 9) {  // 64-bit carry chain add (IADD3.X pairs)
                asm volatile("add.cc.u32 %0, %0, %2; addc.cc.u32 %1, %1, %2;" : "+r"(a[i]), "+r"(b[i]) : "r"(x));
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) r ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int V>
void run(const char* name, int instr_per_step, uint32_t* out, int sms, double mhz_hint) {
    int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<V><<<blocks, threads>>>(out, 12345);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<V><<<blocks, threads>>>(out, 12345);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_instrs = (double)blocks * (threads / 32) * ITERS * (V == 10 ? 16 : NACC) * instr_per_step;
    double per_s = warp_instrs / (ms * 1e-3);
    printf("%-34s %8.3f ms  %8.2f Gwarp-instr/s  %6.3f warp-instr/clk/SM @%.0f MHz  (lane-ops/clk/SM %.1f)\n",
           name, ms, per_s * 1e-9, per_s / sms / (mhz_hint * 1e6), mhz_hint, per_s / sms / (mhz_hint * 1e6) * 32);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double mhz = clk_khz / 1000.0;
    printf("device %s, %d SMs, max clock %.0f MHz (rates are normalised to max clock; real clock may be lower)\n",
           p.name, p.multiProcessorCount, mhz);
    uint32_t* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    int sms = p.multiProcessorCount;
    for (int rep = 0; rep < 2; rep++) {
        run<0>("mad.lo.u32 (IMAD)", 1, out, sms, mhz);
        run<1>("mad.hi.u32 (IMAD.HI)", 1, out, sms, mhz);
        run<2>("mad.wide.u32 (IMAD.WIDE)", 1, out, sms, mhz);
        run<3>("mad.lo.cc+madc.hi.cc pair", 2, out, sms, mhz);
        run<4>("fma.rn.f64 (DFMA)", 1, out, sms, mhz);
        run<5>("IMAD + IADD (2 instr)", 2, out, sms, mhz);
        run<6>("add.u32 (IADD3)", 1, out, sms, mhz);
        run<7>("IMAD.WIDE + add.cc/addc (3 instr)", 3, out, sms, mhz);
        run<8>("IMAD.WIDE + DFMA (2 instr)", 2, out, sms, mhz);
        run<9>("add.cc+addc.cc pair", 2, out, sms, mhz);
        run<10>("IMAD.WIDE.X carry chain (per WIDE)", 1, out, sms, mhz);
    }
    return 0;
}
