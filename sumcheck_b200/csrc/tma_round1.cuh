// Round 1 (no fold) with TMA-staged tables: same arithmetic as round_kernel<NPTS, false>, but no thread ever waits on a
// global load.  The plain kernel issues two LDG.256 per multiplicand and consumes them at once, with 3 resident warps
// per scheduler: 21 % of its warp time was long-scoreboard stall (profiles/).  Here one lane (the duty rotates over the
// four warps) stages every (tile, multiplicand) work item with one cp.async.bulk.tensor — 64 rows x 128 bytes = 128
// pairs, SWIZZLE_128B — R1_SLOTS-1 items ahead; a thread reads its 64-byte pair from shared memory (4 conflict-free
// LDS.128 through the XOR swizzle) and releases the slot through an mbarrier.
#pragma once
#include "kernels.cuh"
#include "tc_fold.cuh"

namespace sck {

#ifndef SC_R1_WIDE_NPTS
#define SC_R1_WIDE_NPTS 5  // NPTS from which the kernel is built for 2 CTAs/SM (255 registers, no spills)
#endif
constexpr uint32_t R1_SLOTS = 4;
constexpr uint32_t R1_THREADS = 128;
constexpr uint32_t R1_TILE_ROWS = 64;                       // 128-byte rows (two pairs each) per work item
constexpr uint32_t R1_TILE_BYTES = R1_TILE_ROWS * 128;      // 8 KiB
constexpr size_t R1_DYN_SMEM = (size_t)R1_SLOTS * R1_TILE_BYTES;

// M as in round_tc_kernel: 0 = any list of products, M > 0 = one product of M multiplicands with a deferred coefficient (unrolled)
template <int NPTS, int M = 0>
__global__ void __launch_bounds__(R1_THREADS, (NPTS >= SC_R1_WIDE_NPTS ? 2 : 3)) round1_tma_kernel(const RoundParams p) {
    extern __shared__ __align__(1024) uint8_t r1_smem[];
    __shared__ uint32_t s_red[32 * NPTS * 8];
    __shared__ bool s_last;
    __shared__ __align__(8) uint64_t s_full[R1_SLOTS], s_empty[R1_SLOTS];

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // (Measured: parking the accumulators in tensor memory — tcgen05.ld/st around every lazy multiply-accumulate — brings the
    // kernel to 128 registers and 4 CTAs/SM, bit-exact, but NOT faster: 1.135 ms against 1.11 ms at nv = 24.  The kernel is
    // bound by the instruction rate of the multiply pipe, not by latency)
    fr::WideAcc accw[NPTS];
#pragma unroll
    for (int t = 0; t < NPTS; t++) fr::wide_zero(accw[t]);
    if (tid == 0) {
        if (tcf::smem_u32(r1_smem) & 1023u) __trap();
#pragma unroll
        for (uint32_t s = 0; s < R1_SLOTS; s++) {
            tcf::mbar_init(&s_full[s], 1);
            tcf::mbar_init(&s_empty[s], R1_THREADS / 32);
        }
        tcf::fence_mbar_init();
    }
    __syncthreads();

    const uint32_t nnz = p.prod_offsets[p.n_products];
    const uint32_t n_tiles = (uint32_t)(p.n_pairs / R1_THREADS);
    const uint32_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const uint32_t Q = my_tiles * nnz;
    // this thread's 64 bytes inside a tile: row tid/2, half tid%2; 16-byte chunk c of the half sits at (4*half + c) ^ (row % 8)
    const uint32_t row = tid >> 1, half = tid & 1u;
    const uint32_t row_off = row * 128u, sw = row & 7u;

    uint32_t tma_q = 0, tma_tile = p.tile_base + blockIdx.x, tma_jj = 0;  // next item to stage (all threads track the cursors)
    auto issue_tma = [&]() {
        if (lane == 0 && warp == (tma_q & 3u)) {
            const uint32_t slot = tma_q % R1_SLOTS;
            if (tma_q >= R1_SLOTS) tcf::mbar_wait(&s_empty[slot], ((tma_q / R1_SLOTS) - 1u) & 1u);  // previous tenant read by all warps
            const uint32_t idx = p.prod_indices[tma_jj];
            tcf::mbar_expect_tx(&s_full[slot], R1_TILE_BYTES);
            tcf::tma_load_tile(r1_smem + (size_t)slot * R1_TILE_BYTES, (const uint8_t*)p.tmaps + (size_t)idx * 128, &s_full[slot],
                               tma_tile * R1_TILE_ROWS);
        }
        tma_q++;
        if (++tma_jj == nnz) {
            tma_jj = 0;
            tma_tile += gridDim.x;
        }
    };
    for (uint32_t s = 0; s + 1 < R1_SLOTS && s < Q; s++) issue_tma();

    uint32_t q = 0;
    auto item = [&](uint32_t k, bool first, bool last, uint32_t kdeg, bool full, Fr (&prod)[NPTS]) {
        if (tma_q < Q) issue_tma();  // item q + R1_SLOTS - 1 into the slot item q - 1 used
        const uint32_t slot = q % R1_SLOTS;
        tcf::mbar_wait(&s_full[slot], (q / R1_SLOTS) & 1u);
        const uint8_t* src = r1_smem + (size_t)slot * R1_TILE_BYTES + row_off;
        Fr v0, v1;
        {
            const uint4 a = *reinterpret_cast<const uint4*>(src + (((4u * half + 0u) ^ sw) << 4));
            const uint4 b = *reinterpret_cast<const uint4*>(src + (((4u * half + 1u) ^ sw) << 4));
            const uint4 c = *reinterpret_cast<const uint4*>(src + (((4u * half + 2u) ^ sw) << 4));
            const uint4 d = *reinterpret_cast<const uint4*>(src + (((4u * half + 3u) ^ sw) << 4));
            v0.l[0] = a.x; v0.l[1] = a.y; v0.l[2] = a.z; v0.l[3] = a.w; v0.l[4] = b.x; v0.l[5] = b.y; v0.l[6] = b.z; v0.l[7] = b.w;
            v1.l[0] = c.x; v1.l[1] = c.y; v1.l[2] = c.z; v1.l[3] = c.w; v1.l[4] = d.x; v1.l[5] = d.y; v1.l[6] = d.z; v1.l[7] = d.w;
        }
        __syncwarp();
        if (lane == 0) tcf::mbar_arrive(&s_empty[slot]);
        RegAccs<NPTS> accs{accw};
        // round 1 sums d+1 points and delivers raw sums to the host: alternative points 0, 1, inf, -1, 2 (kernels.cuh consume_pair_acc)
        consume_pair_acc<NPTS, false, 0, (M > 0), true>(p, k, first, last, kdeg, full, v0, v1, prod, accs);
        q++;
    };
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (M > 0) {
            Fr prod[NPTS];
#pragma unroll
            for (int jj = 0; jj < (M > 0 ? M : 1); jj++) item(0, jj == 0, jj + 1 == M, (uint32_t)jj + 1, true, prod);
        } else {
            for (uint32_t k = 0; k < p.n_products; k++) {
                Fr prod[NPTS];
                const uint32_t j0 = p.prod_offsets[k], j1 = p.prod_offsets[k + 1];
                for (uint32_t jj = j0; jj < j1; jj++) item(k, jj == j0, jj + 1 == j1, jj - j0 + 1, j1 - j0 == p.degree, prod);
            }
        }
    }
    Fr r = fr::zero();  // no challenge yet (round 1)
    finish_round<NPTS>(p, accw, r, s_red, &s_last);
}

}  // namespace sck
