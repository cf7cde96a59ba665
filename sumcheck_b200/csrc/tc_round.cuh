// Fold rounds with TMA-staged tables and the fix_variables fold on the tensor cores (see tc_fold.cuh for the algebra).
//
// One launch per protocol round, as round_kernel<NPTS, true>: same inputs, same outputs, same epilogue (finish_round).
// What changes is how a CTA gets the folded pair (new[2b], new[2b+1]) of every multiplicand:
//   * the table tile (128 rows x 128 bytes = old[4b..4b+3] for 128 consecutive b) is copied HBM -> shared memory by TMA
//     (cp.async.bulk.tensor, SWIZZLE_128B) into a ring of TC_SLOTS slots, TC_SLOTS work items ahead of its use — no
//     thread waits on a global load (the plain kernel spent 20 % of its warp time in long-scoreboard stalls);
//   * one lane (the duty rotates over the four warps) issues 4 tcgen05.mma.kind::i8 per tile (bytes of the tile x the
//     round's constants matrix) into one of two 64-column accumulators in tensor memory, one work item ahead;
//   * thread t (= TMEM lane t = row t of the tile) reads its 2 x 32 column sums, carries/reduces them to the two field
//     elements (tcf::columns_to_fr), stores them (the folded table, 64 B per thread) and feeds the product terms.
// A work item is one (tile, CSR entry of ProverState.list_of_products) — so any list of products works, tables shared
// between products are simply staged again.  The IMAD.WIDE work per pair at degree 3 drops from 981 to 573.
#pragma once
#include "kernels.cuh"
#include "tc_fold.cuh"

namespace sck {

#ifndef SC_TC_SLOTS
#define SC_TC_SLOTS 3
#endif
#ifndef SC_TC_WIDE_NPTS
#define SC_TC_WIDE_NPTS 4  // NPTS from which the kernel is built for 2 CTAs/SM (up to 255 registers) instead of 3 (168):
                           // measured on config 4 (d = 4): 7.92 -> 7.22 ms; at degree 3 two CTAs are 10 % slower
#endif
#ifndef SC_TC_MIN_BLOCKS
#define SC_TC_MIN_BLOCKS 3
#endif
constexpr uint32_t TC_SLOTS = SC_TC_SLOTS;
constexpr uint32_t TC_THREADS = tcf::TILE_ROWS;
constexpr uint32_t TC_TMEM_COLS = 2 * tcf::ACC_COLS;  // power of two >= 32
constexpr size_t TC_DYN_SMEM = (size_t)TC_SLOTS * tcf::TILE_BYTES + tcf::BMAT_BYTES;
constexpr unsigned long long TC_MIN_PAIRS = 1ull << 14;  // smaller rounds are latency-bound: plain kernel

// M = 0: any list of products (CSR).  M > 0: ONE product of M multiplicands with a deferred coefficient — the loops over products
// and multiplicands are unrolled, so first / last / kdeg are compile-time and the per-pair control flow disappears.
template <int NPTS, int M = 0>
__global__ void __launch_bounds__(TC_THREADS, (NPTS >= SC_TC_WIDE_NPTS ? 2 : SC_TC_MIN_BLOCKS)) round_tc_kernel(const RoundParams p) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];  // [TC_SLOTS] tiles, then the constants matrix
    __shared__ uint32_t s_red[32 * NPTS * 8];
    __shared__ bool s_last;
    __shared__ __align__(8) uint64_t s_full[TC_SLOTS], s_done[2], s_empty[2];
    __shared__ uint32_t s_tmem;

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* const bmat = tc_smem + (size_t)TC_SLOTS * tcf::TILE_BYTES;
    fr::WideAcc accw[NPTS];
#pragma unroll
    for (int t = 0; t < NPTS; t++) fr::wide_zero(accw[t]);
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = p.r[i];

    if (tid == 0) {
        if (tcf::smem_u32(tc_smem) & 1023u) __trap();  // SWIZZLE_128B atoms must be 1024-byte aligned
#pragma unroll
        for (uint32_t s = 0; s < TC_SLOTS; s++) tcf::mbar_init(&s_full[s], 1);
        tcf::mbar_init(&s_done[0], 1);
        tcf::mbar_init(&s_done[1], 1);
        tcf::mbar_init(&s_empty[0], TC_THREADS / 32);
        tcf::mbar_init(&s_empty[1], TC_THREADS / 32);
        tcf::fence_mbar_init();
    }
    if (warp == 0) tcf::tmem_alloc(&s_tmem, TC_TMEM_COLS);
    tcf::build_bmat(r, bmat);
    tcf::fence_proxy_async_smem();
    tcf::tc_fence_before();
    __syncthreads();
    tcf::tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t lane_taddr = tmem + ((warp * 32u) << 16);
    const uint32_t tiles_smem = tcf::smem_u32(tc_smem), bmat_smem = tcf::smem_u32(bmat);

    const uint32_t nnz = p.prod_offsets[p.n_products];
    const uint32_t n_tiles = (uint32_t)(p.n_pairs / tcf::TILE_ROWS);
    const uint32_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const uint32_t Q = my_tiles * nnz;  // work items of this CTA, in the order the loops below visit them

    // ---- producer side -----------------------------------------------------------------------------------------------
    // Staging (TMA) and multiplying (tcgen05.mma) item q are done by lane 0 of warp q % 4: the duty rotates so that no warp
    // falls behind the others (with a fixed producer thread the other three warps ran one item ahead and then waited ~900
    // cycles per item for it: 14 % of all warp time in the first measured version).  All threads track the cursors.
    uint32_t tma_q = 0, tma_tile = blockIdx.x, tma_jj = 0;  // next item to stage
    auto issue_tma = [&]() {
        if (lane == 0 && warp == (tma_q & 3u)) {
            const uint32_t slot = tma_q % TC_SLOTS;
            const uint32_t idx = p.prod_indices[tma_jj];
            tcf::mbar_expect_tx(&s_full[slot], tcf::TILE_BYTES);
            tcf::tma_load_tile(tc_smem + (size_t)slot * tcf::TILE_BYTES, (const uint8_t*)p.tmaps + (size_t)idx * 128, &s_full[slot],
                               tma_tile * tcf::TILE_ROWS);
        }
        tma_q++;
        if (++tma_jj == nnz) {
            tma_jj = 0;
            tma_tile += gridDim.x;
        }
    };
    auto issue_mma = [&](uint32_t q) {
        if (lane == 0 && warp == (q & 3u)) {
            const uint32_t slot = q % TC_SLOTS, a = q & 1u;
            tcf::mbar_wait(&s_full[slot], (q / TC_SLOTS) & 1u);              // the tile has landed
            if (q >= 2) tcf::mbar_wait(&s_empty[a], ((q >> 1) - 1u) & 1u);   // item q-2 has been read out of this accumulator
            tcf::tc_fence_after();
            tcf::issue_fold_mma(tiles_smem + slot * tcf::TILE_BYTES, bmat_smem, tmem + a * tcf::ACC_COLS);
            tcf::umma_commit(&s_done[a]);
        }
    };
    for (uint32_t s = 0; s < TC_SLOTS && s < Q; s++) issue_tma();
    issue_mma(0);

    uint32_t q = 0;
    // one work item: wait for its accumulator, read the two folded elements out of tensor memory, store them, multiply them in
    auto item = [&](uint32_t k, uint32_t jj, bool first, bool last, uint32_t kdeg, bool full, unsigned long long b, Fr (&prod)[NPTS]) {
        if (q + 1 < Q) issue_mma(q + 1);  // one item ahead: overlaps this item's arithmetic
        const uint32_t a = q & 1u;
        tcf::mbar_wait(&s_done[a], (q >> 1) & 1u);
        tcf::tc_fence_after();
        // the tile's shared-memory slot is free again (its MMAs completed): stage item q + TC_SLOTS into it
        if (tma_q < Q) issue_tma();
        uint32_t S[32];
        tcf::tmem_ld32(lane_taddr + a * tcf::ACC_COLS, S);
        tcf::tmem_ld_wait();
        const Fr v0 = tcf::columns_to_fr(S);
        tcf::tmem_ld32(lane_taddr + a * tcf::ACC_COLS + 32, S);
        tcf::tmem_ld_wait();
        tcf::tc_fence_before();
        __syncwarp();
        if (lane == 0) tcf::mbar_arrive(&s_empty[a]);
        const Fr v1 = tcf::columns_to_fr(S);
        if (p.write_fold && p.prod_first[jj]) {
            uint32_t* dst = p.tab_out[p.prod_indices[jj]] + b * 16;
            fr::store(dst, v0);
            fr::store(dst + 8, v1);
        }
        RegAccs<NPTS> accs{accw};
        // these rounds always skip P(1) and deliver raw sums to the host: alternative points (kernels.cuh consume_pair_acc)
        consume_pair_acc<NPTS, false, 1, (M > 0), true>(p, k, first, last, kdeg, full, v0, v1, prod, accs);
        q++;
    };
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long b = (unsigned long long)tile * tcf::TILE_ROWS + tid;
        if (M > 0) {
            Fr prod[NPTS];
#pragma unroll
            for (int jj = 0; jj < (M > 0 ? M : 1); jj++) item(0, (uint32_t)jj, jj == 0, jj + 1 == M, (uint32_t)jj + 1, true, b, prod);
        } else {
            for (uint32_t k = 0; k < p.n_products; k++) {
                Fr prod[NPTS];
                const uint32_t j0 = p.prod_offsets[k], j1 = p.prod_offsets[k + 1];
                for (uint32_t jj = j0; jj < j1; jj++) item(k, jj, jj == j0, jj + 1 == j1, jj - j0 + 1, j1 - j0 == p.degree, b, prod);
            }
        }
    }
    tcf::tc_fence_before();
    __syncthreads();
    if (warp == 0) tcf::tmem_dealloc(tmem, TC_TMEM_COLS);
    finish_round<NPTS>(p, accw, r, s_red, &s_last);
}

}  // namespace sck
