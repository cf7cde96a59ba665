// Fused tail of the sumcheck prover: all remaining small rounds in one single-CTA launch with the Fiat-Shamir
// transcript on the device.  Compiled in its own translation unit (tail.cu) with FR_COMPACT so that the round loop
// fits the instruction cache.
#pragma once
#include "kernels.cuh"
#include "tail_params.cuh"

namespace sck {

// tail variant: Montgomery values go to global memory, canonical limbs only to shared memory
template <int NPTS>
__device__ __forceinline__ void publish_canon_then_evals(const RoundParams& p, const Fr (&acc)[NPTS], const Fr& r, uint32_t* scratch,
                                                         uint32_t* canon_smem) {
    Fr claim = claim_from_prev(p.prev_evals, p.lagrange, r, p.degree, scratch);
    if (threadIdx.x == 0) {
        Fr c = fr::load(p.coeffs);
#pragma unroll
        for (int t = 0; t < NPTS; t++) {
            Fr v = p.defer_coeff ? fr::mul(acc[t], c) : acc[t];
            const uint32_t slot = (t == 0) ? 0u : (uint32_t)t + 1u;
            Fr cv = to_canonical(v);
            fr::store(p.evals_out + (size_t)slot * 8, v);
#pragma unroll
            for (int i = 0; i < 8; i++) canon_smem[slot * 8 + i] = cv.l[i];
            if (t == 0) {
                Fr p1 = fr::sub(claim, v);
                Fr c1 = to_canonical(p1);
                fr::store(p.evals_out + 8, p1);
#pragma unroll
                for (int i = 0; i < 8; i++) canon_smem[8 + i] = c1.l[i];
            }
        }
    }
}

// ---- fused tail --------------------------------------------------------------------------------------------------
// Once a round has at most TAIL_PAIRS output pairs, one CTA runs ALL remaining rounds in a single launch, including the
// Fiat-Shamir transcript (ml_sumcheck/mod.rs:59-64: prove_round -> feed(prover_msg) -> sample_round) on the device,
// so the ~nv/2 smallest rounds cost no launches, no PCIe round-trips and no host synchronisation.
// Every tail round folds (the tail starts at round >= 2) and uses the P(1)-from-claim shortcut (NPTS = degree).
template <int NPTS>
__global__ void __launch_bounds__(TAIL_THREADS, 1) tail_kernel(const TailParams tp) {
    __shared__ uint32_t s_red[32 * NPTS * 8];
    __shared__ uint32_t s_canon[(MAX_NPTS + 1) * 8];
    __shared__ uint32_t s_r[8];
    __shared__ b2w::WState s_ws;
    __shared__ __align__(16) uint32_t s_foldC[64];
    const uint32_t npts_msg = tp.rp.degree + 1;
    if (threadIdx.x == 0) b2w::from_state(&s_ws, tp.st_in);
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = tp.rp.r[i];
    int cur = tp.cur;
    RoundParams p = tp.rp;
    for (uint32_t rd = 0; rd < tp.n_rounds; rd++) {
        const int nxt = (cur == 1) ? 2 : 1;
        long long tk0 = clock64(), tk1 = 0, tk2 = 0, tk3 = 0;
        p.tab_in = (const uint32_t* const*)tp.ptrs[cur];
        p.tab_out = tp.ptrs[nxt];
        p.n_pairs = tp.n_pairs_first >> rd;
        p.evals_out = tp.evals_all + (size_t)rd * npts_msg * 8;
        p.canon_out = nullptr;  // canonical limbs only go to shared memory here
        if (rd > 0) p.prev_evals = tp.evals_all + (size_t)(rd - 1) * npts_msg * 8;
        fr::WideAcc accw[NPTS];
#pragma unroll
        for (int t = 0; t < NPTS; t++) fr::wide_zero(accw[t]);
        // the first tail round reads tables written by an earlier launch; later ones read what this CTA just wrote
        prepare_fold_consts(r, s_foldC);
        if (rd == 0)
            accumulate_pairs<NPTS, true, true>(p, s_foldC, threadIdx.x, blockDim.x, accw);
        else
            accumulate_pairs<NPTS, true, false>(p, s_foldC, threadIdx.x, blockDim.x, accw);
        tk1 = clock64();
        Fr acc[NPTS];
#pragma unroll
        for (int t = 0; t < NPTS; t++) acc[t] = fr::wide_reduce(accw[t]);
        block_reduce<NPTS>(acc, s_red);
        tk2 = clock64();
        if (threadIdx.x < 32) {
            publish_canon_then_evals<NPTS>(p, acc, r, s_red, s_canon);
            tk3 = clock64();
            if (threadIdx.x == 0) {
                // rng.feed(&prover_msg): u64 length, then d+1 canonical 32-byte little-endian integers
                b2w::absorb_word(&s_ws, (uint64_t)npts_msg);
                for (uint32_t w = 0; w < npts_msg * 4; w++)
                    b2w::absorb_word(&s_ws, (uint64_t)s_canon[2 * w] | ((uint64_t)s_canon[2 * w + 1] << 32));
                uint64_t rr[4];
                b2w::sample_fr(&s_ws, rr);  // sample_round
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    s_r[2 * i] = (uint32_t)rr[i];
                    s_r[2 * i + 1] = (uint32_t)(rr[i] >> 32);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) tp.chal_all[(size_t)rd * 8 + i] = s_r[i];
                if (tp.prof) {
                    long long tk4 = clock64();
                    tp.prof[rd * 4 + 0] = tk1 - tk0; tp.prof[rd * 4 + 1] = tk2 - tk1; tp.prof[rd * 4 + 2] = tk3 - tk2; tp.prof[rd * 4 + 3] = tk4 - tk3;
                }
            }
        }
        __syncthreads();  // also orders this round's table stores before the next round's loads (same CTA)
#pragma unroll
        for (int i = 0; i < 8; i++) r.l[i] = s_r[i];
        cur = nxt;
    }
    if (threadIdx.x == 0) b2w::to_state(&s_ws, tp.st_out);
}

}  // namespace sck
