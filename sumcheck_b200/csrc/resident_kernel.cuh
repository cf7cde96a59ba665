// Resident rounds: every protocol round with at most RES_MAX_PAIRS output pairs runs inside ONE cooperative launch that
// stays on the GPU until the proof ends (VERDICT r1 "next" #4).  A launch per round costs ~21 us of kernel plus ~4 us of
// host work however little data the round has (launch, constants, partials round trip, PCIe publication), and a proof has
// nv - 10 such rounds.  Here the Fiat-Shamir transcript stays on the host (ml_sumcheck/mod.rs:59-64: prove_round ->
// feed(prover_msg) -> sample_round; Blake2b over ~100 bytes is ~1 us of CPU) and the two sides talk through mapped pinned
// memory with flag-in-data words ({value, sequence number} in one 8-byte access, kernels.cuh mail_store/mail_load):
//
//   host  -> device   the round's fold constants C[k] = r * 2^(32k+64) mod p (fr.cuh fold_const; 64 limbs), computed by
//                     the host from the challenge it has just drawn; CTA 0 polls them over PCIe and re-publishes them in
//                     device memory for the other CTAs
//   device -> host    the round's sums for P(0), P(2), .., P(d) as UNREDUCED 17-limb integers (the lazily accumulated products,
//                     added limb-wise over the threads: kernels.cuh block_sum_wide); the host Montgomery-reduces them and
//                     finishes the message (deferred coefficient, P(1) from the claim, canonical forms — host_fr.h)
//
// Per round every active CTA folds + sums its pairs (accumulate_pairs, the same arithmetic as round_kernel<NPTS, true>),
// the last CTA to arrive adds the per-CTA sums and publishes.  Inside a pair the two folds of a table and the d points of
// a multiplicand go through ILP-interleaved out-of-line routines (fr.cuh mul_round_const_x2, mul_lazy_n).  A round's tables are written by other CTAs of the same
// launch, so they are read with ld.global.cg (L2) and every thread fences its stores before the block's arrival.
// No grid-wide barrier is needed: a CTA proceeds to round k+1 only when it sees that round's constants, and the host
// sends those only after the LAST arrival of round k — every store of round k happens-before every load of round k+1.
// CTAs beyond the round's need exit for good (the grid shrinks with the tables).
#pragma once
#include "kernels.cuh"
#include "tail_params.cuh"

namespace sck {

template <int NPTS>
__global__ void __launch_bounds__(RES_THREADS, 2) resident_kernel(const ResidentParams q) {
    constexpr int NW = NPTS * WL;
    __shared__ unsigned long long s_part[(RES_THREADS / 32) * NW], s_tot[NW];
    __shared__ uint32_t s_out[NW];
    __shared__ uint32_t s_rows[16 * NW];  // sharded: the ranks' sums (n_ranks <= 16)
    __shared__ __align__(16) uint32_t s_foldC[RES_CONST_WORDS];
    __shared__ uint32_t s_slots[(RES_THREADS / 32) * 32 * 8];  // fine-grained rounds: one folded element per lane
    __shared__ volatile int s_flag;
    const uint32_t tid = threadIdx.x, c = blockIdx.x;
    // CTAs a round needs: RES_THREADS pairs per CTA, or (fine-grained) RES_THREADS >> lpp_log2
    auto ctas_for = [&](uint32_t rd, bool& fine) {
        const unsigned long long n = q.n_pairs_first >> rd;
        fine = q.fine_max_pairs && n <= q.fine_max_pairs;
        const unsigned long long per = fine ? (unsigned long long)(RES_THREADS >> q.lpp_log2) : (unsigned long long)RES_THREADS;
        const unsigned long long want = (n + per - 1) / per;
        return want < (unsigned long long)gridDim.x ? (want ? (uint32_t)want : 1u) : gridDim.x;
    };
    RoundParams p = q.rp;
    int cur = q.cur;
    if (tid == 0) s_flag = 0;
    __syncthreads();
    for (uint32_t rd = 0; rd < q.n_rounds; rd++) {
        const unsigned long long n_pairs = q.n_pairs_first >> rd;
        bool fine;
        const uint32_t n_act = ctas_for(rd, fine);
        if (c >= n_act) {
            // no pairs for this CTA in this round; the fine-grained rounds that follow may need it again (they use more CTAs
            // per pair), otherwise it leaves for good
            uint32_t later = 0;
            for (uint32_t r2 = rd + 1; r2 < q.n_rounds; r2++) {
                bool f2;
                const uint32_t a2 = ctas_for(r2, f2);
                later = a2 > later ? a2 : later;
            }
            if (c >= later) return;
            cur = (cur == 1) ? 2 : 1;
            continue;
        }
        const uint32_t seq = q.seq0 + rd;
        const long long tk0 = clock64();
        // ---- this round's fold constants: CTA 0 from the host (PCIe reads), the others from CTA 0's copy in device memory
        if (tid < RES_CONST_WORDS) {
            const uint32_t* src = (c == 0 ? q.h_consts : q.d_bcast) + 2 * tid;
            uint32_t v = 0, f = 0;
            const long long t0 = clock64();
            for (;;) {
                mail_load(src, v, f);
                if (f == seq) break;
                if (tid == 0) {  // the host abandoned the proof (an error path): leave at once, and tell the other CTAs
                    uint32_t a;
                    if (c == 0) {
                        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(a) : "l"(q.h_abort) : "memory");
                        if (a == q.seq0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(q.d_abort), "r"(a) : "memory");
                    } else {
                        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(a) : "l"(q.d_abort) : "memory");
                    }
                    if (a == q.seq0) {
                        s_flag = -1;
                        break;
                    }
                }
                if (s_flag < 0 || clock64() - t0 > q.timeout) {  // ... or it went away: leave instead of hanging the GPU
                    s_flag = -1;
                    break;
                }
            }
            s_foldC[tid] = v;
            if (c == 0 && n_act > 1 && f == seq) mail_store(q.d_bcast + 2 * tid, v, seq);
        }
        __syncthreads();
        if (s_flag < 0) {
            if (tid == 0) *q.h_error = 1;
            return;
        }
        const long long tk1 = clock64();
        const int nxt = (cur == 1) ? 2 : 1;
        p.tab_in = (const uint32_t* const*)q.ptrs[cur];
        p.tab_out = q.ptrs[nxt];
        p.n_pairs = n_pairs;
        fr::WideAcc accw[NPTS];
        long long tk2;
        if (fine) {
            // several lanes per pair; a CTA covers RES_THREADS >> lpp_log2 pairs per sweep
            const uint32_t per_warp = 32u >> q.lpp_log2, per_cta = (RES_THREADS / 32) * per_warp;
            fr::WideAcc mine, part;
            fr::wide_zero(mine);
            int my_pt = -1;
            for (unsigned long long b0 = (unsigned long long)c * per_cta; b0 < n_pairs; b0 += (unsigned long long)n_act * per_cta) {
                int pt;
                accumulate_fine<NPTS>(p, s_foldC, s_slots + (tid >> 5) * 32 * 8, b0 + (tid >> 5) * per_warp + ((tid & 31) >> q.lpp_log2), q.lpp_log2,
                                      part, pt);
                if (pt >= 0) {
                    my_pt = pt;
                    wide_add_limbs(mine, part.l);
                }
                __syncwarp();
            }
            __threadfence();
            tk2 = clock64();
            block_sum_wide_sel<NPTS>(mine, my_pt, s_part, s_tot, s_out, n_act == 1);
        } else {
#pragma unroll
            for (int t = 0; t < NPTS; t++) fr::wide_zero(accw[t]);
            accumulate_pairs<NPTS, true, false, true, true>(p, s_foldC, (unsigned long long)c * RES_THREADS + tid, (unsigned long long)n_act * RES_THREADS, accw);
            __threadfence();  // this thread's folded-table stores are visible GPU-wide before the CTA reports its arrival
            tk2 = clock64();
            block_sum_wide<NPTS>(accw, s_part, s_tot, s_out, n_act == 1);  // the CTA's sums as NPTS 17-limb integers (no Montgomery reduction)
        }
        bool publisher = true;
        if (n_act > 1) {
            // several CTAs: every CTA adds its per-limb sums into the grid's totals with 64-bit reductions (no carries yet: a
            // limb sum stays below 2^48), the last CTA to arrive carries them out and clears the totals for their next use
            unsigned long long* tot = q.totals + (size_t)(rd & 1u) * NW;
            if (tid < NW) {
                atomicAdd(tot + tid, s_tot[tid]);
                __threadfence();
            }
            __syncthreads();
            if (tid == 0) {
                const unsigned int ticket = atomicAdd(q.counters + rd, 1u);
                s_flag = (ticket == n_act - 1) ? 1 : 0;
            }
            __syncthreads();
            publisher = s_flag == 1;
            if (publisher) {
                __threadfence();
                carry_wide<NPTS, true>(tot, s_out);
                if (tid < NW) tot[tid] = 0;  // this buffer serves round rd + 2 next; its CTAs start after this round is published
            }
        }
        const long long tk3 = clock64();
        if (publisher && tid < 32) {
            if (p.peer_mail) {  // sharded: all-to-all of the unreduced sums over NVLink peer memory, then the global sums
                p.mail_seq = q.mail_seq0 + rd;
                p.mail_slot = p.mail_seq % MAIL_SLOTS;
                exchange_wide<NPTS>(p, s_out, s_rows);
            }
            // one 8-byte {limb, seq} store per limb straight into mapped host memory: no fence, no separate flag
            for (uint32_t w = tid; w < (uint32_t)NW; w += 32) mail_store(q.h_sums + 2 * w, s_out[w], seq);
        }
        if (q.prof && c == 0 && tid == 0) {
            const long long tk4 = clock64();
            q.prof[rd * 4 + 0] = tk1 - tk0; q.prof[rd * 4 + 1] = tk2 - tk1; q.prof[rd * 4 + 2] = tk3 - tk2; q.prof[rd * 4 + 3] = tk4 - tk3;
        }
        __syncthreads();  // the shared arrays and s_flag are reused by the next round
        cur = nxt;
    }
}

}  // namespace sck
