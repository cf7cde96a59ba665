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
//   device -> host    the round's raw sums P(0), P(2), .., P(d) (kernels.cuh RoundParams::raw_out semantics); the host
//                     finishes the message (deferred coefficient, P(1) from the claim, canonical forms — host_fr.h)
//
// Per round every active CTA folds + sums its pairs (accumulate_pairs, the same arithmetic as round_kernel<NPTS, true>),
// the last CTA to arrive adds the per-CTA partials and publishes.  A round's tables are written by other CTAs of the same
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
    __shared__ uint32_t s_red[32 * NPTS * 8];
    __shared__ __align__(16) uint32_t s_foldC[RES_CONST_WORDS];
    __shared__ volatile int s_flag;
    const uint32_t tid = threadIdx.x, c = blockIdx.x;
    RoundParams p = q.rp;
    int cur = q.cur;
    if (tid == 0) s_flag = 0;
    __syncthreads();
    for (uint32_t rd = 0; rd < q.n_rounds; rd++) {
        const unsigned long long n_pairs = q.n_pairs_first >> rd;
        const unsigned long long want = (n_pairs + RES_THREADS - 1) / RES_THREADS;
        const uint32_t n_act = want < (unsigned long long)gridDim.x ? (want ? (uint32_t)want : 1u) : gridDim.x;
        if (c >= n_act) return;  // this CTA has no pairs in this round, nor in any later one
        const uint32_t seq = q.seq0 + rd;
        const long long tk0 = clock64();
        // ---- this round's fold constants: CTA 0 from the host (PCIe reads), the others from CTA 0's copy in device memory
        if (tid < RES_CONST_WORDS) {
            const uint32_t* src = (c == 0 ? q.h_consts : q.d_bcast) + 2 * tid;
            uint32_t v = 0, f = 0;
            const long long t0 = clock64();
            for (;;) {
                if (q.flags & 2u) {
                    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(f) : "l"(src) : "memory");
                } else {
                    mail_load(src, v, f);
                }
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
#pragma unroll
        for (int t = 0; t < NPTS; t++) fr::wide_zero(accw[t]);
        accumulate_pairs<NPTS, true, false>(p, s_foldC, (unsigned long long)c * RES_THREADS + tid, (unsigned long long)n_act * RES_THREADS, accw);
        __threadfence();  // this thread's folded-table stores are visible GPU-wide before the CTA reports its arrival
        const long long tk2 = clock64();
        Fr acc[NPTS];
#pragma unroll
        for (int t = 0; t < NPTS; t++) acc[t] = fr::wide_reduce(accw[t]);
        block_reduce<NPTS>(acc, s_red);
        bool publisher = true;
        if (n_act > 1) {
            uint32_t* part = q.partials + (size_t)(rd & 1u) * gridDim.x * NPTS * 8;
            if (tid == 0) {
#pragma unroll
                for (int t = 0; t < NPTS; t++) fr::store(part + ((size_t)c * NPTS + t) * 8, acc[t]);
                __threadfence();
                const unsigned int ticket = atomicAdd(q.counters + rd, 1u);
                s_flag = (ticket == n_act - 1) ? 1 : 0;
            }
            __syncthreads();
            publisher = s_flag == 1;
            if (publisher) {  // last arrival: add the per-CTA partials (prover.rs:138-148, the rayon reduce)
                __threadfence();
#pragma unroll
                for (int t = 0; t < NPTS; t++) acc[t] = fr::zero();
                for (uint32_t g = tid; g < n_act; g += RES_THREADS) {
#pragma unroll
                    for (int t = 0; t < NPTS; t++) acc[t] = fr::add(acc[t], load_cg(part + ((size_t)g * NPTS + t) * 8));
                }
                block_reduce<NPTS>(acc, s_red);
            }
        }
        const long long tk3 = clock64();
        if (publisher && tid < 32) {
            if (p.peer_mail) {  // sharded: all-to-all of the partial sums over NVLink peer memory, then the global sums
                p.mail_seq = q.mail_seq0 + rd;
                p.mail_slot = p.mail_seq % MAIL_SLOTS;
                exchange_partials<NPTS>(p, acc, s_red);
            }
            if (tid == 0) {
#pragma unroll
                for (int t = 0; t < NPTS; t++)
#pragma unroll
                    for (int i = 0; i < 8; i++) s_red[t * 8 + i] = acc[t].l[i];
            }
            __syncwarp();
            // one 8-byte {limb, seq} store per limb straight into mapped host memory: no fence, no separate flag
            for (uint32_t w = tid; w < (uint32_t)NPTS * 8; w += 32) {
                if (q.flags & 2u) {
                    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(q.h_sums + 2 * w), "r"(s_red[w]), "r"(seq) : "memory");
                } else {
                    mail_store(q.h_sums + 2 * w, s_red[w], seq);
                }
            }
            if (q.flags & 1u) __threadfence_system();
        }
        if (q.prof && c == 0 && tid == 0) {
            const long long tk4 = clock64();
            q.prof[rd * 4 + 0] = tk1 - tk0; q.prof[rd * 4 + 1] = tk2 - tk1; q.prof[rd * 4 + 2] = tk3 - tk2; q.prof[rd * 4 + 3] = tk4 - tk3;
        }
        __syncthreads();  // s_red / s_foldC / s_flag are reused by the next round
        cur = nxt;
    }
}

}  // namespace sck
