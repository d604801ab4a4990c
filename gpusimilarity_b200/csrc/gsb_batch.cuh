// Multi-query scan with POPC: one pass over the database scores every 32-row batch against up to
// kMaxBatchQueries queries held in shared memory (BASELINE config "batched 1024 queries, top-100";
// the reference has no batching, gpusim.cpp:407-414 serves one query per request).  Used for small
// batches (fewer than 6 queries); larger ones go to the bit-sliced kernel in gsb_sliced.cuh, which
// shares the candidate lists, select rounds and the end of the launch with this file
// (batch_select_round, batch_finish).
//
// The database is read once per launch, so the kernel is bound by the POPC pipe (16 lanes/clk/SM:
// 32 POPC per row and query), not by HBM.  Data path and work distribution are the single-query
// kernel's (per-warp TMA rings, guided batch claiming); what changes is the select state: every
// (CTA, query) pair owns a candidate list in global memory with its fill counter and threshold
// in shared memory.  After warm-up a row beats a query's threshold with probability ~k/rows seen,
// so appends are rare; lists are cut back with the same one-pass histogram select, staged
// through shared memory.  At the end every CTA sorts its lists, a grid-wide arrival counter lines
// the CTAs up, and the per-query merges are spread over the CTAs.
#pragma once

#include "gsb_kernels.cuh"

namespace gsb
{

constexpr uint32_t kMaxBatchQueries = 256;
constexpr uint32_t kBatchListCap = 2048; // entries per (CTA, query) candidate list
constexpr uint32_t kMaxBatchK = 512;

struct BatchCtrl {
    unsigned int ticket;      // CTAs that finished scanning + sorting their lists
    unsigned int next_batch;  // dynamic work distribution
    unsigned int done;        // CTAs that finished their merges (last one resets the block)
    unsigned int error;       // kErr* bits raised by any CTA (reported through out_n, then cleared)
};

struct BatchParams {
    const uint8_t* tiles;
    uint64_t n_rows, row_base;
    uint32_t n_batches, batch_stride, batch_bytes, stage_bytes, stages;
    uint32_t k;
    float cutoff;
    uint32_t nq;
    const uint32_t* queries;            // [nq][W] device memory
    unsigned long long* cand;           // [grid][nq][kBatchListCap]
    unsigned long long* qlists;         // [grid][nq][k] sorted per-CTA results
    uint32_t* qcounts;                  // [grid][nq]
    unsigned long long* surv_acc;       // [nq] zero on entry, zero again on exit
    BatchCtrl* ctrl;
    unsigned long long* out_keys;       // [nq][k]
    uint32_t* out_n;                    // [nq]
    unsigned long long* out_survivors;  // [nq]
    unsigned long long spin_timeout_ns; // wall-clock bound of the grid-wide arrival spin
    uint32_t metric;                    // kMetric* (the bit-sliced kernel's filter knows Tanimoto only)
    float alpha, beta;
};

// Cut back every candidate list of this CTA that is more than half full (or, with exact_all, cut
// every list to its exact sorted top k).  All NT threads; lists are staged through cs.buf.
// tau_raised(j) runs on thread 0 after s_tau[j] went up; lists with skip(j) are left alone.
template <int NT, class F, class S>
__device__ void batch_select_round(const CandShared& cs, unsigned long long* my_cand, unsigned int* s_cnt,
                                   unsigned long long* s_tau, uint32_t nq, uint32_t k, bool exact_all, uint32_t tid,
                                   F tau_raised, S skip)
{
    cta_sync<NT>();
    for (uint32_t j = 0; j < nq; j++) {
        const uint32_t n = s_cnt[j] < kBatchListCap ? s_cnt[j] : kBatchListCap;
        if ((!exact_all && n <= kBatchListCap / 2) || skip(j))
            continue;
        unsigned long long* list = my_cand + (uint64_t) j * kBatchListCap;
        for (uint32_t i = tid; i < n; i += NT)
            cs.buf[i] = list[i];
        if (tid == 0) {
            *cs.count = n;
            *cs.tau = 0;
        }
        cand_compact<NT>(cs, k, nullptr, tid, exact_all);
        const uint32_t kept = *cs.count;
        for (uint32_t i = tid; i < kept; i += NT)
            list[i] = cs.buf[i];
        if (tid == 0) {
            s_cnt[j] = kept;
            if (*cs.tau > s_tau[j]) {
                s_tau[j] = *cs.tau;
                tau_raised(j);
            }
        }
        cta_sync<NT>();
    }
}

// End of a multi-query launch, all NT threads of every CTA: exact sorted top-k of every list to
// global memory, grid-wide arrival (cooperative launch: all CTAs resident), per-query merges
// spread over the CTAs, and the last CTA leaves the control block clean.  Lists with sorted(j) are
// already exact and sorted.
template <int NT, class S>
__device__ void batch_finish(const BatchParams& p, const CandShared& cs, unsigned long long* my_cand,
                             unsigned int* s_cnt, unsigned long long* s_tau, const unsigned long long* s_surv,
                             unsigned int* s_alive, uint32_t tid, S sorted)
{
    const uint32_t nq = p.nq;
    const bool drop_zero = p.cutoff > 0.0f;
    __threadfence_block();
    batch_select_round<NT>(cs, my_cand, s_cnt, s_tau, nq, p.k, true, tid, [](uint32_t) {}, sorted);
    // a warp per list (one list after the other with the whole CTA is a chain of nq dependent
    // global load -> store round trips: 0.3 ms of a 1024-query pass)
    for (uint32_t j = tid >> 5; j < nq; j += NT / 32) {
        const uint32_t n = s_cnt[j], lane = tid & 31;
        const unsigned long long* list = my_cand + (uint64_t) j * kBatchListCap;
        unsigned long long* dst = p.qlists + ((uint64_t) blockIdx.x * nq + j) * p.k;
        for (uint32_t i0 = 0; i0 < n; i0 += 128) {
            unsigned long long key[4];
#pragma unroll
            for (int e = 0; e < 4; e++)
                key[e] = i0 + e * 32 + lane < n ? list[i0 + e * 32 + lane] : 0ull;
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (i0 + e * 32 + lane < n)
                    dst[i0 + e * 32 + lane] = key[e];
        }
        if (lane == 0) {
            p.qcounts[blockIdx.x * nq + j] = n;
            if (drop_zero && s_surv[j])
                atomicAdd(&p.surv_acc[j], s_surv[j]);
        }
    }
    __threadfence();
    cta_sync<NT>();
    // ---- grid-wide arrival (the grid is persistent: one CTA per SM, all resident)
    if (tid == 0) {
        if (*cs.error)
            atomicOr(&p.ctrl->error, *cs.error);
        atomicAdd(&p.ctrl->ticket, 1u);
        const unsigned long long t0 = global_ns();
        while (*reinterpret_cast<volatile unsigned int*>(&p.ctrl->ticket) < gridDim.x) {
            // (CTAs finish far apart here: the host passes ten times the single-query bound)
            if (global_ns() - t0 > p.spin_timeout_ns) { // not co-resident: report it, never hang or trap
                atomicOr(&p.ctrl->error, kErrGridBarrier);
                break;
            }
        }
        __threadfence();
    }
    cta_sync<NT>();
    // ---- merges, spread over the CTAs: query j is merged by CTA j % grid
    for (uint32_t j = blockIdx.x; j < nq; j += gridDim.x) {
        merge_lists<NT>(cs, p.qlists + (uint64_t) j * p.k, p.qcounts + j, gridDim.x, nq * p.k, p.k, 0ull, s_alive, tid,
                        nq);
        const uint32_t n = *cs.count;
        for (uint32_t i = tid; i < p.k; i += NT)
            p.out_keys[(uint64_t) j * p.k + i] = i < n ? cs.buf[i] : 0ull;
        if (tid == 0) {
            const unsigned int err = *cs.error | *reinterpret_cast<volatile unsigned int*>(&p.ctrl->error);
            p.out_n[j] = err ? kCountError : n;
            p.out_survivors[j] = drop_zero ? *reinterpret_cast<volatile unsigned long long*>(&p.surv_acc[j]) : p.n_rows;
        }
        cta_sync<NT>();
    }
    // ---- the last CTA to finish leaves the control block and accumulators clean
    __threadfence();
    cta_sync<NT>();
    if (tid == 0) {
        const unsigned d = atomicAdd(&p.ctrl->done, 1u);
        *s_alive = (d == gridDim.x - 1) ? 1u : 0u;
    }
    cta_sync<NT>();
    if (*s_alive) {
        for (uint32_t j = tid; j < nq; j += NT)
            p.surv_acc[j] = 0;
        if (tid == 0) {
            p.ctrl->next_batch = 0;
            p.ctrl->done = 0;
            p.ctrl->error = 0;
            __threadfence();
            p.ctrl->ticket = 0;
        }
    }
}

template <int W, int CW>
__global__ void __launch_bounds__(CW * 32, 1) scan_batch_kernel(const __grid_constant__ BatchParams p)
{
    constexpr int L = W / 4;
    constexpr int NT = CW * 32;
    constexpr uint32_t kIterBytes = 512;
    constexpr uint32_t kChunk = 4, kEnd = 0xffffffffu; // small chunks: a batch is a lot of work here

    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_full[CW * kMaxStages];
    __shared__ uint32_t s_bid[CW * kMaxStages];
    __shared__ unsigned long long s_stage_tau;
    __shared__ unsigned int s_stage_count, s_epoch_req, s_done, s_alive, s_dummy_epoch, s_error;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t S = p.stages, nq = p.nq;
    // shared memory carve-up
    uint8_t* my_ring = smem + (size_t) warp * S * p.stage_bytes;
    uint8_t* cursor = smem + (size_t) CW * S * p.stage_bytes;
    uint32_t* s_q = reinterpret_cast<uint32_t*>(cursor);                        // [nq][W]
    cursor += (size_t) kMaxBatchQueries * W * 4;
    unsigned long long* s_buf = reinterpret_cast<unsigned long long*>(cursor);  // select staging
    cursor += (size_t) kBatchListCap * 8;
    unsigned int* s_hist = reinterpret_cast<unsigned int*>(cursor);
    cursor += (size_t) kBuckets * 4;
    unsigned long long* s_tau = reinterpret_cast<unsigned long long*>(cursor);  // [nq]
    cursor += (size_t) kMaxBatchQueries * 8;
    unsigned long long* s_surv = reinterpret_cast<unsigned long long*>(cursor); // [nq]
    cursor += (size_t) kMaxBatchQueries * 8;
    unsigned int* s_cnt = reinterpret_cast<unsigned int*>(cursor);              // [nq]
    cursor += (size_t) kMaxBatchQueries * 4;
    unsigned int* s_popq = reinterpret_cast<unsigned int*>(cursor);             // [nq]

    uint64_t* my_full = s_full + warp * kMaxStages;
    uint32_t* my_bid = s_bid + warp * kMaxStages;
    CandShared cs;
    cs.buf = s_buf;
    cs.cap = kBatchListCap;
    cs.count = &s_stage_count;
    cs.tau = &s_stage_tau;
    cs.epoch_req = &s_dummy_epoch;
    cs.hist = s_hist;
    cs.error = &s_error;
    cs.counted = nullptr;
    unsigned long long* my_cand = p.cand + (uint64_t) blockIdx.x * nq * kBatchListCap;
    const uint32_t high_water = kBatchListCap - 32u * CW - 64u;

    for (uint32_t i = tid; i < nq * W; i += NT)
        s_q[i] = p.queries[i];
    for (uint32_t j = tid; j < nq; j += NT) {
        s_tau[j] = 0;
        s_surv[j] = 0;
        s_cnt[j] = 0;
    }
    if (tid == 0) {
        s_epoch_req = 0;
        s_done = 0;
        s_alive = 0;
        s_error = 0;
    }
    if (lane == 0) {
        for (uint32_t s = 0; s < S; s++)
            mbar_init(&my_full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    for (uint32_t j = tid; j < nq; j += NT) {
        uint32_t pc = 0;
        for (int w = 0; w < W; w++)
            pc += __popc(s_q[j * W + w]);
        s_popq[j] = pc;
    }

    // ---- work distribution and TMA issue: as in scan_topk_kernel
    const uint32_t n_warps_total = gridDim.x * CW;
    uint32_t cur = 0, cur_end = 0, nxt = 0, nxt_size = 0; // lane 0 only
    auto claim = [&](uint32_t progress) {
        const uint32_t remaining = progress < p.n_batches ? p.n_batches - progress : 0;
        uint32_t size = remaining / (4u * n_warps_total);
        size = size < 1u ? 1u : (size > kChunk ? kChunk : size);
        nxt = atomicAdd(&p.ctrl->next_batch, size);
        nxt_size = size;
    };
    auto issue = [&](uint32_t s) {
        if (cur == cur_end) {
            cur = nxt;
            cur_end = nxt + nxt_size < p.n_batches ? nxt + nxt_size : p.n_batches;
            if (cur < p.n_batches)
                claim(cur_end);
        }
        if (cur >= p.n_batches) {
            cur = cur_end = p.n_batches;
            my_bid[s] = kEnd;
            return;
        }
        my_bid[s] = cur;
        mbar_arrive_expect_tx(&my_full[s], p.batch_bytes);
        tma_bulk_g2s(my_ring + (size_t) s * p.stage_bytes, p.tiles + (uint64_t) cur * p.batch_stride,
                     p.batch_bytes, &my_full[s]);
        cur++;
    };
    if (lane == 0) {
        claim(0);
        for (uint32_t s = 0; s < S; s++)
            issue(s);
    }
    __syncthreads();

    auto select_round = [&]() {
        batch_select_round<NT>(cs, my_cand, s_cnt, s_tau, nq, p.k, false, tid, [](uint32_t) {},
                               [](uint32_t) { return false; });
    };

    const bool drop_zero = p.cutoff > 0.0f;
    const uint32_t row_in_batch = (lane % L) * (32 / L) + lane / L;
    const uint32_t row_id_base = static_cast<uint32_t>(p.row_base) + row_in_batch;
    const uint32_t q_off = (lane % L) * 4;
    uint32_t my_epoch = 0, stage = 0, phase = 0;

    for (;;) {
        if (warp_uniform_ld(&s_epoch_req) > my_epoch) {
            select_round();
            my_epoch++;
        }
        const uint32_t bid = *reinterpret_cast<volatile uint32_t*>(&my_bid[stage]);
        if (bid == kEnd)
            break;
        const uint8_t* sp = my_ring + (size_t) stage * p.stage_bytes;
        mbar_wait(&my_full[stage], phase);
        const uint4* src = reinterpret_cast<const uint4*>(sp) + lane;
        uint4 d[L];
#pragma unroll
        for (int i = 0; i < L; i++)
            d[i] = src[i * (kIterBytes / 16)];
        const uint32_t popd = reinterpret_cast<const uint16_t*>(sp + (size_t) kBatchRows * (W * 4))[row_in_batch];
        __syncwarp();
        if (lane == 0)
            issue(stage);
        __syncwarp();
        if (++stage == S) {
            stage = 0;
            phase ^= 1u;
        }
        const bool valid = bid * kBatchRows + row_in_batch < p.n_rows;
        const uint32_t row_id = bid * kBatchRows + row_id_base;

        for (uint32_t j = 0; j < nq; j++) {
            const uint4 q = *reinterpret_cast<const uint4*>(&s_q[j * W + q_off]);
            uint32_t v[L];
#pragma unroll
            for (int i = 0; i < L; i++)
                v[i] = __popc(d[i].x & q.x) + __popc(d[i].y & q.y) + __popc(d[i].z & q.z) + __popc(d[i].w & q.w);
            const uint32_t common = transpose_reduce<L>(v, lane);
            float score = similarity(p.metric, p.alpha, p.beta, common, s_popq[j], popd);
            score = (score >= p.cutoff) ? score : 0.0f; // reference .cu:102
            const bool survivor = valid && (!drop_zero || score != 0.0f);
            if (drop_zero) {
                const unsigned sv = __ballot_sync(0xffffffffu, survivor);
                if (sv && lane == 0)
                    atomicAdd(&s_surv[j], static_cast<unsigned long long>(__popc(sv)));
            }
            const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(score)) << 32) |
                                           static_cast<unsigned long long>(0xffffffffu - row_id);
            const bool pass = survivor && key > *reinterpret_cast<volatile unsigned long long*>(&s_tau[j]);
            const unsigned m = __ballot_sync(0xffffffffu, pass);
            if (m) {
                unsigned base = 0;
                if (lane == 0) {
                    base = atomicAdd(&s_cnt[j], __popc(m));
                    if (base + __popc(m) > high_water)
                        atomicMax(&s_epoch_req, my_epoch + 1u);
                }
                base = __shfl_sync(0xffffffffu, base, 0);
                if (pass) {
                    const unsigned idx = base + __popc(m & ((1u << lane) - 1u));
                    if (idx < kBatchListCap)
                        my_cand[(uint64_t) j * kBatchListCap + idx] = key;
                    else
                        atomicOr(&s_error, kErrOverflow);
                }
            }
        }
    }
    // ---- drain: serve select requests until every warp of the CTA is out of work
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        atomicAdd(&s_done, 1u);
    }
    for (;;) {
        if (warp_uniform_ld(&s_epoch_req) > my_epoch) {
            select_round();
            my_epoch++;
            continue;
        }
        if (warp_uniform_ld(&s_done) == CW) {
            __threadfence_block();
            if (warp_uniform_ld(&s_epoch_req) > my_epoch)
                continue;
            break;
        }
    }
    batch_finish<NT>(p, cs, my_cand, s_cnt, s_tau, s_surv, &s_alive, tid, [](uint32_t) { return false; });
}

// Merge of all-gathered per-rank batch records: one CTA per query.  Record layout per rank:
// [nq][k] keys, then [nq] survivors, then [nq] counts (u64 each).
GSB_KERNEL void __launch_bounds__(kMergeThreads, 1)
merge_batch_kernel(const unsigned long long* records, uint32_t n_ranks, uint32_t nq, uint32_t k, uint32_t cap,
                   uint32_t* out_rows, float* out_scores, uint32_t* out_n, unsigned long long* out_approx)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ unsigned long long s_tau;
    __shared__ unsigned int s_count, s_epoch_req, s_alive, s_error;
    CandShared cs;
    cs.buf = reinterpret_cast<unsigned long long*>(smem);
    cs.cap = cap;
    cs.count = &s_count;
    cs.tau = &s_tau;
    cs.epoch_req = &s_epoch_req;
    cs.hist = reinterpret_cast<unsigned int*>(cs.buf + cap);
    cs.error = &s_error;
    cs.counted = nullptr;
    const uint32_t tid = threadIdx.x, j = blockIdx.x;
    if (tid == 0)
        s_error = 0;
    const uint64_t rec_len = (uint64_t) nq * (k + 2);
    merge_lists<kMergeThreads>(cs, records + (uint64_t) j * k, nullptr, n_ranks, static_cast<uint32_t>(rec_len), k, 0ull,
                               &s_alive, tid);
    const uint32_t n = s_count;
    for (uint32_t i = tid; i < k; i += kMergeThreads) {
        const unsigned long long key = i < n ? cs.buf[i] : 0ull;
        out_rows[(uint64_t) j * k + i] = 0xffffffffu - static_cast<uint32_t>(key & 0xffffffffu);
        out_scores[(uint64_t) j * k + i] = __uint_as_float(static_cast<uint32_t>(key >> 32));
    }
    if (tid == 0) {
        unsigned long long total = 0;
        for (uint32_t r = 0; r < n_ranks; r++)
            total += records[r * rec_len + (uint64_t) nq * k + j];
        out_n[j] = s_error ? kCountError : n;
        out_approx[j] = total;
    }
}

} // namespace gsb
