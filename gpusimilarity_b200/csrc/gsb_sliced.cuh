// Bit-sliced multi-query scan for 1024-bit rows (BASELINE config "batched 1024 queries, top-100").
//
// scan_batch_kernel (gsb_batch.cuh) pays 32 AND + 32 POPC per row and query and is bound by the
// POPC pipe (16 lanes/clk/SM).  This kernel gets the same common-bit counts — hence bit-identical
// scores (reference TanimotoFunctor, fingerprintdb_cuda.cu:89-103) — from ~5 LOP3-class
// instructions per SET BIT of the query and 32 rows:
//
//   * a tile is 32 consecutive 32-row batches (1024 rows).  The CTA pulls the tile into shared
//     memory with TMA bulk copies and transposes every batch in place (32x32 bit-matrix transposes
//     in registers): word T[pos] of a batch then holds bit `pos` of its 32 rows;
//   * lane l of a warp works on batch l of the tile, all lanes on the same query: the warp walks
//     the query's list of set-bit positions (shared memory, broadcast reads), every lane loads
//     T_l[pos] (conflict free by construction, gsb_sliced_math.h) and adds it into bit-sliced
//     counters with carry-save adders; after the list the counter columns are the common-bit
//     counts of 32 x 32 rows;
//   * a row can only beat the query's running threshold tau if common >= m, where m follows from
//     score <= common / (popc(q) + pd_min - common) with pd_min the smallest row popcount of the
//     lane's batch — one bit-sliced compare; only the rare rows that pass are scored exactly
//     (tanimoto_div, popcount trailer) and appended to the query's candidate list, with the same
//     key, cutoff and survivor rules as the other kernels.  Select rounds, per-CTA lists and the grid-wide merge are shared with
//     scan_batch_kernel (batch_select_round / batch_finish).
//
// The host (gsb_api.cu) first runs the kernel over a strided ~1.5 % sample of the tiles to get a
// threshold per query, then over all tiles starting from those thresholds, so that almost no row
// of the full pass takes the exact path.  Dense queries are just longer lists (no fallback).
#pragma once

#include "gsb_batch.cuh"
#include "gsb_sliced_math.h"

namespace gsb
{

constexpr uint32_t kMaxSlicedQueries = 1024;
constexpr uint32_t kSlicedListEntries = 20480; // u16 list entries of one query block in shared memory
constexpr uint32_t kSlicedTileBytes = kSlicedTileBatches * kSlicedRegionBytes;
constexpr uint32_t kSlicedPerQueryBytes = 8 + 8 + 4 + 4 + 4 + 2 + 2 + 2;

// Query blocks: consecutive queries whose lists fit the shared-memory list area together.
struct SlicedMeta {
    uint32_t n_blocks;
    uint32_t total_entries;
    uint32_t blk_start[kMaxSlicedQueries + 1];
};

struct SlicedParams {
    BatchParams b;                       // database, k, cutoff, nq, candidate lists, outputs
    const uint16_t* lists;               // set-bit entries of all queries (sliced_entry), padded to groups of 8
    const uint32_t* lofs;                // [nq] first entry of each query's list
    const uint16_t* ngrp;                // [nq] groups of 8 entries
    const uint16_t* popq;                // [nq] query popcounts
    const SlicedMeta* meta;
    const unsigned long long* tau_init;  // [nq] starting thresholds, or nullptr
    uint32_t n_claims, tile_step;        // tiles scanned: claim c -> tile c * tile_step
};

__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Lists, popcounts and query blocks from the raw queries ([nq][32] words): one thread per query.
__global__ void __launch_bounds__(kMaxSlicedQueries, 1)
sliced_build_lists_kernel(const uint32_t* __restrict__ queries, uint32_t nq, uint16_t* lists, uint32_t* lofs,
                          uint16_t* ngrp, uint16_t* popq, SlicedMeta* meta)
{
    __shared__ uint32_t s_scan[kMaxSlicedQueries];
    const uint32_t j = threadIdx.x;
    uint32_t pc = 0;
    if (j < nq)
        for (uint32_t w = 0; w < 32; w++)
            pc += __popc(queries[j * 32 + w]);
    const uint32_t padded = (pc + kSlicedGroup - 1) / kSlicedGroup * kSlicedGroup;
    s_scan[j] = j < nq ? padded : 0u;
    __syncthreads();
    for (uint32_t d = 1; d < kMaxSlicedQueries; d <<= 1) {
        const uint32_t v = j >= d ? s_scan[j - d] : 0u;
        __syncthreads();
        s_scan[j] += v;
        __syncthreads();
    }
    if (j < nq) {
        const uint32_t off = s_scan[j] - padded;
        lofs[j] = off;
        ngrp[j] = static_cast<uint16_t>(padded / kSlicedGroup);
        popq[j] = static_cast<uint16_t>(pc);
        uint16_t* dst = lists + off;
        uint32_t n = 0;
        for (uint32_t w = 0; w < 32; w++) {
            uint32_t x = queries[j * 32 + w];
            while (x) {
                const uint32_t b = __ffs(x) - 1;
                x &= x - 1;
                dst[n++] = sliced_entry(w * 32 + b);
            }
        }
        for (; n < padded; n++)
            dst[n] = sliced_entry(kSlicedZeroPos);
    }
    if (j == 0) {
        uint32_t nb = 0, used = 0;
        meta->blk_start[0] = 0;
        for (uint32_t q = 0; q < nq; q++) {
            const uint32_t len = s_scan[q] - (q ? s_scan[q - 1] : 0u);
            if (used + len > kSlicedListEntries && used > 0) {
                meta->blk_start[++nb] = q;
                used = 0;
            }
            used += len;
        }
        meta->blk_start[++nb] = nq;
        meta->n_blocks = nb;
        meta->total_entries = nq ? s_scan[nq - 1] : 0u;
    }
}

// Thresholds for the full pass from the sample pass: the k-th key of the sample stays eligible.
__global__ void sliced_seed_tau_kernel(const unsigned long long* keys, const uint32_t* counts, uint32_t nq, uint32_t k,
                                       unsigned long long* tau)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nq)
        tau[j] = (counts[j] >= k && k > 0 && keys[(uint64_t) j * k + k - 1] > 0) ? keys[(uint64_t) j * k + k - 1] - 1ull : 0ull;
}

template <int CW>
__global__ void __launch_bounds__(CW * 32, 1) scan_sliced_kernel(const __grid_constant__ SlicedParams sp)
{
    constexpr int NT = CW * 32;
    constexpr uint32_t kFull = 0xffffffffu;
    const BatchParams& p = sp.b;

    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_full[CW];
    __shared__ unsigned long long s_stage_tau;
    __shared__ unsigned int s_stage_count, s_dummy_epoch, s_alive, s_next_q, s_need_select;
    __shared__ unsigned int s_claim[2];
    __shared__ float s_pdmin[kSlicedTileBatches]; // smallest row popcount of each batch of the tile

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nq = p.nq, nqp = (nq + 7u) & ~7u;
    // shared memory carve-up
    uint8_t* tile = smem;                                                        // 32 batch regions
    uint16_t* s_pd = reinterpret_cast<uint16_t*>(smem + kSlicedTileBytes);       // [1024] row popcounts of the tile
    uint16_t* s_list = s_pd + kSlicedTileBatches * kBatchRows;                   // [kSlicedListEntries]
    uint8_t* cursor = reinterpret_cast<uint8_t*>(s_list + kSlicedListEntries);
    unsigned long long* s_tau = reinterpret_cast<unsigned long long*>(cursor);   // [nq] keys <= tau are out
    cursor += (size_t) nqp * 8;
    unsigned long long* s_surv = reinterpret_cast<unsigned long long*>(cursor);  // [nq]
    cursor += (size_t) nqp * 8;
    unsigned int* s_cnt = reinterpret_cast<unsigned int*>(cursor);               // [nq] list fill
    cursor += (size_t) nqp * 4;
    uint32_t* s_lofs = reinterpret_cast<uint32_t*>(cursor);                      // [nq]
    cursor += (size_t) nqp * 4;
    float* s_tq = reinterpret_cast<float*>(cursor);                              // [nq] filter: see sliced_tq
    cursor += (size_t) nqp * 4;
    uint16_t* s_m = reinterpret_cast<uint16_t*>(cursor);                         // [nq] filter: common >= m
    cursor += (size_t) nqp * 2;
    uint16_t* s_popq = reinterpret_cast<uint16_t*>(cursor);
    cursor += (size_t) nqp * 2;
    uint16_t* s_ngrp = reinterpret_cast<uint16_t*>(cursor);
    // the select rounds run between tiles and stage through the (then dead) tile buffer
    CandShared cs;
    cs.buf = reinterpret_cast<unsigned long long*>(tile);
    cs.cap = kBatchListCap;
    cs.count = &s_stage_count;
    cs.tau = &s_stage_tau;
    cs.epoch_req = &s_dummy_epoch;
    cs.hist = reinterpret_cast<unsigned int*>(tile + (size_t) kBatchListCap * 8);
    unsigned long long* my_cand = p.cand + (uint64_t) blockIdx.x * nq * kBatchListCap;

    const bool drop_zero = p.cutoff > 0.0f; // reference .cu:265
    // With a cutoff every row at or above it must be seen (survivor count), so the filter follows
    // the cutoff; without one it follows the query's threshold.
    auto update_filter = [&](uint32_t j) {
        const float ts = drop_zero ? p.cutoff : __uint_as_float(static_cast<uint32_t>((s_tau[j] + 1ull) >> 32));
        const uint32_t m = sliced_filter_min(ts, s_popq[j], [](uint32_t c, uint32_t u) { return tanimoto_div(c, u); });
        s_m[j] = static_cast<uint16_t>(m > 0xffffu ? 0xffffu : m);
        s_tq[j] = sliced_tq(ts);
    };

    for (uint32_t j = tid; j < nq; j += NT) {
        s_tau[j] = sp.tau_init ? sp.tau_init[j] : 0ull;
        s_surv[j] = 0;
        s_cnt[j] = 0;
        s_lofs[j] = sp.lofs[j];
        s_popq[j] = sp.popq[j];
        s_ngrp[j] = sp.ngrp[j];
    }
    if (tid == 0) {
        s_alive = 0;
        s_need_select = 0;
        s_claim[0] = atomicAdd(&p.ctrl->next_batch, 1u);
    }
    if (lane == 0) {
        mbar_init(&s_full[warp], 1);
        mbar_fence_init();
    }
    __syncthreads();
    for (uint32_t j = tid; j < nq; j += NT)
        update_filter(j);
    const uint32_t n_blocks = sp.meta->n_blocks;
    auto load_lists = [&](uint32_t blk) { // all threads; the caller syncs
        const uint32_t q0 = sp.meta->blk_start[blk], q1 = sp.meta->blk_start[blk + 1];
        if (q1 <= q0)
            return;
        const uint32_t first = s_lofs[q0], last = s_lofs[q1 - 1] + s_ngrp[q1 - 1] * kSlicedGroup;
        const uint4* src = reinterpret_cast<const uint4*>(sp.lists + first);
        uint4* dst = reinterpret_cast<uint4*>(s_list);
        for (uint32_t i = tid; i < (last - first) / 8; i += NT)
            dst[i] = src[i];
    };
    if (n_blocks == 1)
        load_lists(0);

    const uint8_t* my_T = tile + sliced_lane_base(lane);
    const uint32_t row_base32 = static_cast<uint32_t>(p.row_base);
    uint32_t phase = 0;

    // One query against the 1024 rows of the tile; NP = counter planes above "fours".
    auto run_query = [&](auto np_tag, uint32_t j, uint32_t list_base, uint32_t b0, uint32_t nb_tile, float pdmin) {
        constexpr int NP = decltype(np_tag)::value;
        const uint32_t pq = s_popq[j], m = s_m[j];
        if (m > pq)
            return; // no row can reach this query's threshold any more
        // this lane's bound: its batch has no row with fewer than pdmin set bits
        const uint32_t ml = max(m, sliced_lane_min(s_tq[j], static_cast<float>(pq) + pdmin));
        const uint32_t ng = s_ngrp[j];
        const uint4* lp = reinterpret_cast<const uint4*>(s_list + (s_lofs[j] - list_base));
        SlicedCount<NP> cnt;
#pragma unroll 2
        for (uint32_t g = 0; g < ng; g++) {
            const uint4 e = lp[g];
            const uint32_t x0 = *reinterpret_cast<const uint32_t*>(my_T + (e.x & 0xffffu));
            const uint32_t x1 = *reinterpret_cast<const uint32_t*>(my_T + (e.x >> 16));
            const uint32_t x2 = *reinterpret_cast<const uint32_t*>(my_T + (e.y & 0xffffu));
            const uint32_t x3 = *reinterpret_cast<const uint32_t*>(my_T + (e.y >> 16));
            const uint32_t x4 = *reinterpret_cast<const uint32_t*>(my_T + (e.z & 0xffffu));
            const uint32_t x5 = *reinterpret_cast<const uint32_t*>(my_T + (e.z >> 16));
            const uint32_t x6 = *reinterpret_cast<const uint32_t*>(my_T + (e.w & 0xffffu));
            const uint32_t x7 = *reinterpret_cast<const uint32_t*>(my_T + (e.w >> 16));
            cnt.add8(x0, x1, x2, x3, x4, x5, x6, x7);
        }
        uint32_t ge = cnt.at_least_lane(ml);
        if (lane >= nb_tile || (ml >> (3 + NP)) != 0)
            ge = 0; // ragged last tile: this lane has no batch; or a bound no count of this width reaches
        unsigned hit = __ballot_sync(kFull, ge != 0);
        // exact path, one batch with candidates at a time: lane r takes row r of that batch
        while (hit) {
            const uint32_t src = __ffs(hit) - 1;
            hit &= hit - 1;
            const uint32_t g = __shfl_sync(kFull, ge, src);
            uint32_t common = 0;
#pragma unroll
            for (int pl = 0; pl < 3 + NP; pl++)
                common |= ((__shfl_sync(kFull, cnt.plane(pl), src) >> lane) & 1u) << pl;
            const uint32_t pd = s_pd[src * kBatchRows + lane];
            const uint32_t row_local = (b0 + src) * kBatchRows + lane;
            const bool valid = ((g >> lane) & 1u) != 0 && row_local < p.n_rows;
            // reference .cu:100-102: IEEE divide, then the cutoff test (NaN -> 0)
            float score = tanimoto_div(common, pq + pd - common);
            score = (score >= p.cutoff) ? score : 0.0f;
            const bool survivor = valid && (!drop_zero || score != 0.0f); // .cu:265-271
            if (drop_zero) {
                const unsigned sv = __ballot_sync(kFull, survivor);
                if (sv && lane == 0)
                    atomicAdd(&s_surv[j], static_cast<unsigned long long>(__popc(sv)));
            }
            const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(score)) << 32) |
                                           static_cast<unsigned long long>(0xffffffffu - (row_base32 + row_local));
            const bool pass = survivor && key > *reinterpret_cast<volatile unsigned long long*>(&s_tau[j]);
            const unsigned pm = __ballot_sync(kFull, pass);
            if (pm) {
                unsigned base = 0;
                if (lane == 0) {
                    base = atomicAdd(&s_cnt[j], __popc(pm));
                    if (base + __popc(pm) > kBatchListCap / 2)
                        s_need_select = 1; // cut the list back before the next tile (a tile adds <= 1024)
                }
                base = __shfl_sync(kFull, base, 0);
                if (pass) {
                    const unsigned idx = base + __popc(pm & ((1u << lane) - 1u));
                    if (idx >= kBatchListCap)
                        __trap(); // unreachable: lists are cut to <= 1024 entries between tiles
                    my_cand[(uint64_t) j * kBatchListCap + idx] = key;
                }
            }
        }
    };

    for (uint32_t it = 0;; it++) {
        // everybody is done with the tile buffer (queries / select round of the previous tile)
        fence_proxy_async_smem();
        cta_sync<NT>();
        const uint32_t claim = s_claim[it & 1];
        if (claim >= sp.n_claims)
            break;
        const uint32_t b0 = claim * sp.tile_step * kSlicedTileBatches;
        const uint32_t nb_tile = p.n_batches - b0 < kSlicedTileBatches ? p.n_batches - b0 : kSlicedTileBatches;
        // ---- phase A: TMA the tile in (every warp its own batches), transpose in place
        const uint32_t n_mine = warp < nb_tile ? (nb_tile - warp + CW - 1) / CW : 0u;
        if (lane == 0 && n_mine) {
            mbar_arrive_expect_tx(&s_full[warp], n_mine * p.batch_bytes);
            for (uint32_t b = warp; b < nb_tile; b += CW)
                tma_bulk_g2s(tile + (size_t) b * kSlicedRegionBytes, p.tiles + (uint64_t)(b0 + b) * p.batch_stride,
                             p.batch_bytes, &s_full[warp]);
        }
        if (tid == 0) {
            s_claim[(it + 1) & 1] = atomicAdd(&p.ctrl->next_batch, 1u); // next tile: hides the atomic's latency
            s_next_q = 0;
        }
        if (n_blocks > 1)
            load_lists(0);
        if (n_mine) {
            mbar_wait(&s_full[warp], phase);
            phase ^= 1u;
            for (uint32_t b = warp; b < nb_tile; b += CW) {
                uint8_t* region = tile + (size_t) b * kSlicedRegionBytes;
                const uint32_t* raw = reinterpret_cast<const uint32_t*>(region);
                uint32_t x[32]; // lane = word column, x[r] = that word of row r
#pragma unroll
                for (int r = 0; r < 32; r++)
                    x[r] = raw[r * 32 + lane];
                const uint16_t pd = reinterpret_cast<const uint16_t*>(region + (size_t) kBatchRows * 128)[lane];
                const uint32_t pd_lo = __reduce_min_sync(kFull, static_cast<uint32_t>(pd));
                if (lane == 0)
                    s_pdmin[b] = static_cast<float>(pd_lo);
                __syncwarp(); // in place: every lane has read the batch before anyone overwrites it
                transpose32(x);
                uint32_t* T = reinterpret_cast<uint32_t*>(tile + sliced_lane_base(b));
#pragma unroll
                for (int bb = 0; bb < 32; bb++)
                    T[lane * 32 + ((bb + lane) & 31)] = x[bb]; // == sliced_word_index(lane * 32 + bb)
                if (lane == 0)
                    T[kSlicedZeroPos] = 0u;
                s_pd[b * kBatchRows + lane] = pd;
            }
        }
        cta_sync<NT>();
        // ---- phase B: warps take queries one at a time
        const float my_pdmin = s_pdmin[lane];
        for (uint32_t blk = 0; blk < n_blocks; blk++) {
            if (blk > 0) {
                cta_sync<NT>(); // every warp has left the previous block
                load_lists(blk);
                if (tid == 0)
                    s_next_q = sp.meta->blk_start[blk];
                cta_sync<NT>();
            }
            const uint32_t q_end = n_blocks == 1 ? nq : sp.meta->blk_start[blk + 1];
            const uint32_t list_base = n_blocks == 1 ? 0u : s_lofs[sp.meta->blk_start[blk]];
            for (;;) {
                uint32_t j = 0;
                if (lane == 0)
                    j = atomicAdd(&s_next_q, 1u);
                j = __shfl_sync(kFull, j, 0);
                if (j >= q_end)
                    break;
                if (s_ngrp[j] <= 15) // <= 120 set bits: counts fit 7 planes
                    run_query(std::integral_constant<int, 4>{}, j, list_base, b0, nb_tile, my_pdmin);
                else
                    run_query(std::integral_constant<int, 8>{}, j, list_base, b0, nb_tile, my_pdmin);
            }
        }
        cta_sync<NT>();
        // ---- lists that passed 1024 entries are cut back (staged through the tile buffer)
        if (*reinterpret_cast<volatile unsigned int*>(&s_need_select)) {
            batch_select_round<NT>(cs, my_cand, s_cnt, s_tau, nq, p.k, false, tid,
                                   [&](uint32_t j) { update_filter(j); });
            if (tid == 0)
                s_need_select = 0;
        }
    }
    batch_finish<NT>(p, cs, my_cand, s_cnt, s_tau, s_surv, &s_alive, tid);
}

} // namespace gsb
