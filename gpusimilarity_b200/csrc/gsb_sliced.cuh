// Bit-sliced multi-query scan for rows of 128 to 1024 bits (BASELINE config "batched 1024 queries,
// top-100"; the description below is for 1024-bit rows, narrower rows only make the tile smaller).
//
// scan_batch_kernel (gsb_batch.cuh) pays 32 AND + 32 POPC per row and query and is bound by the
// POPC pipe (16 lanes/clk/SM).  This kernel gets the same common-bit counts — hence bit-identical
// scores (reference TanimotoFunctor, fingerprintdb_cuda.cu:89-103) — from ~5.5 instructions per SET
// BIT of the query and 1024 rows (one list entry, one address add, one shared-memory word and
// 2.25 LOP3 per lane and 32 rows):
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
//     key, cutoff and survivor rules as the other kernels.  Select rounds, per-CTA lists and the
//     grid-wide merge are shared with scan_batch_kernel (batch_select_round / batch_finish);
//   * thresholds are shared by the whole grid: every candidate is counted in a global score
//     histogram of its query; each CTA turns the histograms of "its" queries into thresholds
//     (floor of the bucket where the count from the top reaches k), publishes them, and picks up
//     everybody else's.  After the first few tiles almost no row takes the exact path.  A CTA's
//     first tile is a 128-row mini tile so that the warm-up (threshold 0: every row is a
//     candidate) stays cheap.
//
// Dense queries are just longer lists (no fallback).  One CTA per SM, 32 warps, ~215 KB of shared
// memory: the tile (135 KB, single buffer: the next tile is prefetched into L2 meanwhile), the row
// popcounts, the lists of up to 1024 queries (in blocks of 40 KB) and 36 bytes of state per query.
#pragma once

#include "gsb_batch.cuh"
#include "gsb_sliced_math.h"

namespace gsb
{

constexpr uint32_t kMaxSlicedQueries = 1024;
constexpr uint32_t kSlicedListEntries = 20480; // u16 list entries of one query block in shared memory
// The tile buffer: 32 batch regions, and never less than the staging area of the select rounds
// (candidate keys + bucket histogram), which runs between tiles and borrows it.
__host__ __device__ constexpr uint32_t sliced_tile_bytes(uint32_t words)
{
    return kSlicedTileBatches * sliced_region_bytes(words) > kBatchListCap * 8u + kBuckets * 4u
               ? kSlicedTileBatches * sliced_region_bytes(words)
               : kBatchListCap * 8u + kBuckets * 4u;
}
constexpr uint32_t kSlicedMiniBatches = 4;     // batches of a CTA's very first (warm-up) tile
constexpr uint32_t kSlicedWarmupTiles = 4;     // tiles of a CTA after which threshold sharing is overlapped
constexpr uint32_t kSlicedPruneMin = 64;       // lists longer than this are pruned when their threshold rises
constexpr uint32_t kSlicedWarpSortMax = 64;    // final lists up to this length are sorted by one warp

// Per-query constants and filter state, one 16-byte shared-memory load per (query, tile).
struct SlicedQuery {
    uint32_t lofs;  // first list entry (global numbering)
    float tq;       // filter factor, see sliced_tq
    uint16_t m;     // filter: common >= m (batch-independent part)
    uint16_t popq;  // query popcount
    uint16_t ngrp;  // list length in groups of 8
    uint16_t flags; // kSlicedDirty: tau rose since the list was last pruned; kSlicedSorted: list is final
};
constexpr uint16_t kSlicedDirty = 1, kSlicedSorted = 2;
constexpr uint32_t kSlicedPerQueryBytes = 8 + 8 + 4 + sizeof(SlicedQuery);

// Query blocks: consecutive queries whose lists fit the shared-memory list area together.
struct SlicedMeta {
    uint32_t n_blocks;
    uint32_t blk_start[kMaxSlicedQueries + 1];
};

struct SlicedParams {
    BatchParams b;                       // database, k, cutoff, nq, candidate lists, outputs
    const uint16_t* lists;               // set-bit entries of all queries (sliced_entry), padded to groups of 8
    const uint32_t* lofs;                // [nq] first entry of each query's list
    const uint16_t* ngrp;                // [nq] groups of 8 entries
    const uint16_t* popq;                // [nq] query popcounts
    const SlicedMeta* meta;
    unsigned int* ghist;                 // [nq][kSlicedHistBuckets] scores of all candidates so far; zero on entry
    unsigned long long* gtau;            // [nq] thresholds shared by all CTAs; starting values on entry
    // claims [0, n_mini) are mini tiles: batches [0, kSlicedMiniBatches) of tile c, the warm-up of
    // CTA c; claim n_mini + t is tile t (without those batches where t < n_mini)
    uint32_t n_claims, n_mini;
};

__device__ __forceinline__ void fence_proxy_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// Ask L2 for bytes that a later TMA copy will pull into shared memory.
__device__ __forceinline__ void prefetch_l2_bulk(const void* gmem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
}
// 32-bit shared-memory load at base + off.  The add is written as a multiply-add so that it can go
// to the FMA pipe: the logic (ALU) pipe is what bounds the counting loop.
__device__ __forceinline__ uint32_t lds_u32(uint32_t base, uint32_t off)
{
    uint32_t addr, v;
    asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(addr) : "r"(off), "r"(base));
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// Lists, popcounts and query blocks from the raw queries ([nq][words]): one thread per query.
GSB_KERNEL void __launch_bounds__(kMaxSlicedQueries, 1)
sliced_build_lists_kernel(const uint32_t* __restrict__ queries, uint32_t nq, uint32_t words, uint16_t* lists,
                          uint32_t* lofs, uint16_t* ngrp, uint16_t* popq, SlicedMeta* meta)
{
    __shared__ uint32_t s_scan[kMaxSlicedQueries];
    const uint32_t j = threadIdx.x;
    uint32_t pc = 0;
    if (j < nq)
        for (uint32_t w = 0; w < words; w++)
            pc += __popc(queries[j * words + w]);
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
        for (uint32_t w = 0; w < words; w++) {
            uint32_t x = queries[j * words + w];
            while (x) {
                const uint32_t b = __ffs(x) - 1;
                x &= x - 1;
                dst[n++] = sliced_entry(w * 32 + b);
            }
        }
        for (; n < padded; n++)
            dst[n] = sliced_zero_entry(words);
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
    }
}

// Descending sort of list[0, n), n <= 64, by one warp (keys are distinct: they carry the row);
// keeps the best min(n, k).
__device__ __forceinline__ void sliced_warp_sort(unsigned long long* list, uint32_t n, uint32_t k, uint32_t lane)
{
    const unsigned long long k0 = lane < n ? list[lane] : 0ull, k1 = lane + 32 < n ? list[lane + 32] : 0ull;
    uint32_t r0 = 0, r1 = 0;
    for (uint32_t i = 0; i < n; i++) {
        const unsigned long long other = __shfl_sync(0xffffffffu, i < 32 ? k0 : k1, i & 31);
        r0 += other > k0;
        r1 += other > k1;
    }
    __syncwarp();
    if (lane < n && r0 < k)
        list[r0] = k0;
    if (lane + 32 < n && r1 < k)
        list[r1] = k1;
}

// W = 32-bit words per row (4, 8, 16 or 32), CW = warps per CTA.
template <int W, int CW>
__global__ void __launch_bounds__(CW * 32, 1) scan_sliced_kernel(const __grid_constant__ SlicedParams sp)
{
    constexpr int NT = CW * 32;
    constexpr uint32_t kRegion = sliced_region_bytes(W);  // shared memory per batch
    constexpr uint32_t kTileBytes = sliced_tile_bytes(W);
    constexpr uint32_t kGang = 32 / W;                    // batches one warp transposes at a time
    constexpr uint32_t kFull = 0xffffffffu;
    constexpr uint32_t kTauPerThread = (kMaxSlicedQueries + NT - 1) / NT;
    const BatchParams& p = sp.b;

    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_full[CW];
    __shared__ unsigned long long s_stage_tau;
    __shared__ unsigned int s_stage_count, s_dummy_epoch, s_alive, s_next_q, s_need_select, s_error;
    __shared__ unsigned int s_claim[3]; // tile claims of this, the next and the next-but-one iteration
    __shared__ float s_pdmin[kSlicedTileBatches]; // smallest row popcount of each batch of the tile

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nq = p.nq, nqp = (nq + 7u) & ~7u;
    // shared memory carve-up
    uint8_t* tile = smem;                                                        // 32 batch regions
    uint16_t* s_pd = reinterpret_cast<uint16_t*>(smem + kTileBytes);       // [1024] row popcounts of the tile
    uint16_t* s_list = s_pd + kSlicedTileBatches * kBatchRows;                   // [kSlicedListEntries]
    uint8_t* cursor = reinterpret_cast<uint8_t*>(s_list + kSlicedListEntries);
    SlicedQuery* s_qc = reinterpret_cast<SlicedQuery*>(cursor);                  // [nq]
    cursor += (size_t) nqp * sizeof(SlicedQuery);
    unsigned long long* s_tau = reinterpret_cast<unsigned long long*>(cursor);   // [nq] keys <= tau are out
    cursor += (size_t) nqp * 8;
    unsigned long long* s_surv = reinterpret_cast<unsigned long long*>(cursor);  // [nq]
    cursor += (size_t) nqp * 8;
    unsigned int* s_cnt = reinterpret_cast<unsigned int*>(cursor);               // [nq] list fill
    // the select rounds run between tiles and stage through the (then dead) tile buffer
    CandShared cs;
    cs.buf = reinterpret_cast<unsigned long long*>(tile);
    cs.cap = kBatchListCap;
    cs.count = &s_stage_count;
    cs.tau = &s_stage_tau;
    cs.epoch_req = &s_dummy_epoch;
    cs.hist = reinterpret_cast<unsigned int*>(tile + (size_t) kBatchListCap * 8);
    cs.error = &s_error;
    cs.counted = nullptr;
    unsigned long long* my_cand = p.cand + (uint64_t) blockIdx.x * nq * kBatchListCap;

    const bool drop_zero = p.cutoff > 0.0f; // reference .cu:265
    // With a cutoff every row at or above it must be seen (survivor count), so the filter follows
    // the cutoff; without one it follows the query's threshold.
    auto update_filter = [&](uint32_t j) {
        const float ts = drop_zero ? p.cutoff : __uint_as_float(static_cast<uint32_t>((s_tau[j] + 1ull) >> 32));
        const uint32_t m = sliced_filter_min(ts, s_qc[j].popq, [](uint32_t c, uint32_t u) { return tanimoto_div(c, u); });
        s_qc[j].m = static_cast<uint16_t>(m > 0xffffu ? 0xffffu : m);
        s_qc[j].tq = sliced_tq(ts);
    };

    for (uint32_t j = tid; j < nq; j += NT) {
        s_tau[j] = sp.gtau[j];
        s_surv[j] = 0;
        s_cnt[j] = 0;
        SlicedQuery qc;
        qc.lofs = sp.lofs[j];
        qc.popq = sp.popq[j];
        qc.ngrp = sp.ngrp[j];
        qc.flags = 0;
        qc.m = 0;
        qc.tq = 0.0f;
        s_qc[j] = qc;
        update_filter(j);
    }
    if (tid == 0) {
        s_alive = 0;
        s_need_select = 0;
        s_error = 0;
        // claims below n_mini are the mini tiles, one per CTA and handed out statically (every CTA
        // warms up on its own); the tiles proper are claimed dynamically, two iterations ahead
        s_claim[0] = blockIdx.x < sp.n_mini ? blockIdx.x : sp.n_mini + atomicAdd(&p.ctrl->next_batch, 1u);
        s_claim[1] = sp.n_mini + atomicAdd(&p.ctrl->next_batch, 1u);
    }
    if (lane == 0) {
        mbar_init(&s_full[warp], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const uint32_t n_blocks = sp.meta->n_blocks;
    auto load_lists = [&](uint32_t blk) { // all threads; the caller syncs
        const uint32_t q0 = sp.meta->blk_start[blk], q1 = sp.meta->blk_start[blk + 1];
        if (q1 <= q0)
            return;
        const uint32_t first = s_qc[q0].lofs, last = s_qc[q1 - 1].lofs + s_qc[q1 - 1].ngrp * kSlicedGroup;
        const uint4* src = reinterpret_cast<const uint4*>(sp.lists + first);
        uint4* dst = reinterpret_cast<uint4*>(s_list);
        for (uint32_t i = tid; i < (last - first) / 8; i += NT)
            dst[i] = src[i];
    };
    if (n_blocks == 1)
        load_lists(0);

    const uint32_t my_T32 = smem_u32(tile + sliced_lane_base(lane, W)); // this lane's transposed batch
    const uint32_t row_base32 = static_cast<uint32_t>(p.row_base);
    uint32_t phase = 0;

    // Common-bit counts of one query for the 32 rows of this lane's batch; NP = counter planes
    // above "fours".  Entries are read one by one (16-bit broadcast loads: no unpacking arithmetic,
    // the ALU pipe is the bound of this loop).
    auto count_query = [&](auto& cnt, const SlicedQuery qc, uint32_t list_base) {
        const uint16_t* lp = s_list + (qc.lofs - list_base);
        const uint32_t ng = qc.ngrp;
        uint32_t g = 0;
        for (; g + 2 <= ng; g += 2, lp += 2 * kSlicedGroup) {
            uint32_t x[2 * kSlicedGroup];
#pragma unroll
            for (int i = 0; i < static_cast<int>(2 * kSlicedGroup); i++)
                x[i] = lds_u32(my_T32, lp[i]);
            cnt.add16(x);
        }
        if (g < ng) {
            uint32_t x[kSlicedGroup];
#pragma unroll
            for (int i = 0; i < static_cast<int>(kSlicedGroup); i++)
                x[i] = lds_u32(my_T32, lp[i]);
            cnt.add8(x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);
        }
    };
    // Filter, then the exact path for the rows of lanes [lane_lo, lane_hi) that pass it.
    auto finish_query = [&](const auto& cnt, uint32_t j, const SlicedQuery qc, uint32_t b0, uint32_t lane_lo,
                            uint32_t lane_hi, float pdmin) {
        constexpr int NP = std::remove_reference_t<decltype(cnt)>::kPlanes - 3;
        const uint32_t pq = qc.popq;
        // this lane's bound: its batch has no row with fewer than pdmin set bits
        const uint32_t ml = max(static_cast<uint32_t>(qc.m), sliced_lane_min(qc.tq, static_cast<float>(pq) + pdmin));
        uint32_t ge = cnt.at_least_lane(ml);
        if (lane < lane_lo || lane >= lane_hi || (ml >> (3 + NP)) != 0)
            ge = 0; // no batch for this lane in this tile; or a bound no count of this width reaches
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
                    if (idx < kBatchListCap) // (lists are cut to <= 1024 entries between tiles)
                        my_cand[(uint64_t) j * kBatchListCap + idx] = key;
                    else
                        atomicOr(&s_error, kErrOverflow);
                    // every candidate is counted once in the grid-wide score histogram of its query
                    const uint32_t bucket = sliced_bucket(__float_as_uint(score));
                    const unsigned peers = __match_any_sync(pm, bucket);
                    if (lane == static_cast<uint32_t>(__ffs(peers) - 1))
                        atomicAdd(&sp.ghist[(uint64_t) j * kSlicedHistBuckets + bucket], static_cast<unsigned int>(__popc(peers)));
                }
            }
        }
    };

    for (uint32_t it = 0;; it++) {
        // everybody is done with the tile buffer (queries / select round of the previous tile)
        fence_proxy_async_smem();
        cta_sync<NT>();
        const uint32_t claim = s_claim[it % 3];
        if (claim >= sp.n_claims)
            break;
        const bool mini = claim < sp.n_mini;
        const uint32_t t_idx = mini ? claim : claim - sp.n_mini;
        const uint32_t b0 = t_idx * kSlicedTileBatches;
        uint32_t nb_tile = p.n_batches - b0 < kSlicedTileBatches ? p.n_batches - b0 : kSlicedTileBatches;
        uint32_t lane_lo = 0;
        if (mini)
            nb_tile = nb_tile < kSlicedMiniBatches ? nb_tile : kSlicedMiniBatches;
        else if (t_idx < sp.n_mini)
            lane_lo = kSlicedMiniBatches; // those batches were this tile's mini tile
        // ---- phase A: TMA the tile in (every warp its own batches), transpose in place.  A warp
        // handles gangs of 32 / W batches: lane = (batch of the gang, word column).
        uint32_t n_mine = 0;
        for (uint32_t g0 = warp * kGang; g0 < nb_tile; g0 += CW * kGang)
            n_mine += nb_tile - g0 < kGang ? nb_tile - g0 : kGang;
        if (lane == 0 && n_mine) {
            mbar_arrive_expect_tx(&s_full[warp], n_mine * p.batch_bytes);
            for (uint32_t g0 = warp * kGang; g0 < nb_tile; g0 += CW * kGang)
                for (uint32_t b = g0; b < g0 + kGang && b < nb_tile; b++)
                    tma_bulk_g2s(tile + (size_t) b * kRegion, p.tiles + (uint64_t)(b0 + b) * p.batch_stride,
                                 p.batch_bytes, &s_full[warp]);
        }
        if (tid == 0) {
            s_claim[(it + 2) % 3] = sp.n_mini + atomicAdd(&p.ctrl->next_batch, 1u); // claimed two tiles ahead
            s_next_q = 0;
        }
        // the next tile starts its way from HBM into L2 now: it has this whole tile's time to arrive
        {
            const uint32_t nclaim = s_claim[(it + 1) % 3];
            if (lane == 0 && nclaim < sp.n_claims) {
                const bool nmini = nclaim < sp.n_mini;
                const uint32_t nb0 = (nmini ? nclaim : nclaim - sp.n_mini) * kSlicedTileBatches;
                uint32_t nnb = p.n_batches - nb0 < kSlicedTileBatches ? p.n_batches - nb0 : kSlicedTileBatches;
                if (nmini)
                    nnb = nnb < kSlicedMiniBatches ? nnb : kSlicedMiniBatches;
                for (uint32_t b = warp; b < nnb; b += CW)
                    prefetch_l2_bulk(p.tiles + (uint64_t)(nb0 + b) * p.batch_stride, p.batch_bytes);
            }
        }
        if (n_blocks > 1)
            load_lists(0);
        if (n_mine) {
            mbar_wait(&s_full[warp], phase);
            phase ^= 1u;
            const uint32_t sub = lane / W, col = lane % W;
            for (uint32_t g0 = warp * kGang; g0 < nb_tile; g0 += CW * kGang) {
                const uint32_t b = g0 + sub; // this lane's batch of the gang
                const bool have = b < nb_tile;
                uint8_t* region = tile + (size_t) b * kRegion;
                const uint32_t* raw = reinterpret_cast<const uint32_t*>(region);
                uint32_t x[32]; // x[r] = word `col` of row r
#pragma unroll
                for (int r = 0; r < 32; r++)
                    x[r] = have ? raw[r * W + col] : 0u;
                // row popcounts of every batch of the gang: lane r holds row r's
                uint16_t pd[kGang];
#pragma unroll
                for (uint32_t s = 0; s < kGang; s++) {
                    const uint32_t bs = g0 + s;
                    pd[s] = bs < nb_tile
                                ? reinterpret_cast<const uint16_t*>(tile + (size_t) bs * kRegion + (size_t) kBatchRows * W * 4)[lane]
                                : static_cast<uint16_t>(0);
                    const uint32_t pd_lo = __reduce_min_sync(kFull, static_cast<uint32_t>(pd[s]));
                    if (lane == 0 && bs < nb_tile)
                        s_pdmin[bs] = static_cast<float>(pd_lo);
                }
                __syncwarp(); // in place: every lane has read its batch before anyone overwrites it
                transpose32(x);
                if (have) {
                    uint32_t* T = reinterpret_cast<uint32_t*>(tile + sliced_lane_base(b, W));
#pragma unroll
                    for (int bb = 0; bb < 32; bb++)
                        T[col * 32 + ((bb + col) & 31)] = x[bb]; // == sliced_word_index(col * 32 + bb)
                    if (col == 0)
                        T[sliced_zero_index(W)] = 0u;
                }
#pragma unroll
                for (uint32_t s = 0; s < kGang; s++)
                    if (g0 + s < nb_tile)
                        s_pd[(g0 + s) * kBatchRows + lane] = pd[s];
            }
        }
        cta_sync<NT>();
        // ---- share thresholds across the grid.  This CTA turns the global score histograms of "its"
        // queries (j = CTA, CTA + grid, ...) into thresholds: the floor of the bucket where the count
        // of candidates from the top reaches k is a lower bound of the query's k-th best score over
        // everything the grid has scanned so far.  Every thread then reads the published threshold
        // of "its" query.  During warm-up (the first tiles of a CTA) this happens after the tile, so
        // that the very next tile already profits; later at the start of a tile, where the loads
        // overlap with the counting and the values are adopted one tile late.
        unsigned long long tau_seen[kTauPerThread];
        auto share_thresholds = [&]() {
            for (uint32_t j = blockIdx.x + warp * gridDim.x; j < nq; j += CW * gridDim.x) {
                const uint4* row = reinterpret_cast<const uint4*>(sp.ghist + (uint64_t) j * kSlicedHistBuckets) + lane * 8;
                uint32_t c[32]; // this lane's 32 consecutive buckets
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint4 v = __ldcg(row + i);
                    c[4 * i] = v.x, c[4 * i + 1] = v.y, c[4 * i + 2] = v.z, c[4 * i + 3] = v.w;
                }
                uint32_t mine = 0;
#pragma unroll
                for (int i = 0; i < 32; i++)
                    mine += c[i];
                uint32_t incl = mine; // candidates in this lane's buckets and above
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_down_sync(kFull, incl, d);
                    if (lane + d < 32)
                        incl += v;
                }
                const unsigned reach = __ballot_sync(kFull, incl >= p.k);
                if (reach == 0)
                    continue; // fewer than k candidates so far
                if (lane == 31u - __clz(reach)) {
                    uint32_t acc = incl - mine, bstar = 0;
                    bool found = false;
#pragma unroll
                    for (int i = 31; i >= 0; i--) {
                        acc += c[i];
                        if (!found && acc >= p.k) {
                            found = true;
                            bstar = lane * 32 + i;
                        }
                    }
                    if (bstar >= 1) // candidates must beat tau strictly; the bucket floor itself stays eligible
                        atomicMax(&sp.gtau[j], (static_cast<unsigned long long>(sliced_bucket_floor_bits(bstar)) << 32) - 1ull);
                }
            }
#pragma unroll
            for (uint32_t i = 0; i < kTauPerThread; i++)
                tau_seen[i] = tid + i * NT < nq ? __ldcg(&sp.gtau[tid + i * NT]) : 0ull;
        };
        const bool warming_up = it < kSlicedWarmupTiles;
        if (!warming_up)
            share_thresholds();
        // ---- phase B: warps take queries a few at a time
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
            const uint32_t list_base = n_blocks == 1 ? 0u : s_qc[sp.meta->blk_start[blk]].lofs;
            for (;;) {
                uint32_t j = 0, take = 1;
                if (lane == 0) { // four at a time while plenty are left, then one by one (balanced finish)
                    const uint32_t seen = *reinterpret_cast<volatile unsigned int*>(&s_next_q);
                    take = seen + 8u * CW < q_end ? 4u : 1u;
                    j = atomicAdd(&s_next_q, take);
                }
                j = __shfl_sync(kFull, j, 0);
                take = __shfl_sync(kFull, take, 0);
                if (j >= q_end)
                    break;
                const uint32_t j_end = j + take < q_end ? j + take : q_end;
                for (; j < j_end; j++) {
                    const SlicedQuery qc = s_qc[j];
                    if (qc.m > qc.popq)
                        continue; // no row can reach this query's threshold any more
                    if (qc.ngrp <= 15) { // <= 120 set bits: counts fit 7 planes
                        SlicedCount<4> cnt;
                        count_query(cnt, qc, list_base);
                        finish_query(cnt, j, qc, b0, lane_lo, nb_tile, my_pdmin);
                    } else {
                        SlicedCount<8> cnt;
                        count_query(cnt, qc, list_base);
                        finish_query(cnt, j, qc, b0, lane_lo, nb_tile, my_pdmin);
                    }
                }
            }
        }
        cta_sync<NT>();
        // ---- between tiles: adopt the shared thresholds
        if (warming_up)
            share_thresholds();
#pragma unroll
        for (uint32_t i = 0; i < kTauPerThread; i++) {
            const uint32_t j = tid + i * NT;
            if (j < nq && tau_seen[i] > s_tau[j]) {
                s_tau[j] = tau_seen[i];
                update_filter(j);
                s_qc[j].flags |= kSlicedDirty;
            }
        }
        cta_sync<NT>();
        // ---- lists whose threshold rose drop the entries that fell below it (one warp per list)
        for (uint32_t j = warp; j < nq; j += CW) {
            const uint32_t n = s_cnt[j];
            if (n <= kSlicedPruneMin || !(s_qc[j].flags & kSlicedDirty))
                continue;
            const unsigned long long tau = s_tau[j];
            unsigned long long* list = my_cand + (uint64_t) j * kBatchListCap;
            uint32_t out = 0;
            // (the list lives in global memory: four independent loads per lane in flight, then the in-place
            // compaction of those 128 entries — writes never pass what has been read)
            for (uint32_t base = 0; base < n; base += 128) {
                unsigned long long key[4];
#pragma unroll
                for (int e = 0; e < 4; e++)
                    key[e] = base + e * 32 + lane < n ? list[base + e * 32 + lane] : 0ull;
                __syncwarp();
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const bool keep = key[e] > tau;
                    const unsigned km = __ballot_sync(kFull, keep);
                    if (keep)
                        list[out + __popc(km & ((1u << lane) - 1u))] = key[e];
                    out += __popc(km);
                }
            }
            __syncwarp();
            if (lane == 0) {
                s_cnt[j] = out;
                s_qc[j].flags &= static_cast<uint16_t>(~kSlicedDirty);
            }
        }
        // ---- lists that still hold more than 1024 entries are cut back (staged through the tile buffer)
        if (*reinterpret_cast<volatile unsigned int*>(&s_need_select)) {
            batch_select_round<NT>(cs, my_cand, s_cnt, s_tau, nq, p.k, false, tid,
                                   [&](uint32_t j) { update_filter(j); }, [](uint32_t) { return false; });
            if (tid == 0)
                s_need_select = 0;
        }
    }
    // ---- final per-CTA lists: short ones are sorted by one warp each, the rest by the whole CTA
    for (uint32_t j = warp; j < nq; j += CW) {
        const uint32_t n = s_cnt[j];
        if (n > kSlicedWarpSortMax)
            continue;
        sliced_warp_sort(my_cand + (uint64_t) j * kBatchListCap, n, p.k, lane);
        if (lane == 0) {
            s_cnt[j] = n < p.k ? n : p.k;
            s_qc[j].flags |= kSlicedSorted;
        }
    }
    cta_sync<NT>();
    batch_finish<NT>(p, cs, my_cand, s_cnt, s_tau, s_surv, &s_alive, tid,
                     [&](uint32_t j) { return (s_qc[j].flags & kSlicedSorted) != 0; });
}

} // namespace gsb
