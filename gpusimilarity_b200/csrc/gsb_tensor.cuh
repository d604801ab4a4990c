// Tensor-core multi-query scan for 1024-bit rows: the common-bit counts of 128 queries x 128 rows
// per tile come out of tcgen05.mma (kind::i8) instead of the CUDA cores, so the cost of a pass no
// longer depends on how many bits the queries have set (the bit-sliced kernel, gsb_sliced.cuh, pays
// per set bit and wins for sparse queries; the host picks by the batch's set-bit total).
//
//   common = popc(q & d)  (reference TanimotoFunctor, fingerprintdb_cuda.cu:97-98)
//          = dot product of the two bit vectors: D[128 queries][128 rows] += A[128 x K] * B[K x 128]
// with K = 1024 bit positions expanded to bytes (gsb_tensor_math.h: 2^plane on the row side,
// 2^(7-plane) on the query side, D = 128 * common exactly in s32).  HBM keeps the packed rows
// (130 B per row and pass); the expansion happens on the way through shared memory.
//
// One CTA per SM, 20 warps, warp-specialised:
//   warp 0        TMA producer: 128-row tiles (4 batches of 32 rows + popcount trailers) into a raw ring
//   warps 4-11    expanders: a warp keeps 16 raw rows in registers (lane = row word) and writes the
//                 four quarter-K slabs of the tile (bit planes 2g, 2g+1) as the B operand: canonical
//                 K-major core matrices without swizzle, 8-byte stores, bank-conflict free
//   warp 1        MMA issuer (one thread): 8 x tcgen05.mma M128 N128 K32 per slab, A = the queries
//                 resident in tensor memory (256 columns, written once), D double-buffered in the
//                 other 256 columns; tcgen05.commit releases slabs and publishes accumulators
//   warps 12-19   epilogue: tcgen05.ld the accumulators (lane = query, column = row), one FMA +
//                 compare per value against the query's running threshold (tc_filter_*); the rare
//                 values that pass are scored exactly (tanimoto_div, popcount trailer) and appended
//                 to the query's candidate list — same keys, cutoff and survivor rules, select
//                 rounds and end of launch as the other multi-query kernels (gsb_batch.cuh)
//   warp 2        turns the grid-wide candidate score histograms of "its" queries into thresholds
//                 and publishes them (as share_thresholds of the bit-sliced kernel)
// A launch takes up to 128 queries; larger batches are one launch per 128 queries.
#pragma once

#include "gsb_tensor_params.h"

namespace gsb
{

// setmaxnreg per warpgroup (48 / 56 / 160) compiles and its toy form runs (tools/probe/setmaxnreg_probe.cu),
// but this kernel hangs with it on the B200 (first launch never returns); off until that is understood.
#ifndef GSB_TC_SETMAXNREG
#define GSB_TC_SETMAXNREG 0
#endif
#if GSB_TC_SETMAXNREG
#define TC_SETMAXNREG(what) asm volatile("setmaxnreg." what ";")
#else
#define TC_SETMAXNREG(what) do { } while (0)
#endif

// ---- tcgen05 wrappers (PTX ISA 8.6+, sm_100a) --------------------------------------------------
__device__ __forceinline__ void tc_fence_before()
{
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after()
{
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]; u8 x u8 -> s32
__device__ __forceinline__ void tc_mma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
// arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&a)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(a[0]),
                 "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7])
                 : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
                 "%14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ bool tc_elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "elect.sync _|p, 0xffffffff;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(pred));
    return pred != 0;
}
// Shared-memory matrix descriptor: K-major, no swizzle (core matrix = 8 rows x 16 bytes, contiguous),
// leading byte offset = distance of the two 16-byte chunks of a K-step, stride byte offset =
// distance of consecutive 8-row groups; bits 46-47 = descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr)
{
    return static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4) | (static_cast<uint64_t>(kTcLbo >> 4) << 16) |
           (static_cast<uint64_t>(kTcSbo >> 4) << 32) | (1ull << 46);
}
// Instruction descriptor, kind::i8: D = s32 (bits 4-5 = 2), A and B unsigned 8 bit (bits 7-12 = 0),
// both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t kTcIdesc = (2u << 4) | ((kTcTileRows >> 3) << 17) | ((kTcQueries >> 4) << 24);

// Pipeline wait with a way out: never hangs and never traps.  A wait that outlasts the launch's spin
// bound raises kErrPipeline (reported like a grid-barrier timeout: every count becomes
// GSB_COUNT_ERROR) and sets the abort flag; every role leaves its loop at its next wait (false is
// returned), touches no pipeline barrier again, and the CTA runs to its end: the context stays usable.
// (inlined at every wait: ptxas cannot give a function called from warpgroups with different setmaxnreg
// limits one register allocation — C7600 —, and the loop is a dozen instructions)
__device__ __forceinline__ bool tc_wait_slow(uint64_t* bar, uint32_t parity, volatile unsigned int* abort_flag,
                                         unsigned int* error, unsigned long long timeout_ns)
{
    const unsigned long long t0 = global_ns();
    for (uint32_t spins = 0;; spins++) {
        if (mbar_try_wait(bar, parity))
            return true;
        if ((spins & 63u) == 63u) {
            if (*abort_flag)
                return false;
            if (global_ns() - t0 > timeout_ns) {
                atomicOr(error, kErrPipeline);
                *abort_flag = 1;
                return false;
            }
        }
    }
}
__device__ __forceinline__ bool tc_wait(uint64_t* bar, uint32_t parity, volatile unsigned int* abort_flag,
                                        unsigned int* error, unsigned long long timeout_ns)
{
    if (mbar_try_wait(bar, parity))
        return true;
    return tc_wait_slow(bar, parity, abort_flag, error, timeout_ns);
}

// Everything the exact path of one (query, row) pair needs; lives in local memory, the path is rare.
struct TcExact {
    const BatchParams* p;
    unsigned long long* list;       // this CTA's candidate list of the query
    volatile unsigned long long* tau;
    unsigned long long* surv;
    unsigned int* cnt;
    unsigned int* need_select;
    unsigned int* error;
    unsigned int* ghist;            // the query's row of the grid-wide histogram
    uint32_t pq, row0, row_base32;  // query popcount, first row of the tile (shard numbering), shard base
    bool drop_zero;
    bool hist_on;                   // count candidates in the grid-wide histogram (off for most CTAs' first tile)
};

__device__ __noinline__ void tc_exact(const TcExact& e, uint32_t d, uint32_t col, float pdf)
{
    const BatchParams& p = *e.p;
    const uint32_t common = d >> 7, pd = static_cast<uint32_t>(pdf), row_local = e.row0 + col;
    if (row_local >= p.n_rows)
        return;
    // reference .cu:100-102: IEEE divide, then the cutoff test (NaN -> 0)
    float score = tanimoto_div(common, e.pq + pd - common);
    score = (score >= p.cutoff) ? score : 0.0f;
    if (e.drop_zero) { // .cu:265-271
        if (score == 0.0f)
            return;
        atomicAdd(e.surv, 1ull);
    }
    const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(score)) << 32) |
                                   static_cast<unsigned long long>(0xffffffffu - (e.row_base32 + row_local));
    if (!(key > *e.tau))
        return;
    const unsigned idx = atomicAdd(e.cnt, 1u);
    if (idx + 1 > kBatchListCap / 2)
        *reinterpret_cast<volatile unsigned int*>(e.need_select) = 1; // cut the list back at the next tile boundary
    if (idx < kBatchListCap)
        e.list[idx] = key;
    else
        atomicOr(e.error, kErrOverflow);
    if (e.hist_on)
        atomicAdd(&e.ghist[sliced_bucket(__float_as_uint(score))], 1u);
}

// Level 2 of the epilogue filter for one 16-column chunk that passed level 1: the per-row test,
// then the exact path.  Out of line on purpose (see tc_wait).
__device__ __noinline__ void tc_level2(const TcExact& e, float slope, float thr, const float* pdf, uint32_t col0,
                                       uint32_t d0, uint32_t d1, uint32_t d2, uint32_t d3, uint32_t d4, uint32_t d5,
                                       uint32_t d6, uint32_t d7, uint32_t d8, uint32_t d9, uint32_t d10, uint32_t d11,
                                       uint32_t d12, uint32_t d13, uint32_t d14, uint32_t d15)
{
    const uint32_t d[16] = {d0, d1, d2, d3, d4, d5, d6, d7, d8, d9, d10, d11, d12, d13, d14, d15};
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const float pdv = pdf[i];
        if (__fmaf_rn(slope, pdv, __uint_as_float(kTcMagicBits | d[i])) >= thr)
            tc_exact(e, d[i], col0 + i, pdv);
    }
}

__global__ void __launch_bounds__(kTcThreads, 1) scan_tensor_kernel(const __grid_constant__ TensorParams tp)
{
    constexpr uint32_t kFull = 0xffffffffu;
    const BatchParams& p = tp.b;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_raw_full[kTcRawRing], s_raw_empty[kTcRawRing], s_slab_full[kTcSlabRing],
        s_slab_empty[kTcSlabRing], s_tmem_full[2], s_tmem_empty[2], s_pd_full[kTcPdRing], s_drain;
    __shared__ uint32_t s_tmem_base;
    __shared__ unsigned long long s_stage_tau;
    __shared__ unsigned int s_stage_count, s_dummy_epoch, s_alive, s_need_select, s_error, s_done, s_abort, s_any_dirty;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // site = which wait (0 raw_empty, 1 tmem_empty, 2 slab_full, 3 raw_full, 4 slab_empty, 5 pd_full, 6 tmem_full)
    long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool alive = true; // false once a pipeline wait of this thread gave up
#define TC_WAIT(bar, parity, site)                                                               \
    do {                                                                                         \
        if ((tp.fault == 5 || tp.fault == 8) && (site) != 0 && (site) != 3) { /* timing experiment: free-running roles */ \
        } else {                                                                                 \
            const long long t_w = tp.dbg ? clock64() : 0;                                        \
            alive = tc_wait(bar, parity, &s_abort, &s_error, p.spin_timeout_ns);                 \
            if (tp.dbg)                                                                          \
                dbg_acc[site] += clock64() - t_w;                                                \
        }                                                                                        \
        if (warp >= kTcExpWarp0 || warp == 1) /* whole warps wait there: one outcome for all lanes */ \
            alive = __all_sync(0xffffffffu, alive);                                              \
    } while (0)
    const uint32_t nq = p.nq;
    // shared memory carve-up
    uint8_t* slabs = smem;
    uint8_t* raw = slabs + kTcSlabRing * kTcSlabBytes;
    uint8_t* stage = raw + kTcRawRing * kTcRawStage;
    float* s_pdf = reinterpret_cast<float*>(stage + kBatchListCap * 8u + kBuckets * 4u); // [kTcPdRing][128]
    float* s_pdmin = s_pdf + kTcPdRing * kTcTileRows;                                    // [kTcPdRing][8 chunks of 16 rows]
    unsigned long long* s_tau = reinterpret_cast<unsigned long long*>(s_pdmin + kTcPdRing * 8u); // [128]
    unsigned long long* s_surv = s_tau + kTcQueries;
    unsigned int* s_cnt = reinterpret_cast<unsigned int*>(s_surv + kTcQueries);
    unsigned int* s_flags = s_cnt + kTcQueries;
    CandShared cs;
    cs.buf = reinterpret_cast<unsigned long long*>(stage);
    cs.cap = kBatchListCap;
    cs.count = &s_stage_count;
    cs.tau = &s_stage_tau;
    cs.epoch_req = &s_dummy_epoch;
    cs.hist = reinterpret_cast<unsigned int*>(stage + kBatchListCap * 8u);
    cs.error = &s_error;
    cs.counted = nullptr;
    unsigned long long* my_cand = p.cand + (uint64_t) blockIdx.x * nq * kBatchListCap;

    // this CTA's tiles: blockIdx.x, blockIdx.x + grid, ... (every role derives the same sequence)
    const uint32_t n_local = blockIdx.x < tp.n_tiles ? (tp.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;

    for (uint32_t j = tid; j < kTcQueries; j += kTcThreads) {
        s_tau[j] = 0;
        s_surv[j] = 0;
        s_cnt[j] = 0;
        s_flags[j] = 0;
    }
    if (tid == 0) {
        for (uint32_t i = 0; i < kTcRawRing; i++) {
            mbar_init(&s_raw_full[i], 1);
            mbar_init(&s_raw_empty[i], kTcExpWarps);
        }
        for (uint32_t i = 0; i < kTcSlabRing; i++) {
            mbar_init(&s_slab_full[i], kTcExpWarps);
            mbar_init(&s_slab_empty[i], 1);
        }
        for (uint32_t i = 0; i < 2; i++) {
            mbar_init(&s_tmem_full[i], 1);
            mbar_init(&s_tmem_empty[i], kTcEpiWarps);
        }
        for (uint32_t i = 0; i < kTcPdRing; i++)
            mbar_init(&s_pd_full[i], kTcExpWarps);
        mbar_init(&s_drain, 1);
        mbar_fence_init();
        s_alive = 0;
        s_need_select = 0;
        s_error = 0;
        s_done = 0;
        s_abort = 0;
        s_any_dirty = 0;
    }
    if (warp == 1) { // tensor memory: all 512 columns (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "n"(kTcTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;

    // ---- the queries become the A operand: thread = query = lane of tensor memory
    uint32_t pq = 0;
    const uint32_t ew = warp - kTcEpiWarp0;           // epilogue warp number (valid for warps >= 12)
    const uint32_t quarter = warp & 3u;               // the tensor-memory lanes a warp may touch: 32 * (warp % 4)
    const uint32_t qj = quarter * 32u + lane;         // the query of an epilogue thread
    if (warp >= kTcEpiWarp0) {
        uint32_t wq[32];
#pragma unroll
        for (int i = 0; i < 32; i++)
            wq[i] = qj < nq ? __ldg(&p.queries[(size_t) qj * 32 + i]) : 0u;
#pragma unroll
        for (int i = 0; i < 32; i++)
            pq += __popc(wq[i]);
        if (ew < 4) {
#pragma unroll
            for (uint32_t ks = 0; ks < kTcSlabs * kTcSlabSteps; ks++) {
                uint32_t a[8];
#pragma unroll
                for (uint32_t u = 0; u < 8; u++) {
                    uint32_t iw = 0, plane = 0;
                    tc_a_source(ks, u, &iw, &plane);
                    a[u] = tc_query_word(wq[iw], plane);
                }
                tc_st8(tmem + ((quarter * 32u) << 16) + ks * 8u, a);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // (GSB_TC_SETMAXNREG: registers follow the roles per warpgroup, 4 x 48 + 8 x 56 + 8 x 160 = 20 x 96)
    // (each setmaxnreg sits at the head of its warpgroup's branch: that is how ptxas ties the new limit to the code)

    const long long dbg_t0 = tp.dbg ? clock64() : 0;
    if (warp < kTcExpWarp0) {
    //TC_SETMAXNREG("dec.sync.aligned.u32 48");
    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            for (uint32_t n = 0; n < n_local && alive; n++) {
                const uint32_t t = blockIdx.x + n * gridDim.x, slot = n % kTcRawRing, use = n / kTcRawRing;
                TC_WAIT(&s_raw_empty[slot], (use & 1u) ^ 1u, 0);
                if (!alive)
                    break;
                const uint32_t b0 = t * kTcTileBatches;
                const uint32_t nb = p.n_batches - b0 < kTcTileBatches ? p.n_batches - b0 : kTcTileBatches;
                mbar_arrive_expect_tx(&s_raw_full[slot], nb * kTcRawBatch);
                if (tp.fault == 1 && blockIdx.x == 0 && n == 0)
                    continue;
                for (uint32_t b = 0; b < nb; b++)
                    tma_bulk_g2s(raw + slot * kTcRawStage + b * kTcRawBatch, p.tiles + (uint64_t)(b0 + b) * p.batch_stride,
                                 kTcRawBatch, &s_raw_full[slot]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp walks the loop (every address is warp-uniform: uniform registers, no
        // per-lane waterfall around the tensor-core instructions); one elected lane issues.
        const bool leader = tc_elect_one();
        const uint32_t tmem_u = __shfl_sync(kFull, tmem, 0);
        const uint32_t slabs_u = __shfl_sync(kFull, smem_u32(slabs), 0);
        for (uint32_t n = 0; n < n_local && alive; n++) {
            const uint32_t buf = n & 1u;
            TC_WAIT(&s_tmem_empty[buf], ((n >> 1) & 1u) ^ 1u, 1); // the epilogue has drained this accumulator tile
            if (!alive)
                break;
            tc_fence_after();
            const uint32_t d_tmem = tmem_u + kTcTmemD + buf * kTcTileRows;
            for (uint32_t g = 0; g < kTcSlabs; g++) {
                const uint32_t s = n * kTcSlabs + g, slot = s % kTcSlabRing;
                TC_WAIT(&s_slab_full[slot], (s / kTcSlabRing) & 1u, 2);
                if (!alive)
                    break;
                tc_fence_after();
                const uint32_t slab_addr = slabs_u + slot * kTcSlabBytes;
                const long long t_a = tp.dbg ? clock64() : 0;
                if (leader) {
#pragma unroll
                    for (uint32_t st = 0; st < kTcSlabSteps; st++)
                        tc_mma_i8_ts(d_tmem, tmem_u + (g * kTcSlabSteps + st) * 8u, tc_smem_desc(slab_addr + 2u * st * kTcLbo),
                                     kTcIdesc, (g | st) != 0u ? 1u : 0u);
                    tc_commit(&s_slab_empty[slot]); // the slab may be overwritten once these MMAs have read it
                }
                __syncwarp();
                if (tp.dbg)
                    dbg_acc[0] += clock64() - t_a;
            }
            if (!alive)
                break;
            if (leader)
                tc_commit(&s_tmem_full[buf]);
            __syncwarp();
        }
        if (!alive || tp.fault == 5 || tp.fault == 8) { // gave up: let the MMAs in flight finish before tensor memory is released
            if (leader) {
                tc_commit(&s_drain);
                const unsigned long long t0 = global_ns();
                while (!mbar_try_wait(&s_drain, 0) && global_ns() - t0 < p.spin_timeout_ns) {
                }
            }
            __syncwarp();
        }
    } else if (warp == 2) {
        // ================= threshold publisher =================
        // The floor of the histogram bucket where the count of candidates from the top reaches k is a
        // lower bound of the query's k-th best score over everything the grid has scanned so far.
        while (*reinterpret_cast<volatile unsigned int*>(&s_done) == 0) {
            for (uint32_t j = blockIdx.x; j < nq; j += gridDim.x) {
                const uint4* row = reinterpret_cast<const uint4*>(tp.ghist + (uint64_t) j * kSlicedHistBuckets) + lane * 8;
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
                        atomicMax(&tp.gtau[j], (static_cast<unsigned long long>(sliced_bucket_floor_bits(bstar)) << 32) - 1ull);
                }
            }
            __nanosleep(2000);
        }
    }
    } else if (warp < kTcEpiWarp0) {
        //TC_SETMAXNREG("dec.sync.aligned.u32 56");
        // ================= expanders =================
        const uint32_t e = warp - kTcExpWarp0;         // rows [16e, 16e+16) of the tile
        const uint32_t batch = e >> 1, r_in_batch = (e & 1u) * 16u;
        for (uint32_t n = 0; n < n_local && alive; n++) {
            const uint32_t t = blockIdx.x + n * gridDim.x, slot = n % kTcRawRing;
            const uint32_t b0 = t * kTcTileBatches;
            const uint32_t nb = p.n_batches - b0 < kTcTileBatches ? p.n_batches - b0 : kTcTileBatches;
            TC_WAIT(&s_raw_full[slot], (n / kTcRawRing) & 1u, 3);
            if (!alive)
                break;
            const uint8_t* rb = raw + slot * kTcRawStage + batch * kTcRawBatch;
            const long long t_r = tp.dbg ? clock64() : 0;
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; i++)
                w[i] = batch < nb ? reinterpret_cast<const uint32_t*>(rb)[(r_in_batch + i) * 32u + lane] : 0u;
            if (lane < 16) {
                const uint32_t pd = batch < nb ? reinterpret_cast<const uint16_t*>(rb + kBatchRows * 128u)[r_in_batch + lane] : 0u;
                s_pdf[(n % kTcPdRing) * kTcTileRows + e * 16u + lane] = static_cast<float>(pd);
            }
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&s_pd_full[n % kTcPdRing]);
            if (tp.dbg)
                dbg_acc[5] += clock64() - t_r;
            for (uint32_t g = 0; g < kTcSlabs; g++) {
                const uint32_t s = n * kTcSlabs + g, sslot = s % kTcSlabRing;
                TC_WAIT(&s_slab_empty[sslot], ((s / kTcSlabRing) & 1u) ^ 1u, 4);
                if (!alive)
                    break;
                uint8_t* dst = slabs + sslot * kTcSlabBytes + tc_slab_offset(e * 16u, lane);
                const uint32_t m0 = 0x01010101u << (2u * g), m1 = m0 << 1;
                const long long t_a = tp.dbg ? clock64() : 0;
                if (tp.fault != 2 && tp.fault != 8) { // (faults 2, 3, 5, 8: timing experiments with void results)
#pragma unroll
                    for (int i = 0; i < 16; i++) // row 16e+i: + (i >> 3) row groups, + (i & 7) rows
                        *reinterpret_cast<uint2*>(dst + (i >> 3) * kTcSbo + (i & 7) * 16) = make_uint2(w[i] & m0, w[i] & m1);
                }
                const long long t_b = tp.dbg ? clock64() : 0;
                if (tp.fault != 2 && tp.fault != 8)
                    fence_proxy_async_smem(); // the tensor core reads the slab through the async proxy
                __syncwarp();
                if (tp.dbg) {
                    dbg_acc[0] += t_b - t_a;
                    dbg_acc[1] += clock64() - t_b;
                }
                if (lane == 0) {
                    mbar_arrive(&s_slab_full[sslot]);
                    if (g == 0)
                        mbar_arrive(&s_raw_empty[slot]); // the raw rows are in registers
                }
            }
        }
    } else {
        TC_SETMAXNREG("inc.sync.aligned.u32 160");
        // ================= epilogue =================
        const uint32_t etid = tid - kTcEpiWarp0 * 32u;   // 0..255
        const uint32_t half = ew >> 2;                    // columns [64 half, 64 half + 64)
        const bool drop_zero = p.cutoff > 0.0f;
        const bool live = qj < nq;
        TcExact ex;
        ex.p = &p;
        ex.list = my_cand + (uint64_t) qj * kBatchListCap;
        ex.tau = &s_tau[qj];
        ex.surv = &s_surv[qj];
        ex.cnt = &s_cnt[qj];
        ex.need_select = &s_need_select;
        ex.error = &s_error;
        ex.ghist = tp.ghist + (uint64_t) qj * kSlicedHistBuckets;
        ex.pq = pq;
        ex.row_base32 = static_cast<uint32_t>(p.row_base);
        ex.drop_zero = drop_zero;
        unsigned long long tau_cached = 0;
        float slope = 0.0f, thr = 0.0f;
        // With a cutoff every row at or above it must be seen (survivor count), so the filter follows
        // the cutoff; without one it follows the query's threshold.
        auto refresh_filter = [&]() {
            const float ts = drop_zero ? p.cutoff : __uint_as_float(static_cast<uint32_t>((tau_cached + 1ull) >> 32));
            const float tq = sliced_tq(ts);
            slope = tc_filter_slope(tq);
            thr = live ? tc_filter_threshold(tq, pq) : __int_as_float(0x7f000000); // idle lanes never pass
        };
        refresh_filter();
        for (uint32_t n = 0; n < n_local && alive; n++) {
            const uint32_t t = blockIdx.x + n * gridDim.x, buf = n & 1u;
            if (tp.fault == 8)
                continue; // (timing experiment: tensor pipe and TMA alone)
            ex.row0 = t * kTcTileRows;
            // Every value of a CTA's first tile is a candidate (threshold 0).  The first thresholds need
            // k of them per query, not 128 per CTA: only the first 16 CTAs count theirs, the other
            // 130 x 128 x nq same-address atomics on a few histogram buckets are skipped.
            ex.hist_on = n != 0 || blockIdx.x < kTcWarmCtas || tp.variant == 1;
            const bool maint = n < 4 || (n & 3u) == 3u;
            const unsigned long long g_seen = (maint && live && half == 0) ? __ldcg(&tp.gtau[qj]) : 0ull;
            TC_WAIT(&s_pd_full[n % kTcPdRing], (n / kTcPdRing) & 1u, 5);
            if (!alive)
                break;
            // the row popcounts of this half tile: 16 broadcast loads, in flight while the tile is computed
            const float* pdf = s_pdf + (n % kTcPdRing) * kTcTileRows + half * 64u;
            float4 pnext[4]; // (chunk c + 1 is loaded while chunk c is tested)
#pragma unroll
            for (int i = 0; i < 4; i++)
                pnext[i] = *reinterpret_cast<const float4*>(pdf + i * 4);
            TC_WAIT(&s_tmem_full[buf], (n >> 1) & 1u, 6);
            if (!alive)
                break;
            tc_fence_after();
            const uint32_t taddr = tmem + ((quarter * 32u) << 16) + kTcTmemD + buf * kTcTileRows + half * 64u;
            // The half tile comes out of tensor memory 16 columns at a time, the next chunk (and its
            // row popcounts) in flight while the current one is tested: 32 + 32 registers instead of
            // 64 + 64, which keeps the loop free of spills at 96 registers per thread.
            uint32_t v[2][16];
            const long long t_a = tp.dbg ? clock64() : 0;
            tc_ld16(taddr, v[0]);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // (after a pipeline timeout the accumulators are meaningless: no candidates from them)
            const bool aborted = *reinterpret_cast<volatile unsigned int*>(&s_abort) != 0 || tp.fault == 2 || tp.fault == 3 || tp.fault == 5;
            // Level 1: the per-row test itself, folded into one running maximum per 16-column chunk
            // (per value: OR of the magic exponent, FMA with the row's popcount, FMNMX; no memory
            // operation on the dependent path), one compare per chunk.  Level 2 (out of line) only
            // for the chunks in which some row can really beat the query's threshold: it finds the
            // row and takes the exact path.
#pragma unroll
            for (int c = 0; c < 4; c++) {
                float4 pcur[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    pcur[i] = pnext[i];
                    if (c < 3)
                        pnext[i] = *reinterpret_cast<const float4*>(pdf + (c + 1) * 16 + i * 4);
                }
                if (c < 3)
                    tc_ld16(taddr + (c + 1) * 16u, v[(c + 1) & 1]);
                float m = -3.0e38f;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float4 f = pcur[i];
                    m = fmaxf(m, __fmaf_rn(slope, f.x, __uint_as_float(kTcMagicBits | v[c & 1][4 * i])));
                    m = fmaxf(m, __fmaf_rn(slope, f.y, __uint_as_float(kTcMagicBits | v[c & 1][4 * i + 1])));
                    m = fmaxf(m, __fmaf_rn(slope, f.z, __uint_as_float(kTcMagicBits | v[c & 1][4 * i + 2])));
                    m = fmaxf(m, __fmaf_rn(slope, f.w, __uint_as_float(kTcMagicBits | v[c & 1][4 * i + 3])));
                }
                if (m >= thr && !aborted) { // rare: a call, so that the hot loop stays small
                    if (tp.dbg)
                        dbg_acc[3]++; // (GSB_TC_DEBUG: chunks of lane 0's query that pass level 1)
                    tc_level2(ex, slope, thr, pdf + c * 16, half * 64u + c * 16u, v[c & 1][0], v[c & 1][1], v[c & 1][2],
                              v[c & 1][3], v[c & 1][4], v[c & 1][5], v[c & 1][6], v[c & 1][7], v[c & 1][8], v[c & 1][9],
                              v[c & 1][10], v[c & 1][11], v[c & 1][12], v[c & 1][13], v[c & 1][14], v[c & 1][15]);
                }
                if (c < 3)
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c == 2) { // the last chunk is in registers: the next-but-one tile may start
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0)
                        mbar_arrive(&s_tmem_empty[buf]);
                }
            }
            if (tp.dbg)
                dbg_acc[1] += clock64() - t_a; // (GSB_TC_DEBUG: epilogue slot 1 = tensor-memory loads + filter)
            // ---- every fourth tile (every tile during warm-up), the eight epilogue warps only: adopt
            // the shared thresholds, cut lists back.  A list grows by at most 128 entries per tile and
            // is cut once it is half full, so four tiles between two looks cannot overflow it.
            if (!maint)
                continue;
            const long long t_m = tp.dbg ? clock64() : 0;
            auto maint_done = [&]() {
                if (tp.dbg)
                    dbg_acc[4] += clock64() - t_m; // (GSB_TC_DEBUG: epilogue slot 4 = maintenance)
            };
            // thread (half 0, lane) owns query qj: adopt the grid-wide threshold if it rose.  (The
            // partner thread of the other half tile may still be in tc_exact and read s_tau[qj] while
            // it is written here: one aligned 64-bit store, and either value is a valid lower bound.)
            const bool adopted = live && half == 0 && g_seen > s_tau[qj];
            if (adopted)
                s_tau[qj] = g_seen;
            unsigned dirty_mask = __ballot_sync(kFull, adopted); // (queries 32 quarter + bit; half-0 warps)
            if (dirty_mask && lane == 0)
                *reinterpret_cast<volatile unsigned int*>(&s_any_dirty) = n + 1u; // (a stamp, never cleared)
            cta_sync<kTcEpiThreads>(); // every candidate of the tiles so far is in its list, every new threshold in s_tau
            if (*reinterpret_cast<volatile unsigned int*>(&s_need_select)) {
                batch_select_round<kTcEpiThreads>(cs, my_cand, s_cnt, s_tau, nq, p.k, false, etid,
                                                  [&](uint32_t) { // (the list was just cut and holds its new tau: no prune)
                                                      *reinterpret_cast<volatile unsigned int*>(&s_any_dirty) = n + 1u;
                                                  },
                                                  [](uint32_t) { return false; });
                if (etid == 0)
                    s_need_select = 0;
                cta_sync<kTcEpiThreads>();
            }
            if (*reinterpret_cast<volatile unsigned int*>(&s_any_dirty) != n + 1u) {
                maint_done();
                continue; // no threshold moved: nothing to prune, no filter to refresh
            }
            // lists whose threshold rose drop the entries that fell below it (the warp that owns the query)
            while (dirty_mask) {
                const uint32_t j = quarter * 32u + (__ffs(dirty_mask) - 1u);
                dirty_mask &= dirty_mask - 1u;
                const uint32_t cnt = s_cnt[j];
                if (cnt > kTcPruneMin) {
                    const unsigned long long tau = s_tau[j];
                    unsigned long long* list = my_cand + (uint64_t) j * kBatchListCap;
                    uint32_t out = 0;
                    // (the list lives in global memory: four independent loads per lane in flight, then the in-place
                    // compaction of those 128 entries — writes never pass what has been read)
                    for (uint32_t base = 0; base < cnt; base += 128) {
                        unsigned long long key[4];
#pragma unroll
                        for (int e = 0; e < 4; e++)
                            key[e] = base + e * 32 + lane < cnt ? list[base + e * 32 + lane] : 0ull;
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
                    if (lane == 0)
                        s_cnt[j] = out;
                }
                __syncwarp();
            }
            if (live) {
                const unsigned long long tau_now = s_tau[qj];
                if (tau_now != tau_cached) {
                    tau_cached = tau_now;
                    refresh_filter();
                }
            }
            cta_sync<kTcEpiThreads>(); // (pruned counts are final before the next tile appends)
            maint_done();
        }
        if (etid == 0)
            *reinterpret_cast<volatile unsigned int*>(&s_done) = 1;
        // ---- final per-CTA lists: short ones are sorted by one warp each, the rest by the eight warps together
        cta_sync<kTcEpiThreads>();
        for (uint32_t j = ew; j < nq; j += kTcEpiWarps) {
            const uint32_t cnt = s_cnt[j];
            if (cnt > kTcWarpSortMax)
                continue;
            sliced_warp_sort(my_cand + (uint64_t) j * kBatchListCap, cnt, p.k, lane);
            if (lane == 0) {
                s_cnt[j] = cnt < p.k ? cnt : p.k;
                s_flags[j] |= kSlicedSorted;
            }
        }
        cta_sync<kTcEpiThreads>();
        batch_finish<kTcEpiThreads>(p, cs, my_cand, s_cnt, s_tau, s_surv, &s_alive, etid,
                                    [&](uint32_t j) { return (s_flags[j] & kSlicedSorted) != 0; });
    }
    if (tp.dbg && lane == 0) {
        dbg_acc[7] = clock64() - dbg_t0;
        for (int i = 0; i < 8; i++)
            tp.dbg[((uint64_t) blockIdx.x * kTcWarps + warp) * 8 + i] = static_cast<unsigned long long>(dbg_acc[i]);
    }
    // ---- teardown: every MMA has completed (the epilogue waited for all accumulator tiles)
    tc_fence_before();
    __syncthreads();
#undef TC_WAIT
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTcTmemCols) : "memory");
}

} // namespace gsb
