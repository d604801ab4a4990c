// Host side of the C ABI declared in include/gpusim_b200.h: database life cycle, HBM layout,
// launch planning and result marshalling around the kernels in gsb_kernels.cuh.
//
// Reference interfaces replaced (see the header for the per-function map):
//   FingerprintDB ctor / copyToGPU / search / search_storage / getFingerprint
//   (fingerprintdb_cuda.cu:117-381), search_cpu / fold_data (fingerprintdb_cuda.cpp:20-69),
//   get_gpu_count / get_next_gpu / get_available_gpu_memory (fingerprintdb_cuda.cu:33-68,401-413).
#include "gsb_batch.cuh"
#include "gsb_kernels.cuh"
#include "gsb_sliced.cuh"
#include "gsb_tensor_params.h"

#include "../../include/gpusim_b200.h"
#include "gsb_internal.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <sys/mman.h>

#if defined(__x86_64__)
#include <immintrin.h>
#define GSB_CPU_RELAX() _mm_pause()
#else
#define GSB_CPU_RELAX() std::this_thread::yield()
#endif

namespace
{

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_reset_epoch{0}; // bumped by gsb_devices_reset

int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

#define GSB_CUDA(expr)                                                                           \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return fail(GSB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));      \
    } while (0)

// No exception may cross the C ABI (std::bad_alloc from a huge request would otherwise end in
// std::terminate inside the host process).
#define GSB_TRY try {
#define GSB_CATCH                                                                                \
    }                                                                                            \
    catch (const std::bad_alloc&) { return fail(GSB_ERR_NOMEM, "out of host memory"); }          \
    catch (const std::exception& e) { return fail(GSB_ERR_INVALID, std::string("exception: ") + e.what()); } \
    catch (...) { return fail(GSB_ERR_INVALID, "unknown exception"); }

int env_int(const char* name, int dflt)
{
    const char* v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}

uint32_t pow2ceil(uint64_t v)
{
    uint32_t p = 1;
    while (p < v)
        p <<= 1;
    return p;
}

// ------------------------------------------------------------------------------ HBM layout
// The database lives in HBM as 32-row batches so that one TMA bulk copy moves one batch:
//   batch b = rows [32 b, 32 b + 32), row-major, dev_words 32-bit words per row,
//             then (rowpop layout only) 32 u16 row popcounts.
// Without the trailer the batches are back to back, i.e. the plain row-major matrix.
struct Layout {
    uint32_t words = 0;      // logical words per row (after folding)
    uint32_t dev_words = 0;  // words per row in HBM: power of two in [4, 128], zero padded
    uint32_t tile_rows = 32; // rows per batch
    uint32_t row_bytes = 0;
    uint32_t tile_bytes = 0;  // bytes per bulk copy
    uint32_t tile_stride = 0; // distance between batches
    uint32_t unit_batches = 1;      // batches per work unit / TMA copy of the single-query scan (~4 KB)
    uint32_t unit_bytes = 0;        // unit_batches * tile_stride
    uint32_t stage_bytes = 0;       // ring slot of the single-query scan (one unit)
    uint32_t batch_stage_bytes = 0; // ring slot of the multi-query scan (one batch)
    bool rowpop = false;
};

int make_layout(uint32_t words, Layout* out)
{
    Layout l;
    l.words = words;
    l.dev_words = std::max<uint32_t>(4, pow2ceil(words));
    if (l.dev_words > gsb::kMaxWords)
        return fail(GSB_ERR_INVALID, "fingerprints wider than 4096 bits are not supported");
    l.row_bytes = l.dev_words * 4;
    l.tile_rows = gsb::kBatchRows;
    l.rowpop = env_int("GSB_ROWPOP", 1) != 0;
    l.tile_bytes = l.tile_rows * l.row_bytes + (l.rowpop ? l.tile_rows * 2 : 0);
    l.tile_stride = l.tile_bytes; // multiple of 16
    l.unit_batches = l.dev_words >= 32 ? 1 : 32 / l.dev_words;
    l.unit_bytes = l.unit_batches * l.tile_stride;
    l.stage_bytes = (l.unit_bytes + 127) / 128 * 128;
    l.batch_stage_bytes = (l.tile_bytes + 127) / 128 * 128;
    *out = l;
    return GSB_OK;
}

// Bytes of a shard's batches, padded to whole work units (the padding reads as empty rows).
size_t padded_tile_bytes(const Layout& l, uint32_t n_tiles)
{
    const uint32_t units = std::max<uint32_t>(1, (n_tiles + l.unit_batches - 1) / l.unit_batches);
    return static_cast<size_t>(units) * l.unit_bytes;
}

// Select state of one launch of the single-query kernel.  A shard owns TWO sets and alternates
// between them: with programmatic dependent launch the scan of query i+1 starts while the last
// CTA of query i is still sorting / exchanging / merging out of the other set.
struct SelectSet {
    gsb::ScanCtrl* ctrl = nullptr;
    unsigned int* ghist = nullptr;
    unsigned int* ehist = nullptr; // grid-wide histogram of every appended candidate (threshold sharing)
    unsigned long long* gfinal = nullptr;
    unsigned long long* cta_keys = nullptr;
    uint32_t* cta_counts = nullptr;
    uint64_t cta_keys_cap = 0; // entries
    unsigned long long* tail_lists = nullptr; // [grid][cap]: select without a grid barrier (GSB_TAIL)
    uint32_t* tail_counts = nullptr;          // [4096]
    uint64_t tail_cap = 0;                    // entries
};

// Result record of one host-buffer query in MAPPED pinned host memory: the last CTA of the launch
// stores [k keys][survivors][n][done] straight into it (no cudaMemcpy, no stream synchronize: the
// host polls the done word).  kSlots - 1 records serve gsb_db_search_async tickets, the last one
// the synchronous calls.
constexpr int kAsyncDepth = 4;
constexpr int kSlots = kAsyncDepth + 1;
constexpr int kSyncSlot = kAsyncDepth;
struct ResultSlot {
    unsigned long long* host = nullptr;
    unsigned long long* dev = nullptr; // the same memory as the device sees it
    uint32_t cap = 0;                  // keys
};

struct Workspace {
    cudaStream_t stream = nullptr;
    SelectSet sets[2];
    uint64_t scan_launches = 0;          // picks the set of the next launch
    cudaStream_t last_stream = nullptr;  // stream of the latest launch that used this workspace
    bool launched = false;
    cudaEvent_t order_event = nullptr;   // orders launches that arrive on different streams
    ResultSlot slots[kSlots];
    int max_grid = 0;
    // multi-query kernel scratch (allocated on first use)
    gsb::BatchCtrl* bctrl = nullptr;
    unsigned long long* bcand = nullptr;   // [grid][bnq_cap][kBatchListCap]
    unsigned long long* bqlists = nullptr; // [grid][bnq_cap][bk_cap]
    uint32_t* bqcounts = nullptr;
    unsigned long long* bsurv = nullptr;
    uint32_t* bqueries = nullptr;          // [bnq_cap][dev_words]
    unsigned long long* bout = nullptr;    // [bnq_cap][bk_cap + 2] results (+ survivors, n)
    unsigned long long* bout_host = nullptr;
    uint32_t bk_cap = 0, bnq_cap = 0;
    int bgrid = 0;
    // bit-sliced multi-query kernel: query lists, shared thresholds and score histograms
    uint16_t* slists = nullptr;            // [kMaxSlicedQueries][1024] set-bit entries
    uint32_t* slofs = nullptr;
    uint16_t* sngrp = nullptr;
    uint16_t* spopq = nullptr;
    gsb::SlicedMeta* smeta = nullptr;
    unsigned long long* stau = nullptr;    // [kMaxSlicedQueries] thresholds shared by the CTAs of a pass
    unsigned int* shist = nullptr;         // [kMaxSlicedQueries][kSlicedHistBuckets] candidate score histograms
};

struct Shard {
    int device = 0;
    uint64_t row_base = 0;
    uint64_t n_rows = 0;
    uint32_t n_tiles = 0;
    uint8_t* tiles = nullptr;
    size_t bytes = 0;
    Workspace ws;
    uint64_t epoch = 0; // g_reset_epoch when the handles were made: a device reset leaves them dangling
};

struct Plan {
    int grid = 0;
    int warps = 0;
    uint32_t stages = 0, cap = 0, smem = 0;
};

} // namespace

struct gsb_db {
    int fp_bits = 0;
    uint32_t words = 0; // unfolded words per row
    uint64_t count = 0;
    // unfolded rows on the host, chunk by chunk as they arrived (empty for device-generated
    // shards): used by the upload, getFingerprint, search_cpu and the fold re-score
    struct HostChunk {
        GsbHostBuf bytes; // pinned + mapped when a CUDA device is present (gsb_internal.h)
        uint64_t row0 = 0, n_rows = 0;
    };
    std::vector<HostChunk> host;
    const uint32_t* host_row(uint64_t row) const
    {
        size_t lo = 0, hi = host.size();
        while (hi - lo > 1) {
            const size_t mid = (lo + hi) / 2;
            if (host[mid].row0 <= row)
                lo = mid;
            else
                hi = mid;
        }
        return reinterpret_cast<const uint32_t*>(host[lo].bytes.data()) + (row - host[lo].row0) * words;
    }
    unsigned fold_factor = 1;
    Layout layout;
    std::vector<Shard> shards;
    bool uploaded = false;
    uint64_t synth_seed = 0;
    // similarity metric (gsb_db_set_metric); Tanimoto unless told otherwise
    uint32_t metric = gsb::kMetricTanimoto;
    float alpha = 1.0f, beta = 1.0f;
    // folded databases: the unfolded host rows registered with CUDA (mapped) for the device re-score
    bool host_registered = false;
    gsb::RescoreChunk* d_chunks = nullptr; // chunk table on the re-score device (shard 0's)
    unsigned long long* rescore_host = nullptr, *rescore_dev = nullptr; // mapped pinned: candidates in, keys out
    size_t rescore_cap = 0;
    uint64_t rescore_epoch = 0; // g_reset_epoch when the above were made
    // queries in flight through the host-buffer API (gsb_db_search_async / _wait)
    struct Pending {
        uint64_t ticket = 0; // 0 = free
        uint64_t seq = 0;    // completion word value of the launches
        uint32_t k = 0;
        float cutoff = 0.0f;
        bool deferred = false; // folded / very large k: computed inside gsb_db_search_wait
        std::vector<int32_t> query;
    };
    mutable Pending pending[kAsyncDepth];
    mutable uint64_t next_ticket = 1;
    mutable uint64_t done_seq = 0;
    mutable std::mutex mu; // enqueue side: launch planning and workspace growth are not re-entrant
};

namespace
{

int smem_limit(int device, int* out)
{
    int v = 0;
    GSB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    *out = v;
    return GSB_OK;
}

// ---- ordering of launches that share a device --------------------------------------------------
// Every scan kernel here lines its CTAs up on a grid-wide arrival counter, so all of them must be
// resident together.  On ONE stream that holds by construction (a launch owns every SM it needs
// before the next one starts; with programmatic dependent launch the successor only takes SMs the
// predecessor has left).  Launches on DIFFERENT streams of a device could interleave their CTAs
// and starve each other, and launches that share a shard's workspace must not overlap at all.
// Both are ruled out lazily: when a launch arrives on another stream than the work still in flight,
// an event recorded on that stream's tail is waited for first.  No event is recorded on the steady
// one-stream path (an event between two kernels would also switch off their overlap).
struct DeviceGate {
    std::mutex mu;
    struct Entry {
        cudaStream_t stream;
        int grid;
    };
    std::vector<Entry> inflight;
    cudaEvent_t event = nullptr;
};
DeviceGate g_gates[64];

int order_after(cudaStream_t later, cudaStream_t earlier, cudaEvent_t* ev)
{
    const cudaError_t q = cudaStreamQuery(earlier);
    if (q != cudaErrorNotReady) { // idle (or a stream its owner has destroyed): nothing to wait for
        cudaGetLastError();
        return GSB_OK;
    }
    cudaGetLastError();
    if (!*ev)
        GSB_CUDA(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    GSB_CUDA(cudaEventRecord(*ev, earlier));
    GSB_CUDA(cudaStreamWaitEvent(later, *ev, 0));
    return GSB_OK;
}

// Called (under the database mutex) right before a grid-barrier kernel of `grid` CTAs is launched
// on `st` for workspace `ws` of device `dev`.
int gate_launch(int dev, Workspace& ws, cudaStream_t st, int grid)
{
    if (ws.launched && ws.last_stream != st) { // same workspace, other stream: never overlap
        int rc = order_after(st, ws.last_stream, &ws.order_event);
        if (rc)
            return rc;
    }
    ws.launched = true;
    ws.last_stream = st;
    if (dev < 0 || dev >= 64)
        return GSB_OK;
    DeviceGate& g = g_gates[dev];
    std::lock_guard<std::mutex> lock(g.mu);
    int others = 0;
    for (const auto& e : g.inflight)
        if (e.stream != st)
            others += e.grid;
    if (others + grid > ws.max_grid) { // cannot all be resident together: run after them
        for (const auto& e : g.inflight)
            if (e.stream != st) {
                int rc = order_after(st, e.stream, &g.event);
                if (rc)
                    return rc;
            }
        g.inflight.clear();
    }
    for (auto& e : g.inflight)
        if (e.stream == st) {
            e.grid = std::max(e.grid, grid);
            return GSB_OK;
        }
    g.inflight.push_back({st, grid});
    return GSB_OK;
}

void gate_forget(int dev, cudaStream_t st)
{
    if (dev < 0 || dev >= 64)
        return;
    DeviceGate& g = g_gates[dev];
    std::lock_guard<std::mutex> lock(g.mu);
    for (size_t i = 0; i < g.inflight.size(); i++)
        if (g.inflight[i].stream == st) {
            g.inflight.erase(g.inflight.begin() + i);
            break;
        }
}

template <int W, bool RP, int CW> int launch_scan_t(const gsb::ScanParams& p, const Plan& plan, cudaStream_t st)
{
    static thread_local int configured[64] = {0};
    static thread_local uint64_t configured_epoch = 0;
    if (configured_epoch != g_reset_epoch.load()) { // a device reset forgets function attributes
        std::memset(configured, 0, sizeof(configured));
        configured_epoch = g_reset_epoch.load();
    }
    int dev = 0;
    GSB_CUDA(cudaGetDevice(&dev));
    auto kernel = gsb::scan_topk_kernel<W, RP, CW>;
    if (dev < 64 && configured[dev] < static_cast<int>(plan.smem)) {
        GSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem));
        configured[dev] = plan.smem;
    }
    if (env_int("GSB_PDL", 1)) {
        // Programmatic dependent launch: the scan of this query may start on SMs the previous query
        // (same stream, other control set) has already left.  Residency of the whole grid is what a
        // cooperative launch would check: grid <= SMs and one CTA fits an SM (make_plan), and
        // gate_launch keeps other streams' grids off the device meanwhile.
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(plan.grid);
        cfg.blockDim = dim3(CW * 32);
        cfg.dynamicSmemBytes = plan.smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        GSB_CUDA(cudaLaunchKernelEx(&cfg, kernel, p));
    } else {
        // cooperative launch: the runtime itself guarantees that every CTA is resident
        void* args[] = {const_cast<gsb::ScanParams*>(&p)};
        GSB_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kernel), dim3(plan.grid), dim3(CW * 32), args,
                                             plan.smem, st));
    }
    g_launches++;
    return GSB_OK;
}

template <int W, bool RP> int launch_scan_w(const gsb::ScanParams& p, const Plan& plan, cudaStream_t st)
{
#define GSB_WCASE(CW)                                                                            \
    case CW:                                                                                     \
        return launch_scan_t<W, RP, CW>(p, plan, st);
    switch (plan.warps) {
        GSB_WCASE(4)
        GSB_WCASE(8)
        GSB_WCASE(12)
        GSB_WCASE(16)
#undef GSB_WCASE
    default:
        return fail(GSB_ERR_INVALID, "GSB_WARPS must be 4, 8, 12 or 16");
    }
}

int launch_scan(const Layout& l, const gsb::ScanParams& p, const Plan& plan, cudaStream_t st)
{
#define GSB_CASE(W)                                                                              \
    case W:                                                                                      \
        return l.rowpop ? launch_scan_w<W, true>(p, plan, st) : launch_scan_w<W, false>(p, plan, st);
    switch (l.dev_words) {
        GSB_CASE(4)
        GSB_CASE(8)
        GSB_CASE(16)
        GSB_CASE(32)
        GSB_CASE(64)
        GSB_CASE(128)
    default:
        return fail(GSB_ERR_INVALID, "unsupported device row width");
    }
#undef GSB_CASE
}

// Warps per CTA, ring depth per warp and candidate capacity for this k.  Rules the kernel
// relies on: cap is a power of two; cap >= k + 4*threads (the select is requested 2*threads below
// the top, and there must be room to refill after it); cap >= k + grid (merge rounds).
int make_plan(const Layout& l, const Shard& sh, uint32_t k, Plan* out)
{
    int smem_max = 0;
    int rc = smem_limit(sh.device, &smem_max);
    if (rc)
        return rc;
    const int budget = smem_max - 2048; // static shared memory + alignment slack
    const int pref = env_int("GSB_WARPS", 16);
    if (pref != 4 && pref != 8 && pref != 12 && pref != 16)
        return fail(GSB_ERR_INVALID, "GSB_WARPS must be 4, 8, 12 or 16");
    // preferred shape first, then fewer stages, then fewer warps (wide rows, large k)
    for (int warps = pref; warps >= 4; warps -= 4) {
        // ring depth: enough copies in flight per SM (~64 KB) to cover HBM latency — narrow (folded)
        // rows make small batches and need a deeper ring; 1024-bit rows need two stages
        const int auto_stages = 1 + static_cast<int>((65536 + warps * l.stage_bytes - 1) / (warps * l.stage_bytes));
        const int want = std::min(gsb::kMaxStages, std::max(2, env_int("GSB_STAGES", auto_stages)));
        int grid = sh.ws.max_grid;
        if (const int g = env_int("GSB_GRID", 0))
            grid = std::min(g, grid); // never more CTAs than can be resident together
        const uint32_t n_units = (sh.n_tiles + l.unit_batches - 1) / l.unit_batches;
        const uint32_t n_super = (n_units + warps - 1) / warps;
        grid = std::max(1, std::min<int>(grid, n_super ? n_super : 1));
        const uint32_t threads = warps * 32;
        const uint64_t need = std::max<uint64_t>(static_cast<uint64_t>(k) + 4ull * threads, k + (uint64_t) grid);
        const uint32_t cap = std::max<uint32_t>(pow2ceil(need), static_cast<uint32_t>(env_int("GSB_MIN_CAP", 4096)));
        for (int st = want; st >= 2; st--) {
            const uint64_t smem = static_cast<uint64_t>(warps) * st * l.stage_bytes +
                                  static_cast<uint64_t>(cap) * 8 + gsb::kBuckets * 4;
            if (smem <= static_cast<uint64_t>(budget)) {
                out->grid = grid;
                out->warps = warps;
                out->stages = st;
                out->cap = cap;
                out->smem = static_cast<uint32_t>(smem);
                return GSB_OK;
            }
        }
    }
    return fail(GSB_ERR_INVALID, "max_return_count " + std::to_string(k) +
                                     " is too large for the fused in-kernel select");
}

int set_init(SelectSet& set)
{
    GSB_CUDA(cudaMalloc(&set.ctrl, sizeof(gsb::ScanCtrl)));
    GSB_CUDA(cudaMemset(set.ctrl, 0, sizeof(gsb::ScanCtrl)));
    GSB_CUDA(cudaMalloc(&set.ghist, gsb::kBuckets * 4));
    GSB_CUDA(cudaMemset(set.ghist, 0, gsb::kBuckets * 4));
    GSB_CUDA(cudaMalloc(&set.ehist, gsb::kBuckets * 4));
    GSB_CUDA(cudaMemset(set.ehist, 0, gsb::kBuckets * 4));
    GSB_CUDA(cudaMalloc(&set.gfinal, 32768 * 8));
    GSB_CUDA(cudaMalloc(&set.cta_counts, sizeof(uint32_t) * 4096));
    return GSB_OK;
}

int ws_init(Shard& sh)
{
    Workspace& ws = sh.ws;
    sh.epoch = g_reset_epoch.load();
    GSB_CUDA(cudaSetDevice(sh.device));
    GSB_CUDA(cudaStreamCreateWithFlags(&ws.stream, cudaStreamNonBlocking));
    for (SelectSet& set : ws.sets) {
        int rc = set_init(set);
        if (rc)
            return rc;
    }
    int sms = 0;
    GSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, sh.device));
    ws.max_grid = sms; // persistent: one CTA per SM (shared memory allows no more), all resident
    return GSB_OK;
}

// After a launch reported an error word (grid barrier / peer flag timeout) its control sets are
// suspect: drain the device and zero them, so that the next search starts from a clean state.
int ws_recover(Shard& sh)
{
    GSB_CUDA(cudaSetDevice(sh.device));
    GSB_CUDA(cudaDeviceSynchronize());
    for (SelectSet& set : sh.ws.sets) {
        GSB_CUDA(cudaMemset(set.ctrl, 0, sizeof(gsb::ScanCtrl)));
        GSB_CUDA(cudaMemset(set.ghist, 0, gsb::kBuckets * 4));
        GSB_CUDA(cudaMemset(set.ehist, 0, gsb::kBuckets * 4));
    }
    if (sh.ws.bctrl)
        GSB_CUDA(cudaMemset(sh.ws.bctrl, 0, sizeof(gsb::BatchCtrl)));
    if (sh.ws.bsurv)
        GSB_CUDA(cudaMemset(sh.ws.bsurv, 0, static_cast<size_t>(sh.ws.bnq_cap) * 8));
    GSB_CUDA(cudaDeviceSynchronize());
    return GSB_OK;
}

// GSB_TAIL=1 (opt-in): the single-query kernel's grid-wide select without a grid barrier — every CTA
// leaves its candidates in global memory and goes, the one with the last ticket selects
// (gsb_kernels.cuh, tail_select).  Needs no co-resident grid, but measures slower than the default
// form (all CTAs line up on an arrival counter, each adds its share of the final list): the barrier
// costs a CTA 7 us, the lone last CTA needs 15 - 80 us for what 148 CTAs do in 10
// (profiles/r02_fixed_cost.md).
bool tail_mode()
{
    return env_int("GSB_TAIL", 0) != 0;
}

// slot < 0: results go to caller-provided (device) buffers, no record needed.
int ws_reserve(Shard& sh, uint32_t k, int grid, int slot, uint32_t cap)
{
    Workspace& ws = sh.ws;
    GSB_CUDA(cudaSetDevice(sh.device));
    const uint64_t need = static_cast<uint64_t>(grid) * std::max<uint32_t>(k, 1);
    const uint64_t need_tail = tail_mode() ? static_cast<uint64_t>(grid) * cap : 0;
    for (SelectSet& set : ws.sets) {
        if (need_tail > set.tail_cap) {
            if (set.tail_lists) {
                GSB_CUDA(cudaStreamSynchronize(ws.stream)); // a launch may still be using the old one
                GSB_CUDA(cudaFree(set.tail_lists));
            }
            set.tail_lists = nullptr;
            set.tail_cap = 0;
            GSB_CUDA(cudaMalloc(&set.tail_lists, need_tail * 8));
            set.tail_cap = need_tail;
            if (!set.tail_counts)
                GSB_CUDA(cudaMalloc(&set.tail_counts, sizeof(uint32_t) * 4096));
        }
        if (need > set.cta_keys_cap) {
            if (set.cta_keys) {
                GSB_CUDA(cudaStreamSynchronize(ws.stream)); // a launch may still be using the old one
                GSB_CUDA(cudaFree(set.cta_keys));
            }
            set.cta_keys = nullptr;
            set.cta_keys_cap = 0;
            GSB_CUDA(cudaMalloc(&set.cta_keys, need * 8));
            set.cta_keys_cap = need;
        }
    }
    if (slot >= 0 && k > ws.slots[slot].cap) {
        ResultSlot& rs = ws.slots[slot];
        if (rs.host) {
            GSB_CUDA(cudaStreamSynchronize(ws.stream));
            GSB_CUDA(cudaFreeHost(rs.host));
        }
        rs = ResultSlot();
        const uint32_t cap = pow2ceil(std::max<uint32_t>(k, 1024));
        void* h = nullptr;
        GSB_CUDA(cudaHostAlloc(&h, (static_cast<size_t>(cap) + 4) * 8, cudaHostAllocMapped | cudaHostAllocPortable));
        std::memset(h, 0, (static_cast<size_t>(cap) + 4) * 8);
        void* d = nullptr;
        GSB_CUDA(cudaHostGetDevicePointer(&d, h, 0));
        rs.host = static_cast<unsigned long long*>(h);
        rs.dev = static_cast<unsigned long long*>(d);
        rs.cap = cap;
    }
    return GSB_OK;
}

void ws_free(Shard& sh)
{
    if (sh.epoch != g_reset_epoch.load()) { // cudaDeviceReset took every stream, event and allocation
        sh = Shard();                       // with it; the stale handles must not reach the runtime again
        return;
    }
    cudaSetDevice(sh.device);
    Workspace& ws = sh.ws;
    if (ws.stream) {
        cudaStreamSynchronize(ws.stream);
        gate_forget(sh.device, ws.stream);
        cudaStreamDestroy(ws.stream);
    }
    if (ws.order_event)
        cudaEventDestroy(ws.order_event);
    for (SelectSet& set : ws.sets) {
        cudaFree(set.ctrl);
        cudaFree(set.ghist);
        cudaFree(set.ehist);
        cudaFree(set.gfinal);
        cudaFree(set.cta_keys);
        cudaFree(set.cta_counts);
        cudaFree(set.tail_lists);
        cudaFree(set.tail_counts);
    }
    for (ResultSlot& rs : ws.slots)
        if (rs.host)
            cudaFreeHost(rs.host);
    cudaFree(ws.bctrl);
    cudaFree(ws.bcand);
    cudaFree(ws.bqlists);
    cudaFree(ws.bqcounts);
    cudaFree(ws.bsurv);
    cudaFree(ws.bqueries);
    cudaFree(ws.bout);
    cudaFree(ws.slists);
    cudaFree(ws.slofs);
    cudaFree(ws.sngrp);
    cudaFree(ws.spopq);
    cudaFree(ws.stau);
    cudaFree(ws.shist);
    cudaFree(ws.smeta);
    if (ws.bout_host)
        cudaFreeHost(ws.bout_host);
    cudaFree(sh.tiles);
    sh = Shard();
}

// FoldFingerprintFunctorCPU (reference calculation_functors.cpp:22-41): bit pos -> pos % new
// size with the in-word position kept, i.e. OR of `factor` contiguous word segments.
void fold_row(const uint32_t* in, uint32_t words, uint32_t factor, uint32_t* out)
{
    const uint32_t nw = words / factor;
    for (uint32_t w = 0; w < nw; w++)
        out[w] = 0;
    for (uint32_t w = 0; w < words; w++)
        out[w % nw] |= in[w];
}

void parallel_for(uint64_t n, const std::function<void(uint64_t, uint64_t)>& fn)
{
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    if (const int e = env_int("GSB_HOST_THREADS", 0))
        nt = e;
    if (n < 65536 || nt == 1) {
        fn(0, n);
        return;
    }
    std::vector<std::thread> pool;
    const uint64_t per = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++) {
        const uint64_t lo = std::min(n, t * per), hi = std::min(n, lo + per);
        if (lo >= hi)
            break;
        pool.emplace_back([&fn, lo, hi]() { fn(lo, hi); });
    }
    for (auto& th : pool)
        th.join();
}

// Copy rows [row0, row0+n) of the (possibly folded/padded) host matrix into a shard's tiles.
int upload_rows(const gsb_db* db, Shard& sh)
{
    const Layout& l = db->layout;
    GSB_CUDA(cudaSetDevice(sh.device));
    sh.n_tiles = static_cast<uint32_t>((sh.n_rows + l.tile_rows - 1) / l.tile_rows);
    sh.bytes = padded_tile_bytes(l, sh.n_tiles);
    if (gsb_device_free_bytes(sh.device) <= sh.bytes)
        return fail(GSB_ERR_NOMEM, "Can't find a GPU with enough memory to copy data.");
    GSB_CUDA(cudaMalloc(&sh.tiles, sh.bytes));
    GSB_CUDA(cudaMemsetAsync(sh.tiles, 0, sh.bytes, sh.ws.stream));

    // Pipelined ingest: pieces of the host rows go through two pinned staging buffers (host
    // threads copy into one while the other is on the wire), land in device staging, and one kernel
    // folds (reference FoldFingerprintFunctorCPU, calculation_functors.cpp:22-41 — "port folding to
    // the GPU" is on the reference's own to-do list), zero-pads to the device row width, writes the
    // 32-row batches and their popcount trailers.  The reference copies from pageable memory
    // (.cu:181-182) after folding on the CPU (fingerprintdb_cuda.cpp:56-69).
    const uint32_t src_words = db->words;
    const uint32_t f = db->fold_factor;
    const uint64_t src_row_bytes = static_cast<uint64_t>(src_words) * 4;
    const uint64_t piece_rows = std::max<uint64_t>(32, ((64ull << 20) / src_row_bytes) / 32 * 32);
    struct Staging { // released on every exit path
        uint8_t* pinned[2] = {nullptr, nullptr};
        uint8_t* staged[2] = {nullptr, nullptr};
        cudaEvent_t done[2] = {nullptr, nullptr};
        ~Staging()
        {
            for (int i = 0; i < 2; i++) {
                if (pinned[i])
                    cudaFreeHost(pinned[i]);
                cudaFree(staged[i]);
                if (done[i])
                    cudaEventDestroy(done[i]);
            }
        }
    } stg;
    uint8_t** pinned = stg.pinned;
    uint8_t** staged = stg.staged;
    cudaEvent_t* done = stg.done;
    // Host chunks are pinned from birth when a device is present (GsbHostBuf): the copy engine reads
    // them in place.  Only plain-memory chunks (pinning failed) go through pinned staging buffers.
    bool need_host_staging = false;
    for (const gsb_db::HostChunk& hc : db->host)
        need_host_staging |= !hc.bytes.pinned();
    for (int i = 0; i < 2; i++) {
        if (need_host_staging)
            GSB_CUDA(cudaMallocHost(&pinned[i], piece_rows * src_row_bytes));
        GSB_CUDA(cudaMalloc(&staged[i], piece_rows * src_row_bytes));
        GSB_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    }
    int slot = 0;
    for (const gsb_db::HostChunk& hc : db->host) {
        // rows of this host chunk that belong to the shard
        const uint64_t lo_row = std::max<uint64_t>(hc.row0, sh.row_base);
        const uint64_t hi_row = std::min<uint64_t>(hc.row0 + hc.n_rows, sh.row_base + sh.n_rows);
        for (uint64_t g0 = lo_row; g0 < hi_row; g0 += piece_rows, slot ^= 1) {
            const uint64_t rows = std::min<uint64_t>(piece_rows, hi_row - g0);
            GSB_CUDA(cudaEventSynchronize(done[slot])); // the previous use of this slot has been consumed
            const uint8_t* src = hc.bytes.data() + (g0 - hc.row0) * src_row_bytes;
            if (!hc.bytes.pinned()) {
                uint8_t* dst = pinned[slot];
                parallel_for(rows, [=](uint64_t lo, uint64_t hi) {
                    std::memcpy(dst + lo * src_row_bytes, src + lo * src_row_bytes, (hi - lo) * src_row_bytes);
                });
                src = dst;
            }
            GSB_CUDA(cudaMemcpyAsync(staged[slot], src, rows * src_row_bytes, cudaMemcpyHostToDevice, sh.ws.stream));
            const uint64_t threads = rows * l.dev_words;
            gsb::ingest_rows_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, sh.ws.stream>>>(
                reinterpret_cast<const uint32_t*>(staged[slot]), rows, g0 - sh.row_base, src_words, f, sh.tiles,
                l.tile_stride, l.dev_words, l.rowpop ? 1 : 0);
            g_launches++;
            GSB_CUDA(cudaEventRecord(done[slot], sh.ws.stream));
        }
    }
    GSB_CUDA(cudaStreamSynchronize(sh.ws.stream));
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

uint64_t spin_timeout_ns()
{
    // wall-clock bound of the in-kernel spins (grid barrier, peer flags); GSB_SPIN_TIMEOUT_MS for tests
    return static_cast<uint64_t>(std::max(1, env_int("GSB_SPIN_TIMEOUT_MS", 10000))) * 1000000ull;
}

// Parameters of the next single-query launch over this shard; takes the next control set.
void fill_params(const gsb_db* db, Shard& sh, const Plan& plan, uint32_t k, float cutoff, gsb::ScanParams* p)
{
    const Layout& l = db->layout;
    const SelectSet& set = sh.ws.sets[sh.ws.scan_launches++ & 1];
    std::memset(p, 0, sizeof(*p));
    p->tiles = sh.tiles;
    p->n_rows = sh.n_rows;
    p->row_base = sh.row_base;
    p->n_units = (sh.n_tiles + l.unit_batches - 1) / l.unit_batches;
    p->batch_stride = l.tile_stride;
    p->unit_bytes = l.unit_bytes;
    p->stage_bytes = l.stage_bytes;
    p->stages = plan.stages;
    p->cap = plan.cap;
    p->k = k;
    p->cutoff = cutoff;
    p->key_ceiling = ~0ull;
    p->cta_keys = set.cta_keys;
    p->cta_counts = set.cta_counts;
    p->ctrl = set.ctrl;
    p->ghist = set.ghist;
    p->ehist = env_int("GSB_SHARE_HIST", 0) ? set.ehist : nullptr; // opt-in: fewer selects, same throughput (profiles/r02_fixed_cost.md)
    p->gfinal = set.gfinal;
    if (tail_mode() && set.tail_lists && static_cast<uint64_t>(plan.grid) * plan.cap <= set.tail_cap) {
        p->tail_lists = set.tail_lists;
        p->tail_counts = set.tail_counts;
    }
    p->spin_timeout_ns = spin_timeout_ns();
    p->metric = db->metric;
    p->alpha = db->alpha;
    p->beta = db->beta;
}

struct Cand {
    float score;
    uint32_t row;
};

inline Cand decode(unsigned long long key)
{
    Cand c;
    const uint32_t bits = static_cast<uint32_t>(key >> 32);
    std::memcpy(&c.score, &bits, 4);
    c.row = 0xffffffffu - static_cast<uint32_t>(key & 0xffffffffu);
    return c;
}

void print_debug_times(const std::vector<unsigned long long>& h, int grid)
{
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < grid; c++)
        t0 = std::min(t0, h[c * 8]);
    const char* names[7] = {"start", "first data (warp 0)", "warp0 out of work", "cta out of work", "global histogram read",
                            "ticket", "final sort done"};
    for (int st = 0; st < 7; st++) {
        unsigned long long lo = ~0ull, hi = 0;
        double sum = 0;
        int n = 0;
        for (int c = 0; c < grid; c++) {
            const unsigned long long v = h[c * 8 + st];
            if (!v)
                continue;
            lo = std::min(lo, v - t0), hi = std::max(hi, v - t0), sum += double(v - t0), n++;
        }
        if (n)
            std::fprintf(stderr, "[gsb dbg] %-18s n=%3d min %8.1f us avg %8.1f us max %8.1f us\n", names[st], n, lo / 1e3,
                         sum / n / 1e3, hi / 1e3);
    }
    unsigned long long emax = 0, smax = 0;
    double ssum = 0;
    for (int c = 0; c < grid; c++) {
        emax = std::max(emax, h[c * 8 + 7] & 0xffull);
        smax = std::max(smax, h[c * 8 + 7] >> 8);
        ssum += double(h[c * 8 + 7] >> 8);
    }
    std::fprintf(stderr, "[gsb dbg] selects per CTA during the scan: max %llu; time inside them (warp 0): avg %.1f us max %.1f us\n",
                 emax, ssum / grid / 1e3, smax / 1e3);
}

// One launch per shard with a host-resident query (it travels as a kernel parameter); the last CTA
// of every launch stores the shard's record into result slot `slot` (mapped pinned memory) and then
// the completion value `seq`.  Nothing here waits for the device.
int enqueue_scan(const gsb_db* db, const uint32_t* q_dev_words, uint32_t k, float cutoff,
                 unsigned long long key_ceiling, int slot, uint64_t seq)
{
    for (size_t i = 0; i < db->shards.size(); i++) {
        Shard& sh = const_cast<Shard&>(db->shards[i]);
        if (sh.n_rows == 0)
            continue;
        Plan plan;
        int rc = make_plan(db->layout, sh, k, &plan);
        if (rc)
            return rc;
        rc = ws_reserve(sh, k, plan.grid, slot, plan.cap);
        if (rc)
            return rc;
        GSB_CUDA(cudaSetDevice(sh.device));
        gsb::ScanParams p;
        fill_params(db, sh, plan, k, cutoff, &p);
        std::memcpy(p.q_host, q_dev_words, db->layout.dev_words * 4);
        p.key_ceiling = key_ceiling;
        unsigned long long* rec = sh.ws.slots[slot].dev;
        p.out_keys = rec;
        p.out_survivors = rec + k;
        p.out_n = reinterpret_cast<uint32_t*>(rec + k + 1);
        p.out_done = rec + k + 2;
        p.out_done_value = seq;
        unsigned long long* dbg = nullptr;
        if (env_int("GSB_DEBUG_TIMES", 0)) {
            GSB_CUDA(cudaMalloc(&dbg, plan.grid * 64));
            GSB_CUDA(cudaMemsetAsync(dbg, 0, plan.grid * 64, sh.ws.stream));
            p.dbg = dbg;
        }
        rc = gate_launch(sh.device, sh.ws, sh.ws.stream, plan.grid);
        if (rc)
            return rc;
        rc = launch_scan(db->layout, p, plan, sh.ws.stream);
        if (rc)
            return rc;
        if (dbg) { // developer aid: phase timeline of the launch (ns since the earliest CTA start)
            std::vector<unsigned long long> h(plan.grid * 8);
            GSB_CUDA(cudaMemcpyAsync(h.data(), dbg, h.size() * 8, cudaMemcpyDeviceToHost, sh.ws.stream));
            GSB_CUDA(cudaStreamSynchronize(sh.ws.stream));
            cudaFree(dbg);
            print_debug_times(h, plan.grid);
        }
    }
    return GSB_OK;
}

// Wait until `word` (mapped pinned memory, written by the device) holds `value`.  Spins first
// (GSB_POLL_SPIN_US, default 20 ms: a query over a full HBM takes about that long), then sleeps in
// short steps; the stream is probed now and then so that a failed launch ends the wait.
int wait_word(const unsigned long long* word, unsigned long long value, cudaStream_t stream)
{
    const volatile unsigned long long* w = word;
    const auto t0 = std::chrono::steady_clock::now();
    const long long spin_us = env_int("GSB_POLL_SPIN_US", 20000);
    const long long limit_s = std::max(1, env_int("GSB_WAIT_TIMEOUT_S", 120));
    for (uint64_t spins = 1; *w != value; spins++) {
        GSB_CPU_RELAX();
        if ((spins & 0xfff) != 0)
            continue;
        const auto us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        if (stream) {
            const cudaError_t e = cudaStreamQuery(stream);
            if (e == cudaSuccess) { // everything queued has run: the word is there, or never will be
                if (*w == value)
                    break;
                return fail(GSB_ERR_CUDA, "the search kernel finished without publishing its result");
            }
            if (e != cudaErrorNotReady)
                return fail(GSB_ERR_CUDA, std::string("search kernel failed: ") + cudaGetErrorString(e));
        }
        if (us > limit_s * 1000000ll)
            return fail(GSB_ERR_CUDA, "timed out waiting for the search kernel");
        if (us > spin_us)
            std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return GSB_OK;
}

const char* device_error_text(uint32_t)
{
    return "the search kernel reported a failure (grid barrier or peer flag timed out, or a candidate buffer "
           "overflowed); the workspace was reset";
}

// Results of enqueue_scan(slot, seq): waits for every shard's record and merges them.
int collect_scan(const gsb_db* db, uint32_t k, int slot, uint64_t seq, std::vector<unsigned long long>* keys,
                 uint64_t* survivors)
{
    keys->clear();
    *survivors = 0;
    bool failed = false;
    for (size_t i = 0; i < db->shards.size(); i++) {
        Shard& sh = const_cast<Shard&>(db->shards[i]);
        if (sh.n_rows == 0)
            continue;
        const unsigned long long* rec = sh.ws.slots[slot].host;
        int rc = wait_word(rec + k + 2, seq, sh.ws.stream);
        if (rc)
            return rc;
        const uint32_t n = *reinterpret_cast<const uint32_t*>(rec + k + 1);
        if (n == gsb::kCountError) {
            failed = true;
            ws_recover(sh);
            continue;
        }
        *survivors += rec[k];
        if (db->shards.size() == 1) {
            keys->assign(rec, rec + n);
        } else {
            // merge of per-shard sorted lists on the host (reference .cu:366 sorts on the host too)
            const size_t old = keys->size();
            keys->insert(keys->end(), rec, rec + n);
            std::inplace_merge(keys->begin(), keys->begin() + old, keys->end(), std::greater<unsigned long long>());
            if (keys->size() > k)
                keys->resize(k);
        }
    }
    if (failed)
        return fail(GSB_ERR_CUDA, device_error_text(0));
    return GSB_OK;
}

int scan_all_shards(const gsb_db* db, const uint32_t* q_dev_words, uint32_t k, float cutoff,
                    unsigned long long key_ceiling, std::vector<unsigned long long>* keys, uint64_t* survivors)
{
    const uint64_t seq = ++db->done_seq;
    int rc = enqueue_scan(db, q_dev_words, k, cutoff, key_ceiling, kSyncSlot, seq);
    if (rc)
        return rc;
    return collect_scan(db, k, kSyncSlot, seq, keys, survivors);
}

// Largest k one fused launch can select on every shard of this database.
uint32_t max_fused_k(const gsb_db* db)
{
    for (uint32_t cap = 32768; cap >= 2048; cap >>= 1) {
        const uint32_t k = cap - 4u * 2u * 128u - 4096u / 8u;
        bool ok = true;
        for (const Shard& sh : db->shards) {
            Plan plan;
            if (sh.n_rows && make_plan(db->layout, sh, k, &plan) != GSB_OK)
                ok = false;
        }
        if (ok)
            return k;
    }
    return 0;
}

// Top-k_total keys over all shards.  One fused launch per shard when k_total fits the in-kernel
// select; otherwise peeling passes: every pass selects the best `kmax` keys strictly below the
// worst key of the previous pass (keys are unique, so the passes tile the ranking exactly).
int scan_topk(const gsb_db* db, const uint32_t* q_dev_words, uint64_t k_total, float cutoff,
              std::vector<unsigned long long>* keys, uint64_t* survivors)
{
    Plan probe;
    bool fits = k_total <= 0xffffffffull;
    for (const Shard& sh : db->shards)
        if (fits && sh.n_rows && make_plan(db->layout, sh, static_cast<uint32_t>(k_total), &probe) != GSB_OK)
            fits = false;
    if (fits)
        return scan_all_shards(db, q_dev_words, static_cast<uint32_t>(k_total), cutoff, ~0ull, keys, survivors);
    const uint32_t kmax = max_fused_k(db);
    if (kmax == 0)
        return fail(GSB_ERR_INVALID, "no launch shape fits this database");
    keys->clear();
    unsigned long long ceiling = ~0ull;
    std::vector<unsigned long long> pass;
    while (keys->size() < k_total) {
        uint64_t surv = 0;
        const uint32_t want = static_cast<uint32_t>(std::min<uint64_t>(kmax, k_total - keys->size()));
        int rc = scan_all_shards(db, q_dev_words, want, cutoff, ceiling, &pass, &surv);
        if (rc)
            return rc;
        if (ceiling == ~0ull)
            *survivors = surv;
        keys->insert(keys->end(), pass.begin(), pass.end());
        if (pass.size() < want)
            break; // the database is exhausted
        ceiling = pass.back();
    }
    return GSB_OK;
}

// ---- multi-query kernel ------------------------------------------------------------------
// Which multi-query kernel serves a batch.  GSB_BATCH_KERNEL: 0 = none (loop over the single-query
// kernel), 1 = automatic (default), 2 = POPC kernel, 3 = bit-sliced kernel wherever it applies,
// 4 = tensor-core kernel wherever it applies.
enum BatchKernel { kBatchNone = 0, kBatchPopc = 1, kBatchSliced = 2, kBatchTensor = 3 };

// Cost model of the automatic choice between the bit-sliced and the tensor-core kernel, in
// milliseconds per 10^9 rows on one B200 (profiles/r02_tensor.md: fits of the dense-query sweep): the
// bit-sliced kernel pays per set bit of the batch, the tensor-core kernel per group of 128 queries;
// they meet at ~140 set bits per query.
constexpr double kSlicedMsBase = 47.0, kSlicedMsPerSetBit = 7.9e-3, kTensorMsPerGroup = 164.0;

// set_bits: set bits of all queries of the batch together (-1 = not known: queries in device memory)
BatchKernel batch_kernel_choice(const gsb_db* db, uint32_t k, int n_queries, float cutoff, int64_t set_bits = -1)
{
    const int mode = env_int("GSB_BATCH_KERNEL", 1);
    if (mode == 0 || !db->layout.rowpop || db->layout.dev_words > 32 || k < 1 || k > gsb::kMaxBatchK || n_queries < 2)
        return kBatchNone;
    if (mode == 2)
        return kBatchPopc;
    const bool tensor_ok = db->metric == gsb::kMetricTanimoto && db->layout.dev_words == 32;
    if (mode == 4 && tensor_ok)
        return kBatchTensor;
    if (mode == 3 && db->metric == gsb::kMetricTanimoto)
        return kBatchSliced;
    if (db->metric != gsb::kMetricTanimoto)
        return kBatchPopc; // the filter bound of the other two kernels is Tanimoto's
    // The bit-sliced kernel pays a transposition per tile: measured cross-over between 4 and 8
    // queries (profiles/r01_sweep.md).  (A low positive cutoff sends many rows through its exact
    // path, but even with every row on that path it stays ahead of the POPC kernel.)
    (void) cutoff;
    if (n_queries < 6)
        return kBatchPopc;
    if (tensor_ok && set_bits >= 0 && mode == 1 && env_int("GSB_TENSOR_AUTO", 1)) {
        const double sliced = kSlicedMsBase + kSlicedMsPerSetBit * static_cast<double>(set_bits);
        const double tensor = kTensorMsPerGroup * ((n_queries + gsb::kTcQueries - 1) / gsb::kTcQueries);
        if (tensor < sliced)
            return kBatchTensor;
    }
    return kBatchSliced;
}

uint32_t batch_max_queries(BatchKernel which)
{
    return which == kBatchPopc ? gsb::kMaxBatchQueries : gsb::kMaxSlicedQueries;
}

struct BatchPlan {
    int grid = 0, warps = 16;
    uint32_t stages = 2, smem = 0;
};

int make_batch_plan(const Layout& l, const Shard& sh, BatchPlan* out)
{
    int smem_max = 0;
    int rc = smem_limit(sh.device, &smem_max);
    if (rc)
        return rc;
    const uint64_t fixed = static_cast<uint64_t>(gsb::kMaxBatchQueries) * l.dev_words * 4 + gsb::kBatchListCap * 8ull +
                           gsb::kBuckets * 4ull + gsb::kMaxBatchQueries * (8ull + 8 + 4 + 4);
    for (int warps = 16; warps >= 8; warps -= 8) {
        const uint64_t smem = static_cast<uint64_t>(warps) * 2 * l.batch_stage_bytes + fixed;
        if (smem + 2048 <= static_cast<uint64_t>(smem_max)) {
            out->warps = warps;
            out->stages = 2;
            out->smem = static_cast<uint32_t>(smem);
            int sms = 0;
            GSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, sh.device));
            const uint32_t n_super = (sh.n_tiles + warps - 1) / warps;
            out->grid = std::max(1, std::min<int>(sms, n_super ? n_super : 1)); // all CTAs resident: grid barrier
            if (const int g = env_int("GSB_GRID", 0))
                out->grid = std::max(1, std::min(out->grid, g));
            return GSB_OK;
        }
    }
    return fail(GSB_ERR_INVALID, "no batch launch shape fits");
}

int batch_reserve(Shard& sh, const Layout& l, uint32_t k, int grid, uint32_t nq)
{
    Workspace& ws = sh.ws;
    GSB_CUDA(cudaSetDevice(sh.device));
    if (!ws.bctrl) {
        GSB_CUDA(cudaMalloc(&ws.bctrl, sizeof(gsb::BatchCtrl)));
        GSB_CUDA(cudaMemset(ws.bctrl, 0, sizeof(gsb::BatchCtrl)));
    }
    if (grid > ws.bgrid || k > ws.bk_cap || nq > ws.bnq_cap) {
        cudaFree(ws.bcand);
        cudaFree(ws.bqlists);
        cudaFree(ws.bqcounts);
        cudaFree(ws.bout);
        cudaFree(ws.bsurv);
        cudaFree(ws.bqueries);
        if (ws.bout_host)
            cudaFreeHost(ws.bout_host);
        ws.bcand = ws.bqlists = ws.bout = ws.bout_host = ws.bsurv = nullptr;
        ws.bqcounts = ws.bqueries = nullptr;
        const int g = std::max(grid, ws.bgrid);
        const uint32_t kc = std::max(k, ws.bk_cap);
        // candidate lists are 16 KB per (CTA, query): grow in steps of 256 queries
        const size_t nqc = std::max<uint32_t>((nq + 255u) / 256u * 256u, ws.bnq_cap);
        ws.bgrid = 0;
        ws.bk_cap = ws.bnq_cap = 0;
        GSB_CUDA(cudaMalloc(&ws.bcand, static_cast<size_t>(g) * nqc * gsb::kBatchListCap * 8));
        GSB_CUDA(cudaMalloc(&ws.bqlists, static_cast<size_t>(g) * nqc * kc * 8));
        GSB_CUDA(cudaMalloc(&ws.bqcounts, static_cast<size_t>(g) * nqc * 4));
        GSB_CUDA(cudaMalloc(&ws.bout, nqc * (kc + 2ull) * 8));
        GSB_CUDA(cudaMallocHost(&ws.bout_host, nqc * (kc + 2ull) * 8));
        GSB_CUDA(cudaMalloc(&ws.bsurv, nqc * 8));
        GSB_CUDA(cudaMemset(ws.bsurv, 0, nqc * 8));
        GSB_CUDA(cudaMalloc(&ws.bqueries, nqc * l.dev_words * 4));
        ws.bgrid = g;
        ws.bk_cap = kc;
        ws.bnq_cap = static_cast<uint32_t>(nqc);
    }
    return GSB_OK;
}

int sliced_reserve(Shard& sh)
{
    Workspace& ws = sh.ws;
    GSB_CUDA(cudaSetDevice(sh.device));
    const size_t nq = gsb::kMaxSlicedQueries;
    // every buffer on its own: a failed allocation leaves the others in place for the next call
    if (!ws.slists)
        GSB_CUDA(cudaMalloc(&ws.slists, nq * 1024 * sizeof(uint16_t)));
    if (!ws.slofs)
        GSB_CUDA(cudaMalloc(&ws.slofs, nq * sizeof(uint32_t)));
    if (!ws.sngrp)
        GSB_CUDA(cudaMalloc(&ws.sngrp, nq * sizeof(uint16_t)));
    if (!ws.spopq)
        GSB_CUDA(cudaMalloc(&ws.spopq, nq * sizeof(uint16_t)));
    if (!ws.stau)
        GSB_CUDA(cudaMalloc(&ws.stau, nq * sizeof(unsigned long long)));
    if (!ws.shist)
        GSB_CUDA(cudaMalloc(&ws.shist, nq * gsb::kSlicedHistBuckets * sizeof(unsigned int)));
    if (!ws.smeta)
        GSB_CUDA(cudaMalloc(&ws.smeta, sizeof(gsb::SlicedMeta)));
    return GSB_OK;
}

template <int W> int launch_batch_t(const gsb::BatchParams& p, const BatchPlan& plan, cudaStream_t st)
{
    auto go = [&](auto kernel) -> int {
        GSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.smem));
        // cooperative launch: the runtime refuses the launch unless every CTA can be resident,
        // which the grid-wide arrival counter in the kernel relies on
        void* args[] = {const_cast<gsb::BatchParams*>(&p)};
        GSB_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kernel), dim3(plan.grid), dim3(plan.warps * 32),
                                             args, plan.smem, st));
        g_launches++;
        return GSB_OK;
    };
    return plan.warps == 16 ? go(gsb::scan_batch_kernel<W, 16>) : go(gsb::scan_batch_kernel<W, 8>);
}

int launch_batch(const Layout& l, const gsb::BatchParams& p, const BatchPlan& plan, cudaStream_t st)
{
    switch (l.dev_words) {
    case 4:
        return launch_batch_t<4>(p, plan, st);
    case 8:
        return launch_batch_t<8>(p, plan, st);
    case 16:
        return launch_batch_t<16>(p, plan, st);
    case 32:
        return launch_batch_t<32>(p, plan, st);
    default:
        return fail(GSB_ERR_INVALID, "the multi-query kernel handles rows up to 1024 bits");
    }
}

void fill_batch_params(const gsb_db* db, const Shard& sh, uint32_t nq, uint32_t k, float cutoff,
                       const uint32_t* d_queries, unsigned long long* out_keys, uint32_t* out_n,
                       unsigned long long* out_surv, gsb::BatchParams* out)
{
    const Layout& l = db->layout;
    gsb::BatchParams& p = *out;
    std::memset(&p, 0, sizeof(p));
    p.tiles = sh.tiles;
    p.n_rows = sh.n_rows;
    p.row_base = sh.row_base;
    p.n_batches = sh.n_tiles;
    p.batch_stride = l.tile_stride;
    p.batch_bytes = l.tile_bytes;
    p.stage_bytes = l.batch_stage_bytes;
    p.k = k;
    p.cutoff = cutoff;
    p.nq = nq;
    p.queries = d_queries;
    p.cand = sh.ws.bcand;
    p.qlists = sh.ws.bqlists;
    p.qcounts = sh.ws.bqcounts;
    p.surv_acc = sh.ws.bsurv;
    p.ctrl = sh.ws.bctrl;
    p.out_keys = out_keys;
    p.out_n = out_n;
    p.out_survivors = out_surv;
    p.spin_timeout_ns = 10ull * spin_timeout_ns(); // CTAs finish far apart in the multi-query kernels
    p.metric = db->metric;
    p.alpha = db->alpha;
    p.beta = db->beta;
}

// Bit-sliced kernel over one shard (gsb_sliced.cuh): query lists, then one pass.  Everything is
// queued on `st`.
int sliced_launch_shard(const gsb_db* db, Shard& sh, cudaStream_t st, const uint32_t* d_queries, uint32_t nq, uint32_t k,
                        float cutoff, unsigned long long* out_keys, uint32_t* out_n, unsigned long long* out_surv)
{
    int smem_max = 0, sms = 0;
    int rc = smem_limit(sh.device, &smem_max);
    if (rc)
        return rc;
    GSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, sh.device));
    const uint32_t nqp = (nq + 7u) & ~7u;
    const uint32_t dev_words = db->layout.dev_words;
    const uint32_t smem = gsb::sliced_tile_bytes(dev_words) +
                          gsb::kSlicedTileBatches * gsb::kBatchRows * 2 + gsb::kSlicedListEntries * 2 +
                          nqp * gsb::kSlicedPerQueryBytes;
    if (smem + 2048 > static_cast<uint32_t>(smem_max))
        return fail(GSB_ERR_INVALID, "the bit-sliced multi-query kernel does not fit this device's shared memory");
    const uint32_t n_tiles = (sh.n_tiles + gsb::kSlicedTileBatches - 1) / gsb::kSlicedTileBatches;
    int grid = std::max(1, std::min<int>(sms, n_tiles)); // all CTAs resident: grid-wide arrival counter
    if (const int g = env_int("GSB_GRID", 0))
        grid = std::max(1, std::min(grid, g));
    rc = batch_reserve(sh, db->layout, k, grid, nq);
    if (rc)
        return rc;
    rc = sliced_reserve(sh);
    if (rc)
        return rc;
    Workspace& ws = sh.ws;
    GSB_CUDA(cudaSetDevice(sh.device));
    rc = gate_launch(sh.device, ws, st, grid); // before anything of this pass touches the workspace
    if (rc)
        return rc;
    gsb::sliced_build_lists_kernel<<<1, gsb::kMaxSlicedQueries, 0, st>>>(d_queries, nq, dev_words, ws.slists, ws.slofs,
                                                                         ws.sngrp, ws.spopq, ws.smeta);
    g_launches++;
    GSB_CUDA(cudaGetLastError());
    // warps per CTA (one CTA per SM): 32 = one batch of a tile per warp in the transposition and eight
    // warps per scheduler to hide the shared-memory and ALU latencies of the counting loop
    const int warps = env_int("GSB_SLICED_WARPS", 32);
    void* kernel = nullptr;
    if (dev_words == 32)
        kernel = warps == 16   ? reinterpret_cast<void*>(gsb::scan_sliced_kernel<32, 16>)
                 : warps == 24 ? reinterpret_cast<void*>(gsb::scan_sliced_kernel<32, 24>)
                 : warps == 32 ? reinterpret_cast<void*>(gsb::scan_sliced_kernel<32, 32>)
                               : nullptr;
    else if (warps == 32) // narrower rows: the default shape only
        kernel = dev_words == 16  ? reinterpret_cast<void*>(gsb::scan_sliced_kernel<16, 32>)
                 : dev_words == 8 ? reinterpret_cast<void*>(gsb::scan_sliced_kernel<8, 32>)
                 : dev_words == 4 ? reinterpret_cast<void*>(gsb::scan_sliced_kernel<4, 32>)
                                  : nullptr;
    if (!kernel)
        return fail(GSB_ERR_INVALID, "GSB_SLICED_WARPS must be 16, 24 or 32 (32 for rows narrower than 1024 bits)");
    GSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    gsb::SlicedParams sp;
    std::memset(&sp, 0, sizeof(sp));
    sp.lists = ws.slists;
    sp.lofs = ws.slofs;
    sp.ngrp = ws.sngrp;
    sp.popq = ws.spopq;
    sp.meta = ws.smeta;
    sp.ghist = ws.shist;
    sp.gtau = ws.stau;
    fill_batch_params(db, sh, nq, k, cutoff, d_queries, out_keys, out_n, out_surv, &sp.b);
    // claims [0, n_mini) are each CTA's warm-up mini tile, then the tiles themselves
    sp.n_mini = std::min<uint32_t>(static_cast<uint32_t>(grid), n_tiles);
    sp.n_claims = sp.n_mini + n_tiles;
    GSB_CUDA(cudaMemsetAsync(ws.shist, 0, static_cast<size_t>(nq) * gsb::kSlicedHistBuckets * sizeof(unsigned int), st));
    GSB_CUDA(cudaMemsetAsync(ws.stau, 0, static_cast<size_t>(nq) * sizeof(unsigned long long), st));
    void* args[] = {&sp};
    GSB_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(warps * 32), args, smem, st));
    g_launches++;
    return GSB_OK;
}

// Tensor-core kernel over one shard (gsb_tensor.cuh): one launch per group of 128 queries, all
// queued on `st`.  The groups share the per-shard workspace; launches of one stream run one after
// the other and the last CTA of each leaves the control block clean for the next.
int tensor_launch_shard(const gsb_db* db, Shard& sh, cudaStream_t st, const uint32_t* d_queries, uint32_t nq, uint32_t k,
                        float cutoff, unsigned long long* out_keys, uint32_t* out_n, unsigned long long* out_surv)
{
    int smem_max = 0, sms = 0;
    int rc = smem_limit(sh.device, &smem_max);
    if (rc)
        return rc;
    GSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, sh.device));
    const uint32_t smem = gsb::tc_smem_bytes();
    if (db->layout.dev_words != 32 || !db->layout.rowpop || db->layout.tile_bytes != gsb::kTcRawBatch)
        return fail(GSB_ERR_INVALID, "the tensor-core multi-query kernel handles 1024-bit rows with popcount trailers");
    if (smem + 2048 > static_cast<uint32_t>(smem_max))
        return fail(GSB_ERR_INVALID, "the tensor-core multi-query kernel does not fit this device's shared memory");
    const uint32_t n_tiles = (sh.n_tiles + gsb::kTcTileBatches - 1) / gsb::kTcTileBatches;
    int grid = std::max(1, std::min<int>(sms, n_tiles)); // all CTAs resident: grid-wide arrival counter
    if (const int g = env_int("GSB_GRID", 0))
        grid = std::max(1, std::min(grid, g));
    rc = batch_reserve(sh, db->layout, k, grid, std::min<uint32_t>(nq, gsb::kTcQueries));
    if (rc)
        return rc;
    rc = sliced_reserve(sh); // thresholds and score histograms are the bit-sliced kernel's
    if (rc)
        return rc;
    Workspace& ws = sh.ws;
    GSB_CUDA(cudaSetDevice(sh.device));
    rc = gate_launch(sh.device, ws, st, grid);
    if (rc)
        return rc;
    GSB_CUDA(cudaMemsetAsync(ws.shist, 0, static_cast<size_t>(nq) * gsb::kSlicedHistBuckets * sizeof(unsigned int), st));
    GSB_CUDA(cudaMemsetAsync(ws.stau, 0, static_cast<size_t>(nq) * sizeof(unsigned long long), st));
    for (uint32_t q0 = 0; q0 < nq; q0 += gsb::kTcQueries) {
        const uint32_t n = std::min<uint32_t>(gsb::kTcQueries, nq - q0);
        gsb::TensorParams tp;
        std::memset(&tp, 0, sizeof(tp));
        fill_batch_params(db, sh, n, k, cutoff, d_queries + static_cast<size_t>(q0) * 32, out_keys + static_cast<size_t>(q0) * k,
                          out_n + q0, out_surv + q0, &tp.b);
        tp.ghist = ws.shist + static_cast<size_t>(q0) * gsb::kSlicedHistBuckets;
        tp.gtau = ws.stau + q0;
        tp.n_tiles = n_tiles;
        tp.fault = static_cast<uint32_t>(env_int("GSB_TC_FAULT", 0));
        tp.variant = static_cast<uint32_t>(env_int("GSB_TC_VARIANT", 0));
        unsigned long long* dbg = nullptr;
        const size_t dbg_n = static_cast<size_t>(grid) * gsb::kTcWarps * 8;
        if (env_int("GSB_TC_DEBUG", 0)) {
            GSB_CUDA(cudaMalloc(&dbg, dbg_n * 8));
            GSB_CUDA(cudaMemsetAsync(dbg, 0, dbg_n * 8, st));
        }
        tp.dbg = dbg;
        GSB_CUDA(gsb::tensor_kernel_launch(tp, grid, st));
        g_launches++;
        if (dbg) { // per-role wait times (clocks), averaged over the CTAs
            std::vector<unsigned long long> h(dbg_n);
            GSB_CUDA(cudaStreamSynchronize(st));
            GSB_CUDA(cudaMemcpy(h.data(), dbg, dbg_n * 8, cudaMemcpyDeviceToHost));
            cudaFree(dbg);
            // (slots 0/1 double as phase timers: expanders = stores / fence, MMA = issue, epilogue = tmem loads / filter)
            const char* site[8] = {"raw_empty|stores|issue|ldtm", "tmem_empty|fence|filter", "slab_full|epi arrive", "raw_full|mma fences", "slab_empty|commit|maint", "pd_full|raw phase", "tmem_full", "role total"};
            for (int w : {0, 1, 4, 11, 12, 19}) {
                std::fprintf(stderr, "[gsb tc dbg] warp %2d:", w);
                for (int i = 0; i < 8; i++) {
                    double sum = 0;
                    for (int c = 0; c < grid; c++)
                        sum += static_cast<double>(h[(static_cast<size_t>(c) * gsb::kTcWarps + w) * 8 + i]);
                    if (sum > 0)
                        std::fprintf(stderr, " %s %.2f", site[i], sum / grid / std::max<uint32_t>(1, (n_tiles + grid - 1) / grid));
                }
                std::fprintf(stderr, "  (clocks per tile)\n");
            }
        }
    }
    return GSB_OK;
}

// One pass over one shard for nq <= batch_max_queries(which) queries already in device memory.
// Results: out_keys [nq][k], out_survivors [nq], out_n [nq].
int batch_launch_shard(const gsb_db* db, Shard& sh, cudaStream_t st, BatchKernel which, const uint32_t* d_queries,
                       uint32_t nq, uint32_t k, float cutoff, unsigned long long* out_keys, uint32_t* out_n,
                       unsigned long long* out_surv)
{
    if (which == kBatchSliced)
        return sliced_launch_shard(db, sh, st, d_queries, nq, k, cutoff, out_keys, out_n, out_surv);
    if (which == kBatchTensor)
        return tensor_launch_shard(db, sh, st, d_queries, nq, k, cutoff, out_keys, out_n, out_surv);
    BatchPlan plan;
    int rc = make_batch_plan(db->layout, sh, &plan);
    if (rc)
        return rc;
    rc = batch_reserve(sh, db->layout, k, plan.grid, nq);
    if (rc)
        return rc;
    gsb::BatchParams p;
    fill_batch_params(db, sh, nq, k, cutoff, d_queries, out_keys, out_n, out_surv, &p);
    p.stages = plan.stages;
    GSB_CUDA(cudaSetDevice(sh.device));
    rc = gate_launch(sh.device, sh.ws, st, plan.grid);
    if (rc)
        return rc;
    return launch_batch(db->layout, p, plan, st);
}

void to_dev_words(const gsb_db* db, const int32_t* query, std::vector<uint32_t>* out);

// Host-buffer batch over every shard: groups of up to batch_max_queries(which) queries per pass.
// merged[q] receives the best k_scan keys of query q over all shards, approx[q] its survivor count.
int search_batch_kernel_path(const gsb_db* db, BatchKernel which, const int32_t* query_words, int n_queries,
                             uint32_t k_scan, float cutoff, std::vector<std::vector<unsigned long long>>* merged_out,
                             std::vector<uint64_t>* approx_out)
{
    const Layout& l = db->layout;
    const uint32_t words = db->words, k = k_scan;
    const int group = static_cast<int>(batch_max_queries(which));
    merged_out->assign(n_queries, {});
    approx_out->assign(n_queries, 0);
    std::vector<uint32_t> one;
    for (int q0 = 0; q0 < n_queries; q0 += group) {
        const uint32_t nq = static_cast<uint32_t>(std::min<int>(group, n_queries - q0));
        std::vector<uint32_t> padded(static_cast<size_t>(nq) * l.dev_words, 0u);
        for (uint32_t j = 0; j < nq; j++) { // zero padded to the device width, folded like the rows
            to_dev_words(db, query_words + static_cast<size_t>(q0 + j) * words, &one);
            std::memcpy(padded.data() + static_cast<size_t>(j) * l.dev_words, one.data(), l.dev_words * 4);
        }
        for (size_t i = 0; i < db->shards.size(); i++) {
            Shard& sh = const_cast<Shard&>(db->shards[i]);
            if (sh.n_rows == 0)
                continue;
            // reserve for the largest grid a launch can use (one CTA per SM), so that the launch
            // itself never re-allocates the buffers the queries are being copied into
            int sms = 0;
            GSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, sh.device));
            int rc = batch_reserve(sh, l, k, sms, nq);
            if (rc)
                return rc;
            GSB_CUDA(cudaSetDevice(sh.device));
            rc = gate_launch(sh.device, sh.ws, sh.ws.stream, 0); // the copy below already touches the workspace
            if (rc)
                return rc;
            GSB_CUDA(cudaMemcpyAsync(sh.ws.bqueries, padded.data(), padded.size() * 4, cudaMemcpyHostToDevice,
                                     sh.ws.stream));
            unsigned long long* keys = sh.ws.bout;
            unsigned long long* surv = keys + static_cast<size_t>(nq) * k;
            uint32_t* cnt = reinterpret_cast<uint32_t*>(surv + nq);
            rc = batch_launch_shard(db, sh, sh.ws.stream, which, sh.ws.bqueries, nq, k, cutoff, keys, cnt, surv);
            if (rc)
                return rc;
            GSB_CUDA(cudaMemcpyAsync(sh.ws.bout_host, sh.ws.bout, (static_cast<size_t>(nq) * (k + 2ull)) * 8,
                                     cudaMemcpyDeviceToHost, sh.ws.stream));
        }
        bool failed = false;
        for (size_t i = 0; i < db->shards.size(); i++) {
            Shard& sh = const_cast<Shard&>(db->shards[i]);
            if (sh.n_rows == 0)
                continue;
            GSB_CUDA(cudaSetDevice(sh.device));
            GSB_CUDA(cudaStreamSynchronize(sh.ws.stream));
            const unsigned long long* keys = sh.ws.bout_host;
            const unsigned long long* surv = keys + static_cast<size_t>(nq) * k;
            const uint32_t* cnt = reinterpret_cast<const uint32_t*>(surv + nq);
            for (uint32_t j = 0; j < nq; j++) {
                if (cnt[j] == gsb::kCountError) {
                    failed = true;
                    break;
                }
                (*approx_out)[q0 + j] += surv[j];
                auto& m = (*merged_out)[q0 + j];
                const size_t old = m.size();
                m.insert(m.end(), keys + static_cast<size_t>(j) * k, keys + static_cast<size_t>(j) * k + cnt[j]);
                std::inplace_merge(m.begin(), m.begin() + old, m.end(), std::greater<unsigned long long>());
                if (m.size() > k)
                    m.resize(k);
            }
            if (failed)
                ws_recover(sh);
        }
        if (failed)
            return fail(GSB_ERR_CUDA, device_error_text(0));
    }
    return GSB_OK;
}

void to_dev_words(const gsb_db* db, const int32_t* query, std::vector<uint32_t>* out)
{
    out->assign(db->layout.dev_words, 0u);
    if (db->fold_factor == 1) {
        std::memcpy(out->data(), query, db->words * 4);
    } else {
        fold_row(reinterpret_cast<const uint32_t*>(query), db->words, db->fold_factor, out->data());
    }
}

// Host twin of the kernels' epilogue (gsb::similarity): reference calculation_functors.cpp:8-19 /
// fingerprintdb_cuda.cu:387-399 for Tanimoto; Dice and Tversky as defined in gsb_kernels.cuh (f32,
// every operation rounded — this file is compiled with -ffp-contract=off).
float score_cpu(const gsb_db* db, const uint32_t* q, const uint32_t* d, uint32_t words)
{
    int pq = 0, pd = 0, common = 0;
    for (uint32_t i = 0; i < words; i++) {
        pq += __builtin_popcount(q[i]);
        pd += __builtin_popcount(d[i]);
        common += __builtin_popcount(q[i] & d[i]);
    }
    if (db->metric == gsb::kMetricTanimoto)
        return static_cast<float>(common) / static_cast<float>(pq + pd - common);
    if (db->metric == gsb::kMetricDice)
        return static_cast<float>(2 * common) / static_cast<float>(pq + pd);
    const float c = static_cast<float>(common);
    const float t1 = db->alpha * static_cast<float>(pq - common);
    const float t2 = db->beta * static_cast<float>(pd - common);
    const float den = (t1 + t2) + c;
    return c / den;
}

// ---- folded search, second stage on the device ------------------------------------------------
void unregister_host_rows(gsb_db* db)
{
    const bool stale = db->rescore_epoch != g_reset_epoch.load(); // a device reset freed all of it already
    if (db->host_registered && !stale)
        for (auto& hc : db->host)
            if (!hc.bytes.pinned())
                cudaHostUnregister(hc.bytes.data());
    db->host_registered = false;
    if (db->d_chunks && !stale)
        cudaFree(db->d_chunks);
    db->d_chunks = nullptr;
    if (db->rescore_host && !stale)
        cudaFreeHost(db->rescore_host);
    db->rescore_host = db->rescore_dev = nullptr;
    db->rescore_cap = 0;
    cudaGetLastError();
}

// Page-lock and map the unfolded host chunks so that the re-score kernel can read candidate rows
// in place.  Failure is not an error (the re-score then runs on the host, as in the reference).
void register_host_rows(gsb_db* db)
{
    if (db->host_registered || db->host.empty() || db->shards.empty() || env_int("GSB_FOLD_RESCORE_HOST", 0))
        return;
    db->rescore_epoch = g_reset_epoch.load();
    cudaSetDevice(db->shards[0].device);
    std::vector<gsb::RescoreChunk> table;
    size_t done = 0;
    bool ok = true;
    for (auto& hc : db->host) {
        if (!hc.bytes.pinned() && // (pinned chunks are mapped from birth)
            cudaHostRegister(hc.bytes.data(), hc.bytes.size(), cudaHostRegisterPortable | cudaHostRegisterMapped) !=
                cudaSuccess) {
            ok = false;
            break;
        }
        done++;
        void* d = nullptr;
        if (cudaHostGetDevicePointer(&d, hc.bytes.data(), 0) != cudaSuccess) {
            ok = false;
            break;
        }
        table.push_back({static_cast<const uint32_t*>(d), hc.row0, hc.n_rows});
    }
    if (ok && cudaMalloc(&db->d_chunks, table.size() * sizeof(gsb::RescoreChunk)) == cudaSuccess &&
        cudaMemcpy(db->d_chunks, table.data(), table.size() * sizeof(gsb::RescoreChunk), cudaMemcpyHostToDevice) ==
            cudaSuccess) {
        db->host_registered = true;
        return;
    }
    for (size_t i = 0; i < done; i++)
        if (!db->host[i].bytes.pinned())
            cudaHostUnregister(db->host[i].bytes.data());
    cudaGetLastError();
    std::fprintf(stderr, "[gpusim_b200] could not map the host rows for the device re-score; folded searches "
                         "re-score on the host\n");
}

// keys: candidates of the folded scan in canonical order.  On return keys holds, best first, the
// re-scored keys ((score bits + 1) << 32 | 0xFFFFFFFF - candidate index; high word 0 = NaN).
int rescore_on_device(gsb_db* db, const uint32_t* qfull, const std::vector<unsigned long long>& cand,
                      std::vector<unsigned long long>* out)
{
    const size_t n = cand.size();
    out->clear();
    if (n == 0)
        return GSB_OK;
    Shard& sh = db->shards[0];
    GSB_CUDA(cudaSetDevice(sh.device));
    if (2 * n + 1 > db->rescore_cap) {
        if (db->rescore_host) {
            GSB_CUDA(cudaStreamSynchronize(sh.ws.stream));
            GSB_CUDA(cudaFreeHost(db->rescore_host));
        }
        db->rescore_host = db->rescore_dev = nullptr;
        db->rescore_cap = 0;
        const size_t cap = 2 * static_cast<size_t>(pow2ceil(n)) + 1;
        void* h = nullptr;
        void* d = nullptr;
        GSB_CUDA(cudaHostAlloc(&h, cap * 8, cudaHostAllocMapped | cudaHostAllocPortable));
        GSB_CUDA(cudaHostGetDevicePointer(&d, h, 0));
        db->rescore_host = static_cast<unsigned long long*>(h);
        db->rescore_dev = static_cast<unsigned long long*>(d);
        db->rescore_cap = cap;
    }
    std::memcpy(db->rescore_host, cand.data(), n * 8);
    gsb::RescoreParams p;
    std::memset(&p, 0, sizeof(p));
    p.cand = db->rescore_dev;
    p.out = db->rescore_dev + n;
    p.n = static_cast<uint32_t>(n);
    p.words = db->words;
    p.n_chunks = static_cast<uint32_t>(db->host.size());
    p.chunks = db->d_chunks;
    p.metric = db->metric;
    p.alpha = db->alpha;
    p.beta = db->beta;
    std::memcpy(p.q, qfull, db->words * 4);
    const unsigned blocks = static_cast<unsigned>((n * 32 + 255) / 256);
    gsb::rescore_kernel<<<blocks, 256, 0, sh.ws.stream>>>(p);
    g_launches++;
    GSB_CUDA(cudaGetLastError());
    GSB_CUDA(cudaStreamSynchronize(sh.ws.stream));
    out->assign(db->rescore_host + n, db->rescore_host + 2 * n);
    std::sort(out->begin(), out->end(), std::greater<unsigned long long>());
    return GSB_OK;
}

int fetch_rows(const gsb_db* db, const std::vector<uint64_t>& rows, std::vector<uint32_t>* out);

} // namespace

// Pinned buffers are plain (huge-page advised) memory that is page-locked with cudaHostRegister, not
// cudaHostAlloc memory: the rows must outlive a cudaDeviceReset (gsb_devices_reset re-registers
// every live buffer from this table; cudaHostAlloc memory would be unmapped with the context).
namespace {
std::mutex g_hostbuf_mu;
std::map<void*, size_t> g_hostbufs; // live registered buffers
} // namespace

bool GsbHostBuf::allocate(size_t n)
{
    release();
    if (n == 0)
        return true;
    void* p = nullptr;
    const size_t align = n >= (2u << 20) ? (2u << 20) : 4096;
    if (posix_memalign(&p, align, n) != 0)
        return false;
#ifdef MADV_HUGEPAGE
    if (align > 4096)
        madvise(p, n, MADV_HUGEPAGE); // 512 x fewer pages to fault and lock
#endif
    m_p = static_cast<uint8_t*>(p);
    m_n = n;
    m_pinned = false;
    if (env_int("GSB_PINNED_HOST", 1) && gsb_device_count() > 0) {
        if (cudaHostRegister(p, n, cudaHostRegisterPortable | cudaHostRegisterMapped) == cudaSuccess) {
            m_pinned = true;
            std::lock_guard<std::mutex> lock(g_hostbuf_mu);
            g_hostbufs[p] = n;
        } else {
            cudaGetLastError(); // e.g. the locked-memory limit: plain memory and the staged upload instead
        }
    }
    return true;
}

void GsbHostBuf::release()
{
    if (m_p) {
        if (m_pinned) {
            {
                std::lock_guard<std::mutex> lock(g_hostbuf_mu);
                g_hostbufs.erase(m_p);
            }
            if (cudaHostUnregister(m_p) != cudaSuccess)
                cudaGetLastError();
        }
        std::free(m_p);
    }
    m_p = nullptr;
    m_n = 0;
    m_pinned = false;
}

int gsb_db_create_adopt(std::vector<GsbHostBuf>&& chunks, int fp_bits, uint64_t fp_count, gsb_db** out)
{
    if (!out)
        return fail(GSB_ERR_INVALID, "null argument");
    if (fp_bits <= 0 || fp_bits % 32 != 0 || fp_bits / 32 > GSB_MAX_WORDS)
        return fail(GSB_ERR_INVALID, "fp_bitcount must be a multiple of 32 in [32, 4096]");
    if (fp_count > 0xfffffffeull) // results carry 32-bit row ids (the .fsim format caps N at 2^31-1)
        return fail(GSB_ERR_INVALID, "more than 2^32-2 rows");
    const uint64_t row_bytes = fp_bits / 8;
    uint64_t total = 0;
    for (const auto& c : chunks) {
        if (c.size() % row_bytes != 0)
            return fail(GSB_ERR_CORRUPT, "Mismatch between FP count and data, potential database corruption.");
        total += c.size() / row_bytes;
    }
    if (total != fp_count) // reference fingerprintdb_cuda.cu:153-156
        return fail(GSB_ERR_CORRUPT, "Mismatch between FP count and data, potential database corruption.");
    std::unique_ptr<gsb_db> db(new gsb_db);
    db->fp_bits = fp_bits;
    db->words = fp_bits / 32;
    db->count = fp_count;
    uint64_t row0 = 0;
    for (auto& c : chunks) {
        if (c.empty())
            continue;
        gsb_db::HostChunk hc;
        hc.n_rows = c.size() / row_bytes;
        hc.row0 = row0;
        hc.bytes = std::move(c);
        row0 += hc.n_rows;
        db->host.push_back(std::move(hc));
    }
    *out = db.release();
    return GSB_OK;
}

// ====================================================================================== C ABI
extern "C" {

const char* gsb_last_error(void) { return g_err.c_str(); }
const char* gsb_version(void) { return "gpusim_b200 0.1 (sm_100a)"; }
uint64_t gsb_launch_count(void) { return g_launches.load(); }

int gsb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

uint64_t gsb_device_free_bytes(int device)
{
    size_t free_b = 0, total = 0;
    int cur = 0;
    cudaGetDevice(&cur);
    if (cudaSetDevice(device) != cudaSuccess || cudaMemGetInfo(&free_b, &total) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    cudaSetDevice(cur);
    return free_b;
}

uint64_t gsb_available_device_bytes(void)
{
    uint64_t sum = 0;
    for (int d = 0; d < gsb_device_count(); d++)
        sum += gsb_device_free_bytes(d);
    return sum;
}

uint64_t gsb_layout_bytes(int fp_bits, uint64_t rows, unsigned fold_factor)
{
    if (fp_bits <= 0 || fp_bits % 32 != 0)
        return 0;
    const uint32_t words = static_cast<uint32_t>(fp_bits / 32);
    unsigned f = std::max(1u, fold_factor);
    while (words % f != 0) // copyToGPU's normalisation, reference .cu:170-173
        f++;
    Layout l;
    if (make_layout(words / f, &l) != GSB_OK)
        return 0;
    const uint64_t tiles = (rows + l.tile_rows - 1) / l.tile_rows;
    return tiles * l.tile_stride + l.unit_bytes;
}

int gsb_devices_reset(void)
{
    const int n = gsb_device_count();
    for (int d = 0; d < n; d++) {
        if (cudaSetDevice(d) != cudaSuccess || cudaDeviceReset() != cudaSuccess) {
            const cudaError_t e = cudaGetLastError();
            return fail(GSB_ERR_CUDA, std::string("cudaDeviceReset: ") + cudaGetErrorString(e));
        }
        std::lock_guard<std::mutex> lock(g_gates[d < 64 ? d : 63].mu);
        g_gates[d < 64 ? d : 63].inflight.clear();
        g_gates[d < 64 ? d : 63].event = nullptr;
    }
    cudaGetLastError();
    g_reset_epoch++;
    if (n > 0) { // the reset dropped every page-lock: put the host rows back under it
        cudaSetDevice(0);
        std::lock_guard<std::mutex> lock(g_hostbuf_mu);
        for (auto& kv : g_hostbufs)
            if (cudaHostRegister(kv.first, kv.second, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) {
                const cudaError_t e = cudaGetLastError();
                return fail(GSB_ERR_CUDA, std::string("re-registering host rows after the reset: ") + cudaGetErrorString(e));
            }
    }
    return GSB_OK;
}

int gsb_next_device(uint64_t required_bytes, int* device)
{
    static std::atomic<unsigned> next{0};
    const int n = gsb_device_count();
    if (n == 0)
        return fail(GSB_ERR_CUDA, "no CUDA device");
    for (int i = 0; i < n; i++) {
        const int gpu = static_cast<int>(next++ % static_cast<unsigned>(n));
        if (gsb_device_free_bytes(gpu) > required_bytes) {
            *device = gpu;
            return GSB_OK;
        }
    }
    return fail(GSB_ERR_NOMEM, "Can't find a GPU with enough memory to copy data.");
}

int gsb_db_create(const void* const* chunk_ptrs, const uint64_t* chunk_bytes, int n_chunks,
                  int fp_bits, uint64_t fp_count, gsb_db** out)
{
    if (!out || n_chunks < 0 || (n_chunks > 0 && (!chunk_ptrs || !chunk_bytes)))
        return fail(GSB_ERR_INVALID, "null argument");
    GSB_TRY
    std::vector<GsbHostBuf> chunks(n_chunks);
    for (int c = 0; c < n_chunks; c++) { // the reference copies the chunk bytes too (.cu:123-125)
        if (!chunks[c].allocate(chunk_bytes[c]))
            return fail(GSB_ERR_NOMEM, "out of host memory");
        const uint8_t* src = static_cast<const uint8_t*>(chunk_ptrs[c]);
        uint8_t* dst = chunks[c].data();
        parallel_for(chunk_bytes[c] >> 12, [=](uint64_t lo, uint64_t hi) {
            std::memcpy(dst + (lo << 12), src + (lo << 12), (hi - lo) << 12);
        });
        const uint64_t tail = chunk_bytes[c] & ~0xfffull;
        std::memcpy(dst + tail, src + tail, chunk_bytes[c] - tail);
    }
    return gsb_db_create_adopt(std::move(chunks), fp_bits, fp_count, out);
    GSB_CATCH
}

int gsb_db_create_synthetic_sharded(const int* devices, int n_devices, int fp_bits, uint64_t n_rows,
                                    uint64_t row_base, uint64_t seed, uint32_t plant_period, gsb_db** out)
{
    GSB_TRY
    if (!out || !devices || n_devices < 1)
        return fail(GSB_ERR_INVALID, "null argument");
    if (fp_bits <= 0 || fp_bits % 32 != 0 || fp_bits / 32 > GSB_MAX_WORDS)
        return fail(GSB_ERR_INVALID, "fp_bitcount must be a multiple of 32 in [32, 4096]");
    if (row_base + n_rows > 0xffffffffull)
        return fail(GSB_ERR_INVALID, "row ids must fit 32 bits");
    for (int i = 0; i < n_devices; i++)
        if (devices[i] < 0 || devices[i] >= gsb_device_count())
            return fail(GSB_ERR_CUDA, "no such CUDA device");
    // (unique_ptr with a destroying deleter: a failure half way releases the shards made so far)
    std::unique_ptr<gsb_db, void (*)(gsb_db*)> db(new gsb_db, gsb_db_destroy);
    db->fp_bits = fp_bits;
    db->words = fp_bits / 32;
    db->count = n_rows;
    db->synth_seed = seed;
    int rc = make_layout(db->words, &db->layout);
    if (rc)
        return rc;
    if (db->layout.dev_words != db->words)
        return fail(GSB_ERR_INVALID, "synthetic shards need a power-of-two word count >= 4");
    const Layout& l = db->layout;
    const uint64_t per = (n_rows + n_devices - 1) / n_devices;
    db->shards.resize(n_devices);
    for (int i = 0; i < n_devices; i++) { // contiguous, equal row ranges; generation runs on all devices at once
        Shard& sh = db->shards[i];
        sh.device = devices[i];
        const uint64_t off = std::min<uint64_t>(n_rows, static_cast<uint64_t>(i) * per);
        sh.row_base = row_base + off;
        sh.n_rows = std::min<uint64_t>(n_rows - off, per);
        rc = ws_init(sh);
        if (rc)
            return rc;
        sh.n_tiles = static_cast<uint32_t>((sh.n_rows + l.tile_rows - 1) / l.tile_rows);
        sh.bytes = padded_tile_bytes(l, sh.n_tiles);
        if (gsb_device_free_bytes(sh.device) <= sh.bytes)
            return fail(GSB_ERR_NOMEM, "Can't find a GPU with enough memory to copy data.");
        GSB_CUDA(cudaSetDevice(sh.device));
        GSB_CUDA(cudaMalloc(&sh.tiles, sh.bytes));
        GSB_CUDA(cudaMemsetAsync(sh.tiles, 0, sh.bytes, sh.ws.stream)); // padding batches must read as empty rows
        const uint64_t n_words = static_cast<uint64_t>(sh.n_tiles) * l.tile_rows * l.dev_words;
        if (n_words) {
            const uint64_t blocks = (n_words + 255) / 256;
            if (blocks > 0x7fffffffull)
                return fail(GSB_ERR_INVALID, "shard too large");
            gsb::synth_fill_kernel<<<static_cast<unsigned>(blocks), 256, 0, sh.ws.stream>>>(
                sh.tiles, sh.n_rows, sh.row_base, sh.n_tiles, l.tile_rows, l.tile_stride, l.dev_words, seed, plant_period);
            g_launches++;
            if (l.rowpop) {
                const uint64_t rows = static_cast<uint64_t>(sh.n_tiles) * l.tile_rows;
                gsb::tile_popcount_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, sh.ws.stream>>>(
                    sh.tiles, sh.n_tiles, l.tile_rows, l.tile_stride, l.dev_words);
                g_launches++;
            }
        }
    }
    for (Shard& sh : db->shards) {
        GSB_CUDA(cudaSetDevice(sh.device));
        GSB_CUDA(cudaStreamSynchronize(sh.ws.stream));
        GSB_CUDA(cudaGetLastError());
    }
    db->uploaded = true;
    *out = db.release();
    return GSB_OK;
    GSB_CATCH
}

int gsb_db_create_synthetic(int device, int fp_bits, uint64_t n_rows, uint64_t row_base, uint64_t seed,
                            uint32_t plant_period, gsb_db** out)
{
    return gsb_db_create_synthetic_sharded(&device, 1, fp_bits, n_rows, row_base, seed, plant_period, out);
}

int gsb_db_upload(gsb_db* db, const int* devices, int n_devices, unsigned fold_factor)
{
    if (!db)
        return fail(GSB_ERR_INVALID, "null database");
    std::lock_guard<std::mutex> lock(db->mu);
    if (db->host.empty() && db->count > 0)
        return fail(GSB_ERR_STATE, "device-generated shards are already uploaded");
    const int n_vis = gsb_device_count();
    if (n_vis == 0)
        return fail(GSB_ERR_CUDA, "no CUDA device: the GPU search path has no CPU fallback");
    for (auto& sh : db->shards)
        ws_free(sh);
    db->shards.clear();
    db->uploaded = false;
    for (auto& pd : db->pending)
        pd.ticket = 0;
    unregister_host_rows(db);
    // reference fingerprintdb_cuda.cu:170-173
    unsigned f = std::max(1u, fold_factor);
    while (db->words % f != 0)
        f++;
    db->fold_factor = f;
    int rc = make_layout(db->words / f, &db->layout);
    if (rc)
        return rc;
    std::vector<int> devs;
    if (devices && n_devices > 0) {
        for (int i = 0; i < n_devices; i++) {
            if (devices[i] < 0 || devices[i] >= n_vis)
                return fail(GSB_ERR_INVALID, "no such CUDA device");
            devs.push_back(devices[i]);
        }
    } else {
        // as many devices as needed: keep shards >= 1 GiB so tiny databases stay on one GPU
        const uint64_t bytes = db->count * db->layout.row_bytes;
        const int want = static_cast<int>(std::min<uint64_t>(n_vis, std::max<uint64_t>(1, bytes >> 30)));
        for (int i = 0; i < want; i++) {
            int d = 0;
            rc = gsb_next_device(0, &d);
            if (rc)
                return rc;
            devs.push_back(d);
        }
    }
    const uint64_t per = (db->count + devs.size() - 1) / devs.size();
    db->shards.resize(devs.size());
    for (size_t i = 0; i < devs.size(); i++) {
        Shard& sh = db->shards[i];
        sh.device = devs[i];
        sh.row_base = std::min<uint64_t>(db->count, i * per);
        sh.n_rows = std::min<uint64_t>(db->count - sh.row_base, per);
        rc = ws_init(sh);
        if (rc)
            return rc;
    }
    if (devs.size() == 1) {
        rc = upload_rows(db, db->shards[0]);
        if (rc)
            return rc;
    } else {
        // every device has its own PCIe link: one host thread per shard keeps them all busy
        std::vector<int> rcs(devs.size(), GSB_OK);
        std::vector<std::string> errs(devs.size());
        std::vector<std::thread> pool;
        for (size_t i = 0; i < devs.size(); i++)
            pool.emplace_back([db, i, &rcs, &errs]() {
                rcs[i] = upload_rows(db, db->shards[i]);
                if (rcs[i])
                    errs[i] = g_err; // thread-local: carry it over to the caller's thread
            });
        for (auto& th : pool)
            th.join();
        for (size_t i = 0; i < devs.size(); i++)
            if (rcs[i])
                return fail(rcs[i], errs[i]);
    }
    if (f > 1) // second stage of folded searches reads the full rows in place (mapped host memory)
        register_host_rows(db);
    db->uploaded = true;
    return GSB_OK;
}

void gsb_db_destroy(gsb_db* db)
{
    if (!db)
        return;
    unregister_host_rows(db);
    for (auto& sh : db->shards)
        ws_free(sh);
    delete db;
}

uint64_t gsb_db_count(const gsb_db* db) { return db ? db->count : 0; }
int gsb_db_fp_bits(const gsb_db* db) { return db ? db->fp_bits : 0; }
uint64_t gsb_db_data_bytes(const gsb_db* db) { return db ? db->count * db->words * 4ull : 0; }
unsigned gsb_db_fold_factor(const gsb_db* db) { return db ? db->fold_factor : 0; }
int gsb_db_shard_count(const gsb_db* db) { return db ? static_cast<int>(db->shards.size()) : 0; }

int gsb_db_get_fingerprint(const gsb_db* db, uint64_t row, int32_t* out_words)
{
    if (!db || !out_words)
        return fail(GSB_ERR_INVALID, "null argument");
    if (row >= db->count)
        return fail(GSB_ERR_INVALID, "row out of range");
    if (!db->host.empty()) {
        std::memcpy(out_words, db->host_row(row), db->words * 4);
        return GSB_OK;
    }
    std::vector<uint32_t> tmp;
    int rc = fetch_rows(db, {row}, &tmp);
    if (rc)
        return rc;
    std::memcpy(out_words, tmp.data(), db->words * 4);
    return GSB_OK;
}

} // extern "C"

namespace
{
// reference .cu:284-287: candidates to pull back from the (folded) scan
uint64_t scan_candidates(const gsb_db* db, uint32_t k)
{
    const unsigned f = db->fold_factor;
    uint64_t k_scan = k;
    if (f > 1)
        k_scan = static_cast<uint64_t>(k) * f * static_cast<uint64_t>(std::log2(2.0 * f));
    return std::min<uint64_t>(k_scan, db->count);
}

// Canonical keys of the scan -> caller's arrays.  Unfolded: decode.  Folded: second stage, the
// candidates are scored again with their full fingerprints (reference .cu:307-331), ordered by the
// new score (stable over the candidate order == top_results_bubble_sort, .cpp:92-103), cut to k and
// at the first score below the cutoff (:323-326).
int finish_search(const gsb_db* db, const int32_t* query_words, uint32_t k, float cutoff,
                  std::vector<unsigned long long>& keys, uint32_t* out_rows, float* out_scores, uint32_t* out_n)
{
    if (db->fold_factor == 1) {
        const uint32_t n = static_cast<uint32_t>(std::min<uint64_t>(keys.size(), k));
        for (uint32_t i = 0; i < n; i++) {
            const Cand c = decode(keys[i]);
            out_rows[i] = c.row;
            out_scores[i] = c.score;
        }
        *out_n = n;
        return GSB_OK;
    }
    const uint32_t* qfull = reinterpret_cast<const uint32_t*>(query_words);
    std::vector<Cand> cand(keys.size());
    if (db->host_registered) {
        std::vector<unsigned long long> rescored;
        int rc = rescore_on_device(const_cast<gsb_db*>(db), qfull, keys, &rescored);
        if (rc)
            return rc;
        for (size_t i = 0; i < rescored.size(); i++) {
            const uint32_t idx = 0xffffffffu - static_cast<uint32_t>(rescored[i] & 0xffffffffu);
            const uint32_t hi = static_cast<uint32_t>(rescored[i] >> 32);
            const uint32_t bits = hi ? hi - 1u : 0x7fc00000u;
            cand[i].row = decode(keys[idx]).row;
            std::memcpy(&cand[i].score, &bits, 4);
        }
    } else {
        for (size_t i = 0; i < keys.size(); i++) {
            cand[i].row = decode(keys[i]).row;
            cand[i].score = score_cpu(db, qfull, db->host_row(cand[i].row), db->words);
        }
        auto nan_last = [](float sc) { return sc != sc ? -1.0f : sc; };
        std::stable_sort(cand.begin(), cand.end(),
                         [&](const Cand& a, const Cand& b) { return nan_last(a.score) > nan_last(b.score); });
    }
    uint32_t n = static_cast<uint32_t>(std::min<uint64_t>(cand.size(), k));
    for (uint32_t i = 0; i < n; i++) {
        if (cand[i].score < cutoff) { // .cu:323-326
            n = i;
            break;
        }
        out_rows[i] = cand[i].row;
        out_scores[i] = cand[i].score;
    }
    *out_n = n;
    return GSB_OK;
}

int check_search_args(const gsb_db* db, const int32_t* query_words, int n_words)
{
    if (!db || !query_words)
        return fail(GSB_ERR_INVALID, "null argument");
    if (n_words != static_cast<int>(db->words))
        return fail(GSB_ERR_INVALID, "query width does not match the database");
    if (gsb_device_count() == 0)
        return fail(GSB_ERR_CUDA, "no CUDA device: the GPU search path has no CPU fallback");
    if (!db->uploaded)
        return fail(GSB_ERR_STATE, "database is not on the GPU: call gsb_db_upload first");
    return GSB_OK;
}

// One launch serves the query (unfolded rows and a k the in-kernel select can hold)?
bool one_launch_query(const gsb_db* db, uint32_t k)
{
    if (db->fold_factor != 1 || k == 0)
        return false;
    Plan probe;
    for (const Shard& sh : db->shards)
        if (sh.n_rows && make_plan(db->layout, sh, k, &probe) != GSB_OK)
            return false;
    return true;
}

int search_sync_locked(const gsb_db* db, const int32_t* query_words, uint32_t k, float cutoff, uint32_t* out_rows,
                       float* out_scores, uint32_t* out_n, uint64_t* out_approx)
{
    std::vector<uint32_t> q;
    to_dev_words(db, query_words, &q);
    const uint64_t k_scan = scan_candidates(db, k);
    std::vector<unsigned long long> keys;
    uint64_t survivors = 0;
    int rc = scan_topk(db, q.data(), std::max<uint64_t>(k_scan, 1), cutoff, &keys, &survivors);
    if (rc)
        return rc;
    if (keys.size() > k_scan)
        keys.resize(k_scan);
    if (out_approx)
        *out_approx = survivors;
    return finish_search(db, query_words, k, cutoff, keys, out_rows, out_scores, out_n);
}
} // namespace

extern "C" {

int gsb_db_search(const gsb_db* db, const int32_t* query_words, int n_words, uint32_t k, float cutoff,
                  uint32_t* out_rows, float* out_scores, uint32_t* out_n, uint64_t* out_approx)
{
    GSB_TRY
    if (!out_n || (k > 0 && (!out_rows || !out_scores)))
        return fail(GSB_ERR_INVALID, "null argument");
    int rc = check_search_args(db, query_words, n_words);
    if (rc)
        return rc;
    std::lock_guard<std::mutex> lock(db->mu);
    return search_sync_locked(db, query_words, k, cutoff, out_rows, out_scores, out_n, out_approx);
    GSB_CATCH
}

int gsb_db_search_async(const gsb_db* db, const int32_t* query_words, int n_words, uint32_t k, float cutoff,
                        uint64_t* ticket)
{
    GSB_TRY
    if (!ticket)
        return fail(GSB_ERR_INVALID, "null argument");
    int rc = check_search_args(db, query_words, n_words);
    if (rc)
        return rc;
    std::lock_guard<std::mutex> lock(db->mu);
    const uint64_t t = db->next_ticket;
    gsb_db::Pending& pd = db->pending[t % kAsyncDepth];
    if (pd.ticket != 0)
        return fail(GSB_ERR_STATE, "too many searches in flight: wait for ticket " + std::to_string(pd.ticket) + " first");
    pd.k = k;
    pd.cutoff = cutoff;
    pd.deferred = !one_launch_query(db, k);
    pd.query.assign(query_words, query_words + n_words);
    if (!pd.deferred) {
        std::vector<uint32_t> q;
        to_dev_words(db, query_words, &q);
        pd.seq = ++db->done_seq;
        rc = enqueue_scan(db, q.data(), k, cutoff, ~0ull, static_cast<int>(t % kAsyncDepth), pd.seq);
        if (rc)
            return rc;
    }
    pd.ticket = t;
    db->next_ticket++;
    *ticket = t;
    return GSB_OK;
    GSB_CATCH
}

int gsb_db_search_wait(const gsb_db* db, uint64_t ticket, uint32_t* out_rows, float* out_scores, uint32_t* out_n,
                       uint64_t* out_approx)
{
    GSB_TRY
    if (!db || !out_n || ticket == 0)
        return fail(GSB_ERR_INVALID, "null argument");
    gsb_db::Pending& pd = db->pending[ticket % kAsyncDepth];
    gsb_db::Pending mine;
    {
        std::lock_guard<std::mutex> lock(db->mu);
        if (pd.ticket != ticket)
            return fail(GSB_ERR_STATE, "no such search in flight");
        mine = pd;
    }
    if (mine.k > 0 && (!out_rows || !out_scores))
        return fail(GSB_ERR_INVALID, "null argument");
    int rc;
    if (mine.deferred) { // folded rows / very large k: several launches, done here
        std::lock_guard<std::mutex> lock(db->mu);
        rc = search_sync_locked(db, mine.query.data(), mine.k, mine.cutoff, out_rows, out_scores, out_n, out_approx);
    } else {
        // the record is this ticket's until the slot is released below: no lock while waiting
        std::vector<unsigned long long> keys;
        uint64_t survivors = 0;
        rc = collect_scan(db, mine.k, static_cast<int>(ticket % kAsyncDepth), mine.seq, &keys, &survivors);
        if (rc == GSB_OK) {
            if (out_approx)
                *out_approx = survivors;
            rc = finish_search(db, mine.query.data(), mine.k, mine.cutoff, keys, out_rows, out_scores, out_n);
        }
    }
    std::lock_guard<std::mutex> lock(db->mu);
    pd.ticket = 0;
    return rc;
    GSB_CATCH
}

int gsb_db_search_batch(const gsb_db* db, const int32_t* query_words, int n_words, int n_queries, uint32_t k,
                        float cutoff, uint32_t* out_rows, float* out_scores, uint32_t* out_n,
                        uint64_t* out_approx)
{
    GSB_TRY
    if (n_queries < 0)
        return fail(GSB_ERR_INVALID, "negative query count");
    if (n_queries == 0)
        return GSB_OK;
    if (!out_n || (k > 0 && (!out_rows || !out_scores)))
        return fail(GSB_ERR_INVALID, "null argument");
    int rc = check_search_args(db, query_words, n_words);
    if (rc)
        return rc;
    std::lock_guard<std::mutex> lock(db->mu);
    // folded rows: the scan pulls k * F * floor(log2 2F) candidates per query (reference .cu:284-287)
    const uint64_t k_scan = db->fold_factor == 1 ? k : scan_candidates(db, k);
    int64_t set_bits = 0; // (of the unfolded queries: an upper bound of what the scan sees)
    for (size_t i = 0; i < static_cast<size_t>(n_queries) * n_words; i++)
        set_bits += __builtin_popcount(static_cast<uint32_t>(query_words[i]));
    const BatchKernel which = k_scan <= gsb::kMaxBatchK
                                  ? batch_kernel_choice(db, static_cast<uint32_t>(k_scan), n_queries, cutoff, set_bits)
                                  : kBatchNone;
    if (which != kBatchNone) {
        std::vector<std::vector<unsigned long long>> merged;
        std::vector<uint64_t> approx;
        rc = search_batch_kernel_path(db, which, query_words, n_queries, static_cast<uint32_t>(k_scan), cutoff, &merged,
                                      &approx);
        if (rc)
            return rc;
        for (int qi = 0; qi < n_queries; qi++) {
            rc = finish_search(db, query_words + static_cast<size_t>(qi) * n_words, k, cutoff, merged[qi],
                               out_rows + static_cast<size_t>(qi) * k, out_scores + static_cast<size_t>(qi) * k, out_n + qi);
            if (rc)
                return rc;
            if (out_approx)
                out_approx[qi] = approx[qi];
        }
        return GSB_OK;
    }
    // No multi-query kernel for this shape (rows wider than 1024 bits, more than 512 candidates per
    // query, plain layout): one scan per query.  gsb_db_batch_mode() tells callers beforehand.
    if (n_queries >= 8) {
        static std::atomic<bool> warned{false};
        if (!warned.exchange(true))
            std::fprintf(stderr, "[gpusim_b200] gsb_db_search_batch: no multi-query kernel for this database / k; "
                                 "%d queries run as %d single-query scans (see gsb_db_batch_mode)\n", n_queries, n_queries);
    }
    for (int qi = 0; qi < n_queries; qi++) {
        rc = search_sync_locked(db, query_words + static_cast<size_t>(qi) * n_words, k, cutoff,
                                out_rows + static_cast<size_t>(qi) * k, out_scores + static_cast<size_t>(qi) * k, out_n + qi,
                                out_approx ? out_approx + qi : nullptr);
        if (rc)
            return rc;
    }
    return GSB_OK;
    GSB_CATCH
}

int gsb_db_batch_mode(const gsb_db* db, uint32_t k, int n_queries, float cutoff, int* mode, uint32_t* queries_per_pass)
{
    if (!db || !mode)
        return fail(GSB_ERR_INVALID, "null argument");
    if (!db->uploaded)
        return fail(GSB_ERR_STATE, "the device layout is chosen at upload: call gsb_db_upload first");
    const uint64_t k_scan = db->fold_factor == 1 ? k : scan_candidates(db, k);
    const BatchKernel which =
        k_scan <= gsb::kMaxBatchK ? batch_kernel_choice(db, static_cast<uint32_t>(k_scan), n_queries, cutoff) : kBatchNone;
    *mode = which == kBatchSliced   ? GSB_BATCH_SLICED
            : which == kBatchTensor ? GSB_BATCH_TENSOR
            : which == kBatchPopc   ? GSB_BATCH_POPC
                                    : GSB_BATCH_LOOPED;
    if (queries_per_pass)
        *queries_per_pass = which == kBatchNone ? 1u : batch_max_queries(which);
    return GSB_OK;
}

int gsb_db_search_cpu(const gsb_db* db, const int32_t* query_words, int n_words, uint32_t k, uint32_t* out_rows,
                      float* out_scores, uint32_t* out_n)
{
    if (!db || !query_words || !out_n || (k > 0 && (!out_rows || !out_scores)))
        return fail(GSB_ERR_INVALID, "null argument");
    if (n_words != static_cast<int>(db->words))
        return fail(GSB_ERR_INVALID, "query width does not match the database");
    if (db->host.empty() && db->count > 0)
        return fail(GSB_ERR_STATE, "search_cpu needs host-resident rows");
    const uint32_t* q = reinterpret_cast<const uint32_t*>(query_words);
    std::vector<float> scores(db->count);
    const uint32_t words = db->words;
    float* sp = scores.data();
    for (const gsb_db::HostChunk& hc : db->host) { // every chunk (the reference only looks at the first)
        const uint32_t* rows = reinterpret_cast<const uint32_t*>(hc.bytes.data());
        float* out = sp + hc.row0;
        parallel_for(hc.n_rows, [=](uint64_t lo, uint64_t hi) {
            for (uint64_t r = lo; r < hi; r++)
                out[r] = score_cpu(db, q, rows + r * words, words);
        });
    }
    const uint32_t n = static_cast<uint32_t>(std::min<uint64_t>(k, db->count));
    std::vector<uint32_t> idx(db->count);
    for (uint64_t i = 0; i < db->count; i++)
        idx[i] = static_cast<uint32_t>(i);
    // first k of the stable descending order == the reference's partial bubble sort (NaN last)
    auto better = [sp](uint32_t a, uint32_t b) {
        const float x = sp[a], y = sp[b];
        const bool xn = x != x, yn = y != y;
        if (xn || yn)
            return !xn && yn ? true : (xn && !yn ? false : a < b);
        return x > y || (x == y && a < b);
    };
    std::partial_sort(idx.begin(), idx.begin() + n, idx.end(), better);
    for (uint32_t i = 0; i < n; i++) {
        out_rows[i] = idx[i];
        out_scores[i] = sp[idx[i]];
    }
    *out_n = n;
    return GSB_OK;
}

int gsb_db_search_enqueue(const gsb_db* db, void* stream, const int32_t* h_query, const int32_t* d_query,
                          uint32_t flags, uint32_t k, float cutoff, const gsb_exchange* xchg, const gsb_sink* sink)
{
    GSB_TRY
    if (!db || (!h_query && !d_query) || !sink || !sink->n || !sink->approx || k == 0)
        return fail(GSB_ERR_INVALID, "null argument");
    if (xchg) {
        if (!sink->rows || !sink->scores)
            return fail(GSB_ERR_INVALID, "a fused search needs sink rows and scores");
        if (xchg->world < 2 || xchg->world > gsb::kMaxRanks || xchg->rank >= xchg->world || xchg->seq == 0)
            return fail(GSB_ERR_INVALID, "bad exchange descriptor");
    } else if (!sink->keys) {
        return fail(GSB_ERR_INVALID, "a shard-local search needs sink keys");
    }
    if (!db->uploaded || db->shards.size() != 1)
        return fail(GSB_ERR_STATE, "device search needs exactly one uploaded shard in this process");
    if (db->fold_factor != 1)
        return fail(GSB_ERR_STATE, "device search does not re-score folded databases");
    std::lock_guard<std::mutex> lock(db->mu); // launch planning and workspace growth are not re-entrant
    Shard& sh = const_cast<Shard&>(db->shards[0]);
    Plan plan;
    int rc = make_plan(db->layout, sh, k, &plan);
    if (rc)
        return rc;
    rc = ws_reserve(sh, k, plan.grid, -1, plan.cap);
    if (rc)
        return rc;
    GSB_CUDA(cudaSetDevice(sh.device));
    gsb::ScanParams p;
    fill_params(db, sh, plan, k, cutoff, &p);
    if (h_query) { // travels as a kernel parameter
        std::vector<uint32_t> q;
        to_dev_words(db, h_query, &q);
        std::memcpy(p.q_host, q.data(), db->layout.dev_words * 4);
    } else {
        p.q_dev = reinterpret_cast<const uint32_t*>(d_query);
        p.early_wait = (flags & GSB_QUERY_STABLE) ? 0u : 1u;
    }
    p.out_n = sink->n;
    p.out_survivors = reinterpret_cast<unsigned long long*>(sink->approx);
    p.out_done = reinterpret_cast<unsigned long long*>(sink->done);
    p.out_done_value = sink->done_value;
    if (xchg) {
        p.out_rows = sink->rows;
        p.out_scores = sink->scores;
        p.x_world = xchg->world;
        p.x_rank = xchg->rank;
        p.x_seq = xchg->seq;
        for (uint32_t r = 0; r < xchg->world; r++)
            p.x_peer[r] = xchg->peer_base[r];
    } else {
        p.out_keys = reinterpret_cast<unsigned long long*>(sink->keys);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = gate_launch(sh.device, sh.ws, st, plan.grid);
    if (rc)
        return rc;
    return launch_scan(db->layout, p, plan, st);
    GSB_CATCH
}

int gsb_db_search_device(const gsb_db* db, void* stream, const int32_t* d_query, uint32_t k, float cutoff,
                         gsb_key* d_out_keys, uint32_t* d_out_n, uint64_t* d_out_survivors)
{
    if (!d_query || !d_out_keys || !d_out_n || !d_out_survivors)
        return fail(GSB_ERR_INVALID, "null argument");
    gsb_sink sink;
    std::memset(&sink, 0, sizeof(sink));
    sink.keys = d_out_keys;
    sink.n = d_out_n;
    sink.approx = d_out_survivors;
    return gsb_db_search_enqueue(db, stream, nullptr, d_query, 0, k, cutoff, nullptr, &sink);
}

int gsb_wait_word(const uint64_t* word, uint64_t value, uint64_t timeout_us)
{
    if (!word)
        return fail(GSB_ERR_INVALID, "null argument");
    const volatile uint64_t* w = word;
    const auto t0 = std::chrono::steady_clock::now();
    for (uint64_t spins = 1; *w != value; spins++) {
        GSB_CPU_RELAX();
        if ((spins & 0xfff) == 0 && timeout_us &&
            static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::microseconds>(
                                      std::chrono::steady_clock::now() - t0).count()) > timeout_us)
            return fail(GSB_ERR_CUDA, "timed out waiting for a completion word");
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    return GSB_OK;
}

int gsb_db_set_metric(gsb_db* db, int metric, float alpha, float beta)
{
    if (!db)
        return fail(GSB_ERR_INVALID, "null database");
    if (metric != GSB_METRIC_TANIMOTO && metric != GSB_METRIC_DICE && metric != GSB_METRIC_TVERSKY)
        return fail(GSB_ERR_INVALID, "unknown metric");
    if (metric == GSB_METRIC_TVERSKY && !(alpha >= 0.0f && beta >= 0.0f))
        return fail(GSB_ERR_INVALID, "Tversky weights must be >= 0");
    std::lock_guard<std::mutex> lock(db->mu);
    db->metric = static_cast<uint32_t>(metric);
    db->alpha = metric == GSB_METRIC_TVERSKY ? alpha : 1.0f;
    db->beta = metric == GSB_METRIC_TVERSKY ? beta : 1.0f;
    return GSB_OK;
}

int gsb_db_search_batch_device(const gsb_db* db, void* stream, const int32_t* d_queries, int n_queries, uint32_t k,
                               float cutoff, gsb_key* d_out_keys, uint32_t* d_out_n, uint64_t* d_out_survivors)
{
    if (!db || !d_queries || !d_out_keys || !d_out_n || !d_out_survivors)
        return fail(GSB_ERR_INVALID, "null argument");
    if (!db->uploaded || db->shards.size() != 1)
        return fail(GSB_ERR_STATE, "device search needs exactly one uploaded shard in this process");
    if (!db->layout.rowpop || db->layout.dev_words > 32 || db->layout.dev_words != db->words || db->fold_factor != 1 ||
        k < 1 || k > gsb::kMaxBatchK)
        return fail(GSB_ERR_INVALID, "the multi-query kernels need the default layout, <= 1024 bits, k <= 512");
    BatchKernel which = batch_kernel_choice(db, k, std::max(n_queries, 2), cutoff);
    if (which == kBatchNone)
        which = kBatchPopc; // GSB_BATCH_KERNEL=0 only switches the host-buffer API to looping
    if (n_queries < 1 || n_queries > static_cast<int>(batch_max_queries(which)))
        return fail(GSB_ERR_INVALID, "1..256 queries per call (1..1024 where the bit-sliced kernel applies: "
                                     "6 or more queries)");
    std::lock_guard<std::mutex> lock(db->mu);
    Shard& sh = const_cast<Shard&>(db->shards[0]);
    return batch_launch_shard(db, sh, static_cast<cudaStream_t>(stream), which,
                              reinterpret_cast<const uint32_t*>(d_queries), static_cast<uint32_t>(n_queries), k, cutoff,
                              reinterpret_cast<unsigned long long*>(d_out_keys), d_out_n,
                              reinterpret_cast<unsigned long long*>(d_out_survivors));
}

int gsb_db_batch_max_queries(const gsb_db* db, uint32_t k, int n_queries, float cutoff, uint32_t* out_max)
{
    if (!db || !out_max)
        return fail(GSB_ERR_INVALID, "null argument");
    if (!db->uploaded)
        return fail(GSB_ERR_STATE, "the device layout is chosen at upload: call gsb_db_upload first");
    BatchKernel which = batch_kernel_choice(db, k, std::max(n_queries, 2), cutoff);
    *out_max = batch_max_queries(which == kBatchNone ? kBatchPopc : which);
    return GSB_OK;
}

int gsb_merge_batch_device(int device, void* stream, const gsb_key* d_records, int n_ranks, int n_queries, uint32_t k,
                           uint32_t* d_out_rows, float* d_out_scores, uint32_t* d_out_n, uint64_t* d_out_approx)
{
    if (!d_records || !d_out_rows || !d_out_scores || !d_out_n || !d_out_approx || n_ranks <= 0 || n_queries <= 0 || k == 0)
        return fail(GSB_ERR_INVALID, "null argument");
    GSB_CUDA(cudaSetDevice(device));
    const uint32_t cap = std::max<uint32_t>(4096, pow2ceil(static_cast<uint64_t>(k) + 32ull * n_ranks));
    const size_t smem = static_cast<size_t>(cap) * 8 + gsb::kBuckets * 4;
    GSB_CUDA(cudaFuncSetAttribute(gsb::merge_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    gsb::merge_batch_kernel<<<n_queries, gsb::kMergeThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const unsigned long long*>(d_records), static_cast<uint32_t>(n_ranks),
        static_cast<uint32_t>(n_queries), k, cap, d_out_rows, d_out_scores, d_out_n,
        reinterpret_cast<unsigned long long*>(d_out_approx));
    g_launches++;
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

int gsb_exchange_bytes(uint32_t world, uint32_t k, uint64_t* bytes)
{
    if (!bytes || world == 0 || world > gsb::kMaxRanks || k == 0)
        return fail(GSB_ERR_INVALID, "bad exchange geometry");
    *bytes = (gsb::xchg_flag_offset(1, world, world - 1, k) + 8 + 255) / 256 * 256;
    return GSB_OK;
}

int gsb_db_search_device_fused(const gsb_db* db, void* stream, const int32_t* d_query, uint32_t k, float cutoff,
                               const gsb_exchange* xchg, uint32_t* d_out_rows, float* d_out_scores,
                               uint32_t* d_out_n, uint64_t* d_out_approx)
{
    if (!d_query || !xchg || !d_out_rows || !d_out_scores || !d_out_n || !d_out_approx)
        return fail(GSB_ERR_INVALID, "null argument");
    gsb_sink sink;
    std::memset(&sink, 0, sizeof(sink));
    sink.rows = d_out_rows;
    sink.scores = d_out_scores;
    sink.n = d_out_n;
    sink.approx = d_out_approx;
    return gsb_db_search_enqueue(db, stream, nullptr, d_query, 0, k, cutoff, xchg, &sink);
}

int gsb_merge_device(int device, void* stream, const gsb_key* d_keys, const uint32_t* d_counts, int n_lists,
                     uint32_t list_stride, uint32_t k, uint32_t* d_out_rows, float* d_out_scores, uint32_t* d_out_n)
{
    if (!d_keys || !d_out_rows || !d_out_scores || !d_out_n || n_lists <= 0 || k == 0)
        return fail(GSB_ERR_INVALID, "null argument");
    GSB_CUDA(cudaSetDevice(device));
    const uint32_t cap = std::max<uint32_t>(4096, pow2ceil(static_cast<uint64_t>(k) + 32ull * n_lists));
    const size_t smem = static_cast<size_t>(cap) * 8 + gsb::kBuckets * 4;
    int smem_max = 0;
    int rc = smem_limit(device, &smem_max);
    if (rc)
        return rc;
    if (smem + 1024 > static_cast<size_t>(smem_max))
        return fail(GSB_ERR_INVALID, "k too large for the merge kernel");
    GSB_CUDA(cudaFuncSetAttribute(gsb::merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    gsb::merge_kernel<<<1, gsb::kMergeThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const unsigned long long*>(d_keys), d_counts, static_cast<uint32_t>(n_lists), list_stride,
        k, cap, d_out_rows, d_out_scores, d_out_n);
    g_launches++;
    GSB_CUDA(cudaGetLastError());
    return GSB_OK;
}

int gsb_fold_fingerprint(const int32_t* words, int n_words, int factor, int32_t* out_words)
{
    if (!words || !out_words || n_words <= 0 || factor <= 0 || n_words % factor != 0)
        return fail(GSB_ERR_INVALID, "fold factor must divide the word count");
    fold_row(reinterpret_cast<const uint32_t*>(words), n_words, factor, reinterpret_cast<uint32_t*>(out_words));
    return GSB_OK;
}

int gsb_selftest_division(int device, uint64_t* mismatches)
{
    if (!mismatches)
        return fail(GSB_ERR_INVALID, "null argument");
    GSB_CUDA(cudaSetDevice(device));
    unsigned long long* d = nullptr;
    GSB_CUDA(cudaMalloc(&d, 8));
    GSB_CUDA(cudaMemset(d, 0, 8));
    gsb::selftest_division_kernel<<<1184, 256>>>(8192, d);
    g_launches++;
    unsigned long long h = 0;
    GSB_CUDA(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
    cudaFree(d);
    *mismatches = h;
    return GSB_OK;
}

int gsb_db_scan_info(const gsb_db* db, int shard, uint32_t k, gsb_scan_info* out)
{
    if (!db || !out || shard < 0 || shard >= static_cast<int>(db->shards.size()))
        return fail(GSB_ERR_INVALID, "bad shard");
    const Shard& sh = db->shards[shard];
    Plan plan;
    int rc = make_plan(db->layout, sh, k, &plan);
    if (rc)
        return rc;
    out->device = sh.device;
    out->grid = plan.grid;
    out->block = plan.warps * 32;
    out->stages = plan.stages;
    out->tile_rows = db->layout.tile_rows * db->layout.unit_batches;
    out->tile_bytes = db->layout.unit_bytes;
    out->smem_bytes = plan.smem;
    out->cand_capacity = plan.cap;
    out->shard_rows = sh.n_rows;
    out->db_bytes_per_query = static_cast<uint64_t>(sh.n_tiles) * db->layout.tile_bytes;
    return GSB_OK;
}

} // extern "C"

namespace
{
int fetch_rows(const gsb_db* db, const std::vector<uint64_t>& rows, std::vector<uint32_t>* out)
{
    // rows are global ids; find the shard of each (device-generated databases have one shard)
    const Layout& l = db->layout;
    out->assign(rows.size() * db->words, 0u);
    for (const Shard& sh : db->shards) {
        std::vector<uint64_t> local;
        std::vector<size_t> where;
        for (size_t i = 0; i < rows.size(); i++) {
            if (rows[i] >= sh.row_base && rows[i] < sh.row_base + sh.n_rows) {
                local.push_back(rows[i] - sh.row_base);
                where.push_back(i);
            }
        }
        if (local.empty())
            continue;
        GSB_CUDA(cudaSetDevice(sh.device));
        uint64_t* d_rows = nullptr;
        uint32_t* d_out = nullptr;
        GSB_CUDA(cudaMalloc(&d_rows, local.size() * 8));
        GSB_CUDA(cudaMalloc(&d_out, local.size() * l.dev_words * 4));
        GSB_CUDA(cudaMemcpy(d_rows, local.data(), local.size() * 8, cudaMemcpyHostToDevice));
        const uint32_t n = static_cast<uint32_t>(local.size());
        gsb::gather_rows_kernel<<<(n * l.dev_words + 255) / 256, 256>>>(sh.tiles, l.tile_rows, l.tile_stride,
                                                                        l.dev_words, d_rows, n, d_out);
        g_launches++;
        std::vector<uint32_t> tmp(local.size() * l.dev_words);
        GSB_CUDA(cudaMemcpy(tmp.data(), d_out, tmp.size() * 4, cudaMemcpyDeviceToHost));
        cudaFree(d_rows);
        cudaFree(d_out);
        for (size_t i = 0; i < local.size(); i++)
            std::memcpy(out->data() + where[i] * db->words, tmp.data() + i * l.dev_words, db->words * 4);
    }
    return GSB_OK;
}
} // namespace
