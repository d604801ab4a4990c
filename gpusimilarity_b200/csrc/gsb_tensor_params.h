// Parameters, sizes and the launcher of the tensor-core multi-query kernel (gsb_tensor.cuh): what the
// host side (gsb_api.cu) needs; the kernel itself is compiled in its own translation unit.
#pragma once

#include "gsb_sliced.cuh"
#include "gsb_tensor_math.h"

namespace gsb
{

constexpr uint32_t kTcSlabRing = 4;   // expanded slabs in flight (one tile)
constexpr uint32_t kTcRawRing = 3;    // raw tiles in flight
constexpr uint32_t kTcPdRing = 8;     // tiles whose row popcounts are kept: the expanders run up to 4 tiles ahead of the epilogue
constexpr uint32_t kTcRawBatch = kBatchRows * 128u + kBatchRows * 2u; // 4160: 32 rows + u16 popcounts
constexpr uint32_t kTcRawStage = kTcTileBatches * kTcRawBatch;        // 16640
constexpr int kTcWarps = 20, kTcThreads = kTcWarps * 32;
constexpr int kTcExpWarp0 = 4, kTcExpWarps = 8, kTcEpiWarp0 = 12, kTcEpiWarps = 8, kTcEpiThreads = kTcEpiWarps * 32;
constexpr uint32_t kTcTmemCols = 512, kTcTmemD = 256; // columns [0,256): queries, [256,512): two accumulator tiles
constexpr uint32_t kTcPruneMin = 64, kTcWarpSortMax = 64;
constexpr uint32_t kTcWarmCtas = 16;  // CTAs whose first tile feeds the grid-wide histograms (16 x 128 >= kMaxBatchK candidates)

__host__ __device__ constexpr uint32_t tc_smem_bytes()
{
    return kTcSlabRing * kTcSlabBytes + kTcRawRing * kTcRawStage + kBatchListCap * 8u + kBuckets * 4u +
           kTcPdRing * kTcTileRows * 4u + kTcPdRing * 8u * 4u + kTcQueries * (8u + 8u + 4u + 4u);
}

struct TensorParams {
    BatchParams b;            // database, k, cutoff, nq (<= 128), candidate lists, outputs
    unsigned int* ghist;      // [nq][kSlicedHistBuckets] scores of all candidates so far; zero on entry
    unsigned long long* gtau; // [nq] thresholds shared by all CTAs; zero on entry
    uint32_t n_tiles;         // 128-row tiles of the shard
    uint32_t fault;           // test hook (GSB_TC_FAULT=1): CTA 0 never loads its first tile -> pipeline timeout;
                              // 2, 3, 5, 8: timing experiments with void results (profiles/r02_tensor.md)
    uint32_t variant;         // GSB_TC_VARIANT=1: every CTA's first tile feeds the histograms (timing experiment)
    unsigned long long* dbg;  // GSB_TC_DEBUG=1: [grid][20 warps][8] clocks spent in each pipeline wait, role time
};

// The kernel (gsb_tensor.cuh) is compiled in its own translation unit, gsb_tensor.cu; everybody
// else sees the parameters above and this launcher (cooperative launch: all CTAs resident).
cudaError_t tensor_kernel_launch(const TensorParams& tp, int grid, cudaStream_t st);

} // namespace gsb
