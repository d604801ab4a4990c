// Operand layouts of the tensor-core multi-query scan (gsb_tensor.cuh).  Host and device:
// tests/cpp/test_tensor_math.cpp builds both operands with these helpers on the CPU, reads them
// back the way tcgen05.mma does (canonical K-major, no swizzle) and checks the dot products
// against plain popcounts.
//
// common = popc(q & d) (reference TanimotoFunctor, fingerprintdb_cuda.cu:97-98) is the dot product
// of two {0,1} vectors of 1024 entries.  The tensor cores take 8-bit integers, so every bit becomes
// a byte — but the byte need not be 0/1: bit `plane` (0..7) of a byte of the row word becomes the
// byte value 2^plane on the row side (ONE AND with 0x01010101 << plane per four bits, no shifts,
// no multiplies) and 2^(7 - plane) on the query side, so that every common bit contributes 2^7 and
//     D = sum_k A_k * B_k = 128 * popc(q & d)            (exact: u8 x u8 -> s32)
// The order of the 1024 k-positions is free as long as both operands agree; it is chosen so that
// the row side is written with conflict-free 8-byte stores:
//   * quarter-K slab g (0..3) holds bit planes 2g and 2g+1 of all 32 row words;
//   * K-step s (0..7) of a slab = 32 bytes = two 16-byte chunks c = 2s, 2s+1; chunk c holds, for
//     row words iw = 2c and 2c+1: [iw=2c plane 2g][iw=2c plane 2g+1][iw=2c+1 plane 2g][iw=2c+1 plane 2g+1]
//     (four bytes each: byte b of such a word is bit 8b + plane of the row word).
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define GSB_TC_HD __host__ __device__ __forceinline__
#else
#define GSB_TC_HD inline
#endif

namespace gsb
{

constexpr uint32_t kTcQueries = 128;     // queries per pass: M of the MMA = lanes of tensor memory
constexpr uint32_t kTcTileRows = 128;    // rows per tile: N of the MMA
constexpr uint32_t kTcTileBatches = 4;   // 32-row batches per tile
constexpr uint32_t kTcSlabs = 4;         // quarter-K slabs per tile
constexpr uint32_t kTcSlabSteps = 8;     // K-steps (32 bytes of K) per slab
constexpr uint32_t kTcSlabChunks = 16;   // 16-byte chunks per row and slab
constexpr uint32_t kTcSbo = 128;         // bytes between 8-row groups of one chunk (core matrices are contiguous)
constexpr uint32_t kTcLbo = (kTcTileRows / 8) * kTcSbo + 16; // bytes between chunks: 16 row groups + 16 bytes of bank skew
constexpr uint32_t kTcSlabBytes = kTcSlabChunks * kTcLbo;    // 33024
static_assert(kTcSlabBytes % 128 == 0, "slabs start on 128-byte lines");

// row side: the four bits {plane, plane+8, plane+16, plane+24} of a row word as bytes of value 2^plane
GSB_TC_HD uint32_t tc_row_word(uint32_t w, uint32_t plane)
{
    return w & (0x01010101u << plane);
}
// query side: the same four bits as bytes of value 2^(7 - plane)
GSB_TC_HD uint32_t tc_query_word(uint32_t w, uint32_t plane)
{
    return ((w >> plane) & 0x01010101u) << (7u - plane);
}
// Byte offset inside a slab of the 8 bytes {plane 2g, plane 2g+1} of word iw of tile row r.
GSB_TC_HD uint32_t tc_slab_offset(uint32_t r, uint32_t iw)
{
    return (iw >> 1) * kTcLbo + (r >> 3) * kTcSbo + (r & 7u) * 16u + (iw & 1u) * 8u;
}
// A operand (queries, tensor memory): 32-bit column u (0..7) of K-step ks (0..31) of a query's lane
// is tc_query_word(query word *iw, *plane).
GSB_TC_HD void tc_a_source(uint32_t ks, uint32_t u, uint32_t* iw, uint32_t* plane)
{
    const uint32_t g = ks >> 3, s = ks & 7u, c = 2u * s + (u >> 2);
    *iw = 2u * c + ((u & 3u) >> 1);
    *plane = 2u * g + (u & 1u);
}

// The filter of the epilogue.  D = 128 * common arrives as s32; with the magic exponent 2^23 it is
// a float without a conversion: as_float(0x4B000000 | D) = 2^23 + D.  A row can only reach the
// score ts if common >= tq * (pq + pd) with tq = ts / (1 + ts) rounded down (sliced_tq), i.e.
//     (2^23 + D) - 128 tq * pd  >=  2^23 + 128 tq * pq
// The left side is one FMA (rounded to an integer: error <= 0.5), the right side a per-query
// constant lowered by 2 (the two roundings and the 0.001 slack of sliced_lane_min), so the test
// never rejects a row the exact comparison would accept.
constexpr uint32_t kTcMagicBits = 0x4B000000u;
GSB_TC_HD float tc_filter_slope(float tq)
{
    return -128.0f * tq;
}
GSB_TC_HD float tc_filter_threshold(float tq, uint32_t pq)
{
    const float x = 128.0f * tq * static_cast<float>(pq) - 2.0f;
    const float fl = static_cast<float>(static_cast<int32_t>(x)); // toward zero: >= floor for x < 0 only by < 1 ...
    return 8388608.0f + (fl > x ? fl - 1.0f : fl);                // ... so make it a true floor
}

} // namespace gsb
